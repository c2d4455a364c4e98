#!/bin/bash
# Round 2, thirty-ninth GPU pass (1 GPU): compute-sanitizer over the kernels added at the end of the round (k1_long_plan / zero /
# scan, k1_generic with slabs, k1_khf_queue, k1_kmv_filter / select, k1_length_stats).
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
out=gpurun_out/r02y_sanitizer.txt
echo "compute-sanitizer on a B200, kernels added at the end of round 2 (sliced scan of long sequences, k1_generic slabs, MinHash feed, length statistics)" > $out
run() { tool=$1; shift; sel="$1"; shift
  timeout 48 $S --tool $tool --error-exitcode 9 "$@" python -m pytest tests/test_gpu_parity.py tests/test_minhash.py -m gpu -q -x -k "$sel" > gpurun_out/san2_$tool.log 2>&1; rc=$?
  echo "  $tool  -k \"$sel\"   rc=$rc  $(grep -E 'passed|failed' gpurun_out/san2_$tool.log | tail -1)  $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/san2_$tool.log | tail -1)" >> $out
}
run memcheck "(long_sequences_take_the_sliced_scan and 21-9) or fewer_minimizers or (fed_sketches_equal and 21-9-64) or fixed_length_and_device_offsets"
run racecheck "(long_sequences_take_the_sliced_scan and 11-9) or fewer_minimizers or (fed_sketches_equal and 21-9-1)"
cat $out
