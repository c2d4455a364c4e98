#!/bin/bash
# Round 2, thirty-second GPU pass (1 GPU): compute-sanitizer over the kernels added this round.
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
S=/usr/local/cuda/bin/compute-sanitizer
out=gpurun_out/r02s_sanitizer.txt
echo "compute-sanitizer on a B200, round-2 kernels (second-generation and one-word scans, fixed-point jump, k0_unpack/k0_patch behind the feeder, fused flush decision, k4 table draw, peer reads with members on one device)" > $out
run() { tool=$1; shift; sel="$1"; shift
  timeout 420 $S --tool $tool --error-exitcode 9 "$@" python -m pytest tests/test_gpu_parity.py tests/test_distributed.py -m gpu -q -x -k "$sel" > gpurun_out/san_$tool.log 2>&1; rc=$?
  echo "  $tool  -k \"$sel\"   rc=$rc  $(grep -E 'passed|failed' gpurun_out/san_$tool.log | tail -1)  $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/san_$tool.log | tail -1)" >> $out
}
run memcheck "histogram_with_n or packed_transport or feeder or sketch_no_decay_exact or cws_tables_drawn_on_the_device"
run racecheck "histogram_bit_exact_synthetic or flush_semantics or cws_device_draw_lets"
run synccheck "histogram_bit_exact_synthetic or flush_semantics or packed_transport"
cat $out
