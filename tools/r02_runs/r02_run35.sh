#!/bin/bash
# Round 2, thirty-fifth GPU pass (1 GPU): MinHash feed (k1_minhash.cuh) against the restatement of src/minhash, its cost at the
# C2 shape; reads of a few hundred to a few thousand bases by sliced-scan threshold; the long-sequence tests again (plan kernel
# with prefix sums).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_minhash.py tests/test_ingest_cli.py tests/test_gpu_parity.py -m gpu -q -x --durations=6 -k "minhash or fed or unfed or fewer or khf or long_sequences or chromosome or long_reads" > gpurun_out/pytest_mh.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/pytest_mh.log
timeout 100 python tools/probe_minhash.py > gpurun_out/r02v_minhash.txt 2>&1; echo "probe rc=$?"; cat gpurun_out/r02v_minhash.txt
timeout 200 python tools/probe_long_reads.py > gpurun_out/r02v_long_reads.txt 2>&1; echo "probe rc=$?"; cat gpurun_out/r02v_long_reads.txt
