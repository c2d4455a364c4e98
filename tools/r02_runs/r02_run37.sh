#!/bin/bash
# Round 2, thirty-seventh GPU pass (1 GPU): candidate lists of up to 192 entries for the w = 9 kernels (reads of several hundred
# bases stay with the fast scan) -- stage-1 parity, reads of 300 / 600 / 2000 / 8000 bases resident in HBM by list cap and
# sliced-scan threshold.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --durations=4 -k "minimizer or histogram or long or large_batches or many_reads or device_resident" > gpurun_out/pytest_l.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_l.log
timeout 200 python tools/probe_long_reads.py > gpurun_out/r02x_long_reads.txt 2>&1; echo "probe rc=$?"; cat gpurun_out/r02x_long_reads.txt
