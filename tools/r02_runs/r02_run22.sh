#!/bin/bash
# Round 2, twenty-second GPU pass (1 GPU): device-side table draw -- two rounds of the raw stream, slot range, host-decided attempts.
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cws" --durations=4 > gpurun_out/pytest_cws.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_cws.log | cut -c1-300
