#!/bin/bash
# Round 2, seventeenth GPU pass (1 GPU): CWS tables drawn on the device -- parity against the host generator, wall clock.
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cws_tables_drawn or device_drawn" > gpurun_out/pytest_cws.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_cws.log | cut -c1-300
timeout 400 python tools/probe_cws_device.py > gpurun_out/r02j_cws_device.txt 2>&1; cat gpurun_out/r02j_cws_device.txt | cut -c1-300
timeout 200 python -m pytest tests/test_ingest_cli.py -m gpu -q -x -k "tables_drawn" > gpurun_out/pytest_cws_cli.log 2>&1; echo "pytest cli rc=$?"; tail -5 gpurun_out/pytest_cws_cli.log | cut -c1-300
