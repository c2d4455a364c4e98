#!/bin/bash
# Runs on the GPU box (via gpurun): build check, smoke, GPU parity tests, a short bench.  Logs to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
timeout ${PYTEST_TIMEOUT:-1200} python -m pytest tests -m gpu -q ${PYTEST_ARGS:--x} > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -${PYTEST_TAIL:-40} gpurun_out/pytest.log
timeout 900 python bench.py --steps ${BENCH_STEPS:-20} --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
