#!/bin/bash
# Round 2, thirty-first GPU pass (1 GPU): the feeder without its statistics switch (the default path), two bench lines.
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_distributed.py -m gpu -q -x -k "feeder or packed or async_input or one_device" > gpurun_out/pytest_a.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_a.log
for i in 1 2; do
timeout 150 python bench.py --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r$i.log 2> gpurun_out/bench_r$i.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r$i.log").read().strip().splitlines()[-1])
print("value %.0f M/s %.4f"%(d["value"]/1e6,d["ms_per_step"]), "e2e %.0f M/s %.4f (pack %.4f, %d thr, h2d %.2f MB)"%(d["e2e"]["value"]/1e6,d["e2e"]["ms_per_step"],d["e2e"]["host_pack_ms_per_step"],d["e2e"]["pack_threads"],d["e2e"]["h2d_bytes_per_step"]/1e6), "ascii %.4f"%d["e2e_ascii"]["ms_per_step"])
PY
done
