#!/bin/bash
# Round 2, thirty-fourth GPU pass (1 GPU): sliced scan of long sequences (k1_long.cuh) -- parity of the per-sequence sets and
# spectra against the oracle, time of a 20 Mbp sequence with and without it, and the C2 bench line (the short-read kernels
# only gained a length test).
mkdir -p gpurun_out
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --durations=8 -k "long_sequences or chromosome or long_reads or large_batches" > gpurun_out/pytest_long.log 2>&1; echo "pytest rc=$?"; tail -16 gpurun_out/pytest_long.log
timeout 120 python tools/probe_long.py > gpurun_out/r02u_long.txt 2>&1; echo "probe rc=$?"; cat gpurun_out/r02u_long.txt
timeout 150 python bench.py --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/bench_u.log 2> gpurun_out/bench_u.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_u.log").read().strip().splitlines()[-1])
r=d["roofline"]
print("value %.0f M/s %.4f"%(d["value"]/1e6,d["ms_per_step"]), "serial %.4f"%r["serial_ms_per_step"], "e2e %.0f M/s %.4f"%(d["e2e"]["value"]/1e6,d["e2e"]["ms_per_step"]), {k[:9]:round(v,4) for k,v in r["kernel_ms_per_step"].items()})
PY
