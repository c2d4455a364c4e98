#!/bin/bash
# Round 2, thirty-sixth GPU pass (1 GPU): k1_generic with a table per thread and the scan's size hint -- parity of every
# stage-1 test, 47 000 reads of 400-1200 bases in one batch, reads of 600 / 2000 / 8000 bases by sliced-scan threshold,
# the C2 bench line.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_minhash.py -m gpu -q -x --durations=6 -k "minimizer or histogram or long or large_batches or many_reads or chromosome or fed_sketches_equal" > gpurun_out/pytest_g.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_g.log
timeout 200 python tools/probe_long_reads.py > gpurun_out/r02w_long_reads.txt 2>&1; echo "probe rc=$?"; cat gpurun_out/r02w_long_reads.txt
timeout 150 python bench.py --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/bench_w.log 2> gpurun_out/bench_w.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_w.log").read().strip().splitlines()[-1])
r=d["roofline"]
print("value %.0f M/s %.4f"%(d["value"]/1e6,d["ms_per_step"]), "serial %.4f"%r["serial_ms_per_step"], "e2e %.0f M/s %.4f"%(d["e2e"]["value"]/1e6,d["e2e"]["ms_per_step"]), {k[:9]:round(v,4) for k,v in r["kernel_ms_per_step"].items()})
PY
