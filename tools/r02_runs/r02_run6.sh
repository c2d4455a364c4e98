#!/bin/bash
# Round 2, sixth GPU pass (1 GPU): state of the restored tree -- all gpu tests, the bench line, launch list, full ncu capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest.log
timeout 600 python bench.py --steps 100 --warmup 3 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02e_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02e_bench.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("ms/step %.4f"%d["ms_per_step"], "serial %.4f"%r["serial_ms_per_step"], "e2e %.4f"%d["e2e"]["ms_per_step"], {k[:9]:round(v,4) for k,v in r["kernel_ms_per_step"].items()}, d.get("parity_check"), d["clocks"])
PY
timeout 300 python tools/probe_timeline.py > gpurun_out/r02e_timeline.txt 2>&1; tail -16 gpurun_out/r02e_timeline.txt
TAG=r02e
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k[123]_' -c 400 \
    --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'k1_scan_w9_v2|k1_jump_queue|k2_cms_update|k2_mask|k2_flush|k2_final|k3_filter|k3_resolve' -s 24 -c 9 \
    -f -o gpurun_out/${TAG}_full $CMD > gpurun_out/${TAG}_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/ | tail -20
