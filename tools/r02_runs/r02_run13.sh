#!/bin/bash
# Round 2, thirteenth GPU pass (1 GPU): whole gpu suite after the packed transport / feeder work, the bench line, launch list, full ncu capture.
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/pytest.log
timeout 240 python bench.py --steps 100 --warmup 3 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02f_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02f_bench.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("value %.0f M/s ms/step %.4f"%(d["value"]/1e6,d["ms_per_step"]), "serial %.4f"%r["serial_ms_per_step"], "e2e %.0f M/s %.4f"%(d["e2e"]["value"]/1e6,d["e2e"]["ms_per_step"]), "ascii %.4f"%d["e2e_ascii"]["ms_per_step"], {k[:9]:round(v,4) for k,v in r["kernel_ms_per_step"].items()}, d.get("parity_check"), d["clocks"], "cpu", d["cpu_baseline"]["value"], "launches", d["gpu_launches"])
PY
TAG=r02f
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k[0123]_' -c 600 \
    --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.log 2>&1
echo "launch list rc=$?"
