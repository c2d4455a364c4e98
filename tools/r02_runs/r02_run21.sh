#!/bin/bash
# Round 2, twenty-first GPU pass (1 GPU): whole gpu suite, bench line, launch list and full ncu capture after the tail-block change.
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 700 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/pytest.log
timeout 240 python bench.py --steps 100 --warmup 3 > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02n_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02n_bench.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("value %.0f M/s ms/step %.4f"%(d["value"]/1e6,d["ms_per_step"]), "serial %.4f"%r["serial_ms_per_step"], "e2e %.0f M/s %.4f"%(d["e2e"]["value"]/1e6,d["e2e"]["ms_per_step"]), "ascii %.4f"%d["e2e_ascii"]["ms_per_step"], {k[:9]:round(v,4) for k,v in r["kernel_ms_per_step"].items()}, d.get("parity_check"), "cpu", d["cpu_baseline"]["value"], "launches", d["gpu_launches"])
PY
export HULK_B200_FEEDER=0
TAG=r02n
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k[0123]_' -c 600 \
    --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.log 2>&1
echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on \
    -k regex:'k1_scan_w9_v2|k1_jump_queue' --launch-skip 8 -c 2 \
    -f -o gpurun_out/${TAG}_full $CMD > gpurun_out/${TAG}_full.log 2>&1
echo "full capture rc=$?"
