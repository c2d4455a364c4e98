#!/bin/bash
# Round 2, thirty-eighth GPU pass (1 GPU): the tree as it will be judged (sliced scan, MinHash feed, k1_generic slabs, lists of up
# to 192 entries) -- smoke, whole gpu suite, default bench line, reference arm.
mkdir -p gpurun_out
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 700 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest.log
timeout 200 python bench.py > gpurun_out/r02zz_bench.json 2> gpurun_out/r02zz_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r02zz_bench.err | cut -c1-200
timeout 120 python bench.py --impl reference > gpurun_out/r02zz_bench_reference.json 2> gpurun_out/r02zz_bench_reference.err; echo "reference rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02zz_bench.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("value %.0f M/s ms/step %.4f"%(d["value"]/1e6,d["ms_per_step"]), "serial %.4f"%r["serial_ms_per_step"], "e2e %.0f M/s %.4f"%(d["e2e"]["value"]/1e6,d["e2e"]["ms_per_step"]), "ascii %.4f"%d["e2e_ascii"]["ms_per_step"], {k[:9]:round(v,4) for k,v in r["kernel_ms_per_step"].items()}, d.get("parity_check"), "cpu", d["cpu_baseline"]["value"], "launches", d["gpu_launches"], "k1_issue", round(r["k1_issue"]["frac"],3), "k3 frac", round(r["k3_filter"]["frac"],3))
x=json.loads(open("gpurun_out/r02zz_bench_reference.json").read().strip().splitlines()[-1])
print("reference arm: %.0f reads/s on %d cores (%s)"%(x["value"], x["cpu_baseline"]["cores"], x["cpu_baseline"]["kind"]), "-> e2e ratio %.0f"%(d["e2e"]["value"]/x["value"]))
PY
