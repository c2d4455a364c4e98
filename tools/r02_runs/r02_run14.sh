#!/bin/bash
# Round 2, fourteenth GPU pass (1 GPU): flush decision folded into the bitmap kernel (parity), launch list + full capture (feeder off under ncu).
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_distributed.py -m gpu -q -x -k "not chromosome and not both_screens and not c3_shape" > gpurun_out/pytest_b.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_b.log
timeout 150 python bench.py --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/bench_k2.log 2> gpurun_out/bench_k2.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_k2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_k2.log").read().strip().splitlines()[-1])
r=d["roofline"]
print("value %.0f M/s ms/step %.4f"%(d["value"]/1e6,d["ms_per_step"]), "serial %.4f"%r["serial_ms_per_step"], "e2e %.0f M/s %.4f"%(d["e2e"]["value"]/1e6,d["e2e"]["ms_per_step"]), {k[:9]:round(v,4) for k,v in r["kernel_ms_per_step"].items()}, d.get("parity_check"))
PY
export HULK_B200_FEEDER=0
TAG=r02f
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k[0123]_' -c 600 \
    --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.log 2>&1
echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on \
    -k regex:'k2_cms_update|k2_mask|k2_final|k3_filter|k3_resolve|k0_unpack' --launch-skip 60 -c 8 \
    -f -o gpurun_out/${TAG}_full $CMD > gpurun_out/${TAG}_full.log 2>&1
echo "full capture rc=$?"
