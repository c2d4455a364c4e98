#!/bin/bash
# Round 2, thirty-third GPU pass (1 GPU): duplicate pre-test by the warp (MATCH) -- parity of the per-read sets and spectra, K1 time.
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "minimizer or histogram or long_reads or large_batches or no_decay or intervals or c2_shape or properties" > gpurun_out/pytest_m.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_m.log
for i in 1 2; do
timeout 150 python bench.py --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/bench_m$i.log 2> gpurun_out/bench_m$i.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_m$i.log").read().strip().splitlines()[-1])
r=d["roofline"]
print("value %.0f M/s %.4f"%(d["value"]/1e6,d["ms_per_step"]), "serial %.4f"%r["serial_ms_per_step"], "e2e %.0f M/s %.4f"%(d["e2e"]["value"]/1e6,d["e2e"]["ms_per_step"]), {k[:9]:round(v,4) for k,v in r["kernel_ms_per_step"].items()})
PY
done
export HULK_B200_FEEDER=0
timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:'k1_scan' --launch-skip 8 -c 2 \
    --csv --log-file gpurun_out/r02t_scan.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02t_scan.log 2>&1
grep -E "k1_scan" gpurun_out/r02t_scan.csv | cut -d, -f5,13- | tail -4
