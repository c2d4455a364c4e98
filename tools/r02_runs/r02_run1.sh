#!/bin/bash
# Round 2, first GPU pass: parity of the second-generation scan and the fixed-point jump step, A/B bench lines, ncu.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -s --durations=8 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest.log
B="python bench.py --steps 40 --warmup 3 --no-cpu-baseline"
run() { tag=$1; shift; env "$@" timeout 600 $B > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err; echo "bench $tag rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$tag.log").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("$tag", "ms/step %.4f"%d["ms_per_step"], "serial %.4f"%r["serial_ms_per_step"], "e2e %.4f"%d["e2e"]["ms_per_step"], {k:round(v,4) for k,v in r["kernel_ms_per_step"].items()}, d["clocks"])
except Exception as e:
    print("$tag", "no line", e)
PY
}
run old HULK_B200_K1_V2=0 HULK_B200_JUMP_FX=0
run v2 HULK_B200_K1_V2=1 HULK_B200_JUMP_FX=0
run fx HULK_B200_K1_V2=0 HULK_B200_JUMP_FX=1
run new HULK_B200_K1_V2=1 HULK_B200_JUMP_FX=1
run new_jb2 HULK_B200_JUMP_BATCH=2
run new_jb3 HULK_B200_JUMP_BATCH=3
run new_k1c5 HULK_B200_K1_CTAS=5
run new_jc4 HULK_B200_JUMP_CTAS=4
run new_jc2 HULK_B200_JUMP_CTAS=2

TAG=r02a
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k[123]_' -c 400 \
    --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'k1_scan_w9_v2|k1_jump_queue' -s 8 -c 2 \
    -f -o gpurun_out/${TAG}_full $CMD > gpurun_out/${TAG}_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/ | tail -20
