#!/bin/bash
# Round 2, fifteenth GPU pass (1 GPU): jump walks with lane-private key queues -- parity and A/B.
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
HULK_B200_JUMP_V=2 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "histogram or c2_shape or properties or intervals or no_decay" > gpurun_out/pytest_j.log 2>&1; echo "pytest(v2) rc=$?"; tail -4 gpurun_out/pytest_j.log
HULK_B200_JUMP_V=2 HULK_B200_JUMP_BATCH=3 HULK_B200_JUMP_TAIL=8 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "histogram or intervals" > gpurun_out/pytest_j2.log 2>&1; echo "pytest(v2 b3 t8) rc=$?"; tail -3 gpurun_out/pytest_j2.log
B="python bench.py --steps 60 --warmup 3 --no-cpu-baseline"
run() { tag=$1; shift; env "$@" timeout 150 $B > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err; tail -1 gpurun_out/bench_$tag.err | cut -c1-200; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$tag.log").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("%-12s"%"$tag", "ms/step %.4f"%d["ms_per_step"], "serial %.4f"%r["serial_ms_per_step"], "e2e %.4f"%d["e2e"]["ms_per_step"], {k[:9]:round(v,4) for k,v in r["kernel_ms_per_step"].items()})
except Exception as e:
    print("$tag", "no line", e)
PY
}
run v1 X=1
run v2b4t15 HULK_B200_JUMP_V=2
run v2b3t15 HULK_B200_JUMP_V=2 HULK_B200_JUMP_BATCH=3
run v2b2t15 HULK_B200_JUMP_V=2 HULK_B200_JUMP_BATCH=2
run v2b4t8 HULK_B200_JUMP_V=2 HULK_B200_JUMP_TAIL=8
run v2b4t25 HULK_B200_JUMP_V=2 HULK_B200_JUMP_TAIL=25
run v2b3t8 HULK_B200_JUMP_V=2 HULK_B200_JUMP_BATCH=3 HULK_B200_JUMP_TAIL=8
run v2b3jc4 HULK_B200_JUMP_V=2 HULK_B200_JUMP_BATCH=3 HULK_B200_JUMP_CTAS=4
run v2b4jc4 HULK_B200_JUMP_V=2 HULK_B200_JUMP_CTAS=4
export HULK_B200_FEEDER=0
timeout 200 env HULK_B200_JUMP_V=2 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:'k1_jump' --launch-skip 8 -c 3 \
    --csv --log-file gpurun_out/r02g_jump_v2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02g_jump_v2.log 2>&1
grep -E "k1_jump" gpurun_out/r02g_jump_v2.csv | cut -d, -f5,13- | tail -6
