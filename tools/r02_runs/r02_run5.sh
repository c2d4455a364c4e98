#!/bin/bash
# Round 2, fifth GPU pass (1 GPU): short-lived scan CTAs / small K2 CTAs -- does the flush chain find room now?
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest.log
timeout 300 python tools/probe_timeline.py > gpurun_out/r02d_timeline.txt 2>&1; tail -16 gpurun_out/r02d_timeline.txt
B="python bench.py --steps 60 --warmup 3 --no-cpu-baseline"
run() { tag=$1; shift; env "$@" timeout 600 $B > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err; tail -2 gpurun_out/bench_$tag.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$tag.log").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("%-14s"%"$tag", "ms/step %.4f"%d["ms_per_step"], "serial %.4f"%r["serial_ms_per_step"], "e2e %.4f"%d["e2e"]["ms_per_step"], {k[:9]:round(v,4) for k,v in r["kernel_ms_per_step"].items()}, d.get("parity_check"))
except Exception as e:
    print("$tag", "no line", e)
PY
}
run static X=1
run persist HULK_B200_K1_PERSISTENT=1
run persist3 HULK_B200_K1_PERSISTENT=1 HULK_B200_K1_CTAS=3
run st_k3c1s8 HULK_B200_K3_CTAS=1 HULK_B200_K3_STAGES=8
run st_jc2 HULK_B200_JUMP_CTAS=2
run st_jc4 HULK_B200_JUMP_CTAS=4
run st_nb3 HULK_B200_NBUF=3
run st_smem HULK_B200_JUMP_SMEM=1
