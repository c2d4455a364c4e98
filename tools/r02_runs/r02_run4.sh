#!/bin/bash
# Round 2, fourth GPU pass (1 GPU): group tests in their own process, timeline of the pipelined step, occupancy sweep.
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 900 python -m pytest tests/test_distributed.py tests/test_ingest_cli.py -m gpu -q --durations=5 > gpurun_out/pytest_a.log 2>&1; echo "pytest(group, cli) rc=$?"; tail -12 gpurun_out/pytest_a.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "minimizer or histogram or jump or c2_shape or properties" > gpurun_out/pytest_b.log 2>&1; echo "pytest(parity subset) rc=$?"; tail -3 gpurun_out/pytest_b.log
timeout 300 python tools/probe_timeline.py > gpurun_out/r02c_timeline.txt 2>&1; tail -16 gpurun_out/r02c_timeline.txt
B="python bench.py --steps 40 --warmup 3 --no-cpu-baseline"
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
run base X=1
run smem HULK_B200_JUMP_SMEM=1
run k3c1s8 HULK_B200_K3_CTAS=1 HULK_B200_K3_STAGES=8
run k3c1s4 HULK_B200_K3_CTAS=1 HULK_B200_K3_STAGES=4
run k3c2s8 HULK_B200_K3_CTAS=2 HULK_B200_K3_STAGES=8
run k1c5 HULK_B200_K1_CTAS=5
run k1c3 HULK_B200_K1_CTAS=3
run k1c5k3c1 HULK_B200_K1_CTAS=5 HULK_B200_K3_CTAS=1 HULK_B200_K3_STAGES=8
run jc2 HULK_B200_JUMP_CTAS=2
run jc4k3c1 HULK_B200_JUMP_CTAS=4 HULK_B200_K3_CTAS=1 HULK_B200_K3_STAGES=8
run nb2 HULK_B200_NBUF=2
run nb3 HULK_B200_NBUF=3
run fp32 HULK_B200_K3_FP32=1
