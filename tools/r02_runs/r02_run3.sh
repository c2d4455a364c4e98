#!/bin/bash
# Round 2, third GPU pass (1 GPU): group/peer mechanism with the members on one device, CLI --gpus, K1 after the jump rewrite.
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 900 python -m pytest tests/test_distributed.py tests/test_ingest_cli.py -m gpu -q -x --durations=5 > gpurun_out/pytest_a.log 2>&1; echo "pytest(group, cli) rc=$?"; tail -15 gpurun_out/pytest_a.log
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_smash.py -m gpu -q -x > gpurun_out/pytest_b.log 2>&1; echo "pytest(parity) rc=$?"; tail -5 gpurun_out/pytest_b.log
B="python bench.py --steps 40 --warmup 3 --no-cpu-baseline"
run() { tag=$1; shift; env "$@" timeout 600 $B > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err; echo "bench $tag rc=$?"; tail -2 gpurun_out/bench_$tag.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$tag.log").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("$tag", "ms/step %.4f"%d["ms_per_step"], "serial %.4f"%r["serial_ms_per_step"], "e2e %.4f"%d["e2e"]["ms_per_step"], {k:round(v,4) for k,v in r["kernel_ms_per_step"].items()}, d["clocks"]["sm_mhz"], d.get("parity_check"), "k3 frac %.3f"%r["k3_filter"]["frac"])
except Exception as e:
    print("$tag", "no line", e)
PY
}
run new X=1
run new_jb3 HULK_B200_JUMP_BATCH=3
run new_jb2 HULK_B200_JUMP_BATCH=2
run new_jc4 HULK_B200_JUMP_CTAS=4
