#!/bin/bash
# Round 2, tenth GPU pass (1 GPU): feeder thread again (pinned buffers sized by the calling thread), every step under a short timeout.
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
HULK_B200_FEED_DEBUG=1 timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "feeder" > gpurun_out/pytest_feed.log 2>&1; rc=$?; echo "pytest feeder rc=$rc"; tail -25 gpurun_out/pytest_feed.log | cut -c1-200
if [ $rc -ne 0 ]; then export HULK_B200_FEEDER=0; echo "FEEDER DISABLED for the rest of this pass"; fi
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "packed or async_input or intervals or device_resident" > gpurun_out/pytest_a.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_a.log
for pf in 2048 4096; do echo "prefetch distance $pf"; HULK_B200_PACK_PREFETCH=$pf timeout 120 python tools/probe_pack.py 2>&1 | grep -E "^( 1|16|15|12) threads"; done > gpurun_out/r02h_pack_probe.txt 2>&1; cat gpurun_out/r02h_pack_probe.txt
B="python bench.py --steps 100 --warmup 3 --no-cpu-baseline"
run() { tag=$1; shift; env "$@" timeout 150 $B > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err; echo "rc=$?"; grep -h "\[feed\]" gpurun_out/bench_$tag.err | tail -1; tail -1 gpurun_out/bench_$tag.err | cut -c1-200; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$tag.log").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("%-10s"%"$tag", "ms/step %.4f"%d["ms_per_step"], "e2e %.4f (%.0f M/s, pack %.4f)"%(d["e2e"]["ms_per_step"], d["e2e"]["value"]/1e6, d["e2e"]["host_pack_ms_per_step"]), "ascii %.4f"%d["e2e_ascii"]["ms_per_step"], "enqueue", {k:round(v,4) for k,v in d["host_enqueue_ms_per_step"].items()})
except Exception as e:
    print("$tag", "no line", e)
PY
}
export HULK_B200_FEED_STATS=1
run auto X=1
run auto2 X=1
run t15 HULK_B200_PACK_THREADS=15
run t12 HULK_B200_PACK_THREADS=12
run t10 HULK_B200_PACK_THREADS=10
run t8 HULK_B200_PACK_THREADS=8
run nofeed HULK_B200_FEEDER=0
