#!/bin/bash
# Round 2, twelfth GPU pass (1 GPU): where the calling thread's time goes in the packed e2e step.
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
export HULK_B200_FEED_STATS=1
for t in 15 12; do
HULK_B200_PACK_THREADS=$t timeout 150 python bench.py --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s$t.log 2> gpurun_out/bench_s$t.err; echo "threads $t rc=$?"; grep -h "^\[host\]\|^\[feed\]" gpurun_out/bench_s$t.err | tail -14
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_s$t.log").read().strip().splitlines()[-1])
print("ms/step %.4f"%d["ms_per_step"], "e2e %.4f (%.0f M/s, pack %.4f)"%(d["e2e"]["ms_per_step"], d["e2e"]["value"]/1e6, d["e2e"]["host_pack_ms_per_step"]), "ascii %.4f"%d["e2e_ascii"]["ms_per_step"], "enqueue", {k:round(v,4) for k,v in d["host_enqueue_ms_per_step"].items()})
PY
done
