#!/bin/bash
# Round 2, eighth GPU pass (1 GPU): where the packed e2e step's host time goes -- DRAM bandwidth of the host, packer from DRAM, bench with pack timing.
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
g++ -O2 -mavx2 -pthread -o /tmp/membw tools/membw.cpp && /tmp/membw > gpurun_out/r02g_membw.txt 2>&1; cat gpurun_out/r02g_membw.txt
timeout 600 python tools/probe_pack.py > gpurun_out/r02g_pack_probe.txt 2>&1; cat gpurun_out/r02g_pack_probe.txt
B="python bench.py --steps 100 --warmup 3 --no-cpu-baseline"
run() { tag=$1; shift; env "$@" timeout 600 $B > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err; tail -2 gpurun_out/bench_$tag.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$tag.log").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("%-10s"%"$tag", "ms/step %.4f"%d["ms_per_step"], "e2e %.4f (%.0f M/s, pack %.4f)"%(d["e2e"]["ms_per_step"], d["e2e"]["value"]/1e6, d["e2e"]["host_pack_ms_per_step"]), "ascii %.4f"%d["e2e_ascii"]["ms_per_step"], "enqueue", {k:round(v,4) for k,v in d["host_enqueue_ms_per_step"].items()})
except Exception as e:
    print("$tag", "no line", e)
PY
}
run auto X=1
run auto2 X=1
run t8 HULK_B200_PACK_THREADS=8
run t12 HULK_B200_PACK_THREADS=12
run t14 HULK_B200_PACK_THREADS=14
run t15 HULK_B200_PACK_THREADS=15
