#!/bin/bash
# Round 2, ninth GPU pass (1 GPU): feeder thread (pack + copy off the calling thread), packer with software prefetch.
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "feeder or packed or async_input or intervals or device_resident" > gpurun_out/pytest_a.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_a.log
for pf in 0 2048 8192 16384; do echo "prefetch distance $pf"; HULK_B200_PACK_PREFETCH=$pf timeout 600 python tools/probe_pack.py 2>&1 | grep -E "^( 1|16|15|12) threads"; done > gpurun_out/r02h_pack_probe.txt 2>&1; cat gpurun_out/r02h_pack_probe.txt
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
run t15 HULK_B200_PACK_THREADS=15
run t14 HULK_B200_PACK_THREADS=14
run t12 HULK_B200_PACK_THREADS=12
run nofeed HULK_B200_FEEDER=0
