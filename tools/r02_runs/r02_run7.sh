#!/bin/bash
# Round 2, seventh GPU pass (1 GPU): packed read transport -- parity, host packer throughput on this box, e2e with and without it.
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
lscpu | grep -E 'Model name|^CPU\(s\)|Thread|Core|Socket|NUMA' > gpurun_out/lscpu.txt; cat gpurun_out/lscpu.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "packed or histogram or device_resident or async_input or intervals" > gpurun_out/pytest_a.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_a.log
timeout 300 python tools/probe_pack.py > gpurun_out/r02f_pack_probe.txt 2>&1; cat gpurun_out/r02f_pack_probe.txt
B="python bench.py --steps 100 --warmup 3 --no-cpu-baseline"
run() { tag=$1; shift; env "$@" timeout 600 $B > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err; tail -2 gpurun_out/bench_$tag.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$tag.log").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("%-10s"%"$tag", "ms/step %.4f"%d["ms_per_step"], "e2e %.4f (%.0f M/s, h2d %.2f MB)"%(d["e2e"]["ms_per_step"], d["e2e"]["value"]/1e6, d["e2e"]["h2d_bytes_per_step"]/1e6), "ascii %.4f"%d["e2e_ascii"]["ms_per_step"], "enqueue", {k:round(v,4) for k,v in d["host_enqueue_ms_per_step"].items()}, d.get("parity_check"))
except Exception as e:
    print("$tag", "no line", e)
PY
}
run auto X=1
run t4 HULK_B200_PACK_THREADS=4
run t8 HULK_B200_PACK_THREADS=8
run t12 HULK_B200_PACK_THREADS=12
run t16 HULK_B200_PACK_THREADS=16
run t24 HULK_B200_PACK_THREADS=24
