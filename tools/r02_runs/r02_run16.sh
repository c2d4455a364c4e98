#!/bin/bash
# Round 2, sixteenth GPU pass (2 GPUs): hybrid transport (packed head + letters tail), pack threads per rank; N = 1 and N = 2.
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "feeder or packed or async_input" > gpurun_out/pytest_a.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_a.log
export HULK_B200_FEED_STATS=1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
run() { tag=$1; n=$2; shift; shift; if [ $n -eq 1 ]; then cmd="python bench.py"; else cmd="$TR --nproc-per-node $n bench.py"; fi
  timeout 240 $cmd --gpus $n --steps 100 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err; echo "rc=$?"; grep -h "^\[feed\]" gpurun_out/bench_$tag.err | tail -2 | cut -c1-200; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_$tag.log").read().strip().splitlines() if l.startswith("{")][-1])
    print("%-10s"%"$tag", "N=%d %s"%(d["n_gpus"], d["scaling"]), "value %.0f M/s %.4f"%(d["value"]/1e6, d["ms_per_step"]), "e2e %.0f M/s %.4f (pack %.4f, %d thr, h2d %.2f MB)"%(d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], d["e2e"]["host_pack_ms_per_step"], d["e2e"]["pack_threads"], d["e2e"]["h2d_bytes_per_step"]/1e6), "ascii %.4f"%d["e2e_ascii"]["ms_per_step"], d.get("parity_check"))
except Exception as e:
    print("$tag", "no line", e)
PY
}
run n1 1
run n1b 1
HULK_B200_PACK_FRACTION=1.0 run n1full 1
run n2weak 2
run n2strong 2 --scaling strong
