#!/bin/bash
# Round 2, twenty-sixth GPU pass (4 GPUs): split by measured rates with several ranks on one host; weak and strong lines.
mkdir -p gpurun_out
nproc > gpurun_out/r02q_host.txt
timeout 300 python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
export HULK_B200_FEED_STATS=1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
run() { tag=$1; n=$2; shift; shift; if [ $n -eq 1 ]; then cmd="python bench.py"; else cmd="$TR --nproc-per-node $n bench.py"; fi
  timeout 200 $cmd --gpus $n --steps 100 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err; echo "rc=$?"; grep -h "^\[feed\] packing" gpurun_out/bench_$tag.err | tail -2 | cut -c1-200; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_$tag.log").read().strip().splitlines() if l.startswith("{")][-1])
    open("gpurun_out/r02q_bench_$tag.json","w").write(json.dumps(d)+"\n")
    print("%-10s"%"$tag", "N=%d %s"%(d["n_gpus"], d["scaling"]), "value %.0f M/s %.4f"%(d["value"]/1e6, d["ms_per_step"]), "e2e %.0f M/s %.4f (pack %.4f, %d thr, h2d %.2f MB)"%(d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], d["e2e"]["host_pack_ms_per_step"], d["e2e"]["pack_threads"], d["e2e"]["h2d_bytes_per_step"]/1e6), "ascii %.0f M/s"%(d["e2e_ascii"]["value"]/1e6), d.get("parity_check"))
except Exception as e:
    print("$tag", "no line", e); print(open("gpurun_out/bench_$tag.err").read()[-1200:])
PY
}
run n4weak 4
run n4strong 4 --scaling strong
run n2weak 2
run n1 1
