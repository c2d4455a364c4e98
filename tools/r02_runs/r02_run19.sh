#!/bin/bash
# Round 2, nineteenth GPU pass (8 GPUs): the scaling lines the driver will take (weak), strong scaling at C2, C4 as one job.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02m_topo8.txt 2>&1; nproc >> gpurun_out/r02m_topo8.txt
timeout 300 python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
run() { tag=$1; n=$2; shift; shift; if [ $n -eq 1 ]; then cmd="python bench.py"; else cmd="$TR --nproc-per-node $n bench.py"; fi
  timeout 200 $cmd --gpus $n --steps 100 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err; echo "rc=$?"; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_$tag.log").read().strip().splitlines() if l.startswith("{")][-1])
    open("gpurun_out/r02m_bench_$tag.json","w").write(json.dumps(d)+"\n")
    print("%-10s"%"$tag", "N=%d %s"%(d["n_gpus"], d["scaling"]), "value %.0f M/s %.4f"%(d["value"]/1e6, d["ms_per_step"]), "e2e %.0f M/s %.4f (pack %.4f, %d thr, h2d %.2f MB)"%(d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], d["e2e"]["host_pack_ms_per_step"], d["e2e"]["pack_threads"], d["e2e"]["h2d_bytes_per_step"]/1e6), "ascii %.0f M/s %.4f"%(d["e2e_ascii"]["value"]/1e6, d["e2e_ascii"]["ms_per_step"]), d.get("parity_check"), d.get("host_cpus_per_rank"))
except Exception as e:
    print("$tag", "no line", e); import subprocess; print(open("gpurun_out/bench_$tag.err").read()[-1500:])
PY
}
run n8weak 8
run n8strong 8 --scaling strong
run n4weak 4
run n1 1
timeout 200 $TR --nproc-per-node 8 tools/run_c4_multi.py > gpurun_out/r02m_c4_8gpu.jsonl 2> gpurun_out/r02m_c4_8gpu.err; echo "c4 rc=$?"; cat gpurun_out/r02m_c4_8gpu.jsonl | cut -c1-600; tail -2 gpurun_out/r02m_c4_8gpu.err | cut -c1-300
