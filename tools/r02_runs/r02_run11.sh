#!/bin/bash
# Round 2, eleventh GPU pass (2 GPUs): the multi-GPU path on real peers -- tests (NCCL + peer reads, group API, CLI --gpus), weak and strong scaling lines.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02i_topo.txt 2>&1
timeout 300 python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 400 python -m pytest tests/test_distributed.py -m gpu -q -x --durations=5 > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest(distributed) rc=$?"; tail -12 gpurun_out/pytest_2gpu.log
timeout 200 python -m pytest tests/test_ingest_cli.py -m gpu -q -x -k "gpus" > gpurun_out/pytest_cli2.log 2>&1; echo "pytest(cli --gpus) rc=$?"; tail -3 gpurun_out/pytest_cli2.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
run() { tag=$1; n=$2; shift; shift; if [ $n -eq 1 ]; then cmd="python bench.py"; else cmd="$TR --nproc-per-node $n bench.py"; fi
  timeout 240 $cmd --gpus $n --steps 100 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err; echo "rc=$?"; tail -1 gpurun_out/bench_$tag.err | cut -c1-200; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_$tag.log").read().strip().splitlines() if l.startswith("{")][-1])
    print("%-10s"%"$tag", "N=%d %s"%(d["n_gpus"], d["scaling"]), "value %.0f M/s ms/step %.4f"%(d["value"]/1e6, d["ms_per_step"]), "e2e %.0f M/s %.4f"%(d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"]), "ascii %.4f"%d["e2e_ascii"]["ms_per_step"], d.get("parity_check"), d.get("parity_job"))
except Exception as e:
    print("$tag", "no line", e)
PY
}
run n1 1
run n2weak 2
run n2strong 2 --scaling strong
cp gpurun_out/bench_n1.log gpurun_out/r02i_bench_n1.json; cp gpurun_out/bench_n2weak.log gpurun_out/r02i_bench_n2_weak.json; cp gpurun_out/bench_n2strong.log gpurun_out/r02i_bench_n2_strong.json
