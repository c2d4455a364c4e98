#!/bin/bash
# Round 2, twenty-ninth GPU pass (2 GPUs): BASELINE's C5 grid, strong scaling at G = 2 (the 100 k-read interval split two ways).
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511 --nproc-per-node 2"
for k in 11 21 31; do for s in 128 512 2048; do
  timeout 200 $TR bench.py --gpus 2 --scaling strong --kmer-size $k --sketch-size $s --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/c5g2_${k}_${s}.log 2> gpurun_out/c5g2_${k}_${s}.err || echo "k=$k s=$s failed"
done; done
python - <<'PY' > gpurun_out/r02r_c5_strong_g2.txt
import json
print("BASELINE config C5, strong scaling on 2 x B200 (bench.py --gpus 2 --scaling strong --k K --s S --steps 30 --warmup 5): the 100 k-read interval is")
print("split two ways, spectra summed by peer reads over NVLink inside the flush, slots sharded; round-2 kernels")
print("%4s %5s %12s %10s %12s   %s" % ("k", "s", "reads/s", "ms/step", "e2e reads/s", "parity"))
for k in (11, 21, 31):
    for s in (128, 512, 2048):
        try:
            d = json.loads([l for l in open("gpurun_out/c5g2_%d_%d.log" % (k, s)).read().strip().split("\n") if l.startswith("{")][-1])
            print("%4d %5d %12.1fM %10.4f %11.1fM   %s" % (k, s, d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, d.get("parity_check")))
        except Exception as e:
            print("%4d %5d  failed: %r" % (k, s, e))
PY
cat gpurun_out/r02r_c5_strong_g2.txt
