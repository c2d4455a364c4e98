#!/bin/bash
# BASELINE config C5 on one B200: k in {11,21,31,51} x s in {128,512,2048}, interval 100 k reads (k=51 is rejected
# like the reference rejects it).  Writes gpurun_out/c5_<k>_<s>.log and prints one table.
mkdir -p gpurun_out
for k in 11 21 31; do for s in 128 512 2048; do
  timeout 300 python bench.py --k $k --s $s --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/c5_${k}_${s}.log 2> gpurun_out/c5_${k}_${s}.err || echo "k=$k s=$s failed"
done; done
python - <<'PY'
import json, glob
print("%4s %5s %12s %10s %12s %10s   %s" % ("k", "s", "reads/s", "ms/step", "e2e reads/s", "serial ms", "k1 / k2 / k3a / k3b ms, k3a GB/s of the stored table"))
for k in (11, 21, 31):
    for s in (128, 512, 2048):
        try:
            d = json.loads(open("gpurun_out/c5_%d_%d.log" % (k, s)).read().strip().split("\n")[-1])
            r = d["roofline"]; km = r["kernel_ms_per_step"]
            print("%4d %5d %12.1fM %10.4f %11.1fM %10.4f   %.3f / %.3f / %.3f / %.3f, %.0f" % (
                k, s, d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, r["serial_ms_per_step"],
                km["k1_minimizer_histogram"], km["k2_countmin"], km["k3_filter"], km["k3_resolve"], (r.get("k3_filter") or {}).get("achieved", 0)))
        except Exception as e:
            print("%4d %5d  failed: %r" % (k, s, e))
PY
python - <<'PY'
import sys
sys.path.insert(0, ".")
import hulk_b200
try:
    hulk_b200.HistoSketch(51, 9, 128)
except hulk_b200.HulkError as e:
    print("  51     *  rejected: %s (reference src/minimizer/minimizer.go:65-67)" % e)
PY
