#!/bin/bash
# Runs on the GPU box: `hulk sketch` wall clock on synthetic FASTQ files (plain and gzip), default sketch size and s=512.

mkdir -p gpurun_out /tmp/cli
python - <<'PY'
import sys, os, gzip, time
sys.path.insert(0, ".")
import numpy as np, hulk_b200
n = 2_000_000
t0 = time.time()
reads = hulk_b200.synthetic_reads(n, 150, seed=1)
qual = b"I" * 150
with open("/tmp/cli/r.fq", "wb") as fh:
    buf = bytearray()
    for i in range(n):
        buf += b"@r%d\n" % i + reads[i].tobytes() + b"\n+\n" + qual + b"\n"
        if len(buf) > (64 << 20):
            fh.write(buf); buf = bytearray()
    fh.write(buf)
print("wrote %d reads in %.1fs, %.0f MB" % (n, time.time() - t0, os.path.getsize("/tmp/cli/r.fq") / 1e6))
PY
head -c 400000000 /tmp/cli/r.fq | head -n 2000000 > /tmp/cli/h.fq   # 500k reads
gzip -1 -k -f /tmp/cli/h.fq
for args in "-f /tmp/cli/r.fq -s 50" "-f /tmp/cli/r.fq -s 50 -i 100000" "-f /tmp/cli/h.fq.gz -s 50" "-f /tmp/cli/r.fq -s 512 -i 100000"; do
  { time hulk_b200/bin/hulk sketch $args -o /tmp/cli/out > /tmp/cli/log.txt ; } 2> /tmp/cli/time.txt || true
  echo "== hulk sketch $args"; tail -4 /tmp/cli/log.txt; tr '\n' ' ' < /tmp/cli/time.txt; echo
done
