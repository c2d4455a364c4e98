#!/bin/bash
# Runs on the GPU box: FASTQ reader throughput, one producer thread vs the parallel parse, on a 3.1 GB plain file;
# then `hulk sketch` on the same file in both modes.
mkdir -p /tmp/cli
python - <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np, hulk_b200
n = 2_000_000
reads = hulk_b200.synthetic_reads(n, 150, seed=1)
qual = b"I" * 150
with open("/tmp/cli/r.fq", "wb") as fh:
    buf = bytearray()
    for i in range(n):
        buf += b"@r%d\n" % i + reads[i].tobytes() + b"\n+\n" + qual + b"\n"
        if len(buf) > (64 << 20):
            fh.write(buf); buf = bytearray()
    fh.write(buf)
PY
cat /tmp/cli/r.fq /tmp/cli/r.fq /tmp/cli/r.fq /tmp/cli/r.fq /tmp/cli/r.fq > /tmp/cli/big.fq
ls -la /tmp/cli/big.fq
for mode in 0 1; do
HULK_B200_PARALLEL_READER=$mode python - <<'PY'
import sys, time, os
sys.path.insert(0, ".")
import hulk_b200
for rep in range(2):
    t0 = time.time(); nr = 0
    with hulk_b200.NativeReader(["/tmp/cli/big.fq"]) as rd:
        t1 = time.time()
        for b, offs in rd: nr += len(offs) - 1
    dt = time.time() - t0
    print("parallel=%s: %d reads, open %.3f s, total %.3f s -> %.2f GB/s, %.1f M reads/s" % (
        os.environ["HULK_B200_PARALLEL_READER"], nr, t1 - t0, dt, 3.134/dt, nr/dt/1e6), flush=True)
PY
done
for mode in 0 1; do
  echo "== hulk sketch -f big.fq -s 50 (parallel reader $mode)"
  { time HULK_B200_PARALLEL_READER=$mode HULK_LOG_MICROSECONDS=1 hulk_b200/bin/hulk sketch -f /tmp/cli/big.fq -s 50 -o /tmp/cli/out$mode > /tmp/cli/log$mode.txt ; } 2>&1 | grep real
  grep -E "finding minimizers|generating final|finished in" /tmp/cli/log$mode.txt | cut -c1-90
done
cmp /tmp/cli/out0.json /tmp/cli/out1.json && echo "sketches identical"
