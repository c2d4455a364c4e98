python - <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np, hulk_b200
n = 2_000_000
reads = hulk_b200.synthetic_reads(n, 150, seed=1)
qual = b"I" * 150
with open("/tmp/r.fq", "wb") as fh:
    buf = bytearray()
    for i in range(n):
        buf += b"@r%d\n" % i + reads[i].tobytes() + b"\n+\n" + qual + b"\n"
        if len(buf) > (64 << 20):
            fh.write(buf); buf = bytearray()
    fh.write(buf)
PY
HULK_LOG_MICROSECONDS=1 hulk_b200/bin/hulk sketch -f /tmp/r.fq -s 50 -i 100000 -o /tmp/o | grep -v "processed [0-9]*00000 seq" | awk '{print $2, $3, $4, $5, $6, $7}' | sed -n '18,60p'
