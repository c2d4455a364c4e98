python - <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np, hulk_b200
n = 1_000_000
reads = hulk_b200.synthetic_reads(n, 150, seed=1)
qual = b"I" * 150
with open("/tmp/r.fq", "wb") as fh:
    buf = bytearray()
    for i in range(n):
        buf += b"@r%d\n" % i + reads[i].tobytes() + b"\n+\n" + qual + b"\n"
    fh.write(buf)
PY
for i in 1 2; do HULK_LOG_MICROSECONDS=1 hulk_b200/bin/hulk sketch -f /tmp/r.fq -s 50 -o /tmp/o | grep -v "processed [0-9]*00000 seq" | awk '{print $2, $3, $4, $5, $6, $7}' | sed -n '1p;17,30p'; echo; done
nvidia-smi --query-gpu=persistence_mode --format=csv
