#!/usr/bin/env python3
"""Regenerates the oracle-derived golden files of tests/golden (run from the repo root).

c1_anchors.json   histogram anchors of the reference fixture at k=21, w=9
c1_k21_s50.json   the complete sketch `hulk sketch -f c1_reads.fq.gz -k 21 -s 50` would write, computed by the
                  CPU oracle with Go-compatible CWS tables (race-free flush semantics)
c1_k21_s50_x02_i250.json   same input with -x 0.2 -i 250 (concept drift + four interval flushes)
c1_k21_s50_minhash.json    KMVsketch / KHFsketch (s = 50) fed with every minimizer of the fixture in read order, by the
                  literal restatement of src/minhash (oracle/pyref.py): what `hulk sketch --kmv --khf` writes with the
                  MinHash feed on (HULK_B200_FEED_MINHASH=1)
"""
import gzip
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O, pyref as P  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
reads = [ln.rstrip(b"\n") for i, ln in enumerate(gzip.open(os.path.join(HERE, "c1_reads.fq.gz"), "rb")) if i % 4 == 1]
bases, offs = O.pack_reads(reads)
k, w, s = 21, 9, 50
D = k ** 4
hist, nmin = O.count_reads(k, w, D, bases, offs)
h32 = hist.astype(np.uint32)
nz = np.nonzero(h32)[0]
json.dump({
    "n_reads": len(reads), "n_minimizers": int(nmin), "used_bins": int((h32 != 0).sum()),
    "max_count": int(h32.max()), "argmax": int(h32.argmax()),
    "hist_md5": hashlib.md5(h32.tobytes()).hexdigest(),
    "first_bins": [[int(i), int(h32[i])] for i in nz[:5]], "last_bins": [[int(i), int(h32[i])] for i in nz[-3:]],
}, open(os.path.join(HERE, "c1_anchors.json"), "w"), indent=1)

r, c, b = O.new_cws(s, D)
for name, decay, interval in (("c1_k21_s50.json", 1.0, 0), ("c1_k21_s50_x02_i250.json", 0.2, 250)):
    hs = O.HistoSketch(k, s, D, decay, r, c, b)
    hs.run(w, bases, offs, interval=interval)
    mins, weights = hs.get()
    doc = P.sketch_json("testing/test-reads-small.fq.gz,", k, mins, weights, D, decay != 1.0)
    open(os.path.join(HERE, name), "w").write(doc)
    print(name, P.md5_of_mins(mins))

kmv, khf = P.KMVsketch(k, s), P.KHFsketch(k, s)
for rd in reads:
    for m in sorted(int(x) for x in O.minimizers(k, w, rd)):
        kmv.add_hash(m)
        khf.add_hash(m)
json.dump({"k": k, "w": w, "s": s, "n_added": kmv.multiplicity_sum, "kmv": kmv.get_sketch(), "khf": khf.get_sketch(),
           "kmv_md5": P.md5_of_mins(kmv.get_sketch()), "khf_md5": P.md5_of_mins(khf.get_sketch())},
          open(os.path.join(HERE, "c1_k21_s50_minhash.json"), "w"), indent=1)
print("c1_k21_s50_minhash.json", kmv.multiplicity_sum)
