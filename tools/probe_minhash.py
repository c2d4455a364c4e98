"""Cost of the MinHash feed (k1_minhash.cuh) at the C2 shape: 100 k reads of 150 bases per push, k 21, s 512."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hulk_b200 as hb  # noqa: E402

hb.load()
n, L = 100_000, 150
reads = hb.synthetic_reads(n * 4, L, seed=1).reshape(4, -1)
for label, kmv, khf in (("off", 0, 0), ("khf", 0, 1), ("kmv", 1, 0), ("both", 1, 1)):
    with hb.HistoSketch(21, 9, 512) as hs:
        if kmv or khf:
            hs.enable_minhash(bool(kmv), bool(khf))
        for i in range(3):
            hs.add_reads_fixed(reads[i % 4], n, L)
        hs.sync()
        t0 = time.perf_counter()
        for i in range(20):
            hs.add_reads_fixed(reads[i % 4], n, L)
        hs.sync()
        dt = (time.perf_counter() - t0) / 20
        print("feed %-4s: %.3f ms per 100 k reads (host copy included), %d launches" %
              (label, dt * 1e3, hs.stats()["n_kernel_launches"]), flush=True)
