"""Cost of the MinHash feed (k1_minhash.cuh) at the C2 shape: 100 k reads of 150 bases per push, k 21, s 512, reads
resident in HBM (push_reads_device), 40 pushes between two synchronisations."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hulk_b200 as hb  # noqa: E402

hb.load()
n, L = 100_000, 150
reads = [torch.from_numpy(hb.synthetic_reads(n, L, seed=1, first_read=i * n).reshape(-1).copy()).cuda() for i in range(4)]
torch.cuda.synchronize()
for s in (512, 2048):
    for label, kmv, khf in (("off", 0, 0), ("khf", 0, 1), ("kmv", 1, 0), ("both", 1, 1)):
        with hb.HistoSketch(21, 9, s) as hs:
            if kmv or khf:
                hs.enable_minhash(bool(kmv), bool(khf))
            for i in range(4):
                hs.add_reads_device(reads[i % 4].data_ptr(), None, n, L)
            hs.sync()
            t0 = time.perf_counter()
            for i in range(40):
                hs.add_reads_device(reads[i % 4].data_ptr(), None, n, L)
            hs.sync()
            dt = (time.perf_counter() - t0) / 40
            print("s %4d, feed %-4s: %.4f ms per 100 k reads" % (s, label, dt * 1e3), flush=True)
