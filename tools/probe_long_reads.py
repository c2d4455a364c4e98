"""Reads of a few hundred to a few thousand bases: longest candidate list of the fast kernels (96: the rest goes to
k1_generic, one thread per read) and the sliced scan's threshold."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hulk_b200 as hb  # noqa: E402

hb.load()
rng = np.random.default_rng(4)
for n_reads, L in ((300_000, 300), (200_000, 600), (50_000, 2000), (20_000, 8000)):
    lens = rng.integers(L // 2, L + L // 2, n_reads)
    offs = np.zeros(n_reads + 1, dtype=np.uint64)
    offs[1:] = np.cumsum(lens)
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(offs[-1]))]
    d_b = torch.from_numpy(np.concatenate([bases, np.zeros(64, np.uint8)])).cuda()      # reads resident in HBM
    d_o = torch.from_numpy(offs.astype(np.int64)).cuda()
    torch.cuda.synchronize()
    sums = set()
    for cap, thr in ((96, 1 << 14), (96, 1024), (192, 1024)) if L < 2000 else ((192, 1 << 14), (192, 1024)):
        if L < 1000 and cap == 96 and thr == 1024:
            continue
        os.environ["HULK_B200_LIST_CAP"] = str(cap)
        os.environ["HULK_B200_LONG_MIN"] = str(thr)
        with hb.HistoSketch(21, 9, 4) as hs:
            hs.add_reads_device(d_b.data_ptr(), d_o.data_ptr(), n_reads, 0)
            hs.sync()
            ts = []
            for _ in range(3):
                t0 = time.perf_counter()
                hs.add_reads_device(d_b.data_ptr(), d_o.data_ptr(), n_reads, 0)
                hs.sync()
                ts.append(time.perf_counter() - t0)
            h = hs.histogram()
            sums.add(int((h.astype(np.uint64) * np.arange(1, h.size + 1, dtype=np.uint64)).sum() % (2 ** 61 - 1)))
            print("%7d reads of ~%5d bases (%4d Mbases), lists up to %3d, sliced scan from %5d bases: %.4f s (%.0f Mbases/s)" %
                  (n_reads, L, int(offs[-1]) // 10 ** 6, cap, thr, min(ts), int(offs[-1]) / min(ts) / 1e6), flush=True)
    print("   spectra equal across settings:", len(sums) == 1, flush=True)
