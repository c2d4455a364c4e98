"""Time of one long sequence through push_reads -> histogram, sliced scan (k1_long.cuh) on and off."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hulk_b200 as hb  # noqa: E402

hb.load()
rng = np.random.default_rng(3)
for n in (5_000_000, 20_000_000, 100_000_000):
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)]
    offs = np.array([0, n], dtype=np.uint64)
    for mode in ("1", "0"):
        if mode == "0" and n > 5_000_000:
            continue
        os.environ["HULK_B200_LONG"] = mode
        with hb.HistoSketch(21, 9, 4) as hs:
            hs.add_reads(seq[:100_000], np.array([0, 100_000], dtype=np.uint64))     # allocations, first launches
            hs.sync()
            ts = []
            for _ in range(2):
                t0 = time.perf_counter()
                hs.add_reads(seq, offs)
                hs.sync()
                ts.append(time.perf_counter() - t0)
            h = hs.histogram()
            print("%11d bases, sliced scan %s: %.4f s first, %.4f s second (%.1f Mbases/s), %d minimizers, checksum %d" %
                  (n, "on " if mode == "1" else "off", ts[0], ts[1], n / ts[1] / 1e6, hs.stats()["n_minimizers"],
                   int((h.astype(np.uint64) * np.arange(1, h.size + 1, dtype=np.uint64)).sum() % (2 ** 61 - 1))), flush=True)
