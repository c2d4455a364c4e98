#!/usr/bin/env python3
"""Host-side newCWS draw: sequential vs all cores (bit-identical), on the box's CPUs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, hulk_b200
L = hulk_b200.load()
s, D = 512, 194481
r = np.empty((s, D)); c = np.empty((s, D)); b = np.empty((s, D))
for T in (1, 4, 8, 16, 32):
    t0 = time.time()
    rc = L.hulk_b200_new_cws_parallel(s, D, 0, s, r.ctypes.data, c.ctypes.data, b.ctypes.data, T, 1 << 21)
    print("newCWS s=%d D=%d threads=%d: %.2f s (rc %d)" % (s, D, T, time.time() - t0, rc), flush=True)
print("cores:", os.cpu_count())
