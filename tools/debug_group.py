#!/usr/bin/env python3
"""Debug: a group of contexts on one device, one flush, the flags afterwards."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hulk_b200 as hb
L = hb.load()
k, w, s = 11, 9, 12
D = hb.spectrum_size(k)
rng = np.random.default_rng(1)
r = rng.gamma(2.0, 1.0, (s, D)); c = np.log(rng.gamma(2.0, 1.0, (s, D))); b = rng.random((s, D)) * r
reads = hb.synthetic_reads(3000, 150, seed=1)
G = int(os.environ.get("G", "2"))
g = hb.GroupSketch(k, w, s, 1.0, devices=[0] * G, tables=(r, c, b))
g.add_reads_fixed(reads.reshape(-1), 3000, 150)
print("pushed", flush=True)
g.flush()
print("flush enqueued", flush=True)
try:
    g.sync()
    print("sync ok", flush=True)
except Exception as e:
    print("sync failed:", e, flush=True)
for i in range(G):
    m = L.hulk_b200_group_member(g._g, i)
    out = np.zeros((2, 4, 16), dtype=np.uint32)
    L.hulk_b200_peer_flags(m, out.ctypes.data_as(C.c_void_p))
    print("member", i, "counted", out[0, :, :G].tolist(), "gathered", out[1, :, :G].tolist(), flush=True)
