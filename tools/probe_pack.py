#!/usr/bin/env python3
"""Host packer throughput (hulk_b200_pack_bases) on this box by thread count: one C2 interval (15 MB of bases) per
call, (a) the same interval again and again (cache-resident input), (b) cycling through 128 distinct intervals of a
1.9 GB pinned buffer (input streamed from DRAM, what a real run and bench.py's e2e pass do)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hulk_b200 as hb
L = hb.load()
n, NI = 15_000_000, 128
pin = C.c_void_p(); assert L.hulk_b200_alloc_pinned(C.byref(pin), n * NI) == 0
b = np.ctypeslib.as_array(C.cast(pin, C.POINTER(C.c_uint8)), shape=(NI, n))
b[0] = hb.synthetic_reads(100_000, 150, seed=1).reshape(-1)
for i in range(1, NI):
    b[i] = b[0]
pout = C.c_void_p(); assert L.hulk_b200_alloc_pinned(C.byref(pout), 4 * (n // 4 + 64)) == 0
exc = np.zeros(1 << 20, np.uint32)
ne = C.c_uint64()
ncpu = len(os.sched_getaffinity(0))
print("cpus usable:", ncpu, "of", os.cpu_count())
for th in [1, 2, 4, 8, 12, 14, 15, 16, 24]:
    if th > 2 * ncpu:
        break
    res = []
    for stride in (0, n):
        for i in range(10):
            L.hulk_b200_pack_bases(pin.value + (i % NI) * stride, n, pout.value + (i % 4) * (n // 4 + 64), exc.ctypes.data_as(C.c_void_p), exc.size, C.byref(ne), th)
        reps = 128
        ts = []
        for i in range(reps):
            t0 = time.perf_counter()
            L.hulk_b200_pack_bases(pin.value + (i % NI) * stride, n, pout.value + (i % 4) * (n // 4 + 64), exc.ctypes.data_as(C.c_void_p), exc.size, C.byref(ne), th)
            ts.append(time.perf_counter() - t0)
        ts = np.array(ts)
        res.append("%s: median %.3f ms (%.0f GB/s), mean %.3f, max %.3f" % ("cached" if stride == 0 else "DRAM  ", np.median(ts) * 1e3, n / np.median(ts) / 1e9, ts.mean() * 1e3, ts.max() * 1e3))
    print("%2d threads  %s | %s" % (th, res[0], res[1]))
