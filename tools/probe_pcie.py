#!/usr/bin/env python3
"""Pinned host->device copy bandwidth of the box (the bound of bench.py's e2e number): sizes 15 MB .. 240 MB."""
import torch, time
dev = torch.device("cuda", 0)
for mb in (15, 60, 240):
    n = mb * 1000 * 1000
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("H2D pinned %4d MB: %.3f ms  %.1f GB/s" % (mb, ms, n / ms / 1e6))
    e0.record()
    for _ in range(20):
        h.copy_(d, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("D2H pinned %4d MB: %.3f ms  %.1f GB/s" % (mb, ms, n / ms / 1e6))

# same with cudaHostAlloc'ed memory (the library's allocator) and 1 / 2 / 4 copy streams in flight
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hulk_b200
L = hulk_b200.load()
n = 15 * 1000 * 1000
p = C.c_void_p()
assert L.hulk_b200_alloc_pinned(C.byref(p), 8 * n) == 0
rt = C.CDLL("libcudart.so.12")
d = torch.empty(8 * n, dtype=torch.uint8, device=dev)
for ns in (1, 2, 4):
    streams = [torch.cuda.Stream() for _ in range(ns)]
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(32):
            s = streams[i % ns]
            off = (i % 8) * n
            rt.cudaMemcpyAsync(C.c_void_p(d.data_ptr() + off), C.c_void_p(p.value + off), C.c_size_t(n), 1, C.c_void_p(s.cuda_stream))
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 32
    print("cudaHostAlloc H2D 15 MB x32 on %d stream(s): %.3f ms each  %.1f GB/s" % (ns, dt * 1e3, n / dt / 1e9))
for chunk in (1, 2, 4, 8):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(32 // chunk):
        off = ((i * chunk) % 8) * n
        rt.cudaMemcpyAsync(C.c_void_p(d.data_ptr() + off), C.c_void_p(p.value + off), C.c_size_t(n * chunk), 1, C.c_void_p(0))
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 32
    print("cudaHostAlloc H2D in %d x 15 MB pieces: %.3f ms per 15 MB  %.1f GB/s" % (chunk, dt * 1e3, n / dt / 1e9))
