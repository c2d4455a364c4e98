#!/usr/bin/env python3
"""Where the e2e step time goes: variants of bench.py's host-input step (C2 shape), device-timed."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hulk_b200
from bench import synthetic_reads_torch, synthetic_tables_torch
L = hulk_b200.load()
dev = torch.device("cuda", 0)
k, w, s, I, RL, K = 21, 9, 512, 100_000, 150, 20
D = k ** 4
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    reads = torch.stack([synthetic_reads_torch(torch, I, RL, 1, st * I, dev) for st in range(K)])
    r, c, b = synthetic_tables_torch(torch, s, D, 1234, dev)
stream.synchronize()
hs = hulk_b200.HistoSketch(k, w, s, 1.0, device=0, stream=stream.cuda_stream, async_input=True, input_ready=True)
hs.set_tables_device(r.data_ptr(), c.data_ptr(), b.data_ptr())
del r, c, b
pin = C.c_void_p(); L.hulk_b200_alloc_pinned(C.byref(pin), K * I * RL)
np.ctypeslib.as_array(C.cast(pin, C.POINTER(C.c_uint8)), shape=(K * I * RL,))[:] = reads.reshape(-1).cpu().numpy()
out = C.c_void_p(); L.hulk_b200_alloc_pinned(C.byref(out), 16 * s)

def run(name, push=True, flush=True, snap=True, device_in=False):
    hs.reset()
    for rep in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for st in range(K):
            if device_in:
                hs.add_reads_device(reads[st].data_ptr(), None, I, RL)
            elif push:
                assert L.hulk_b200_push_reads_fixed(hs._ctx, pin.value + st * I * RL, I, RL) == 0
            if flush:
                hs.flush()
            if snap:
                L.hulk_b200_snapshot_async(hs._ctx, out.value, out.value + 8 * s)
        t_host = time.perf_counter() - t0
        e1.record(stream)
        hs.sync()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
    print("%-34s %.4f ms/step (host enqueue %.4f ms/step)" % (name, ms, t_host * 1e3 / K))

run("device input + flush", device_in=True, snap=False)
run("device input + flush + snapshot", device_in=True)
run("host input, push only", flush=False, snap=False)
run("host input + flush", snap=False)
run("host input + flush + snapshot")
