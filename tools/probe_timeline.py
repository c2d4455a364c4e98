#!/usr/bin/env python3
"""Timeline of the overlapped pipeline (C2 shape, device-resident reads): when each kernel class of each
interval starts and ends.  Event records perturb the schedule slightly; read it for structure."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hulk_b200
from bench import synthetic_reads_torch, synthetic_tables_torch
L = hulk_b200.load()
dev = torch.device("cuda", 0)
k, w, s, I, RL, K = 21, 9, 512, 100_000, 150, int(os.environ.get("TIMELINE_STEPS", "12"))
D = k ** 4
stream = torch.cuda.Stream(priority=-1)
with torch.cuda.stream(stream):
    reads = torch.stack([synthetic_reads_torch(torch, I, RL, 1, st * I, dev) for st in range(K)])
    r, c, b = synthetic_tables_torch(torch, s, D, 1234, dev)
stream.synchronize()
hs = hulk_b200.HistoSketch(k, w, s, 1.0, device=0, stream=stream.cuda_stream, async_input=True, input_ready=True)
hs.set_tables_device(r.data_ptr(), c.data_ptr(), b.data_ptr())
for rep in range(2):
    hs.reset()
    hs.profile(rep == 1)
    for st in range(K):
        hs.add_reads_device(reads[st].data_ptr(), None, I, RL)
        hs.flush()
    hs.sync()
rows = np.zeros((4096, 3)); n = C.c_uint64()
assert L.hulk_b200_profile_timeline(hs._ctx, rows.ctypes.data_as(C.c_void_p), 4096, C.byref(n)) == 0
rows = rows[:n.value]
names = ["k1", "k2", "k3a", "k3b"]
per = {c: rows[rows[:, 0] == c] for c in range(4)}
print("interval  " + "  ".join("%-17s" % nm for nm in names) + "   (start-end, us)")
for i in range(K):
    print("%8d  " % i + "  ".join("%7.0f-%-9.0f" % (per[c][i, 1] * 1e3, per[c][i, 2] * 1e3) for c in range(4)))
print("steady-state step: %.1f us" % ((per[3][-1, 2] - per[3][3, 2]) * 1e3 / (K - 4)))
