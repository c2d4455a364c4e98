#!/usr/bin/env python3
"""BASELINE configs C3 and C4 (one GPU's share) as whole jobs on one B200, device-timed, with the reference's CWS tables
(HistoSketch.newCWS, Go math/rand seed 1) drawn on the device and that draw included in the job's wall clock:
   C3: 10^8 x 150 bp reads, k=31, s=1024, concept drift 0.02, no interval (one flush at the end)
   C4: 1.25 x 10^8 x 150 bp reads (1/8 of 10^9), k=21, s=512/8 slots of 512, no interval
Reads are generated on the device in 10 M-read chunks (same counter-based generator as bench.py) and pushed
with hulk_b200_push_reads_device; generation is outside the timed kernels (events bracket push + flush only)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hulk_b200
from bench import synthetic_reads_torch, synthetic_tables_torch

dev = torch.device("cuda", 0)
CH = 10_000_000


def job(name, k, s, slots, decay, n_reads):
    D = k ** 4
    stream = torch.cuda.Stream(priority=-1)
    hs = hulk_b200.HistoSketch(k, 9, s, decay, device=0, slots=slots, stream=stream.cuda_stream, input_ready=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    hs.generate_tables_device()                     # newCWS, rows of this GPU's slots (histosketch.go:95-126)
    t_tables = time.perf_counter() - t0
    t_k = 0.0
    done = 0
    bufs = []
    while done < n_reads:
        n = min(CH, n_reads - done)
        with torch.cuda.stream(stream):
            reads = synthetic_reads_torch(torch, n, 150, 1, done, dev)
        stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        hs.add_reads_device(reads.data_ptr(), None, n, 150)
        hs.sync()                                   # the chunk buffer is reused; counting is what is timed
        e1.record(stream)
        torch.cuda.synchronize()
        t_k += e0.elapsed_time(e1)
        done += n
        del reads
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    hs.flush()
    e1.record(stream)
    mins, weights = hs.finish()
    t_f = e0.elapsed_time(e1)
    st = hs.stats()
    out = {"config": name, "k": k, "s": s, "slots": list(slots), "decay": decay, "reads": n_reads,
           "tables_ms": t_tables * 1e3, "count_ms": t_k, "flush_ms": t_f,
           "reads_per_s_with_table_draw": n_reads / ((t_k + t_f) * 1e-3 + t_tables),
           "reads_per_s": n_reads / ((t_k + t_f) * 1e-3),
           "gbases_per_s": n_reads * 150 / ((t_k + t_f) * 1e-3) / 1e9, "n_minimizers": st["n_minimizers"],
           "md5_mins": hulk_b200.md5_mins(mins)}
    print(json.dumps(out), flush=True)
    hs.close()


job("C3", 31, 1024, (0, 1024), 0.02, 100_000_000)
job("C4 (one of 8 GPUs)", 21, 512, (0, 64), 1.0, 125_000_000)
