#!/usr/bin/env python3
"""BASELINE config C4 as one job over the GPUs of a node: 10^9 x 150 bp reads, k=21, s=512, read-sharded, ONE flush at the
end over the spectrum summed across the GPUs (peer reads over NVLink), slots sharded.  One process per GPU:
   python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_c4_multi.py [total_reads]
Reads are generated on the device in 10 M-read chunks (bench.py's counter-based generator; rank g takes chunk indices
g, g + N, ...) and pushed with hulk_b200_push_reads_device; generation is outside the timed region (CUDA events bracket the
pushes and the flush on every rank, the job time is the maximum over ranks).  The reference's CWS tables are drawn on the
device by every rank for its own slots, timed separately."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import hulk_b200
from bench import synthetic_reads_torch

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
total = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
k, s, CH = 21, 512, 10_000_000
slots = hulk_b200.slot_range(s, world, rank)
stream = torch.cuda.Stream(priority=-1)
hs = hulk_b200.HistoSketch(k, 9, s, 1.0, device=local, slots=slots, stream=stream.cuda_stream, input_ready=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
hs.generate_tables_device()
t_tables = time.perf_counter() - t0
sh = hulk_b200.ShardedSketch(hs, s, world, rank)
n_chunks = (total + CH - 1) // CH
t_k = 0.0
mine = 0
for ci in range(rank, n_chunks, world):
    n = min(CH, total - ci * CH)
    with torch.cuda.stream(stream):
        reads = synthetic_reads_torch(torch, n, 150, 1, ci * CH, dev)
    stream.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    hs.add_reads_device(reads.data_ptr(), None, n, 150)
    hs.sync()
    e1.record(stream)
    torch.cuda.synchronize()
    t_k += e0.elapsed_time(e1)
    mine += n
    del reads
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
sh.flush()
e1.record(stream)
mins, weights = sh.finish()
t_f = e0.elapsed_time(e1)
nmin = sh.total_minimizers()
t = torch.tensor([t_k, t_f, t_tables * 1e3], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    t_k, t_f, t_tab = [float(x) for x in t.tolist()]
    print(json.dumps({"config": "C4", "n_gpus": world, "reads": total, "k": k, "s": s, "count_ms_max_over_ranks": t_k,
                      "flush_ms": t_f, "tables_ms": t_tab, "reads_per_s": total / ((t_k + t_f) * 1e-3),
                      "gbases_per_s": total * 150 / ((t_k + t_f) * 1e-3) / 1e9,
                      "reads_per_s_with_table_draw": total / ((t_k + t_f + t_tab) * 1e-3), "n_minimizers": nmin,
                      "peer_mode": sh.peer_mode, "md5_mins": hulk_b200.md5_mins(mins)}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
