"""Read-sharded, slot-sharded sketching across the GPUs of one node (one process per GPU).

The reference is a single process (goroutine pool only, src/pipeline/boss.go:75-95); this module is the
multi-GPU form of SeqMinimizer.Run + Sketcher.Run (src/pipeline/sketch.go:182-301) described in
SURVEY.md section 8(e):

  * every interval's reads [(f-1)*I, f*I) are split into `world` contiguous chunks, rank g counts chunk g
    into its own uint32 spectrum (stage 1-2 is a commutative integer sum over reads);
  * ONE integer sum of the k^4-bin spectrum per flush (order independent => the summed spectrum is bit-identical
    to the single-GPU one).  With the CUDA engine every rank READS its peers' counting buffers over NVLink inside
    the flush (hulk_b200_peer_connect: CUDA IPC handles exchanged once through the process group, sequence flags in
    device memory, no collective call and no host synchronisation per flush); any other engine (the CPU checker
    of the gloo tests) gets a torch.distributed all-reduce;
  * the count-min update is replicated (tiny), the CWS sweep is sharded by sketch slot: rank g owns slots
    [g*s/world, (g+1)*s/world) and only those rows of the CWS tables;
  * `finish` all-gathers the (min, weight) pairs, so every rank returns the complete s-slot sketch.

The compute engine behind it is anything with the HistoSketch surface of hulk_b200.sketch (add_reads,
histogram_tensor, flush, finish, stats); the collectives are torch.distributed (NCCL on GPUs).  The CPU
tests drive this file at world size 2 over gloo with a checker engine defined in tests/.
"""
from __future__ import annotations

from typing import Iterable, Optional, Tuple

import numpy as np


def slot_range(sketch_size: int, world: int, rank: int) -> Tuple[int, int]:
    """Sketch slots owned by `rank`: contiguous, sizes differ by at most one."""
    return (rank * sketch_size) // world, ((rank + 1) * sketch_size) // world


def chunk_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Reads [begin, end) of an n-read segment that `rank` counts."""
    return (rank * n) // world, ((rank + 1) * n) // world


class ShardedSketch:
    """One rank's share of a sketch that spans `world` processes.

    engine: a HistoSketch-like object created with slots=slot_range(s, world, rank);
            `engine.histogram_tensor()` must return a torch int32 tensor aliasing the engine's spectrum
            on the device the process group communicates on.
    """

    def __init__(self, engine, sketch_size: int, world: int, rank: int, group=None, peers: bool = True):
        import torch.distributed as dist
        self.engine = engine
        self.sketch_size = sketch_size
        self.world, self.rank = world, rank
        self.group = group
        self._dist = dist
        self.slots = slot_range(sketch_size, world, rank)
        self._streams = {}          # CUDA stream handle -> torch ExternalStream
        self.seq_count = 0          # global seqCount (src/pipeline/sketch.go:203)
        # peers: the engine sums the spectra itself; the process group only carries the IPC handles, once
        self.peer_mode = False
        if world > 1 and peers and hasattr(engine, "peer_export") and hasattr(engine, "peer_connect"):
            handles = [None] * world
            dist.all_gather_object(handles, engine.peer_export(), group=group)
            engine.peer_connect(world, rank, handles)
            dist.barrier(group=group)                       # every rank is connected before anyone flushes
            self.peer_mode = True
        self._hist = engine.histogram_tensor() if world > 1 and not self.peer_mode else None

    def _device(self):
        if self._hist is not None:
            return self._hist.device
        import torch
        return torch.device("cuda", self.engine.device) if self.peer_mode else torch.device("cpu")

    def _collective_stream(self, hist):
        """The stream the engine orders its spectrum on (asked every flush: it depends on the engine's mode).
        The collective must be enqueued there: behind the counting kernels, ahead of the flush."""
        if not hist.is_cuda or not hasattr(self.engine, "stream_handle"):
            return None
        import torch
        h = self.engine.stream_handle()
        if h not in self._streams:
            self._streams[h] = torch.cuda.ExternalStream(h, device=hist.device)
        return self._streams[h]

    # stage 1-2 on this rank's chunk of a segment that lies inside one interval
    def add_segment(self, bases: np.ndarray, offsets: np.ndarray):
        n = len(offsets) - 1
        lo, hi = chunk_range(n, self.world, self.rank)
        if hi > lo:
            self.engine.add_reads(bases, offsets[lo:hi + 1])
        self.seq_count += n

    def flush(self):
        """theBoss.Flush (src/pipeline/boss.go:112-128) over the spectrum summed across ranks."""
        if self.world > 1 and not self.peer_mode:
            hist = self.engine.histogram_tensor()      # the engine multi-buffers its spectrum: ask every flush
            stream = self._collective_stream(hist)
            if stream is not None:
                import torch
                with torch.cuda.stream(stream):
                    self._dist.all_reduce(hist, op=self._dist.ReduceOp.SUM, group=self.group)
            else:
                self._dist.all_reduce(hist, op=self._dist.ReduceOp.SUM, group=self.group)
        self.engine.flush()

    def finish(self) -> Tuple[np.ndarray, np.ndarray]:
        """Complete (mins, weights) of all s slots on every rank."""
        import torch
        mins, weights = self.engine.finish()
        if self.world == 1:
            return mins, weights
        # (min, weight) pairs as int64 bit patterns: one object-free all_gather on the group's device
        dev = self._device()
        rows_max = max(b - a for a, b in (slot_range(self.sketch_size, self.world, r) for r in range(self.world)))
        mine = np.zeros((2, rows_max), dtype=np.int64)
        mine[0, :mins.size] = mins.view(np.int64)
        mine[1, :weights.size] = weights.view(np.int64)
        t = torch.from_numpy(mine).to(dev)
        out = [torch.empty_like(t) for _ in range(self.world)]
        self._dist.all_gather(out, t, group=self.group)
        all_mins = np.zeros(self.sketch_size, dtype=np.uint64)
        all_w = np.zeros(self.sketch_size, dtype=np.float64)
        for r, o in enumerate(out):
            a, b = slot_range(self.sketch_size, self.world, r)
            h = o.cpu().numpy()
            all_mins[a:b] = h[0, :b - a].view(np.uint64)
            all_w[a:b] = h[1, :b - a].view(np.float64)
        return all_mins, all_w

    def total_minimizers(self) -> int:
        """theBoss.GetMinimizerCount() over all ranks (src/pipeline/boss.go:93)."""
        import torch
        n = int(self.engine.stats()["n_minimizers"])
        if self.world == 1:
            return n
        t = torch.tensor([n], dtype=torch.int64, device=self._device())
        self._dist.all_reduce(t, group=self.group)
        return int(t.item())


def sketch_reads_sharded(sh: ShardedSketch, batches: Iterable[Tuple[np.ndarray, np.ndarray]], interval: int = 0):
    """SeqMinimizer.Run (src/pipeline/sketch.go:197-224) with every interval segment split across ranks.

    Every rank iterates the same global `batches` (bases uint8[], offsets uint64[n+1]) and counts only its
    chunk of each segment; flushes happen at the same global read counts as in the single-process run,
    so the result equals hulk_b200.sketch.sketch_reads on one GPU."""
    for bases, offsets in batches:
        n = len(offsets) - 1
        done = 0
        while done < n:
            take = n - done
            if interval:
                take = min(take, interval - (sh.seq_count % interval))
            sh.add_segment(bases, offsets[done:done + take + 1])
            done += take
            if interval and sh.seq_count % interval == 0:
                sh.flush()                                    # sketch.go:211-215
    sh.flush()                                                # sketch.go:221
    mins, weights = sh.finish()
    return mins, weights, sh.total_minimizers()
