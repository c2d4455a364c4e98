"""World-size-2 tests of the multi-GPU host logic (hulk_b200/distributed.py) on CPU over gloo.

The compute engine here is a CHECKER built on the CPU oracle (tests may use it; the product never
does): it has the HistoSketch surface ShardedSketch drives.  What is under test is the shard plan,
the per-flush spectrum all-reduce, slot sharding and the final gather: the sharded result must be
identical to the single-process run.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from hulk_b200.distributed import ShardedSketch, chunk_range, sketch_reads_sharded, slot_range  # noqa: E402


def test_shard_plan_covers_everything_once():
    for n, world in [(0, 2), (1, 2), (7, 2), (100000, 8), (512, 3), (50, 8), (5, 8)]:
        cuts = [chunk_range(n, world, r) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
        sizes = [b - a for a, b in cuts]
        assert max(sizes) - min(sizes) <= 1
        assert [slot_range(n, world, r) for r in range(world)] == cuts


class CheckerEngine:
    """HistoSketch-shaped engine on the CPU oracle, owning sketch slots [a, b)."""

    def __init__(self, O, k, w, D, decay, tables, slots):
        import torch
        self.O, self.k, self.w, self.D = O, k, w, D
        a, b = slots
        r, c, bb = tables
        self.hs = O.HistoSketch(k, b - a, D, decay, r[a:b], c[a:b], bb[a:b])
        self.hist = np.zeros(D, dtype=np.int32)
        self._t = torch.from_numpy(self.hist)
        self.n_min = 0

    def add_reads(self, bases, offsets):
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        h, nm = self.O.count_reads(self.k, self.w, self.D, bases, offsets)
        self.hist += h.astype(np.int32)
        self.n_min += nm

    def histogram_tensor(self):
        return self._t

    def flush(self):
        h = self.hist.astype(np.float64)
        if h.any():
            self.hs.flush(h)
        self.hist[:] = 0

    def finish(self):
        return self.hs.get()

    def stats(self):
        return {"n_minimizers": self.n_min}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_case(seed, n, L, ragged):
    from conftest import random_reads
    reads = random_reads(n, L, seed, n_frac=0.01, ragged=ragged)
    offsets = np.zeros(n + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(r) for r in reads], dtype=np.uint64)
    bases = np.frombuffer(b"".join(reads), dtype=np.uint8).copy()
    return bases, offsets


def _worker(rank, world, port, case, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import oracle as O
    k, w, s, decay, interval, seed, n, L, ragged = case
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        D = k ** 4
        rng = np.random.default_rng(99)
        r = rng.gamma(2.0, 1.0, (s, D))
        c = np.log(rng.gamma(2.0, 1.0, (s, D)))
        b = rng.random((s, D)) * r
        bases, offsets = _make_case(seed, n, L, ragged)
        eng = CheckerEngine(O, k, w, D, decay, (r, c, b), slot_range(s, world, rank))
        sh = ShardedSketch(eng, s, world, rank)
        # feed in uneven batches so segments straddle batch and interval boundaries
        cuts = [0, n // 3, n // 3 + 1, n]
        batches = [(bases, offsets[cuts[i]:cuts[i + 1] + 1]) for i in range(len(cuts) - 1)]
        mins, weights, nmin = sketch_reads_sharded(sh, batches, interval)
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), mins=mins, weights=weights, nmin=nmin)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", [
    # k, w, s, decay, interval, seed, n_reads, read_len, ragged
    (7, 5, 11, 1.0, 0, 1, 400, 60, 0),
    (7, 5, 10, 1.0, 150, 2, 500, 60, 9),
    (6, 4, 9, 0.5, 128, 3, 512, 50, 5),
])
def test_world2_sharded_sketch_equals_single_process(oracle, tmp_path, case):
    import torch.multiprocessing as mp
    k, w, s, decay, interval, seed, n, L, ragged = case
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, case, str(tmp_path)), nprocs=world, join=True)

    # single-process oracle run on the same reads
    D = k ** 4
    rng = np.random.default_rng(99)
    r = rng.gamma(2.0, 1.0, (s, D))
    c = np.log(rng.gamma(2.0, 1.0, (s, D)))
    b = rng.random((s, D)) * r
    bases, offsets = _make_case(seed, n, L, ragged)
    ref = oracle.HistoSketch(k, s, D, decay, r, c, b)
    nmin_ref, _ = ref.run(w, bases, offsets, interval=interval)
    mins_ref, weights_ref = ref.get()
    for rank in range(world):
        z = np.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
        assert int(z["nmin"]) == nmin_ref
        assert (z["mins"] == mins_ref).all()
        assert (z["weights"] == weights_ref).all()


# ---- the same sharded run on real GPUs over NCCL (needs >= 2 devices; `gpurun --gpus 2`) --------------
def _gpu_worker(rank, world, port, case, out_dir, peers=True):
    import torch
    import torch.distributed as dist
    import hulk_b200 as hb
    k, w, s, decay, interval, seed, n, L, ragged = case
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        D = hb.spectrum_size(k)
        rng = np.random.default_rng(99)
        r = rng.gamma(2.0, 1.0, (s, D))
        c = np.log(rng.gamma(2.0, 1.0, (s, D)))
        b = rng.random((s, D)) * r
        bases, offsets = _make_case(seed, n, L, ragged)
        a0, a1 = slot_range(s, world, rank)
        with hb.HistoSketch(k, w, s, decay, device=rank, slots=(a0, a1), tables=(r[a0:a1], c[a0:a1], b[a0:a1])) as hs:
            sh = ShardedSketch(hs, s, world, rank, peers=peers)
            assert sh.peer_mode == peers
            cuts = [0, n // 3, n // 3 + 1, n]
            batches = [(bases, offsets[cuts[i]:cuts[i + 1] + 1]) for i in range(len(cuts) - 1)]
            mins, weights, nmin = sketch_reads_sharded(sh, batches, interval)
            # the same job with the engine's internal pipelining switched off and back on: the collective has
            # to follow the engine to whichever stream orders its spectrum
            for overlap in (False, True):
                hs.reset()
                hs.set_overlap(overlap)
                sh.seq_count = 0
                m2, w2, n2 = sketch_reads_sharded(sh, batches, interval)
                assert (m2 == mins).all() and (w2 == weights).all() and n2 == nmin, "overlap=%s differs" % overlap
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), mins=mins, weights=weights, nmin=nmin)
    finally:
        dist.destroy_process_group()


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("peers", [True, False])      # spectra summed over NVLink inside the flush / NCCL all-reduce
@pytest.mark.parametrize("case", [
    (11, 9, 64, 1.0, 3000, 5, 10000, 150, 0),
    (9, 5, 33, 0.3, 2500, 6, 8000, 100, 20),
])
def test_nccl_sharded_sketch_equals_single_process_oracle(oracle, tmp_path, case, peers):
    import torch.multiprocessing as mp
    k, w, s, decay, interval, seed, n, L, ragged = case
    world = min(_n_gpus(), 4) if peers else 2
    mp.spawn(_gpu_worker, args=(world, _free_port(), case, str(tmp_path), peers), nprocs=world, join=True)
    D = k ** 4
    rng = np.random.default_rng(99)
    r = rng.gamma(2.0, 1.0, (s, D))
    c = np.log(rng.gamma(2.0, 1.0, (s, D)))
    b = rng.random((s, D)) * r
    bases, offsets = _make_case(seed, n, L, ragged)
    ref = oracle.HistoSketch(k, s, D, decay, r, c, b)
    nmin_ref, _ = ref.run(w, bases, offsets, interval=interval)
    mins_ref, weights_ref = ref.get()
    for rank in range(world):
        z = np.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
        assert int(z["nmin"]) == nmin_ref
        np.testing.assert_array_equal(z["mins"], mins_ref)
        np.testing.assert_allclose(z["weights"], weights_ref, rtol=1e-9)


# ---- one process driving several GPUs through the C ABI (hulk_b200_group_*) ---------------------------------
@pytest.mark.gpu
@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("case", [
    (11, 9, 64, 1.0, 3000, 5, 10000, 150, 0),
    (21, 9, 50, 1.0, 0, 7, 6000, 150, 30),
    (9, 5, 33, 0.3, 2500, 6, 8000, 100, 20),
])
def test_group_of_gpus_equals_single_process_oracle(oracle, case):
    """hulk_b200_group_*: one handle, the reference's call order, reads split over the GPUs of this process, the
    spectrum summed over NVLink in every flush: same sketch as the oracle's single loop, for 2 .. all GPUs, and the same
    bits as one GPU."""
    import hulk_b200 as hb
    k, w, s, decay, interval, seed, n, L, ragged = case
    D = hb.spectrum_size(k)
    rng = np.random.default_rng(99)
    r = rng.gamma(2.0, 1.0, (s, D))
    c = np.log(rng.gamma(2.0, 1.0, (s, D)))
    b = rng.random((s, D)) * r
    bases, offsets = _make_case(seed, n, L, ragged)
    ref = oracle.HistoSketch(k, s, D, decay, r, c, b)
    nmin_ref, nflush_ref = ref.run(w, bases, offsets, interval=interval)
    mins_ref, weights_ref = ref.get()
    cuts = [0, n // 3, n // 3 + 1, n]
    batches = [(bases, offsets[cuts[i]:cuts[i + 1] + 1]) for i in range(len(cuts) - 1)]
    with hb.HistoSketch(k, w, s, decay, tables=(r, c, b)) as one:
        mins1, weights1, _ = hb.sketch_reads(one, batches, interval=interval)
    for G in sorted({2, _n_gpus()}):
        with hb.GroupSketch(k, w, s, decay, ngpus=G, tables=(r, c, b)) as g:
            for rep in range(2):                       # a second sample through the same group: reset is collective
                mins, weights, st = hb.sketch_reads(g, batches, interval=interval)
                np.testing.assert_array_equal(mins, mins_ref)
                np.testing.assert_allclose(weights, weights_ref, rtol=1e-9)
                np.testing.assert_array_equal(mins, mins1)
                np.testing.assert_array_equal(weights, weights1)
                assert st["n_minimizers"] == nmin_ref and st["n_reads"] == n
                g.reset()


@pytest.mark.gpu
def test_group_of_one_gpu_is_the_plain_context(oracle):
    import hulk_b200 as hb
    k, w, s = 11, 9, 40
    D = hb.spectrum_size(k)
    rng = np.random.default_rng(5)
    r = rng.gamma(2.0, 1.0, (s, D))
    c = np.log(rng.gamma(2.0, 1.0, (s, D)))
    b = rng.random((s, D)) * r
    bases, offsets = _make_case(3, 5000, 150, 0)
    ref = oracle.HistoSketch(k, s, D, 1.0, r, c, b)
    nmin_ref, _ = ref.run(w, bases, offsets, interval=2000)
    with hb.GroupSketch(k, w, s, 1.0, ngpus=1, tables=(r, c, b)) as g:
        mins, weights, st = hb.sketch_reads(g, [(bases, offsets)], interval=2000)
    np.testing.assert_array_equal(mins, ref.get()[0])
    assert st["n_minimizers"] == nmin_ref
    with pytest.raises(hb.HulkError):
        hb.GroupSketch(k, w, 3, 1.0, ngpus=1, devices=[0, 0, 0, 0])      # more GPUs than slots


ONE_DEVICE_SCRIPT = r"""
import json, os, sys
import numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import hulk_b200 as hb
from oracle import oracle as O
from test_distributed import _make_case
k, w, s, decay, interval, seed, n, L, ragged = json.loads(sys.argv[2])
D = hb.spectrum_size(k)
rng = np.random.default_rng(99)
r = rng.gamma(2.0, 1.0, (s, D)); c = np.log(rng.gamma(2.0, 1.0, (s, D))); b = rng.random((s, D)) * r
bases, offsets = _make_case(seed, n, L, ragged)
ref = O.HistoSketch(k, s, D, decay, r, c, b)
nmin_ref, _ = ref.run(w, bases, offsets, interval=interval)
mins_ref, weights_ref = ref.get()
cuts = [0, n // 3, n // 3 + 1, n]
batches = [(bases, offsets[cuts[i]:cuts[i + 1] + 1]) for i in range(len(cuts) - 1)]
with hb.GroupSketch(k, w, s, decay, devices=[0, 0, 0], tables=(r, c, b)) as g:
    assert g.ngpus == 3
    for rep in range(2):
        mins, weights, st = hb.sketch_reads(g, batches, interval=interval)
        np.testing.assert_array_equal(mins, mins_ref)
        np.testing.assert_allclose(weights, weights_ref, rtol=1e-9)
        assert st["n_minimizers"] == nmin_ref and st["n_reads"] == n
        g.reset()
print("ONE_DEVICE_OK")
"""


@pytest.mark.gpu
@pytest.mark.parametrize("case", [
    (11, 9, 64, 1.0, 3000, 5, 10000, 150, 0),
    (9, 5, 33, 0.3, 2500, 6, 8000, 100, 20),
    (21, 9, 12, 1.0, 1000, 8, 3000, 150, 0),
])
def test_group_members_on_one_device_take_the_peer_path(oracle, case):
    """The whole multi-GPU mechanism -- chunked reads, sequence flags, the spectrum summed from the peers' buffers
    inside the flush, sharded slots, collective reset -- with three member contexts on ONE device, so it is checked
    on a single-GPU box as well: same sketch as the oracle's single loop.  (Own process: three contexts' streams on one
    device need more hardware queues than the default eight, or a waiting kernel can sit in front of the very
    signal it waits for -- on separate GPUs each device only carries its own context's streams.)"""
    import json
    import subprocess
    env = dict(os.environ, CUDA_DEVICE_MAX_CONNECTIONS="32")
    p = subprocess.run([sys.executable, "-c", ONE_DEVICE_SCRIPT, ROOT, json.dumps(case)], env=env, capture_output=True,
                       text=True, timeout=600)
    assert p.returncode == 0 and "ONE_DEVICE_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]
    if case[0] == 11:
        # the same with every member's share of a batch travelling packed: three feeder threads, one shared packer pool
        env["HULK_B200_PACK_INPUT"] = "1"
        p = subprocess.run([sys.executable, "-c", ONE_DEVICE_SCRIPT, ROOT, json.dumps(case)], env=env, capture_output=True,
                           text=True, timeout=600)
        assert p.returncode == 0 and "ONE_DEVICE_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]
