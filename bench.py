#!/usr/bin/env python3
"""bench.py -- throughput of the `hulk sketch` hot path on B200 (contract: see the task statement).

Workload (BASELINE.json configs[1], "C2"): synthetic 150 bp reads, k=21, w=9, sketch size 512,
one sketch interval of 100 000 reads per step; `--steps 100` is the whole 10 M-read job.
A step = minimizer/histogram kernel over the interval's reads + the flush (count-min + CWS sweep).

  value   reads/s with the reads already resident in HBM (device timed, CUDA events on the launch stream)
  e2e     reads/s through the C ABI with HOST (pinned) input buffers: every step copies its reads
          host->device and the sketch (mins, weights) device->host inside the timed region
  --impl reference   the CPU oracle (restatement of the Go reference; no Go toolchain exists here)
          timed on the host cores with the same metric/config on a bounded sample per step

N > 1 (torchrun, one process per GPU): every rank sketches its own `interval` reads per step, the
uint32 histograms are NCCL-all-reduced, each rank sweeps its s/N sketch slots (weak scaling in reads).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# one hardware queue per stream (read when the CUDA context is created, i.e. before torch touches the device): the host
# path orders a counting stream behind its input copy with stream memory operations, see hulk_b200_create
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np  # noqa: E402

METRIC = "reads/sec (150 bp, k=21, s=512)"
UNIT = "reads/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--k", "--kmer-size", dest="k", type=int, default=21)
    ap.add_argument("--w", type=int, default=9)
    ap.add_argument("--s", "--sketch-size", dest="s", type=int, default=512,
                    help="(under torchrun use --sketch-size: its own parser takes --s for an abbreviation)")
    ap.add_argument("--interval", type=int, default=100_000)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--decay", type=float, default=1.0)
    ap.add_argument("--reads", default="iid", choices=["iid", "genome"],
                    help="iid: uniform i.i.d. bases (the headline workload).  genome: SURVEY section 8(d)'s realism "
                         "variant -- reads sampled at uniform offsets and strands from a seeded 100 Mbp random genome, so "
                         "minimizers repeat across reads (reported, never the headline)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = every GPU sketches `interval` reads per step (the default the driver's scaling run "
                         "uses); strong = the interval is fixed at `interval` reads and split N ways (BASELINE's C2/C5 "
                         "as written: 12 500 reads per GPU per flush at N = 8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=2)
    return ap.parse_args()


def workload_config(a, n_gpus):
    return {
        "workload": "C2: synthetic %d bp reads, k=%d, w=%d, sketch-size %d, interval=%d reads/step, decay=%g"
                    % (a.read_len, a.k, a.w, a.s, a.interval, a.decay)
                    + (" [realism variant: reads sampled from a seeded %d Mbp random genome, both strands; not the headline]"
                       % (GENOME_BASES // 1_000_000) if a.reads == "genome" else ""),
        "reads": a.reads,
        "k": a.k, "w": a.w, "sketch_size": a.s, "num_bins": a.k ** 4, "interval_reads": a.interval,
        "read_len": a.read_len, "decay_ratio": a.decay,
        "reads_per_step_total": a.interval * (n_gpus if a.scaling == "weak" else 1),
        "parallelism": "reads sharded over %d GPU(s); spectra summed over NVLink inside the flush (peer reads, no "
                       "collective call); CWS slots sharded" % n_gpus if n_gpus > 1 else "single GPU",
        "l2_policy": "inputs larger than L2: each step reads fresh reads and streams the %.0f MB (bf16) CWS screen table"
                     % (2.0 * a.s * a.k ** 4 / n_gpus / 1e6),
        "pipelining": "the spectrum is multi-buffered: intervals i+1.. are counted (k1) while interval i is flushed (k2, k3)",
    }


# ------------------------------------------------------------------------------------------------
# synthetic data on the device (same definition as hulk_b200.seqio.synthetic_reads / BASELINE.md section 3)
# ------------------------------------------------------------------------------------------------
def synthetic_reads_torch(torch, n_reads, read_len, seed, first_read, device):
    wpr = (read_len + 31) // 32

    def lsr(x, n):   # logical shift right on int64
        return (x >> n) & ((1 << (64 - n)) - 1)

    def c64(v):      # python int -> wrapped int64
        v &= (1 << 64) - 1
        return v - (1 << 64) if v >> 63 else v

    idx = (torch.arange(first_read, first_read + n_reads, dtype=torch.int64, device=device)[:, None] * wpr
           + torch.arange(wpr, dtype=torch.int64, device=device)[None, :])
    x = (idx ^ seed) + c64(0x9E3779B97F4A7C15)
    z = x
    z = (z ^ lsr(z, 30)) * c64(0xBF58476D1CE4E5B9)
    z = (z ^ lsr(z, 27)) * c64(0x94D049BB133111EB)
    z = z ^ lsr(z, 31)
    shifts = (torch.arange(32, dtype=torch.int64, device=device) * 2)[None, None, :]
    codes = ((z[:, :, None] >> shifts) & 3).reshape(n_reads, wpr * 32)[:, :read_len]
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    return lut[codes].contiguous()


GENOME_BASES = 100_000_000


def genome_torch(torch, device, seed=7):
    """The realism variant's genome: GENOME_BASES i.i.d. bases from the same counter-based generator."""
    return synthetic_reads_torch(torch, GENOME_BASES // 100, 100, seed, 0, device).reshape(-1)


def genome_reads_torch(torch, genome, n_reads, read_len, seed, first_read, device):
    """Reads `first_read .. first_read + n_reads` of the realism variant: offset and strand of read i come from
    splitmix64(seed ^ i) (so any rank can generate any read), the reverse strand is reverse-complemented."""
    def lsr(x, n):
        return (x >> n) & ((1 << (64 - n)) - 1)

    def c64(v):
        v &= (1 << 64) - 1
        return v - (1 << 64) if v >> 63 else v

    i = torch.arange(first_read, first_read + n_reads, dtype=torch.int64, device=device)
    z = (i ^ seed) + c64(0x9E3779B97F4A7C15)
    z = (z ^ lsr(z, 30)) * c64(0xBF58476D1CE4E5B9)
    z = (z ^ lsr(z, 27)) * c64(0x94D049BB133111EB)
    z = z ^ lsr(z, 31)
    off = lsr(z, 1) % (genome.numel() - read_len + 1)
    rev = (z & 1).bool()
    pos = off[:, None] + torch.arange(read_len, dtype=torch.int64, device=device)[None, :]
    fwd = genome[pos]
    comp = torch.zeros(256, dtype=torch.uint8, device=device)
    comp[torch.tensor(list(b"ACGT"), device=device).long()] = torch.tensor(list(b"TGCA"), dtype=torch.uint8, device=device)
    rc = comp[fwd.long()].flip(1)
    return torch.where(rev[:, None], rc, fwd).contiguous()


def synthetic_tables_torch(torch, s, D, seed, device, slots=None, chunk=32):
    """Random CWS tables with the reference's distributions (r, exp(c) ~ Gamma(2,1); b = U(0,1)*r), rows
    [slots[0], slots[1]) of the s x D tables.  Row chunk j is drawn from its own generator (seed + j), so any
    rank can produce exactly its rows without materialising the rest (k=31, s=2048 is 45 GB in float64)."""
    a, b_ = slots if slots is not None else (0, s)
    r = torch.empty((b_ - a, D), dtype=torch.float64, device=device)
    c = torch.empty_like(r)
    b = torch.empty_like(r)

    def gamma2(shape, g):   # Gamma(2,1) = -ln(U1*U2)
        u = torch.rand(shape, dtype=torch.float64, device=device, generator=g).clamp_min(1e-300)
        v = torch.rand(shape, dtype=torch.float64, device=device, generator=g).clamp_min(1e-300)
        return -(torch.log(u) + torch.log(v))

    for j in range(a // chunk, (b_ + chunk - 1) // chunk):
        g = torch.Generator(device=device)
        g.manual_seed(seed + j)
        rr = gamma2((chunk, D), g)
        cc = torch.log(gamma2((chunk, D), g))
        bb = torch.rand((chunk, D), dtype=torch.float64, device=device, generator=g) * rr
        lo, hi = max(a, j * chunk), min(b_, (j + 1) * chunk)
        r[lo - a:hi - a] = rr[lo - j * chunk:hi - j * chunk]
        c[lo - a:hi - a] = cc[lo - j * chunk:hi - j * chunk]
        b[lo - a:hi - a] = bb[lo - j * chunk:hi - j * chunk]
    return r, c, b


class ClockSampler(threading.Thread):
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.time(), [x.strip() for x in line.split(",")]))
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()

    def summary(self, t0, t1):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, f in self.rows:
            if len(f) < 9:
                continue
            try:
                mx = max(mx, int(float(f[2])))
                if t0 - 0.05 <= t <= t1 + 0.15:
                    sm.append(int(float(f[1])))
                    for nme, v in zip(names, f[5:9]):
                        if v.lower().startswith("active"):
                            reasons.add(nme)
            except ValueError:
                continue
        if not sm:   # timed region shorter than the sampling period: take the nearest samples
            sm = [int(float(f[1])) for _, f in self.rows[-3:] if len(f) >= 9 and f[1].replace(".", "").isdigit()]
        return {"sm_mhz": int(statistics.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle on host cores
# ------------------------------------------------------------------------------------------------
def cpu_arm(a, steps, warmup):
    """Time the CPU restatement on bounded samples of the same workload.

    One CPU step = 1/8 of an interval: interval/8 reads through stage 1-2 plus the flush of a DENSE
    full-interval histogram into s/8 sketch slots (both parts scale linearly, so reads/s of the
    scaled step equals reads/s of the full interval).  Uses every host thread OpenMP gives it."""
    import hulk_b200
    from oracle import oracle as O
    frac = 8
    k, w, L = a.k, a.w, a.read_len
    D = k ** 4
    s_sub = max(1, a.s // frac)
    n_sub = max(1, a.interval // frac)
    threads = O.num_threads()
    rng = np.random.default_rng(1)
    r = rng.gamma(2.0, 1.0, (s_sub, D))
    c = np.log(rng.gamma(2.0, 1.0, (s_sub, D)))
    b = rng.random((s_sub, D)) * r
    # dense histogram of one full interval (untimed set-up)
    full = hulk_b200.synthetic_reads(a.interval, L, seed=1).reshape(-1)
    offs_full = np.arange(a.interval + 1, dtype=np.uint64) * np.uint64(L)
    dense, _ = O.count_reads(k, w, D, full, offs_full)
    hs = O.HistoSketch(k, s_sub, D, a.decay, r, c, b)
    offs = np.arange(n_sub + 1, dtype=np.uint64) * np.uint64(L)
    times = []
    for step in range(warmup + steps):
        reads = hulk_b200.synthetic_reads(n_sub, L, seed=1, first_read=(step % frac) * n_sub).reshape(-1)
        t0 = time.perf_counter()
        O.count_reads(k, w, D, reads, offs)
        hs.flush(dense.copy(), parallel=True)
        dt = time.perf_counter() - t0
        if step >= warmup:
            times.append(dt)
    total = sum(times)
    value = n_sub * len(times) / total
    # the same step on ONE thread (one untimed + one timed step; it is ~threads x slower)
    single = None
    if threads > 1:
        try:
            O.set_threads(1)
            reads = hulk_b200.synthetic_reads(n_sub, L, seed=1).reshape(-1)
            for timed in (False, True):
                t0 = time.perf_counter()
                O.count_reads(k, w, D, reads, offs)
                hs.flush(dense.copy(), parallel=True)
                if timed:
                    single = n_sub / (time.perf_counter() - t0)
        finally:
            O.set_threads(threads)
    sample = ("%d steps of %d reads (1/%d interval) + flush of a dense %d-bin interval histogram into %d of %d "
              "slots; C oracle (restatement of the Go reference, %s), %d threads"
              % (len(times), n_sub, frac, D, s_sub, a.s, oracle_flags(), threads))
    base = {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
            "single_thread_value": single if threads > 1 else value, "host": host_info()}
    return value, total / len(times) * 1e3, base


def oracle_flags():
    """Compiler line of the oracle, read from its Makefile (always print the flags)."""
    try:
        mk = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle", "Makefile")).read()
        fl = [ln.split("=", 1)[1].strip() for ln in mk.splitlines() if ln.startswith("CFLAGS")][:1]
        return " ".join(["gcc"] + fl + ["-fopenmp"])
    except OSError:
        return "gcc"


def host_info():
    model = None
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                model = ln.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    try:
        usable = len(os.sched_getaffinity(0))
    except AttributeError:
        usable = os.cpu_count()
    return {"cpu_model": model, "nproc": os.cpu_count(), "usable_cpus": usable}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(a.steps, 6))
    warm = min(a.warmup, 1)
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the reference arm uses every core this process may run on
    from oracle import oracle as O
    try:
        O.set_threads(len(os.sched_getaffinity(0)))
    except AttributeError:
        O.set_threads(os.cpu_count() or 1)
    value, ms, base = cpu_arm(a, steps, warm)
    base["value"] = value
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(a, 1), "cpu_baseline": base,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "steps_run": steps,
    }))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def parity_check(hulk_b200, dist, world, rank, local):
    """Before anything is timed: a small job (k=11, w=9, s=64, four intervals, ragged tail) sketched by the SAME
    multi-GPU path the timed steps use -- reads split over the ranks, spectra summed in the flush, slots sharded --
    and compared on rank 0 with the CPU oracle's single loop (the checker; src/pipeline/sketch.go:197-224,
    src/histosketch/histosketch.go:129-155).  Returns "ok"; anything else stops the bench."""
    k, w, s, interval, n, L = 11, 9, 64, 4000, 15000, 150
    D = k ** 4
    rng = np.random.default_rng(4242)
    r = rng.gamma(2.0, 1.0, (s, D))
    c = np.log(rng.gamma(2.0, 1.0, (s, D)))
    b = rng.random((s, D)) * r
    reads = hulk_b200.synthetic_reads(n, L, seed=9)
    reads[::97, 40] = ord("N")                                        # the code-4 path too
    bases = reads.reshape(-1)
    offsets = np.arange(n + 1, dtype=np.uint64) * np.uint64(L)
    a0, a1 = hulk_b200.slot_range(s, world, rank)
    with hulk_b200.HistoSketch(k, w, s, 1.0, device=local, slots=(a0, a1), tables=(r[a0:a1], c[a0:a1], b[a0:a1])) as hs:
        sh = hulk_b200.ShardedSketch(hs, s, world, rank)
        mins, weights, nmin = hulk_b200.sketch_reads_sharded(sh, [(bases, offsets)], interval)
        mode = "peer reads over NVLink" if sh.peer_mode else ("all-reduce" if world > 1 else "single GPU")
        if world > 1:
            dist.barrier()                                            # nobody unmaps a buffer a peer may still read
    verdict = "ok"
    if rank == 0:
        try:
            from oracle import oracle as O
            ref = O.HistoSketch(k, s, D, 1.0, r, c, b)
            nmin_ref, _ = ref.run(w, bases, offsets, interval=interval, parallel=True)
            mins_ref, weights_ref = ref.get()
            if not ((mins == mins_ref).all() and np.allclose(weights, weights_ref, rtol=1e-12, atol=0) and nmin == nmin_ref):
                verdict = "MISMATCH against the oracle (%s, %d ranks)" % (mode, world)
        except OSError as e:
            verdict = "unavailable: %r" % (e,)
    if world > 1:
        import torch
        flag = torch.tensor([0 if verdict == "ok" or verdict.startswith("unavailable") else 1], device="cuda")
        dist.broadcast(flag, 0)
        if int(flag.item()):
            raise SystemExit("bench.py: parity check failed: " + verdict)
    elif not (verdict == "ok" or verdict.startswith("unavailable")):
        raise SystemExit("bench.py: parity check failed: " + verdict)
    return {"verdict": verdict, "job": "k=11 w=9 s=64, 15000 reads, interval 4000, %d rank(s), %s" % (world, mode)}


def run_b200(a):
    import torch
    import torch.distributed as dist
    import hulk_b200
    from hulk_b200 import _native as N

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; hulk_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    # NUMA: run this rank (and first-touch its pinned input buffers) on the CPUs next to its GPU
    affinity = None
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64 + 1)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            affinity = len(cpus)
    except Exception:
        pass
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L_ = hulk_b200.load()

    k, w, s, RL = a.k, a.w, a.s, a.read_len
    if a.scaling == "strong" and a.interval % world:
        raise SystemExit("--scaling strong: the interval must divide by the number of GPUs")
    I = a.interval if a.scaling == "weak" else a.interval // world      # reads THIS rank counts per step
    D = k ** 4
    if s % world:
        raise SystemExit("sketch size must divide by the number of GPUs")
    rows = s // world
    slots = (rank * rows, (rank + 1) * rows)
    K, W = a.steps, a.warmup
    stream = torch.cuda.Stream(device=dev, priority=-1)   # the flush chain runs here: ahead of the k1 streams
    n_steps_data = min(K, 128)                         # distinct intervals of reads kept resident (cycled beyond that; >> L2)

    with torch.cuda.stream(stream):
        # reads of this rank: rank-th shard of every interval of the global job
        reads_dev = torch.empty((n_steps_data, I, RL), dtype=torch.uint8, device=dev)
        genome = genome_torch(torch, dev) if a.reads == "genome" else None
        for st in range(n_steps_data):
            first = (st * world + rank) * I
            reads_dev[st] = (synthetic_reads_torch(torch, I, RL, 1, first, dev) if genome is None
                             else genome_reads_torch(torch, genome, I, RL, 1, first, dev))
        del genome
        r_t, c_t, b_t = synthetic_tables_torch(torch, s, D, 1234, dev, slots)
    stream.synchronize()

    parity = parity_check(hulk_b200, dist, world, rank, local)

    hs = hulk_b200.HistoSketch(k, w, s, a.decay, device=local, slots=slots, stream=stream.cuda_stream,
                               async_input=True, input_ready=True)
    hs.set_tables_device(r_t.data_ptr(), c_t.data_ptr(), b_t.data_ptr())
    del r_t, c_t, b_t
    torch.cuda.empty_cache()
    # multi-GPU: the same ShardedSketch the CPU (gloo) tests drive; its flush = spectrum all-reduce + local flush
    sh = hulk_b200.ShardedSketch(hs, s, world, rank)
    assert sh.slots == slots

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device(st):
        hs.add_reads_device(reads_dev[st % n_steps_data].data_ptr(), None, I, RL)
        sh.flush()

    host_enqueue = {}

    def timed(fn, nsteps, tag=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            h0 = time.perf_counter()
            for st in range(nsteps):
                fn(st)
            if tag:
                host_enqueue[tag] = (time.perf_counter() - h0) / nsteps * 1e3     # host time to issue one step (no sync)
            e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- value: inputs resident in HBM ------------------------------------------------------------
    with torch.cuda.stream(stream):
        for st in range(W):
            step_device(st)
    hs.sync()
    hs.reset()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.35)
    launches0 = hs.stats()["n_kernel_launches"]
    t0 = time.time()
    ms_value = timed(step_device, K, "device")
    t1 = time.time()
    hs.sync()                                          # surfaces deferred errors (short reads, sparse flush)
    st_value = hs.stats()
    launches = st_value["n_kernel_launches"] - launches0
    assert st_value["n_flushes"] == K, st_value
    value = world * I * K / (ms_value * 1e-3)
    mins_value, _ = hs.finish()

    # ---- per-kernel durations: the same K steps with the k1/flush overlap switched off, so that the CUDA
    # events around each kernel class (recorded on the stream it is launched on) bracket that kernel alone;
    # in the overlapped pass above the kernels of two intervals share the SMs and a bracket would time both
    hs.reset()
    hs.set_overlap(False)
    hs.profile(True)
    ms_serial = timed(step_device, K)
    hs.sync()
    prof = hs.profile_read()
    hs.profile(False)
    hs.set_overlap(True)
    mins_serial, _ = hs.finish()
    assert (mins_serial == mins_value).all()

    # ---- e2e: host (pinned) inputs through the C ABI, sketch read back every step -----------------
    pin_in = C.c_void_p()
    assert L_.hulk_b200_alloc_pinned(C.byref(pin_in), n_steps_data * I * RL) == 0
    pin_np = np.ctypeslib.as_array(C.cast(pin_in, C.POINTER(C.c_uint8)), shape=(n_steps_data * I * RL,))
    pin_np[:] = reads_dev.reshape(-1).cpu().numpy()
    pin_out = C.c_void_p()
    assert L_.hulk_b200_alloc_pinned(C.byref(pin_out), 16 * rows) == 0
    mins_p = pin_out.value
    weights_p = pin_out.value + 8 * rows

    def step_host(st):
        off = (st % n_steps_data) * I * RL
        rc = L_.hulk_b200_push_reads_fixed(hs._ctx, pin_in.value + off, I, RL)
        assert rc == 0, hs._L.hulk_b200_last_error(hs._ctx)
        sh.flush()
        rc = L_.hulk_b200_snapshot_async(hs._ctx, mins_p, weights_p)
        assert rc == 0

    def run_e2e(tag):
        hs.reset()
        with torch.cuda.stream(stream):
            for st in range(min(W, n_steps_data)):
                step_host(st)
        hs.sync()
        hs.reset()
        b0 = hs.stats()
        ms = timed(step_host, K, tag)
        hs.sync()
        b1 = hs.stats()
        mins = np.ctypeslib.as_array(C.cast(mins_p, C.POINTER(C.c_uint64)), shape=(rows,)).copy()
        return {"value": world * I * K / (ms * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": (b1["h2d_bytes"] - b0["h2d_bytes"]) / K * world,
                "d2h_bytes_per_step": (b1["d2h_bytes"] - b0["d2h_bytes"]) / K * world,
                "ms_per_step": ms / K, "host_pack_ms_per_step": (b1["pack_ns"] - b0["pack_ns"]) / K * 1e-6}, mins

    # ASCII reads over PCIe as they are (the link is the bound: 15 MB per step) ...
    hs.set_input_packing(0)
    e2e_ascii, mins_ascii = run_e2e("e2e_ascii")
    e2e_ascii["transport"] = "ASCII bases, one byte each"
    # ... and the default host path: the same call packs the batch to 2 bits per base on the host's cores INSIDE the
    # timed region (hulk_b200_pack_bases) and ships a quarter of the bytes; the device unpacks (k0_unpack)
    # host threads that pack: this rank's share of the CPUs it may run on, less two for the threads that issue the GPU work
    ncpu = len(os.sched_getaffinity(0))
    share = max(1, min(ncpu, (os.cpu_count() or ncpu) // world))
    pack_threads = int(os.environ.get("HULK_B200_PACK_THREADS", "0")) or max(1, min(32, share - 2 if share > 4 else share - 1))
    hs.set_input_packing(pack_threads)
    e2e, mins_e2e = run_e2e("e2e")
    e2e["transport"] = ("most of every batch as 2 bits per base (+ positions of non-ACGTU bytes), packed from the host's ASCII "
                        "buffer by %d host threads inside the timed region and unpacked on the device, the rest as letters "
                        "while the cores pack: the split follows the measured balance of cores and link, packed share = "
                        "(15e6 - h2d bytes per rank and step) / 11.25e6" % pack_threads)
    e2e["pack_threads"] = pack_threads
    hs.set_input_packing(0)
    assert (mins_ascii == mins_e2e).all()

    # the runs sketched the same reads: identical sketches
    hs_mins, _ = hs.finish()
    assert (hs_mins == mins_e2e).all() and (hs_mins == mins_value).all()
    if sampler:
        sampler.stop()
        clocks = sampler.summary(t0, t1)
    L_.hulk_b200_free_pinned(pin_in)
    L_.hulk_b200_free_pinned(pin_out)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    dom = max(prof, key=lambda n: prof[n]["ms"])
    alg_bytes = {
        "k1_minimizer_histogram": float(I * RL),                    # every base read once
        "k2_countmin": 16.0 * D,                                    # hist zero+read, f write+read
        "k3_filter": 4.0 * rows * D,                                # one fp32 coefficient per (slot, bin)
        "k3_resolve": 4.0 * rows * (D / 512.0),
    }
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
    except Exception:
        pass
    avg_ms = prof[dom]["ms"] / max(1, prof[dom]["launches"])
    achieved = alg_bytes[dom] / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "frac_of_nominal_8000": achieved / 8000.0, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes[dom], "avg_launch_ms": avg_ms,
                "timing": "CUDA events around every launch, second pass of the same %d steps with the k1/flush "
                          "overlap disabled (serial step %.4f ms); the overlapped pass is what `value` reports" % (K, ms_serial / K),
                "kernel_ms_per_step": {n: prof[n]["ms"] / K for n in prof},
                "kernel_share_of_step": {n: prof[n]["ms"] / ms_serial for n in prof},
                "serial_ms_per_step": ms_serial / K,
                }
    # the CWS screen: bytes it really streams (the stored table: 2 B per slot and bin as bfloat16, 4 B with
    # HULK_B200_K3_FP32=1, plus the (1/f) vector once per slot row from L2) against the measured HBM peak FIRST; the contract's
    # dense fp32 stream is given beside it as what the screen stands in for
    if prof["k3_filter"]["ms"] > 0:
        k3_ms = prof["k3_filter"]["ms"] / max(1, prof["k3_filter"]["launches"])
        elem = 4.0 if os.environ.get("HULK_B200_K3_FP32") == "1" else 2.0
        Dp = (D + 511) // 512 * 512
        streamed = elem * rows * Dp
        roofline["k3_filter"] = {"bound": "hbm", "streamed_bytes_per_launch": streamed, "avg_launch_ms": k3_ms,
                                 "achieved": streamed / (k3_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                 "frac": streamed / (k3_ms * 1e-3) / 1e9 / peak,
                                 "contract_bytes_per_launch": alg_bytes["k3_filter"],
                                 "contract_GBps": alg_bytes["k3_filter"] / (k3_ms * 1e-3) / 1e9,
                                 "note": "frac = bytes streamed from HBM / time / measured peak; contract_GBps divides the "
                                         "contract's dense fp32 bytes (which are never streamed) by the same time"}

    # K1 is issue-bound, not HBM-bound: next to the contract's HBM figure, report its
    # warp-instruction rate against the SM sub-partitions' issue rate.  Instruction counts per read are the
    # ncu-measured ones of profiles/traffic.json (they depend on k, w and the read length only).
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        wipr = sum(tj.get("warp_inst_per_read", {}).values())
        if wipr and k == 21 and w == 9 and RL == 150 and prof["k1_minimizer_histogram"]["ms"] > 0:
            k1_ms = prof["k1_minimizer_histogram"]["ms"] / max(1, prof["k1_minimizer_histogram"]["launches"])
            sm_count = torch.cuda.get_device_properties(local).multi_processor_count
            mhz = float(clocks.get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0)
            peak_ips = sm_count * 4 * mhz * 1e6
            ach = wipr * I / (k1_ms * 1e-3)
            roofline["k1_issue"] = {"bound": "issue", "achieved": ach / 1e9, "peak": peak_ips / 1e9, "unit": "G warp-inst/s",
                                    "frac": ach / peak_ips, "warp_inst_per_launch": wipr * I,
                                    "source": "profiles/traffic.json (ncu smsp__inst_executed.sum), 4 issue slots per SM per clock"}
    except Exception:
        pass

    # whole-step figure of SURVEY 8(d): B = sum len + F (4 s D + 16 D) + 16 s bytes over the pipelined step time
    step_bytes = float(I * RL) + 4.0 * rows * D + 16.0 * D + 16.0 * rows / max(1, K)
    step_gbs = step_bytes / (ms_value / K * 1e-3) / 1e9
    elem = 4.0 if os.environ.get("HULK_B200_K3_FP32") == "1" else 2.0
    moved = float(I * RL) + elem * rows * ((D + 511) // 512 * 512) + 40.0 * D      # reads, stored screen table, spectrum + f + 1/f
    moved_gbs = moved / (ms_value / K * 1e-3) / 1e9
    roofline["step"] = {"bytes_moved_per_step": moved, "achieved": moved_gbs, "unit": "GB/s", "frac": moved_gbs / peak,
                        "note": "bytes the pipelined step really moves through HBM over its time; per rank"}
    roofline["step_algorithmic"] = {"bytes_per_step": step_bytes, "achieved": step_gbs, "unit": "GB/s",
                                    "frac": step_gbs / peak, "frac_of_nominal_8000": step_gbs / 8000.0,
                                    "note": "contract bytes of SURVEY 8(d) (dense fp32 CWS stream, never streamed as such: the "
                                            "screen table is stored at half that) over the pipelined step; per rank"}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_value / K, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
        "dtype": "u64 minimizer/jump-hash, u32 histogram, bf16 CWS screen + f64 CWS resolve", "data": "synthetic",
        "config": workload_config(a, world),
        "gbases_per_s": value * RL / 1e9,
        "e2e": e2e,
        "e2e_ascii": e2e_ascii,
        "host_enqueue_ms_per_step": host_enqueue,
        "gpu_launches": int(launches),
        "parity_check": parity["verdict"], "parity_job": parity["job"],
        "host_cpus_per_rank": affinity,
        "clocks": clocks,
        "roofline": roofline,
        "n_minimizers": st_value["n_minimizers"], "n_rescans": st_value["n_rescans"],
    }
    if world == 1 and not a.no_cpu_baseline:
        try:
            _, _, base = cpu_arm(a, a.cpu_steps, 1)
            out["cpu_baseline"] = base
        except Exception as e:  # the oracle is test infrastructure; its absence must not hide the GPU number
            out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
