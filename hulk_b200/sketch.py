"""Host-side mirror of the reference's sketch pipeline over the C ABI (include/hulk_b200.h).

Names follow the reference: `HistoSketch` stands where histosketch.HistoSketch +
kmerspectrum.KmerSpectrum + the boss/minion pool stand in src/pipeline (the GPU library fuses
them), `sketch_reads` is SeqMinimizer.Run + Sketcher.Run (src/pipeline/sketch.go:182-301).
Everything numeric happens in libhulk_b200.so on the GPU; this module only marshals buffers.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Optional, Sequence, Tuple

import numpy as np

from . import _native as N


class HulkError(RuntimeError):
    """A reference-fatal condition (helpers.ErrorCheck -> log.Fatalf, src/helpers/helpers.go:31-35)."""

    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code
        self.message = message


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def spectrum_size(k: int) -> int:
    """spectrumSize := int32(helpers.Pow(k, 4))  (cmd/sketch.go:118)"""
    v = (k ** 4) & 0xFFFFFFFF
    return v - (1 << 32) if v >= (1 << 31) else v


def new_cws(s: int, num_bins: int, slot_begin: int = 0, slot_end: Optional[int] = None):
    """HistoSketch.newCWS tables (src/histosketch/histosketch.go:95-126): r, c, b float64[rows, D]."""
    L = N.load()
    slot_end = s if slot_end is None else slot_end
    rows = slot_end - slot_begin
    r = np.empty((rows, num_bins)); c = np.empty((rows, num_bins)); b = np.empty((rows, num_bins))
    rc = L.hulk_b200_new_cws(s, num_bins, slot_begin, slot_end, _ptr(r), _ptr(c), _ptr(b))
    if rc:
        raise HulkError(rc, L.hulk_b200_strerror(rc).decode())
    return r, c, b


def pack_reads(reads: Sequence[bytes]) -> Tuple[np.ndarray, np.ndarray]:
    offsets = np.zeros(len(reads) + 1, dtype=np.uint64)
    if len(reads):
        offsets[1:] = np.cumsum([len(r) for r in reads], dtype=np.uint64)
    bases = np.frombuffer(b"".join(reads), dtype=np.uint8).copy() if len(reads) else np.zeros(0, np.uint8)
    return bases, offsets


def pack_bases(bases: np.ndarray, n_threads: int = 0, exceptions_cap: Optional[int] = None):
    """hulk_b200_pack_bases: ASCII bases -> (packed uint8[ceil(n/4)], exceptions uint32[], true exception count)."""
    L = N.load()
    bases = np.ascontiguousarray(bases, dtype=np.uint8).reshape(-1)
    n = bases.size
    packed = np.zeros(int(L.hulk_b200_packed_bytes(n)), dtype=np.uint8)
    cap = n if exceptions_cap is None else exceptions_cap
    exc = np.zeros(max(1, cap), dtype=np.uint32)
    n_exc = C.c_uint64()
    rc = L.hulk_b200_pack_bases(_ptr(bases), n, _ptr(packed), _ptr(exc), cap, C.byref(n_exc), n_threads)
    if rc:
        raise HulkError(rc, L.hulk_b200_strerror(rc).decode())
    return packed, exc[:min(cap, n_exc.value)].copy(), n_exc.value


def md5_mins(mins) -> str:
    L = N.load()
    m = np.ascontiguousarray(mins, dtype=np.uint64)
    out = C.create_string_buffer(33)
    L.hulk_b200_md5_mins(_ptr(m), m.size, out)
    return out.value.decode()


def sketch_json(filename: str, k: int, mins, weights, num_bins: int, concept_drift: bool,
                banner_label: str = "blank", kmv=None, khf=None) -> str:
    """The JSON document hulk writes to <outFile>.json (src/sketchio/sketchio.go:78-97).

    `kmv` / `khf`: mins of the optional MinHash signatures of `hulk sketch --kmv / --khf`, appended in that
    order (src/pipeline/sketch.go:227-234,289-294); an EMPTY array raises the reference's "no sketch was
    generated" error (src/sketchio/sketchio.go:59-61), None leaves the signature out."""
    L = N.load()
    m = np.ascontiguousarray(mins, dtype=np.uint64)
    w = np.ascontiguousarray(weights, dtype=np.float64)
    kv = None if kmv is None else np.ascontiguousarray(kmv, dtype=np.uint64)
    kh = None if khf is None else np.ascontiguousarray(khf, dtype=np.uint64)
    keep = np.zeros(1, dtype=np.uint64)              # a non-NULL address for "present but empty"

    def opt(a):
        return (None, 0) if a is None else (_ptr(a if a.size else keep), a.size)

    args = (filename.encode(), banner_label.encode(), k, _ptr(m), _ptr(w), m.size, num_bins, int(concept_drift),
            *opt(kv), *opt(kh))
    need = L.hulk_b200_sketch_json_minhash(None, 0, *args)
    if need < 0:
        raise HulkError(int(need), L.hulk_b200_strerror(int(need)).decode())
    buf = C.create_string_buffer(need + 1)
    L.hulk_b200_sketch_json_minhash(buf, need + 1, *args)
    return buf.raw[:need].decode()


def khf_unfed(s: int) -> np.ndarray:
    """The KHF sketch `hulk sketch --khf` writes: minhash.NewKHFsketch fills s slots with MaxUint64
    (src/minhash/khf.go:20-32) and nothing ever calls AddHash on it (src/pipeline/boss.go:18-19 "not used yet")."""
    return np.full(s, np.iinfo(np.uint64).max, dtype=np.uint64)


class HistoSketch:
    """One sketching context on one GPU.

    Constructor checks mirror minimizer.NewMinimizerSketch (src/minimizer/minimizer.go:62-67),
    kmerspectrum.NewKmerSpectrum (src/kmerspectrum/kmerspectrum.go:33-35) and
    histosketch.NewHistoSketch (src/histosketch/histosketch.go:53-67).
    """

    def __init__(self, k: int = 21, w: int = 9, sketch_size: int = 50, decay_ratio: float = 1.0,
                 num_bins: Optional[int] = None, device: int = 0, slots: Optional[Tuple[int, int]] = None,
                 stream: Optional[int] = None, tables=None, async_input: bool = False,
                 input_ready: bool = False, pack_input: bool = False):
        self._L = N.load()
        self._ctx = C.c_void_p()
        self.k, self.w, self.sketch_size, self.decay_ratio = k, w, sketch_size, decay_ratio
        self.num_bins = spectrum_size(k) if num_bins is None else num_bins
        self.slot_begin, self.slot_end = slots if slots is not None else (0, sketch_size)
        self.device = device
        p = N.Params()
        p.k, p.w, p.sketch_size = k, w, sketch_size
        p.num_bins = self.num_bins
        p.decay_ratio = decay_ratio
        p.device = device
        p.slot_begin, p.slot_end = (slots if slots is not None else (0, 0))
        p.stream = stream
        p.flags = ((N.F_ASYNC_INPUT if async_input else 0) | (N.F_INPUT_READY if input_ready else 0)
                   | (N.F_PACK_INPUT if pack_input else 0))
        ctx = C.c_void_p()
        rc = self._L.hulk_b200_create(C.byref(p), C.byref(ctx))
        if rc:
            raise HulkError(rc, self._L.hulk_b200_last_error(None).decode())
        self._ctx = ctx
        self._keep = []     # host buffers that must outlive asynchronous copies
        if tables is not None:
            self.set_tables(*tables)

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None):
            self._L.hulk_b200_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc:
            raise HulkError(rc, self._L.hulk_b200_last_error(self._ctx).decode())

    @property
    def rows(self) -> int:
        return self.slot_end - self.slot_begin

    @property
    def concept_drift(self) -> bool:
        return self.decay_ratio != 1.0           # histosketch.go:79-81

    # -- CWS tables --------------------------------------------------------------------------
    def set_tables(self, r, c, b):
        r = np.ascontiguousarray(r, dtype=np.float64)
        c = np.ascontiguousarray(c, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        want = (self.rows, self.num_bins)
        for t in (r, c, b):
            if t.shape != want:
                raise HulkError(N.EARG, f"table shape {t.shape}, expected {want}")
        self._check(self._L.hulk_b200_set_cws_tables(self._ctx, _ptr(r), _ptr(c), _ptr(b)))

    def set_tables_device(self, d_r: int, d_c: int, d_b: int):
        """Tables already in device memory (raw device pointers to rows x num_bins float64)."""
        self._check(self._L.hulk_b200_set_cws_tables_device(self._ctx, d_r, d_c, d_b))

    def reset(self):
        """Start a new sample with the same parameters and tables."""
        self._check(self._L.hulk_b200_reset(self._ctx))

    def profile(self, enable: bool):
        self._check(self._L.hulk_b200_profile_enable(self._ctx, int(enable)))

    def set_overlap(self, enable: bool):
        """Counting interval i+1 while interval i is flushed (default on); off = strict call order on one stream."""
        self._check(self._L.hulk_b200_set_overlap(self._ctx, int(enable)))

    def profile_read(self) -> dict:
        pr = N.Profile()
        self._check(self._L.hulk_b200_profile_read(self._ctx, C.byref(pr)))
        names = ("k1_minimizer_histogram", "k2_countmin", "k3_filter", "k3_resolve")
        return {n: {"ms": pr.ms[i], "launches": int(pr.launches[i])} for i, n in enumerate(names)}

    def generate_tables(self, background: bool = False):
        """newCWS with the reference's seeded streams (histosketch.go:95-126).  background=True draws on a
        host thread while reads are pushed; the first flush waits for it."""
        fn = self._L.hulk_b200_generate_cws_tables_async if background else self._L.hulk_b200_generate_cws_tables
        self._check(fn(self._ctx))

    # -- stage 1+2 ---------------------------------------------------------------------------
    def generate_tables_device(self):
        """HistoSketch.newCWS drawn on the GPU (hulk_b200_generate_cws_tables_device)."""
        self._check(self._L.hulk_b200_generate_cws_tables_device(self._ctx))

    def tables(self):
        """The float64 CWS tables as they sit on the device: (r, c, b), rows x num_bins each."""
        shape = (self.slot_end - self.slot_begin, self.num_bins)
        r, c, b = np.zeros(shape), np.zeros(shape), np.zeros(shape)
        self._check(self._L.hulk_b200_get_cws_tables(self._ctx, _ptr(r), _ptr(c), _ptr(b)))
        return r, c, b

    def add_reads(self, bases: np.ndarray, offsets: np.ndarray):
        """theBoss.AddSeq for a batch (src/pipeline/boss.go:24-26)."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self._keep = [bases, offsets]
        self._check(self._L.hulk_b200_push_reads(self._ctx, _ptr(bases), _ptr(offsets), offsets.size - 1))

    def add_seqs(self, reads: Sequence[bytes]):
        self.add_reads(*pack_reads(reads))

    def add_reads_fixed(self, bases: np.ndarray, n_reads: int, read_len: int):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        self._keep = [bases]
        self._check(self._L.hulk_b200_push_reads_fixed(self._ctx, _ptr(bases), n_reads, read_len))

    def set_input_packing(self, n_threads: int):
        """Host batches travel as 2 bits per base (packed on n_threads host threads; < 0: all CPUs; 0: off)."""
        self._check(self._L.hulk_b200_set_input_packing(self._ctx, n_threads))

    def add_reads_packed(self, packed: np.ndarray, exceptions: np.ndarray, offsets: Optional[np.ndarray],
                         n_reads: int, read_len: int = 0):
        """theBoss.AddSeq for a batch the caller packed with pack_bases (hulk_b200_push_reads_packed)."""
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        exceptions = np.ascontiguousarray(exceptions, dtype=np.uint32)
        offs = None if offsets is None else np.ascontiguousarray(offsets, dtype=np.uint64)
        self._keep = [packed, exceptions, offs]
        self._check(self._L.hulk_b200_push_reads_packed(self._ctx, _ptr(packed), _ptr(exceptions), exceptions.size,
                                                        _ptr(offs), n_reads, read_len))

    def add_reads_device(self, d_bases_ptr: int, d_offsets_ptr: Optional[int], n_reads: int, read_len: int = 0):
        self._check(self._L.hulk_b200_push_reads_device(self._ctx, d_bases_ptr, d_offsets_ptr, n_reads, read_len))

    # -- stage 3 -----------------------------------------------------------------------------
    def flush(self):
        """theBoss.Flush (src/pipeline/boss.go:112-128).  Asynchronous; errors surface at sync()/finish()."""
        self._check(self._L.hulk_b200_flush(self._ctx))

    def sync(self):
        self._check(self._L.hulk_b200_sync(self._ctx))

    def finish(self) -> Tuple[np.ndarray, np.ndarray]:
        mins = np.zeros(self.rows, dtype=np.uint64)
        weights = np.zeros(self.rows, dtype=np.float64)
        self._check(self._L.hulk_b200_finish(self._ctx, _ptr(mins), _ptr(weights)))
        return mins, weights

    def stats(self) -> dict:
        st = N.Stats()
        self._check(self._L.hulk_b200_get_stats(self._ctx, C.byref(st)))
        return {n: int(getattr(st, n)) for n, _ in N.Stats._fields_}

    # -- MinHash side sketches (src/minhash; unfed in the reference, src/pipeline/boss.go:18-19) ---------------
    def enable_minhash(self, kmv: bool = True, khf: bool = True):
        """Feed every minimizer to KMVsketch.AddHash / KHFsketch.AddHash as well (before the first read)."""
        self._check(self._L.hulk_b200_minhash_enable(self._ctx, int(kmv), int(khf)))

    def khf(self) -> np.ndarray:
        """KHFsketch.GetSketch (src/minhash/khf.go:58-60)."""
        mins = np.zeros(self.sketch_size, dtype=np.uint64)
        self._check(self._L.hulk_b200_get_khf(self._ctx, _ptr(mins)))
        return mins

    def kmv(self) -> np.ndarray:
        """KMVsketch.GetSketch (src/minhash/kmv.go:160-176): at most sketch_size values, low -> high."""
        mins = np.zeros(self.sketch_size, dtype=np.uint64)
        n = C.c_uint32(0)
        self._check(self._L.hulk_b200_get_kmv(self._ctx, _ptr(mins), C.byref(n)))
        return mins[:n.value].copy()

    # -- multi-GPU plumbing ------------------------------------------------------------------
    def histogram_device_ptr(self) -> int:
        p = C.c_void_p()
        nb = C.c_int32()
        self._check(self._L.hulk_b200_histogram_device_ptr(self._ctx, C.byref(p), C.byref(nb)))
        return p.value

    def histogram_tensor(self):
        """The device spectrum of the interval being counted, as a torch int32 tensor aliasing the context's
        memory (for the per-flush all-reduce of hulk_b200.distributed; uint32 sums wrap identically in two's
        complement).  The spectrum is multi-buffered, so ask again after every flush; work enqueued on the
        context's stream after this call sees every read pushed so far."""
        import torch
        ptr = self.histogram_device_ptr()
        cache = self.__dict__.setdefault("_hist_tensors", {})
        if ptr not in cache:
            class _Dev:
                pass
            v = _Dev()
            v.__cuda_array_interface__ = {"shape": (self.num_bins,), "typestr": "<i4", "data": (ptr, False), "version": 2}
            cache[ptr] = (torch.as_tensor(v, device=torch.device("cuda", self.device)), v)
        return cache[ptr][0]

    def stream_handle(self) -> int:
        p = C.c_void_p()
        self._check(self._L.hulk_b200_stream(self._ctx, C.byref(p)))
        return p.value or 0

    def merge_histogram(self, hist: np.ndarray):
        """Add a partial spectrum counted elsewhere (uint32[num_bins]) to this context's histogram."""
        h = np.ascontiguousarray(hist, dtype=np.uint32)
        if h.size != self.num_bins:
            raise HulkError(N.EARG, "histogram size")
        self._check(self._L.hulk_b200_merge_histogram(self._ctx, _ptr(h)))

    def add_minimizer_count(self, n: int):
        self._check(self._L.hulk_b200_add_minimizer_count(self._ctx, int(n)))

    def peer_export(self) -> bytes:
        """This context's handle for hulk_b200_peer_connect in ANOTHER process (CUDA IPC, one process per GPU)."""
        buf = C.create_string_buffer(N.PEER_HANDLE_BYTES)
        self._check(self._L.hulk_b200_peer_export(self._ctx, buf))
        return buf.raw

    def peer_connect(self, world: int, rank: int, handles: Sequence[bytes]):
        """Join `world` contexts (one per GPU, one per process): from now on a flush works on the SUM of their spectra,
        read from the peers over NVLink inside the flush -- no collective call.  flush() and reset() become collective."""
        blob = b"".join(handles)
        if len(blob) != world * N.PEER_HANDLE_BYTES:
            raise HulkError(N.EARG, "one handle per rank")
        self._check(self._L.hulk_b200_peer_connect(self._ctx, world, rank, blob))
        self.peers = world

    # -- parity taps -------------------------------------------------------------------------
    def histogram(self) -> np.ndarray:
        h = np.zeros(self.num_bins, dtype=np.uint32)
        self._check(self._L.hulk_b200_get_histogram(self._ctx, _ptr(h)))
        return h

    def estimates(self) -> np.ndarray:
        f = np.zeros(self.num_bins, dtype=np.float64)
        self._check(self._L.hulk_b200_get_estimates(self._ctx, _ptr(f)))
        return f

    def cms(self) -> np.ndarray:
        q = np.zeros((7, 2000), dtype=np.float64)
        self._check(self._L.hulk_b200_get_cms(self._ctx, _ptr(q)))
        return q

    def minimizers(self, reads: Sequence[bytes], cap: int = 0):
        """Per-read minimizer sets computed by the device code (list of sorted uint64 arrays)."""
        bases, offsets = pack_reads(reads)
        if cap == 0:
            cap = max(1, max((len(r) for r in reads), default=1))
        out = np.zeros((len(reads), cap), dtype=np.uint64)
        counts = np.zeros(len(reads), dtype=np.uint32)
        self._check(self._L.hulk_b200_minimizers(self._ctx, _ptr(bases), _ptr(offsets), len(reads), _ptr(out),
                                                 cap, _ptr(counts)))
        return [np.sort(out[i, :min(int(counts[i]), cap)]) for i in range(len(reads))], counts

    def jump_hash(self, keys, num_buckets: int) -> np.ndarray:
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        out = np.zeros(keys.size, dtype=np.int32)
        self._check(self._L.hulk_b200_jump_hash(self._ctx, _ptr(keys), keys.size, num_buckets, _ptr(out)))
        return out

    def jump_hash_fx(self, keys, num_buckets: int):
        """jump.Hash through the kernel's fixed-point step; returns (bins, number of exact-division fallbacks)."""
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        out = np.zeros(keys.size, dtype=np.int32)
        amb = C.c_uint64(0)
        self._check(self._L.hulk_b200_jump_hash_fx(self._ctx, _ptr(keys), keys.size, num_buckets, _ptr(out), C.byref(amb)))
        return out, int(amb.value)

    def rcp_selftest(self, q_begin: int, n: int):
        """Largest relative error of the reciprocal seed and of the refined reciprocal over q in [q_begin, q_begin + n)."""
        out = np.zeros(2, dtype=np.float64)
        self._check(self._L.hulk_b200_rcp_selftest(self._ctx, q_begin, n, _ptr(out)))
        return float(out[0]), float(out[1])

    def folded_table(self) -> np.ndarray:
        stride = C.c_uint64()
        self._check(self._L.hulk_b200_get_folded_table(self._ctx, None, C.byref(stride)))
        out = np.zeros((self.rows, stride.value), dtype=np.float32)
        self._check(self._L.hulk_b200_get_folded_table(self._ctx, _ptr(out), C.byref(stride)))
        return out


class GroupSketch:
    """One sketch over several GPUs of THIS process (hulk_b200_group_*): the HistoSketch calls, reads split into
    contiguous chunks per call, slots sharded, the spectrum summed over NVLink inside every flush."""

    def __init__(self, k: int = 21, w: int = 9, sketch_size: int = 50, decay_ratio: float = 1.0,
                 num_bins: Optional[int] = None, devices: Optional[Sequence[int]] = None, ngpus: Optional[int] = None,
                 tables=None, async_input: bool = False):
        self._L = N.load()
        self._g = C.c_void_p()
        self.k, self.w, self.sketch_size, self.decay_ratio = k, w, sketch_size, decay_ratio
        self.num_bins = spectrum_size(k) if num_bins is None else num_bins
        devs = list(devices) if devices is not None else list(range(ngpus or 1))
        self.ngpus = len(devs)
        p = N.Params()
        p.k, p.w, p.sketch_size = k, w, sketch_size
        p.num_bins = self.num_bins
        p.decay_ratio = decay_ratio
        p.flags = N.F_ASYNC_INPUT if async_input else 0
        ids = (C.c_int32 * len(devs))(*devs)
        g = C.c_void_p()
        rc = self._L.hulk_b200_group_create(C.byref(p), ids, len(devs), C.byref(g))
        if rc:
            raise HulkError(rc, self._L.hulk_b200_group_last_error(None).decode())
        self._g = g
        self._keep = []
        if tables is not None:
            self.set_tables(*tables)

    def close(self):
        if getattr(self, "_g", None):
            self._L.hulk_b200_group_destroy(self._g)
            self._g = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc:
            raise HulkError(rc, self._L.hulk_b200_group_last_error(self._g).decode())

    def set_tables(self, r, c, b):
        """The FULL s x D tables; every member takes its rows."""
        r = np.ascontiguousarray(r, dtype=np.float64)
        c = np.ascontiguousarray(c, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        want = (self.sketch_size, self.num_bins)
        for t in (r, c, b):
            if t.shape != want:
                raise HulkError(N.EARG, f"table shape {t.shape}, expected {want}")
        self._check(self._L.hulk_b200_group_set_cws_tables(self._g, _ptr(r), _ptr(c), _ptr(b)))

    def generate_tables_device(self):
        """Every member draws the rows of its slots on its own GPU (hulk_b200_group_generate_cws_tables_device)."""
        self._check(self._L.hulk_b200_group_generate_cws_tables_device(self._g))

    def generate_tables(self, background: bool = False):
        self._check(self._L.hulk_b200_group_generate_cws_tables(self._g, int(background)))

    def add_reads(self, bases: np.ndarray, offsets: np.ndarray):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self._keep = [bases, offsets]
        self._check(self._L.hulk_b200_group_push_reads(self._g, _ptr(bases), _ptr(offsets), offsets.size - 1))

    def add_seqs(self, reads: Sequence[bytes]):
        self.add_reads(*pack_reads(reads))

    def add_reads_fixed(self, bases: np.ndarray, n_reads: int, read_len: int):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        self._keep = [bases]
        self._check(self._L.hulk_b200_group_push_reads_fixed(self._g, _ptr(bases), n_reads, read_len))

    def flush(self):
        self._check(self._L.hulk_b200_group_flush(self._g))

    def sync(self):
        self._check(self._L.hulk_b200_group_sync(self._g))

    def reset(self):
        self._check(self._L.hulk_b200_group_reset(self._g))

    def finish(self) -> Tuple[np.ndarray, np.ndarray]:
        mins = np.zeros(self.sketch_size, dtype=np.uint64)
        weights = np.zeros(self.sketch_size, dtype=np.float64)
        self._check(self._L.hulk_b200_group_finish(self._g, _ptr(mins), _ptr(weights)))
        return mins, weights

    def stats(self) -> dict:
        st = N.Stats()
        self._check(self._L.hulk_b200_group_get_stats(self._g, C.byref(st)))
        return {n: int(getattr(st, n)) for n, _ in N.Stats._fields_}

    def enable_minhash(self, kmv: bool = True, khf: bool = True):
        self._check(self._L.hulk_b200_group_minhash_enable(self._g, int(kmv), int(khf)))

    def khf(self) -> np.ndarray:
        mins = np.zeros(self.sketch_size, dtype=np.uint64)
        self._check(self._L.hulk_b200_group_get_khf(self._g, _ptr(mins)))
        return mins

    def kmv(self) -> np.ndarray:
        mins = np.zeros(self.sketch_size, dtype=np.uint64)
        n = C.c_uint32(0)
        self._check(self._L.hulk_b200_group_get_kmv(self._g, _ptr(mins), C.byref(n)))
        return mins[:n.value].copy()

    @property
    def concept_drift(self) -> bool:
        return self.decay_ratio != 1.0


def sketch_reads(hs: HistoSketch, batches: Iterable[Tuple[np.ndarray, np.ndarray]], interval: int = 0):
    """SeqMinimizer.Run (src/pipeline/sketch.go:197-224): AddSeq every read, Flush after every
    `interval`-th read (0 = never) and once more at the end, with the race-free semantics
    "every read up to the boundary is counted before the flush".

    batches: iterable of (bases uint8[], offsets uint64[n+1]).  Returns (mins, weights, stats)."""
    seq_count = 0
    for bases, offsets in batches:
        n = len(offsets) - 1
        done = 0
        while done < n:
            take = n - done
            if interval:
                take = min(take, interval - (seq_count % interval))
            hs.add_reads(bases, offsets[done:done + take + 1])
            done += take
            seq_count += take
            if interval and seq_count % interval == 0:
                hs.flush()                                    # sketch.go:211-215
    hs.flush()                                                # sketch.go:221
    mins, weights = hs.finish()
    if seq_count == 0:
        raise HulkError(N.EARG, "no sequences received")     # sketch.go:237-239
    return mins, weights, hs.stats()


def load_sketch(path: str, k: int = 21, algo: str = "histosketch"):
    """sketchio.LoadHULKdata + HULKdata.FindSketch (src/sketchio/sketchio.go:98-260): (mins, weights, banner)."""
    L = N.load()
    f = C.c_void_p()
    err = C.create_string_buffer(1024)
    rc = L.hulk_b200_sketch_load(path.encode(), C.byref(f), err, 1024)
    if rc:
        raise HulkError(rc, err.value.decode(errors="replace"))
    try:
        mins, weights, s = C.c_void_p(), C.c_void_p(), C.c_uint32()
        rc = L.hulk_b200_sketch_find(f, k, algo.encode(), C.byref(mins), C.byref(weights), C.byref(s), err, 1024)
        if rc:
            raise HulkError(rc, err.value.decode(errors="replace"))
        m = np.ctypeslib.as_array(C.cast(mins, C.POINTER(C.c_uint64)), shape=(s.value,)).copy()
        w = (np.ctypeslib.as_array(C.cast(weights, C.POINTER(C.c_double)), shape=(s.value,)).copy()
             if weights.value else np.zeros(0))
        return m, w, L.hulk_b200_sketch_banner(f).decode()
    finally:
        L.hulk_b200_sketch_free(f)


def smash(mins: np.ndarray, weights: Optional[np.ndarray] = None, metric: str = "jaccard", device: int = 0) -> np.ndarray:
    """All-pairs similarity (100 - 100 * distance, cmd/smash.go:183-226) of n sketches: mins uint64[n, s],
    weights float64[n, s] (needed for "weightedjaccard").  Returns float64[n, n], computed on the GPU."""
    L = N.load()
    m = np.ascontiguousarray(mins, dtype=np.uint64)
    n, s = m.shape
    weighted = {"jaccard": 0, "weightedjaccard": 1}.get(metric)
    if weighted is None:
        raise HulkError(N.EARG, "supplied distance metric is not available: %s" % metric)
    w = np.ascontiguousarray(weights, dtype=np.float64) if weights is not None else None
    out = np.zeros((n, n), dtype=np.float64)
    rc = L.hulk_b200_smash(_ptr(m), _ptr(w), n, s, weighted, device, _ptr(out))
    if rc:
        raise HulkError(rc, L.hulk_b200_strerror(rc).decode())
    return out
