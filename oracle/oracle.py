"""ctypes loader for the CPU oracle (oracle/hulk_oracle.c, oracle/go_rand.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under hulk_b200/ may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libhulk_oracle.so")

ERRORS = {
    -1: "w must be: 0 < w < 257",
    -2: "k size must be: 0 < k < 32",
    -3: "sequence length must be > 0",
    -4: "sequence length must be >= w + k - 1",
    -5: "oracle buffer too small",
    -6: "not used yet",
    -10: "histosketching only supports k <= 31",
    -11: "decay ratio must be between 0.0 and 1.0",
    -12: "histogram must have at least 2 bins",
}


class OracleError(RuntimeError):
    def __init__(self, code):
        super().__init__(ERRORS.get(code, f"oracle error {code}"))
        self.code = code


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("hulk_oracle.c", "go_rand.c", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    try:
        build()
    except Exception:
        if not os.path.exists(_SO):
            raise
    L = C.CDLL(_SO)
    u8p, u64p, f64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.POINTER(C.c_double)
    L.hulk_oracle_nt4.restype = C.c_uint8
    L.hulk_oracle_nt4.argtypes = [C.c_uint8]
    L.hulk_oracle_hash64.restype = C.c_uint64
    L.hulk_oracle_hash64.argtypes = [C.c_uint64, C.c_uint64]
    L.hulk_oracle_jump.restype = C.c_int32
    L.hulk_oracle_jump.argtypes = [C.c_uint64, C.c_int32]
    L.hulk_oracle_minimizers.restype = C.c_int
    L.hulk_oracle_minimizers.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_int64, C.c_void_p,
                                         C.c_int64, C.POINTER(C.c_int64)]
    L.hulk_oracle_count_reads.restype = C.c_int
    L.hulk_oracle_count_reads.argtypes = [C.c_uint32, C.c_uint32, C.c_int32, C.c_void_p, C.c_void_p,
                                          C.c_int64, C.c_void_p, C.POINTER(C.c_uint64)]
    L.hulk_oracle_hs_new.restype = C.c_int
    L.hulk_oracle_hs_new.argtypes = [C.c_uint32, C.c_uint32, C.c_int32, C.c_double, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    L.hulk_oracle_hs_free.argtypes = [C.c_void_p]
    L.hulk_oracle_hs_add_element.restype = C.c_double
    L.hulk_oracle_hs_add_element.argtypes = [C.c_void_p, C.c_uint64, C.c_double]
    L.hulk_oracle_hs_flush.restype = C.c_int
    L.hulk_oracle_hs_flush.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.hulk_oracle_hs_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.hulk_oracle_hs_get_cms.argtypes = [C.c_void_p, C.c_void_p]
    L.hulk_oracle_run.restype = C.c_int
    L.hulk_oracle_run.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int64,
                                  C.c_uint64, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.hulk_oracle_num_threads.restype = C.c_int
    L.hulk_oracle_set_threads.argtypes = [C.c_int]
    L.go_rand_cooked.argtypes = [C.c_void_p]
    L.go_rand_new.restype = C.c_void_p
    L.go_rand_new.argtypes = [C.c_int64]
    L.go_rand_free.argtypes = [C.c_void_p]
    L.go_rand_int63.restype = C.c_int64
    L.go_rand_int63.argtypes = [C.c_void_p]
    L.go_rand_float64.restype = C.c_double
    L.go_rand_float64.argtypes = [C.c_void_p]
    L.go_rand_intn.restype = C.c_int32
    L.go_rand_intn.argtypes = [C.c_void_p, C.c_int32]
    L.go_rng_gamma_draw.restype = C.c_double
    L.go_rng_gamma_draw.argtypes = [C.c_void_p, C.c_double, C.c_double]
    L.hulk_oracle_new_cws.argtypes = [C.c_uint32, C.c_int32, C.c_uint32, C.c_uint32, C.c_void_p,
                                      C.c_void_p, C.c_void_p]
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def nt4(b: int) -> int:
    return lib().hulk_oracle_nt4(b)


def hash64(key: int, mask: int) -> int:
    return lib().hulk_oracle_hash64(key, mask)


def jump(key: int, n: int) -> int:
    return lib().hulk_oracle_jump(key & 0xFFFFFFFFFFFFFFFF, n)


def minimizers(k: int, w: int, seq: bytes) -> np.ndarray:
    """Per-read minimizer SET in first-insertion order (src/minimizer/minimizer.go:59-204)."""
    buf = np.frombuffer(bytes(seq), dtype=np.uint8) if len(seq) else np.zeros(0, np.uint8)
    out = np.zeros(max(len(seq), 1), dtype=np.uint64)
    n = C.c_int64(0)
    rc = lib().hulk_oracle_minimizers(k, w, _ptr(buf) if len(seq) else None, len(seq), _ptr(out),
                                      out.size, C.byref(n))
    if rc:
        raise OracleError(rc)
    return out[: n.value].copy()


def pack_reads(reads):
    """list[bytes] -> (bases uint8[total], offsets uint64[n+1])"""
    offsets = np.zeros(len(reads) + 1, dtype=np.uint64)
    if reads:
        offsets[1:] = np.cumsum([len(r) for r in reads], dtype=np.uint64)
    bases = np.frombuffer(b"".join(reads), dtype=np.uint8).copy() if reads else np.zeros(0, np.uint8)
    return bases, offsets


def count_reads(k, w, D, bases, offsets, hist=None):
    """Histogram of jump-hashed per-read minimizer sets. Returns (hist float64[D], n_minimizers)."""
    if hist is None:
        hist = np.zeros(D, dtype=np.float64)
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    nm = C.c_uint64(0)
    rc = lib().hulk_oracle_count_reads(k, w, D, _ptr(bases), _ptr(offsets), len(offsets) - 1,
                                       _ptr(hist), C.byref(nm))
    if rc:
        raise OracleError(rc)
    return hist, nm.value


def new_cws(s, D, slot_begin=0, slot_end=None):
    """Go-compatible CWS tables (histosketch.go:95-126): returns r, c, b float64[(rows), D]."""
    slot_end = s if slot_end is None else slot_end
    rows = slot_end - slot_begin
    r = np.zeros((rows, D)); c = np.zeros((rows, D)); b = np.zeros((rows, D))
    lib().hulk_oracle_new_cws(s, D, slot_begin, slot_end, _ptr(r), _ptr(c), _ptr(b))
    return r, c, b


class HistoSketch:
    """Mirror of histosketch.HistoSketch (src/histosketch/histosketch.go:36-170)."""

    def __init__(self, k, s, D, decay, r, c, b):
        self.k, self.s, self.D, self.decay = k, s, D, decay
        self._r = np.ascontiguousarray(r, dtype=np.float64)
        self._c = np.ascontiguousarray(c, dtype=np.float64)
        self._b = np.ascontiguousarray(b, dtype=np.float64)
        h = C.c_void_p()
        rc = lib().hulk_oracle_hs_new(k, s, D, float(decay), _ptr(self._r), _ptr(self._c),
                                      _ptr(self._b), C.byref(h))
        if rc:
            raise OracleError(rc)
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            lib().hulk_oracle_hs_free(self._h)
            self._h = None

    def add_element(self, bin_id, value):
        return lib().hulk_oracle_hs_add_element(self._h, int(bin_id), float(value))

    def flush(self, hist, parallel=False):
        """hist float64[D] is consumed (wiped).  Returns count-min estimates f (nan = empty bin)."""
        assert hist.dtype == np.float64 and hist.size == self.D
        f = np.full(self.D, np.nan)
        rc = lib().hulk_oracle_hs_flush(self._h, _ptr(hist), _ptr(f), int(parallel))
        if rc:
            raise OracleError(rc)
        return f

    def run(self, w, bases, offsets, interval=0, parallel=False):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        nm, nf = C.c_uint64(0), C.c_uint64(0)
        rc = lib().hulk_oracle_run(self._h, w, _ptr(bases), _ptr(offsets), len(offsets) - 1,
                                   int(interval), int(parallel), C.byref(nm), C.byref(nf))
        if rc:
            raise OracleError(rc)
        return nm.value, nf.value

    def get(self):
        mins = np.zeros(self.s, dtype=np.uint64)
        weights = np.zeros(self.s, dtype=np.float64)
        lib().hulk_oracle_hs_get(self._h, _ptr(mins), _ptr(weights))
        return mins, weights

    def cms(self):
        q = np.zeros((7, 2000))
        lib().hulk_oracle_hs_get_cms(self._h, _ptr(q))
        return q


class GoRand:
    """rand.New(rand.NewSource(seed)) of Go's math/rand."""

    def __init__(self, seed=1):
        self._p = C.c_void_p(lib().go_rand_new(seed))

    def __del__(self):
        if getattr(self, "_p", None):
            lib().go_rand_free(self._p)
            self._p = None

    def int63(self):
        return lib().go_rand_int63(self._p)

    def float64(self):
        return lib().go_rand_float64(self._p)

    def intn(self, n):
        return lib().go_rand_intn(self._p, n)

    def gamma(self, alpha, beta):
        return lib().go_rng_gamma_draw(self._p, alpha, beta)


def rng_cooked():
    out = np.zeros(607, dtype=np.uint64)
    lib().go_rand_cooked(_ptr(out))
    return out


def num_threads():
    return lib().hulk_oracle_num_threads()


def set_threads(n):
    lib().hulk_oracle_set_threads(int(n))
