"""Input side of the pipeline: the reference's line reader + FASTQ/FASTA framing, and the
synthetic read generator of BASELINE.md §3.

Framing quirks kept from the reference (src/pipeline/sketch.go:40-161, src/seqio/seqio.go:37-48):
  * lines come from bufio.Scanner: split on "\\n", one trailing "\\r" dropped, final line without
    newline still delivered; `append([]byte(nil), line...)` turns an EMPTY line into nil, and the
    4-slot filler `if l1 == nil ... else if l2 == nil ...` therefore re-fills the same slot;
  * only `l1[0] == '@'` is validated; the quality line is ignored;
  * FASTA mode stops at the first empty line and concatenates sequence lines.
"""
from __future__ import annotations

import gzip
from typing import Iterator, List

import numpy as np


def _lines(path: str) -> Iterator[bytes]:
    opener = gzip.open if path.split(".")[-1] == "gz" else open        # sketch.go:62-70
    with opener(path, "rb") as fh:
        data = fh.read()
    if not data:
        return
    parts = data.split(b"\n")
    if parts and parts[-1] == b"":
        parts.pop()
    for ln in parts:
        if ln.endswith(b"\r"):
            ln = ln[:-1]
        if len(ln) >= 64 * 1024:          # bufio.MaxScanTokenSize: the line and its "\n" must fit the buffer
            raise ValueError("bufio.Scanner: token too long")
        yield ln


def read_fastq(path: str) -> List[bytes]:
    """FastqHandler.Run, FASTQ branch (src/pipeline/sketch.go:139-159): sequences only."""
    reads: List[bytes] = []
    slot = [None, None, None, None]
    for ln in _lines(path):
        line = ln if len(ln) else None
        for i in range(4):
            if slot[i] is None:
                slot[i] = line
                break
        if slot[3] is not None:
            if slot[0][0] != 64:
                raise ValueError("read ID in fastq file does not begin with @: %s" % slot[0].decode(errors="replace"))
            reads.append(slot[1])
            slot = [None, None, None, None]
    return reads


def read_fasta(path: str) -> List[bytes]:
    """FastqHandler.Run, FASTA branch (src/pipeline/sketch.go:102-135)."""
    reads: List[bytes] = []
    header, seq = None, None
    for ln in _lines(path):
        if len(ln) == 0:
            break
        if ln[0] == 62:
            if header is not None:
                reads.append(seq if seq is not None else b"")
            header, seq = ln, None
        else:
            seq = (seq or b"") + ln
    if header is None:
        raise ValueError("no FASTA record")      # the reference panics on l1[0] of a nil slice
    reads.append(seq if seq is not None else b"")
    return reads


class NativeReader:
    """The library's threaded line reader + FASTQ/FASTA framing (hulk_b200_reader_*, csrc/ingest.cpp):
    what the `hulk` front end feeds the GPU from.  Iterating yields (bases uint8[], offsets uint64[n+1])
    batches in input order; the arrays alias the reader's pinned ring and are valid until the next batch."""

    def __init__(self, paths, fasta: bool = False, batch_bytes: int = 0):
        import ctypes as C
        from . import _native as N
        self._C, self._N, self._L = C, N, N.load()
        arr = (C.c_char_p * max(1, len(paths)))(*[p.encode() for p in paths])
        self._rd = C.c_void_p()
        rc = self._L.hulk_b200_reader_open(arr, len(paths), int(fasta), batch_bytes, C.byref(self._rd))
        if rc:
            raise OSError(self._L.hulk_b200_strerror(rc).decode())

    def __iter__(self):
        C = self._C
        while True:
            bases, offsets, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
            rc = self._L.hulk_b200_reader_next(self._rd, C.byref(bases), C.byref(offsets), C.byref(n))
            if rc:
                raise ValueError(self._L.hulk_b200_reader_error(self._rd).decode())
            if n.value == 0:
                return
            offs = np.ctypeslib.as_array(C.cast(offsets, C.POINTER(C.c_uint64)), shape=(n.value + 1,))
            nb = int(offs[-1])
            b = (np.ctypeslib.as_array(C.cast(bases, C.POINTER(C.c_uint8)), shape=(nb,)) if nb
                 else np.zeros(0, np.uint8))
            yield b, offs

    def reads(self) -> List[bytes]:
        out: List[bytes] = []
        for b, offs in self:
            raw = b.tobytes()
            out.extend(raw[int(offs[i]):int(offs[i + 1])] for i in range(len(offs) - 1))
        return out

    def handle(self):
        return self._rd

    def close(self):
        if getattr(self, "_rd", None):
            self._L.hulk_b200_reader_close(self._rd)
            self._rd = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def synthetic_reads(n_reads: int, read_len: int = 150, seed: int = 1, first_read: int = 0) -> np.ndarray:
    """Counter-based synthetic reads: uint8[n_reads, read_len] over "ACGT".

    word = splitmix64(seed XOR (read_idx * ceil(L/32) + j // 32)); base j = "ACGT"[(word >> 2*(j % 32)) & 3]
    """
    wpr = (read_len + 31) // 32
    idx = (np.arange(first_read, first_read + n_reads, dtype=np.uint64)[:, None] * np.uint64(wpr)
           + np.arange(wpr, dtype=np.uint64)[None, :])
    words = _splitmix64(np.uint64(seed) ^ idx)                                  # [n, wpr]
    shifts = (np.arange(32, dtype=np.uint64) * np.uint64(2))[None, None, :]
    codes = ((words[:, :, None] >> shifts) & np.uint64(3)).astype(np.uint8)     # [n, wpr, 32]
    codes = codes.reshape(n_reads, wpr * 32)[:, :read_len]
    return np.frombuffer(b"ACGT", dtype=np.uint8)[codes]
