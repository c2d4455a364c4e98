import gzip
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
# before anything creates a CUDA context: one hardware queue per stream (hulk_b200_create explains why the feeder needs it)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def fixture_reads():
    """The reference's own FASTQ fixture (testing/test-reads-small.fq.gz, 1000 x 100 bp)."""
    reads = []
    with gzip.open(os.path.join(GOLDEN, "c1_reads.fq.gz"), "rb") as fh:
        for i, ln in enumerate(fh):
            if i % 4 == 1:
                reads.append(ln.rstrip(b"\n"))
    return reads


def random_reads(n, length, seed, n_frac=0.0, lower_frac=0.0, ragged=0):
    rng = np.random.default_rng(seed)
    out = []
    alpha = np.frombuffer(b"ACGT", dtype=np.uint8)
    for _ in range(n):
        L = length + (int(rng.integers(0, ragged + 1)) if ragged else 0)
        a = alpha[rng.integers(0, 4, L)].copy()
        if n_frac:
            a[rng.random(L) < n_frac] = ord("N")
        if lower_frac:
            m = rng.random(L) < lower_frac
            a[m] = a[m] | 0x20
        out.append(a.tobytes())
    return out
