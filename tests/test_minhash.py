"""The MinHash side sketches of `hulk sketch --kmv / --khf` (src/minhash), fed from stage 1's minimizers when the
feed is switched on (hulk_b200_minhash_enable); off, the library keeps the reference's behaviour: both stay unfed
(src/pipeline/boss.go:18-19).

CPU: the oracle's restatement of khf.go / kmv.go / heap.go against the package's own test inputs
(src/minhash/minhash_test.go) and against the order-independent definitions the device code uses.
GPU: the CUDA path through the C ABI against that restatement, bit for bit.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, random_reads
from oracle import pyref as P

M64 = (1 << 64) - 1


def khf_numpy(keys: np.ndarray, s: int) -> np.ndarray:
    """KHFsketch over a multiset: slot i = min(hv + i * hv) (uint64 wrap-around), MaxUint64 when nothing was added."""
    out = np.full(s, M64, dtype=np.uint64)
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    if keys.size:
        for i in range(s):
            out[i] = (keys * np.uint64(i + 1)).min()
    return out


def kmv_numpy(keys: np.ndarray, s: int) -> np.ndarray:
    """KMVsketch over a multiset: its s smallest values (duplicates kept), low -> high."""
    return np.sort(np.ascontiguousarray(keys, dtype=np.uint64))[:s]


# ---- CPU ---------------------------------------------------------------------------------------------
def test_minhash_restatement_on_the_reference_test_inputs():
    # src/minhash/minhash_test.go:9-17,20-46: k 7, ten slots, four hash values into two KMV sketches
    hashvalues = [12345, 54321, 9999999, 98765]
    a, b = P.KMVsketch(7, 10), P.KMVsketch(7, 10)
    assert a.s == 10 and a.k == 7                                 # TestMinHashConstructors
    for h in hashvalues:
        a.add_hash(h)
        b.add_hash(h)
    assert a.get_sketch() == sorted(hashvalues)                   # fewer values than slots: all of them, low -> high
    assert a.similarity(b) == 1.0                                 # TestSimilarityEstimates (its 0.5 branch is dead code)
    c = P.KMVsketch(7, 10)
    for h in [12345, 54321, 111111, 222222]:                      # hashvalues2
        c.add_hash(h)
    assert a.similarity(c) == 0.5
    k1, k2 = P.KHFsketch(7, 10), P.KHFsketch(7, 10)
    assert k1.get_sketch() == [M64] * 10                          # khf.go:20-32
    for h in hashvalues:
        k1.add_hash(h)
        k2.add_hash(h)
    assert k1.get_sketch() == [12345 * (i + 1) for i in range(10)]
    assert k1.similarity(k2) == 1.0


@pytest.mark.parametrize("seed,n,s,hi", [(1, 500, 16, 50), (2, 2000, 64, 2 ** 64), (3, 40, 64, 2 ** 64), (4, 300, 8, 3),
                                         (5, 1000, 1, 2 ** 64), (6, 64, 64, 1000)])
def test_minhash_restatement_is_order_independent(seed, n, s, hi):
    # what lets the device compute both per batch and merge: KMV ends up with the s smallest values of the MULTISET
    # (strict '<' against the heap's largest, no duplicate check: kmv.go:57-68), KHF with slot-wise minima
    rng = np.random.default_rng(seed)
    keys = rng.integers(0, hi, n, dtype=np.uint64)
    keys[: n // 10] = rng.integers(2 ** 63, 2 ** 64, n // 10, dtype=np.uint64)        # products that wrap
    kmv, khf = P.KMVsketch(21, s), P.KHFsketch(21, s)
    for x in keys:
        kmv.add_hash(int(x))
        khf.add_hash(int(x))
    assert kmv.multiplicity_sum == n
    np.testing.assert_array_equal(np.array(kmv.get_sketch(), dtype=np.uint64), kmv_numpy(keys, s))
    np.testing.assert_array_equal(np.array(khf.get_sketch(), dtype=np.uint64), khf_numpy(keys, s))
    # KHFsketch.Merge (khf.go:47-55) of two halves = the sketch of the whole
    h1, h2 = P.KHFsketch(21, s), P.KHFsketch(21, s)
    for x in keys[: n // 2]:
        h1.add_hash(int(x))
    for x in keys[n // 2:]:
        h2.add_hash(int(x))
    h1.merge(h2)
    assert h1.get_sketch() == khf.get_sketch()


def test_minhash_golden_of_the_reference_fixture(oracle, fixture_reads):
    # tests/golden/c1_k21_s50_minhash.json (make_golden.py: the literal restatement, read by read) against the
    # order-independent definitions the GPU tests use
    from conftest import GOLDEN
    g = json.load(open(os.path.join(GOLDEN, "c1_k21_s50_minhash.json")))
    keys = np.concatenate([oracle.minimizers(g["k"], g["w"], rd) for rd in fixture_reads]).astype(np.uint64)
    assert keys.size == g["n_added"] == 17040                    # total minimizers of the fixture
    assert kmv_numpy(keys, g["s"]).tolist() == g["kmv"] and khf_numpy(keys, g["s"]).tolist() == g["khf"]
    assert len(set(g["kmv"])) < len(g["kmv"])                    # equal values from different reads are kept (kmv.go:40-71)
    assert P.md5_of_mins(g["kmv"]) == g["kmv_md5"] and P.md5_of_mins(g["khf"]) == g["khf_md5"]


# ---- GPU ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def hb():
    import hulk_b200
    hulk_b200.load()
    return hulk_b200


def _keys(oracle, k, w, reads):
    """the collector's stream (src/pipeline/boss.go:90-95): every member of every read's set"""
    parts = [oracle.minimizers(k, w, r) for r in reads]
    return np.concatenate(parts).astype(np.uint64) if parts else np.zeros(0, np.uint64)


def _tables(s, D, seed):
    rng = np.random.default_rng(seed)
    r = rng.gamma(2.0, 1.0, (s, D))
    return r, np.log(rng.gamma(2.0, 1.0, (s, D))), rng.random((s, D)) * r


@pytest.mark.gpu
def test_unfed_by_default_like_the_reference(hb):
    with hb.HistoSketch(21, 9, 12) as hs:
        hs.add_seqs(random_reads(500, 150, seed=1))
        np.testing.assert_array_equal(hs.khf(), np.full(12, M64, dtype=np.uint64))    # khf.go:20-32, never fed
        assert hs.kmv().size == 0                                                      # an empty heap: sketchio.go:59-61


@pytest.mark.gpu
@pytest.mark.parametrize("k,w,s", [(21, 9, 64), (31, 9, 50), (11, 9, 512), (7, 40, 32), (21, 9, 1)])
def test_fed_sketches_equal_the_restatement(hb, oracle, k, w, s):
    # reads for every stage-1 kernel: the fast scan, k1_generic (several hundred bases), the sliced scan (1024 and more);
    # the same read many times over (equal values at the KMV boundary), N's, three pushes
    base = random_reads(1500, 150, seed=k + s, n_frac=0.01, ragged=100)
    batches = [base + [base[0]] * 40,
               random_reads(30, 500, seed=s, ragged=500) + random_reads(10, 3000, seed=s + 1, ragged=2000)
               + random_reads(2, 30_000, seed=w, n_frac=0.001) + base[:300],
               random_reads(400, 300, seed=3) + [base[1]] * 7]
    batches = [[r for r in b if len(r) >= k + w - 1] for b in batches]
    with hb.HistoSketch(k, w, s) as hs:
        hs.enable_minhash()
        keys = np.zeros(0, np.uint64)
        for b in batches:
            hs.add_seqs(b)
            keys = np.concatenate([keys, _keys(oracle, k, w, b)])
            np.testing.assert_array_equal(hs.khf(), khf_numpy(keys, s))
            np.testing.assert_array_equal(hs.kmv(), kmv_numpy(keys, s))
        assert hs.stats()["n_minimizers"] == keys.size
        # the spectrum is what it is without the feed
        h = hs.histogram()
        ho, _ = oracle.count_reads(k, w, k ** 4, *oracle.pack_reads([r for b in batches for r in b]))
        np.testing.assert_array_equal(h, ho.astype(np.uint32))
        # reset: both back to their constructors' state, and usable again
        hs.reset()
        assert hs.kmv().size == 0 and (hs.khf() == np.uint64(M64)).all()
        hs.add_seqs(batches[2])
        np.testing.assert_array_equal(hs.kmv(), kmv_numpy(_keys(oracle, k, w, batches[2]), s))


@pytest.mark.gpu
def test_fewer_minimizers_than_slots_and_only_one_sketch(hb, oracle):
    reads = [b"ACGTTGCATGCATGCATTACGATCAGCTACGATCAGCATCGACTAGCTA", b"GATTACAGATTACAGGATCCGATTACATTTACGACGATCAGCTACGG"]
    keys = _keys(oracle, 21, 9, reads)
    with hb.HistoSketch(21, 9, 256) as hs:
        hs.enable_minhash(kmv=True, khf=False)
        hs.add_seqs(reads)
        assert keys.size < 256
        np.testing.assert_array_equal(hs.kmv(), np.sort(keys))
        assert (hs.khf() == np.uint64(M64)).all()
    with hb.HistoSketch(21, 9, 256) as hs:
        hs.enable_minhash(kmv=False, khf=True)
        hs.add_seqs(reads)
        np.testing.assert_array_equal(hs.khf(), khf_numpy(keys, 256))
        assert hs.kmv().size == 0
        with pytest.raises(hb.HulkError) as e:                  # the feed is chosen before the first read
            hs.enable_minhash()
        assert e.value.code == -21


@pytest.mark.gpu
def test_fed_sketches_across_intervals_and_input_forms(hb, oracle):
    # flushes move the counting to the next spectrum buffer and its k1 stream: every stream keeps a KMV pool of its own,
    # merged when read; fixed-length and packed batches feed the same queue
    k, w, s = 11, 9, 96
    D = hb.spectrum_size(k)
    tables = _tables(s, D, 3)
    keys = np.zeros(0, np.uint64)
    with hb.HistoSketch(k, w, s, tables=tables) as hs:
        hs.enable_minhash()
        for it in range(5):
            reads = random_reads(3000, 150, seed=100 + it)
            if it == 2:
                hs.add_reads_fixed(np.frombuffer(b"".join(reads), dtype=np.uint8).copy(), len(reads), 150)
            elif it == 3:
                hs.set_input_packing(-1)
                hs.add_seqs(reads)
            else:
                hs.add_seqs(reads)
            hs.flush()
            keys = np.concatenate([keys, _keys(oracle, k, w, reads)])
        mins, _ = hs.finish()
        np.testing.assert_array_equal(hs.khf(), khf_numpy(keys, s))
        np.testing.assert_array_equal(hs.kmv(), kmv_numpy(keys, s))
    with hb.HistoSketch(k, w, s, tables=tables) as ref:          # the histosketch itself does not notice the feed
        for it in range(5):
            ref.add_seqs(random_reads(3000, 150, seed=100 + it))
            ref.flush()
        np.testing.assert_array_equal(ref.finish()[0], mins)


GROUP_SCRIPT = r"""
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import hulk_b200 as hb
from oracle import oracle as O
from conftest import random_reads
from test_minhash import khf_numpy, kmv_numpy, _keys, _tables
k, w, s = 11, 9, 48
tables = _tables(s, hb.spectrum_size(k), 4)
keys = np.zeros(0, np.uint64)
with hb.GroupSketch(k, w, s, 1.0, devices=[0, 0, 0], tables=tables) as g:
    g.enable_minhash()
    for it in range(3):
        reads = random_reads(2500, 150, seed=40 + it, ragged=60)
        g.add_seqs(reads)
        g.flush()
        keys = np.concatenate([keys, _keys(O, k, w, reads)])
    g.sync()
    np.testing.assert_array_equal(g.khf(), khf_numpy(keys, s))
    np.testing.assert_array_equal(g.kmv(), kmv_numpy(keys, s))
print("GROUP_MINHASH_OK")
"""


@pytest.mark.gpu
def test_group_merges_its_members_sketches(oracle):
    # three member contexts on one device (own process: see test_distributed.py), each fed a third of every batch:
    # KHFsketch.Merge slot-wise, KMV as the s smallest of the union
    env = dict(os.environ, CUDA_DEVICE_MAX_CONNECTIONS="32")
    p = subprocess.run([sys.executable, "-c", GROUP_SCRIPT, ROOT], env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "GROUP_MINHASH_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]
