"""Pins the CPU oracle (oracle/hulk_oracle.c) before anything is compared against it.

Sources of truth, in order: the reference's own tests (src/minimizer/minimizer_test.go,
src/kmerspectrum/kmerspectrum_test.go, src/helpers/helpers_test.go), the published vectors of the
un-vendored third-party algorithms (jump consistent hash), the reference's FASTQ fixture, and a
second independent restatement (oracle/pyref.py).  The reference has no golden values for this
path, so fixture anchors are oracle-derived and cross-checked.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, random_reads
from oracle import pyref as P


# ---- reference unit tests restated -------------------------------------------------------------
def test_nt4_table_matches_reference_test(oracle):
    # src/minimizer/minimizer_test.go:14-30: "ACGTN" -> 0,1,2,3,4
    assert [oracle.nt4(b) for b in b"ACGTN"] == [0, 1, 2, 3, 4]
    # full table (src/minimizer/minimizer.go:13-30)
    for b in range(256):
        exp = b if b < 4 else {"A": 0, "C": 1, "G": 2, "T": 3, "U": 3}.get(chr(b).upper(), 4)
        assert oracle.nt4(b) == exp == P.nt4(b), b


def test_minimizer_sketch_is_deterministic_like_reference_test(oracle):
    # src/minimizer/minimizer_test.go:32-61: NewMinimizerSketch(4, 4, "ACTGAAAATTTT") twice -> same set
    a = set(oracle.minimizers(4, 4, b"ACTGAAAATTTT").tolist())
    b = set(oracle.minimizers(4, 4, b"ACTGAAAATTTT").tolist())
    assert a == b and len(a) > 0
    assert a == P.minimizers(4, 4, b"ACTGAAAATTTT")


def test_kmerspectrum_reference_test(oracle):
    # src/kmerspectrum/kmerspectrum_test.go:37-44: AddHash(1), AddHash(1234) into 10 bins -> cardinality 1 then 2
    assert oracle.jump(1, 10) != oracle.jump(1234, 10)


def test_pow_reference_test():
    # src/helpers/helpers_test.go:7-17 and cmd/sketch.go:118
    import hulk_b200
    assert [hulk_b200.spectrum_size(k) for k in (11, 21, 31)] == [14641, 194481, 923521]


# ---- third-party arithmetic ---------------------------------------------------------------------
JUMP_VECTORS = [(1, 1, 0), (42, 57, 43), (0xDEAD10CC, 1, 0), (0xDEAD10CC, 666, 361), (256, 1024, 520)]


def test_jump_hash_published_vectors(oracle):
    for key, n, want in JUMP_VECTORS:
        assert oracle.jump(key, n) == want
        assert P.jump(key, n) == want


def test_jump_hash_consistency_property(oracle):
    # growing the bucket count only ever moves a key to the NEW bucket (Lamping & Veach)
    rng = np.random.default_rng(7)
    for key in rng.integers(0, 2 ** 63, 200, dtype=np.uint64):
        prev = oracle.jump(int(key), 1)
        for n in (2, 3, 10, 100, 2000, 14641, 194481, 923521):
            cur = oracle.jump(int(key), n)
            assert cur == prev or prev < cur < n
            assert 0 <= cur < n
            prev = cur


def test_hash64_anchors_and_invertibility(oracle):
    m21, m31 = (1 << 42) - 1, (1 << 62) - 1
    assert oracle.hash64(0, m21) == 0x1DF06F29BC0
    assert oracle.hash64(1, m21) == 0x69B794F8CE
    assert oracle.hash64(0x123456789, m21) == 0x1AC74BC9DE6
    assert oracle.hash64(m21, m21) == 0xDDF0B551BF
    assert oracle.hash64(0x2AAAAAAAAAAAAAAA & m31, m31) == 0xE9193A09D1C5869
    rng = np.random.default_rng(3)
    for k in (4, 11, 21, 31):
        mask = (1 << (2 * k)) - 1
        keys = [int(x) & mask for x in rng.integers(0, 2 ** 63, 2000, dtype=np.uint64)]
        hs = [oracle.hash64(x, mask) for x in keys]
        assert hs == [P.hash64(x, mask) for x in keys]
        assert len(set(hs)) == len(set(keys))          # a bijection on [0, 4^k)


# ---- minimizers: C oracle vs the independent restatement ----------------------------------------
@pytest.mark.parametrize("k,w", [(21, 9), (31, 9), (11, 9), (4, 4), (21, 1), (15, 32), (5, 9), (7, 40), (3, 7)])
def test_minimizers_two_restatements_agree(oracle, k, w):
    reads = random_reads(40, 90, seed=k * 100 + w, n_frac=0.03, lower_frac=0.2, ragged=60)
    reads += [b"A" * 120, b"ACGT" * 40, b"N" * 100, bytes(range(256)), b"AC" * 70, b"acgtnACGTN" * 13]
    for r in reads:
        if len(r) < k + w - 1:
            continue
        assert set(oracle.minimizers(k, w, r).tolist()) == P.minimizers(k, w, r), (k, w, r)


def test_minimizer_errors(oracle):
    from oracle.oracle import OracleError
    for k, w, seq, code in [(21, 257, b"A" * 400, -1), (32, 9, b"A" * 100, -2), (21, 9, b"", -3),
                            (21, 9, b"A" * 28, -4)]:
        with pytest.raises(OracleError) as e:
            oracle.minimizers(k, w, seq)
        assert e.value.code == code
    assert len(oracle.minimizers(21, 9, b"ACGTTGCAAC" * 3)[:1]) >= 0     # exactly k+w-1 = 29 < 30 ok


def test_kmer_span_low_byte(oracle):
    # minimizer.go:127-131,157: low byte of X is i-w+2 for the first w-1 k-mers, then k
    k, w = 21, 9
    read = random_reads(1, 150, seed=5)[0]
    xs = P.position_values(k, w, read)
    spans = [x & 0xFF for x in xs if x is not None]
    assert spans[:w - 1] == list(range(k - w + 1, k)) and set(spans[w - 1:]) == {k}


# ---- the reference fixture (config C1) ----------------------------------------------------------
def test_fixture_histogram_anchors(oracle, fixture_reads):
    assert len(fixture_reads) == 1000 and all(len(r) == 100 for r in fixture_reads)
    D = 21 ** 4
    hist, nmin = oracle.count_reads(21, 9, D, *oracle.pack_reads(fixture_reads))
    h32 = hist.astype(np.uint32)
    anchors = json.load(open(os.path.join(GOLDEN, "c1_anchors.json")))
    assert nmin == anchors["n_minimizers"] == 17040
    assert int((h32 != 0).sum()) == anchors["used_bins"] == 12212
    assert int(h32.max()) == 15 and int(h32.argmax()) == 2873
    assert hashlib.md5(h32.tobytes()).hexdigest() == anchors["hist_md5"] == "92e0141ea84fe70e74350302b24402a2"
    # independent restatement on a slice of the fixture
    hp = P.histogram(21, 9, D, fixture_reads[:60])
    hc, _ = oracle.count_reads(21, 9, D, *oracle.pack_reads(fixture_reads[:60]))
    assert (np.array(hp) == hc).all()
    # CMS columns of one bin (countmin.go:122-125)
    assert [oracle.jump(12345 * (d + 1), 2000) for d in range(7)] == [1873, 676, 1031, 1236, 1559, 1916, 580]


# ---- count-min + CWS: C oracle vs pyref ----------------------------------------------------------
def _tables(s, D, seed):
    rng = np.random.default_rng(seed)
    r = rng.gamma(2.0, 1.0, (s, D))
    c = np.log(rng.gamma(2.0, 1.0, (s, D)))
    b = rng.random((s, D)) * r
    return r, c, b


@pytest.mark.parametrize("decay", [1.0, 0.02, 0.5, 0.0])
def test_histosketch_two_restatements_agree(oracle, decay):
    k, s, D = 5, 6, 5 ** 4
    r, c, b = _tables(s, D, 11)
    hs = oracle.HistoSketch(k, s, D, decay, r, c, b)
    hp = P.HistoSketch(k, s, D, decay, r, c, b)
    rng = np.random.default_rng(2)
    for flush in range(3):
        hist = rng.integers(0, 4, D).astype(np.float64)
        hp.flush(hist.tolist())
        hs.flush(hist.copy())
        mins, weights = hs.get()
        assert mins.tolist() == hp.sketch
        np.testing.assert_array_equal(weights, np.array(hp.weights))


@pytest.mark.parametrize("decay", [1.0, 0.3])
def test_parallel_flush_is_bit_identical_to_literal_loop(oracle, decay):
    k, s, D = 7, 40, 7 ** 4
    r, c, b = _tables(s, D, 5)
    a = oracle.HistoSketch(k, s, D, decay, r, c, b)
    bb = oracle.HistoSketch(k, s, D, decay, r, c, b)
    rng = np.random.default_rng(9)
    for _ in range(3):
        hist = rng.integers(0, 3, D).astype(np.float64)
        fa = a.flush(hist.copy(), parallel=False)
        fb = bb.flush(hist.copy(), parallel=True)
        np.testing.assert_array_equal(fa, fb)
        np.testing.assert_array_equal(a.get()[0], bb.get()[0])
        np.testing.assert_array_equal(a.get()[1], bb.get()[1])
        np.testing.assert_array_equal(a.cms(), bb.cms())


def test_flush_semantics(oracle):
    from oracle.oracle import OracleError
    k, s, D = 5, 4, 5 ** 4
    r, c, b = _tables(s, D, 1)
    hs = oracle.HistoSketch(k, s, D, 1.0, r, c, b)
    empty = np.zeros(D)
    hs.flush(empty)                                   # boss.go:117: empty spectrum -> no-op
    assert (hs.get()[1] == 1.7976931348623157e308).all() and (hs.get()[0] == 0).all()
    sparse = np.zeros(D)
    sparse[:6] = 1                                    # 6/625 < 1 %
    with pytest.raises(OracleError) as e:
        hs.flush(sparse)
    assert e.value.code == -6                         # kmerspectrum.go:94-96 "not used yet"
    for bad in (-0.1, 1.5):
        with pytest.raises(OracleError):
            oracle.HistoSketch(k, s, D, bad, r, c, b)
    with pytest.raises(OracleError):
        oracle.HistoSketch(k, s, 1, 1.0, r, c, b)


def test_run_intervals(oracle):
    # src/pipeline/sketch.go:197-224: flush every `interval` reads and once at the end
    k, w, s = 7, 5, 8
    D = k ** 4
    r, c, b = _tables(s, D, 4)
    reads = random_reads(300, 80, seed=8)
    bases, offs = oracle.pack_reads(reads)
    hs = oracle.HistoSketch(k, s, D, 1.0, r, c, b)
    nmin, nfl = hs.run(w, bases, offs, interval=100)
    assert nfl == 4                                  # 3 interval flushes + the (empty) final one
    # the same thing by hand
    hs2 = oracle.HistoSketch(k, s, D, 1.0, r, c, b)
    tot = 0
    for i in range(3):
        hist, n = oracle.count_reads(k, w, D, bases, offs[i * 100:i * 100 + 101])
        tot += n
        hs2.flush(hist)
    assert tot == nmin
    np.testing.assert_array_equal(hs.get()[0], hs2.get()[0])
    np.testing.assert_array_equal(hs.get()[1], hs2.get()[1])


def test_long_read_set_path_equals_python_restatement(oracle):
    # reads longer than 8192 bases take the oracle's hashed per-read set (same set as the linear scan)
    from oracle import pyref as P
    rng = np.random.default_rng(5)
    for L in (9000, 20000):
        r = np.frombuffer(b"ACGTN", dtype=np.uint8)[rng.choice(5, L, p=[.24, .25, .25, .25, .01])].tobytes() + b"ACGT" * 300
        a = oracle.minimizers(21, 9, r)
        assert len(a) == len(set(int(x) for x in a))
        assert set(int(x) for x in a) == set(P.minimizers(21, 9, r))
