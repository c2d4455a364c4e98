"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Bars: bit-exact for minimizer values, jump-hash bins, histogram counts, count-min estimates without
decay and sketch `mins`; float64 sketch weights within 1e-12 relative (north_star asks 1e-5);
count-min estimates under decay within 1e-9 relative (pow(w, n) instead of n repeated products).
"""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, random_reads

pytestmark = pytest.mark.gpu

W_RTOL = 1e-12


def _tables(s, D, seed):
    rng = np.random.default_rng(seed)
    r = rng.gamma(2.0, 1.0, (s, D))
    c = np.log(rng.gamma(2.0, 1.0, (s, D)))
    b = rng.random((s, D)) * r
    return r, c, b


@pytest.fixture(scope="module")
def hb():
    import hulk_b200
    hulk_b200.load()
    return hulk_b200


# ---- stage 1: minimizer sets ---------------------------------------------------------------------
EDGE_READS = [b"A" * 150, b"ACGT" * 40, b"N" * 100, bytes(range(256)), b"AC" * 75, b"acgtnACGTN" * 15,
              b"ACGTTGCATGCATGCATTACGATCAGCTACGATCAGCATCGACTAGCTANNNNNNACGATCGACTAGCTAGCATCGATCAGCTAGCTAGCATGC",
              b"GATTACA" * 30, b"T" * 60 + b"N" + b"T" * 60, b"ACGU" * 40, b"\x00\x01\x02\x03" * 40]


@pytest.mark.parametrize("k,w", [(21, 9), (31, 9), (11, 9), (4, 4), (21, 1), (15, 32), (5, 9), (3, 7), (7, 40), (21, 200)])
def test_minimizer_sets_bit_exact(hb, oracle, k, w):
    reads = random_reads(300, 100, seed=k * 1000 + w, n_frac=0.02, lower_frac=0.1, ragged=120) + EDGE_READS
    reads = [r for r in reads if len(r) >= k + w - 1]
    with hb.HistoSketch(k, w, 4) as hs:
        sets, counts = hs.minimizers(reads)
    for r, got, n in zip(reads, sets, counts):
        want = np.sort(oracle.minimizers(k, w, r))
        assert n == want.size, (k, w, r)
        np.testing.assert_array_equal(got, want)


def test_minimizer_sets_on_reference_fixture(hb, oracle, fixture_reads):
    with hb.HistoSketch(21, 9, 4) as hs:
        sets, counts = hs.minimizers(fixture_reads)
    assert int(counts.sum()) == 17040
    for r, got in zip(fixture_reads, sets):
        np.testing.assert_array_equal(got, np.sort(oracle.minimizers(21, 9, r)))


def test_long_reads_take_the_generic_path(hb, oracle):
    # candidate lists longer than the shared-memory capacity are queued for k1_generic
    reads = random_reads(40, 3000, seed=77, n_frac=0.001, ragged=4000) + [b"ACGTAC" * 2000, b"A" * 5000]
    reads += random_reads(100, 150, seed=78)
    with hb.HistoSketch(21, 9, 4) as hs:
        sets, counts = hs.minimizers(reads, cap=3000)
        for r, got in zip(reads, sets):
            np.testing.assert_array_equal(got, np.sort(oracle.minimizers(21, 9, r)))
        hs.add_seqs(reads)
        h = hs.histogram()
    ho, nm = oracle.count_reads(21, 9, 21 ** 4, *oracle.pack_reads(reads))
    np.testing.assert_array_equal(h, ho.astype(np.uint32))


def test_many_reads_of_a_few_hundred_bases(hb, oracle):
    # tens of thousands of reads whose candidate lists overflow (merged pairs, 454, short nanopore): every k1_generic
    # thread reuses its own table, so the number of such reads in a batch does not matter (round 2 ran out of scratch
    # memory at ~33 000 of them)
    reads = random_reads(45_000, 400, seed=91, n_frac=0.001, ragged=800) + random_reads(2000, 150, seed=92)
    with hb.HistoSketch(21, 9, 4) as hs:
        hs.add_seqs(reads)
        h = hs.histogram()
        st = hs.stats()
    ho, nm = oracle.count_reads(21, 9, 21 ** 4, *oracle.pack_reads(reads))
    np.testing.assert_array_equal(h, ho.astype(np.uint32))
    assert st["n_minimizers"] == nm


@pytest.mark.parametrize("k,w", [(21, 9), (31, 9), (11, 9), (7, 40), (21, 200), (30, 256)])
def test_long_sequences_take_the_sliced_scan(hb, oracle, k, w):
    # sequences of 1024 bases or more (--fasta contigs, long reads) are cut into slices that share one set per
    # sequence (k1_long.cuh); next to them: reads for the fast kernels and reads for k1_generic, N runs, lower case,
    # a sequence of exactly the threshold length and one a base short of it
    reads = (random_reads(3, 40_000, seed=k + w, n_frac=0.002, lower_frac=0.1, ragged=30_000)
             + [random_reads(1, 1024, seed=5)[0], random_reads(1, 1023, seed=6)[0], random_reads(1, 16_384, seed=4)[0],
                b"ACGTAC" * 5000,
                b"N" * 9000 + random_reads(1, 12_000, seed=7)[0] + b"N" * 300 + b"acgt" * 2000, b"A" * 20_000]
             + random_reads(50, 400, seed=8, ragged=3000) + random_reads(100, 300, seed=9))
    cap = 70_000
    with hb.HistoSketch(k, w, 4) as hs:
        sets, counts = hs.minimizers(reads, cap=cap)
        for r, got, n in zip(reads, sets, counts):
            want = np.sort(oracle.minimizers(k, w, r))
            assert n == want.size, (k, w, len(r))
            np.testing.assert_array_equal(got, want)
        hs.add_seqs(reads)
        hs.add_seqs(reads[:2])                       # a second batch: tables and slice numbers start over
        h = hs.histogram()
        st = hs.stats()
    ho, nm = oracle.count_reads(k, w, k ** 4, *oracle.pack_reads(reads + reads[:2]))
    np.testing.assert_array_equal(h, ho.astype(np.uint32))
    assert st["n_minimizers"] == nm


def test_long_sequences_fixed_length_and_device_offsets(hb, oracle):
    # the two other ways a batch arrives: fixed-length reads (the host knows the one length) and offsets that only
    # exist on the device (the host cannot rule long sequences out, so the sliced scan's launches are made)
    import torch
    reads = random_reads(6, 20_000, seed=31, n_frac=0.001)
    ho, nm = oracle.count_reads(21, 9, 21 ** 4, *oracle.pack_reads(reads))
    with hb.HistoSketch(21, 9, 4) as hs:
        hs.add_reads_fixed(np.frombuffer(b"".join(reads), dtype=np.uint8).copy(), len(reads), 20_000)
        np.testing.assert_array_equal(hs.histogram(), ho.astype(np.uint32))
        assert hs.stats()["n_minimizers"] == nm
    ragged = random_reads(4, 18_000, seed=32, ragged=9000) + random_reads(200, 150, seed=33, ragged=100)
    bases, offsets = oracle.pack_reads(ragged)
    ho, nm = oracle.count_reads(21, 9, 21 ** 4, bases, offsets)
    with hb.HistoSketch(21, 9, 4) as hs:
        d_b = torch.from_numpy(np.concatenate([bases, np.zeros(64, np.uint8)])).cuda()
        d_o = torch.from_numpy(offsets.astype(np.int64)).cuda()
        torch.cuda.synchronize()
        hs.add_reads_device(d_b.data_ptr(), d_o.data_ptr(), len(ragged), 0)
        np.testing.assert_array_equal(hs.histogram(), ho.astype(np.uint32))
        assert hs.stats()["n_minimizers"] == nm


def test_chromosome_sized_sequence(hb, oracle):
    # --fasta mode hands whole contigs to AddSeq: a 20 Mbp sequence needs a 2^26-entry table, more than the
    # default scratch arena holds next to its neighbours, and must still give the reference's set
    rng = np.random.default_rng(77)
    big = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 20_000_000)].tobytes()
    reads = [big, b"ACGTTGCA" * 40, big[:3_000_000]]
    with hb.HistoSketch(21, 9, 4) as hs:
        hs.add_seqs(reads)
        h = hs.histogram()
        st = hs.stats()
    ho, nm = oracle.count_reads(21, 9, 21 ** 4, *oracle.pack_reads(reads))
    np.testing.assert_array_equal(h, ho.astype(np.uint32))
    assert st["n_minimizers"] == nm and st["n_bases"] == sum(len(r) for r in reads)


def test_read_length_errors_mirror_reference(hb):
    with hb.HistoSketch(21, 9, 4) as hs:
        with pytest.raises(hb.HulkError) as e:
            hs.minimizers([b"ACGT" * 30, b"ACGTACGTACGTACGTACGTACGTACGT"])      # 28 < k + w - 1 = 29
        assert e.value.code == -4 and "read 1" in str(e.value)
        with pytest.raises(hb.HulkError) as e:
            hs.minimizers([b"ACGT" * 30, b""])
        assert e.value.code == -3
    r, c, b = _tables(4, 5 ** 4, 1)
    with hb.HistoSketch(5, 4, 4, tables=(r, c, b)) as hs:
        hs.add_seqs([b"ACGTACGTAC", b"ACGTACG"])                                # second read: 7 < 8
        hs.flush()
        with pytest.raises(hb.HulkError) as e:
            hs.finish()
        assert e.value.code == -4


# ---- stage 2: jump hash + histogram --------------------------------------------------------------
def test_jump_hash_bit_exact(hb, oracle):
    rng = np.random.default_rng(5)
    keys = np.concatenate([rng.integers(0, 2 ** 64, 20000, dtype=np.uint64),
                           np.array([0, 1, 42, 256, 0xDEAD10CC, 2 ** 64 - 1, 2 ** 63, 2 ** 33, 2 ** 33 - 1], dtype=np.uint64)])
    with hb.HistoSketch(21, 9, 4) as hs:
        for n in (1, 2, 10, 57, 666, 1024, 2000, 14641, 194481, 923521, 2 ** 31 - 1):
            got = hs.jump_hash(keys, n)
            want = np.array([oracle.jump(int(x), n) for x in keys], dtype=np.int32)
            np.testing.assert_array_equal(got, want)
        assert hs.jump_hash(np.array([1, 42, 0xDEAD10CC, 0xDEAD10CC, 256], dtype=np.uint64), 1).tolist() == [0] * 5
        assert int(hs.jump_hash(np.array([42], dtype=np.uint64), 57)[0]) == 43
        assert int(hs.jump_hash(np.array([0xDEAD10CC], dtype=np.uint64), 666)[0]) == 361
        assert int(hs.jump_hash(np.array([256], dtype=np.uint64), 1024)[0]) == 520


def test_fixture_histogram_matches_anchor(hb, fixture_reads):
    anchors = json.load(open(os.path.join(GOLDEN, "c1_anchors.json")))
    with hb.HistoSketch(21, 9, 4) as hs:
        hs.add_seqs(fixture_reads)
        h = hs.histogram()
        st = hs.stats()
    assert hashlib.md5(h.tobytes()).hexdigest() == anchors["hist_md5"]
    assert st["n_minimizers"] == anchors["n_minimizers"] and st["n_reads"] == 1000 and st["n_bases"] == 100000


@pytest.mark.parametrize("k,w,L", [(21, 9, 150), (31, 9, 150), (11, 9, 150), (21, 9, 250), (15, 5, 100), (21, 16, 151)])
def test_histogram_bit_exact_synthetic(hb, oracle, k, w, L):
    n = 20000
    reads = hb.synthetic_reads(n, L, seed=2)
    D = hb.spectrum_size(k)
    offs = (np.arange(n + 1, dtype=np.uint64) * np.uint64(L))
    ho, nm = oracle.count_reads(k, w, D, reads.reshape(-1), offs)
    with hb.HistoSketch(k, w, 4) as hs:
        hs.add_reads_fixed(reads.reshape(-1), n, L)
        h1 = hs.histogram()
        assert hs.stats()["n_minimizers"] == nm
    np.testing.assert_array_equal(h1, ho.astype(np.uint32))
    # ragged entry point, pushed in three uneven batches: the histogram is a plain sum
    with hb.HistoSketch(k, w, 4) as hs:
        for a, z in ((0, 7), (7, 12345), (12345, n)):
            hs.add_reads(reads.reshape(-1), offs[a:z + 1])
        np.testing.assert_array_equal(hs.histogram(), h1)


def test_jump_hash_fixed_point_step_bit_exact(hb, oracle):
    """The binning kernel's step for num_buckets <= 2^20 (hd_math.h jump_step_fx) against the oracle, and on two million
    keys against the device's literal jump.Hash (itself pinned above); the true-division fallback must stay rare."""
    rng = np.random.default_rng(6)
    keys = np.concatenate([rng.integers(0, 2 ** 64, 20000, dtype=np.uint64),
                           np.array([0, 1, 42, 256, 0xDEAD10CC, 2 ** 64 - 1, 2 ** 63, 2 ** 33, 2 ** 33 - 1], dtype=np.uint64)])
    big = rng.integers(0, 2 ** 64, 2_000_000, dtype=np.uint64)
    with hb.HistoSketch(21, 9, 4) as hs:
        for n in (1, 2, 10, 57, 666, 1024, 2000, 14641, 194481, 923521, 2 ** 20):
            got, amb = hs.jump_hash_fx(keys, n)
            want = np.array([oracle.jump(int(x), n) for x in keys], dtype=np.int32)
            np.testing.assert_array_equal(got, want)
        for n in (194481, 923521, 2 ** 20, 4 ** 10 - 1):
            got, amb = hs.jump_hash_fx(big, n)
            np.testing.assert_array_equal(got, hs.jump_hash(big, n))
            assert amb < big.size * 20 * 2.0 ** -15, amb        # ~2^-18 per step expected
        with pytest.raises(hb.HulkError):
            hs.jump_hash_fx(keys, 2 ** 20 + 1)


def test_reciprocal_error_budget_of_the_jump_step(hb):
    """EVERY divisor the Lamping-Veach step can meet (q = 1 .. 2^31): the hardware seed is good to 2^-19.9 and one Newton
    step brings it to 2^-39.8 relative (measured on B200: 2^-19.94 and 2^-39.88) -- the figures the fixed-point step's
    ambiguity band (3 units of 2^-20 at x < 2^20) and the bracket's epsilon (2^-39) are sized for."""
    with hb.HistoSketch(21, 9, 4) as hs:
        worst_seed = worst = 0.0
        for q0 in range(1, 2 ** 31, 2 ** 29):
            a, b = hs.rcp_selftest(q0, min(2 ** 29, 2 ** 31 - q0 + 1))
            worst_seed, worst = max(worst_seed, a), max(worst, b)
        print("reciprocal seed error 2^%.2f, refined 2^%.2f" % (np.log2(worst_seed), np.log2(worst)))
        assert worst_seed < 2.0 ** -19.9
        assert worst < 2.0 ** -39.8


def test_histogram_with_n_and_ragged_reads(hb, oracle):
    reads = random_reads(5000, 60, seed=3, n_frac=0.01, lower_frac=0.3, ragged=200)
    bases, offs = oracle.pack_reads(reads)
    ho, nm = oracle.count_reads(21, 9, 21 ** 4, bases, offs)
    with hb.HistoSketch(21, 9, 4) as hs:
        hs.add_reads(bases, offs)
        np.testing.assert_array_equal(hs.histogram(), ho.astype(np.uint32))
        assert hs.stats()["n_minimizers"] == nm


def test_packed_transport_gives_the_same_spectrum(hb, oracle):
    """2 bits per base + code-4 positions (HULK_B200_F_PACK_INPUT, push_reads_packed) against the oracle on the ASCII
    reads: N, IUPAC, lower case, U, raw bytes 0..3 and every other byte value, ragged and fixed-length batches,
    a caller-packed batch entered at a read that does not start on a byte of the packed stream."""
    k, w = 21, 9
    D = k ** 4
    reads = random_reads(6000, 60, seed=5, n_frac=0.01, lower_frac=0.3, ragged=200)
    rng = np.random.default_rng(6)
    weird = np.frombuffer(b"ACGTacgtUuNRYKM\x00\x01\x02\x03\xff", np.uint8)
    reads += [weird[rng.integers(0, weird.size, 151)].tobytes() for _ in range(300)]
    reads += [bytes(range(256)), b"N" * 90, b"\x00\x01\x02\x03" * 40, b"ACGU" * 40]
    bases, offs = oracle.pack_reads(reads)
    ho, nm = oracle.count_reads(k, w, D, bases, offs)
    ho = ho.astype(np.uint32)
    # (a) the library packs a pushed ASCII batch itself
    with hb.HistoSketch(k, w, 4, pack_input=True) as hs:
        hs.add_reads(bases, offs)
        np.testing.assert_array_equal(hs.histogram(), ho)
        st = hs.stats()
        assert st["n_minimizers"] == nm
        assert st["h2d_bytes"] < 0.45 * bases.size + 8 * offs.size          # a quarter of the bases (+ exceptions, offsets)
    # (b) the caller packs; the batch is pushed in two calls, the second entered mid-byte
    packed, exc, n_exc = hb.pack_bases(bases, 3)
    assert n_exc == exc.size > 0
    with hb.HistoSketch(k, w, 4) as hs:
        hs.add_reads_packed(packed, exc, offs, len(reads))
        np.testing.assert_array_equal(hs.histogram(), ho)
    cut = next(i for i in range(2000, len(reads)) if offs[i] % 4 == 1)
    with hb.HistoSketch(k, w, 4) as hs:
        hs.add_reads_packed(packed, exc, offs[:cut + 1], cut)
        o2 = offs[cut:].copy()
        p2, e2, _ = hb.pack_bases(bases[int(o2[0]):], 2)
        hs.add_reads_packed(p2, e2, o2, len(reads) - cut)
        np.testing.assert_array_equal(hs.histogram(), ho)
    # (c) fixed-length reads, the other scan kernels (k = 11: first-generation scan; w = 5: staged tiles)
    n, L = 20000, 150
    fixed = hb.synthetic_reads(n, L, seed=8)
    fixed[::53, 70] = ord("N")
    fixed[::97, 3] = ord("c")
    offs_f = np.arange(n + 1, dtype=np.uint64) * np.uint64(L)
    for kk, ww in [(21, 9), (11, 9), (15, 5)]:
        hf, nmf = oracle.count_reads(kk, ww, kk ** 4, fixed.reshape(-1), offs_f)
        with hb.HistoSketch(kk, ww, 4, pack_input=True) as hs:
            hs.add_reads_fixed(fixed.reshape(-1), n, L)
            np.testing.assert_array_equal(hs.histogram(), hf.astype(np.uint32))
            assert hs.stats()["n_minimizers"] == nmf
        pf, ef, _ = hb.pack_bases(fixed.reshape(-1))
        with hb.HistoSketch(kk, ww, 4) as hs:
            hs.add_reads_packed(pf, ef, None, n, L)
            np.testing.assert_array_equal(hs.histogram(), hf.astype(np.uint32))
    # (d) a batch of mostly foreign bytes is not worth packing: it travels as ASCII, same result
    junk = [bytes(rng.integers(0, 256, 120, dtype=np.uint8)) for _ in range(2000)]
    jb, jo = oracle.pack_reads(junk)
    hj, _ = oracle.count_reads(k, w, D, jb, jo)
    with hb.HistoSketch(k, w, 4, pack_input=True) as hs:
        hs.add_reads(jb, jo)
        np.testing.assert_array_equal(hs.histogram(), hj.astype(np.uint32))
        assert hs.stats()["h2d_bytes"] >= jb.size


def test_feeder_thread_delivers_the_same_batches(hb, oracle):
    """HULK_B200_F_ASYNC_INPUT + HULK_B200_F_PACK_INPUT: a second host thread packs and copies each pushed batch while the
    caller already enqueues its kernels (they wait on the device for the batch's sequence number).  Many batches in a
    row -- more than the staging ring holds -- with flushes in between, ragged and fixed-length, a batch of mostly
    foreign bytes (the feeder ships it as ASCII) and tiny batches (handled by the calling thread) mixed in."""
    k, w, s = 21, 9, 16
    D = k ** 4
    r, c, b = _tables(s, D, 77)
    rng = np.random.default_rng(12)
    batches = []
    for i in range(14):
        if i % 5 == 3:
            reads = [bytes(rng.integers(0, 256, 140, dtype=np.uint8)) for _ in range(1500)]      # junk
        elif i % 5 == 4:
            reads = random_reads(20, 60, seed=100 + i)                                             # tiny: < 4096 bases
        else:
            reads = random_reads(4000, 100, seed=100 + i, n_frac=0.005, lower_frac=0.1, ragged=80)
        batches.append(oracle.pack_reads(reads))
    fixed = hb.synthetic_reads(30000, 150, seed=3)
    fixed[::41, 9] = ord("N")
    ref = oracle.HistoSketch(k, s, D, 1.0, r, c, b)
    hist = np.zeros(D)
    n_min = 0
    with hb.HistoSketch(k, w, s, 1.0, tables=(r, c, b), async_input=True, pack_input=True) as hs:
        keep = []
        for i, (bases, offs) in enumerate(batches):
            hs.add_reads(bases, offs)
            keep.append((bases, offs))
            h, nm = oracle.count_reads(k, w, D, bases, offs)
            hist += h
            n_min += nm
            if i % 4 == 3:
                hs.flush()
                ref.flush(hist)
                hist = np.zeros(D)
        hs.add_reads_fixed(fixed.reshape(-1), 30000, 150)
        h, nm = oracle.count_reads(k, w, D, fixed.reshape(-1), np.arange(30001, dtype=np.uint64) * np.uint64(150))
        hist += h
        n_min += nm
        np.testing.assert_array_equal(hs.histogram(), hist.astype(np.uint32))
        hs.flush()
        ref.flush(hist)
        mins, weights = hs.finish()
        st = hs.stats()
    mins_ref, weights_ref = ref.get()
    np.testing.assert_array_equal(mins, mins_ref)
    np.testing.assert_allclose(weights, weights_ref, rtol=W_RTOL, atol=0)
    assert st["n_minimizers"] == n_min
    assert st["n_packed_batches"] >= 9


def test_device_resident_input_equals_host_input(hb):
    import torch
    n, L = 30000, 150
    reads = hb.synthetic_reads(n, L, seed=4)
    with hb.HistoSketch(21, 9, 4) as a, hb.HistoSketch(21, 9, 4) as b:
        a.add_reads_fixed(reads.reshape(-1), n, L)
        t = torch.from_numpy(reads.reshape(-1).copy()).cuda()
        torch.cuda.synchronize()
        b.add_reads_device(t.data_ptr(), None, n, L)
        np.testing.assert_array_equal(a.histogram(), b.histogram())
        offs = torch.arange(0, (n + 1) * L, L, dtype=torch.int64).cuda()
        torch.cuda.synchronize()
        b.add_reads_device(t.data_ptr(), offs.data_ptr(), n, 0)
        np.testing.assert_array_equal(2 * a.histogram(), b.histogram())


def test_large_batches_are_split_into_bounded_launches(hb, monkeypatch):
    # a push of any size becomes k1 launches of bounded size (here forced down to 1000 reads): same spectrum,
    # same per-read error reporting, for fixed-length, offset-addressed and device-resident input
    import torch
    n, L = 12345, 150
    reads = hb.synthetic_reads(n, L, seed=9)
    ragged = random_reads(3000, 100, seed=10, n_frac=0.01, ragged=80)
    with hb.HistoSketch(21, 9, 4) as ref:
        ref.add_reads_fixed(reads.reshape(-1), n, L)
        ref.add_seqs(ragged)
        want, want_n = ref.histogram(), ref.stats()["n_minimizers"]
    monkeypatch.setenv("HULK_B200_MAX_LAUNCH_READS", "1000")
    with hb.HistoSketch(21, 9, 4) as a:
        a.add_reads_fixed(reads.reshape(-1), n, L)
        a.add_seqs(ragged)
        np.testing.assert_array_equal(a.histogram(), want)
        assert a.stats()["n_minimizers"] == want_n
        sets, counts = a.minimizers(ragged[:2500])
        assert int(counts.sum()) > 0 and len(sets) == 2500
    with hb.HistoSketch(21, 9, 4) as b:
        t = torch.from_numpy(reads.reshape(-1).copy()).cuda()
        bases, offs = hb.pack_reads(ragged)
        tb, to = torch.from_numpy(bases).cuda(), torch.from_numpy(offs.astype(np.int64)).cuda()
        torch.cuda.synchronize()
        b.add_reads_device(t.data_ptr(), None, n, L)
        b.add_reads_device(tb.data_ptr(), to.data_ptr(), len(ragged), 0)
        np.testing.assert_array_equal(b.histogram(), want)
    with hb.HistoSketch(21, 9, 4) as c:
        bad = ragged[:1500] + [b"ACGT"] + ragged[1500:]
        with pytest.raises(hb.HulkError) as e:
            c.add_seqs(bad)
            c.sync()
        assert e.value.code == -4 and "read 1500" in str(e.value)


# ---- stage 3: count-min + CWS ---------------------------------------------------------------------
def _run_both(hb, oracle, k, w, s, decay, reads_batches, tables, check_estimates=True, rtol_f=0.0, parallel=False):
    D = hb.spectrum_size(k)
    r, c, b = tables
    ref = oracle.HistoSketch(k, s, D, decay, r, c, b)
    with hb.HistoSketch(k, w, s, decay, tables=tables) as hs:
        for reads in reads_batches:
            bases, offs = oracle.pack_reads(reads)
            hist, _ = oracle.count_reads(k, w, D, bases, offs)
            f_ref = ref.flush(hist, parallel=parallel)      # parallel: same per-slot order, slots over host threads
            hs.add_reads(bases, offs)
            hs.flush()
            if check_estimates:
                f = hs.estimates()
                assert (np.isnan(f) == np.isnan(f_ref)).all()
                m = ~np.isnan(f)
                if rtol_f == 0.0:
                    np.testing.assert_array_equal(f[m], f_ref[m])
                else:
                    np.testing.assert_allclose(f[m], f_ref[m], rtol=rtol_f, atol=0)
            assert (hs.histogram() == 0).all()                       # wiped (kmerspectrum.go:58-64)
        mins, weights = hs.finish()
        q = hs.cms()
        st = hs.stats()
    mo, wo = ref.get()
    np.testing.assert_array_equal(mins, mo)
    np.testing.assert_allclose(weights, wo, rtol=W_RTOL, atol=0)
    if rtol_f == 0.0:
        np.testing.assert_array_equal(q, ref.cms())
    else:
        np.testing.assert_allclose(q, ref.cms(), rtol=rtol_f, atol=1e-300)
    return st


@pytest.mark.parametrize("k,s", [(11, 64), (21, 50), (7, 33)])
def test_sketch_no_decay_exact(hb, oracle, k, s):
    D = hb.spectrum_size(k)
    tables = _tables(s, D, 21)
    batches = [random_reads(3000, 100, seed=100 + i, n_frac=0.002) for i in range(3)]
    st = _run_both(hb, oracle, k, 9, s, 1.0, batches, tables)
    assert st["n_flushes"] == 3


@pytest.mark.parametrize("decay", [0.02, 0.5, 0.999, 0.0])
def test_sketch_with_concept_drift(hb, oracle, decay):
    k, s = 9, 48
    D = hb.spectrum_size(k)
    tables = _tables(s, D, 22)
    batches = [random_reads(1500, 100, seed=200 + i) for i in range(4)]
    _run_both(hb, oracle, k, 9, s, decay, batches, tables, rtol_f=(0.0 if decay == 0.0 else 1e-9))


@pytest.mark.parametrize("decay", [1.0, 0.3])
def test_pipelined_and_serial_runs_are_identical(hb, decay):
    # many short intervals back to back keep several intervals in flight (counting i+1, i+2 while i is flushed,
    # count-min of i+1 under the CWS sweep of i); switching the overlap off must not change a bit
    k, s, n_int, per = 11, 40, 14, 6000
    D = hb.spectrum_size(k)
    tables = _tables(s, D, 31)
    reads = hb.synthetic_reads(n_int * per, 150, seed=12)
    out = []
    for overlap in (True, False, True):
        with hb.HistoSketch(k, 9, s, decay, tables=tables, async_input=True) as hs:
            hs.set_overlap(overlap)
            for i in range(n_int):
                hs.add_reads_fixed(reads[i * per:(i + 1) * per].reshape(-1), per, 150)
                hs.flush()
            mins, weights = hs.finish()
            out.append((mins, weights, hs.cms(), hs.estimates(), hs.stats()))
    for mins, weights, q, f, st in out[1:]:
        np.testing.assert_array_equal(mins, out[0][0])
        np.testing.assert_array_equal(weights, out[0][1])
        np.testing.assert_array_equal(q, out[0][2])
        np.testing.assert_array_equal(f, out[0][3])
        assert st["n_flushes"] == n_int and st["n_minimizers"] == out[0][4]["n_minimizers"]


def test_c3_shape_k31_drift(hb, oracle):
    # BASELINE config C3 at reduced size: k=31 (D = 923 521, integer-compare scan path), concept drift on
    k, s, decay = 31, 12, 0.02
    D = hb.spectrum_size(k)
    tables = _tables(s, D, 23)
    batches = [[bytes(r) for r in hb.synthetic_reads(15000, 150, seed=31, first_read=15000 * i)] for i in range(3)]
    st = _run_both(hb, oracle, k, 9, s, decay, batches, tables, rtol_f=1e-9)
    assert st["n_flushes"] == 3


def _dense_batches(hb, n_flushes, per, seed):
    return [[bytes(r) for r in hb.synthetic_reads(per, 150, seed=seed, first_read=per * i)] for i in range(n_flushes)]


@pytest.mark.parametrize("fp32", [False, True])
def test_c2_shape_dense_spectrum_three_flushes(hb, oracle, monkeypatch, fp32):
    """BASELINE config C2 at its own shape: k=21, w=9, s=512, three intervals of 100 000 synthetic 150 bp reads each.  Every
    bin of the 194 481-bin spectrum is non-zero in every flush (the steady state the bench times: all screen chunks
    populated, W converged after the first flush), against the oracle's AddElement loop, with either screen."""
    monkeypatch.setenv("HULK_B200_K3_FP32", "1" if fp32 else "0")
    k, s = 21, 512
    D = hb.spectrum_size(k)
    tables = _tables(s, D, 41)
    batches = _dense_batches(hb, 3, 100_000, seed=1)
    st = _run_both(hb, oracle, k, 9, s, 1.0, batches, tables, parallel=True)
    assert st["n_flushes"] == 3 and st["n_adds"] == 3 * D            # dense: every bin in every flush


def test_c3_shape_dense_spectrum_drift(hb, oracle):
    """BASELINE config C3's regime: k=31 (D = 923 521), concept drift 0.02, a dense spectrum (350 000 reads per flush put
    ~9 minimizers into the average bin) over two flushes.  256 of C3's 1024 slots: the float64 tables of all 1024
    would be 22.7 GB of host memory for the oracle; the slots are independent (histosketch.go:135-153)."""
    k, s, decay = 31, 256, 0.02
    D = hb.spectrum_size(k)
    tables = _tables(s, D, 43)
    batches = _dense_batches(hb, 2, 350_000, seed=3)
    st = _run_both(hb, oracle, k, 9, s, decay, batches, tables, rtol_f=1e-9, parallel=True)
    assert st["n_flushes"] == 2 and st["n_adds"] > 1.99 * D


def test_c1_golden_json(hb, fixture_reads):
    # BASELINE config C1: hulk sketch -f testing/test-reads-small.fq.gz -k 21 -s 50, Go-compatible tables
    for name, decay, interval in (("c1_k21_s50.json", 1.0, 0), ("c1_k21_s50_x02_i250.json", 0.2, 250)):
        want = open(os.path.join(GOLDEN, name)).read()
        with hb.HistoSketch(21, 9, 50, decay) as hs:
            hs.generate_tables(background=(interval != 0))     # the drawn-while-counting path gives the same tables
            mins, weights, st = hb.sketch_reads(hs, [hb.pack_reads(fixture_reads)], interval=interval)
        doc = hb.sketch_json("testing/test-reads-small.fq.gz,", 21, mins, weights, 194481, decay != 1.0)
        wj, gj = json.loads(want), json.loads(doc)
        ws, gs = wj["signatures"][0]["Sketch"], gj["signatures"][0]["Sketch"]
        assert gs["mins"] == ws["mins"] and gs["md5sum"] == ws["md5sum"]
        np.testing.assert_allclose(gs["weights"], ws["weights"], rtol=W_RTOL)
        assert st["n_minimizers"] == 17040 and st["n_reads"] == 1000
        if doc != want:      # byte-compatibility up to the last float64 ulp of the GPU's exp/log
            assert [ln for ln in doc.split("\n") if "e" not in ln and "." not in ln] == \
                   [ln for ln in want.split("\n") if "e" not in ln and "." not in ln]


def _ulp_diff(a, b):
    ia, ib = a.view(np.int64), b.view(np.int64)
    return np.abs(ia - ib)


@pytest.mark.parametrize("k,s,slots", [(11, 16, None), (11, 16, (5, 9)), (21, 3, None), (15, 40, (0, 7)), (21, 128, (100, 128))])
def test_cws_tables_drawn_on_the_device(hb, k, s, slots):
    """hulk_b200_generate_cws_tables_device against the host generator (hulk_b200_new_cws): the same accept/reject
    SEQUENCE (a single difference would shift every later entry, so agreement of the last row proves it), values equal
    to the last bit or two (CUDA's exp/log against glibc's), b = U r from the uniform at the element's fixed position."""
    D = k ** 4
    sb, se = slots if slots else (0, s)
    r0, c0, b0 = hb.new_cws(s, D, sb, se)
    with hb.HistoSketch(k, 9, s, 1.0, slots=slots) as hs:
        hs.generate_tables_device()
        r1, c1, b1 = hs.tables()
    assert r1.shape == r0.shape
    # draw = 2 exp(ln(u / (1 - u)) / sqrt 3): one ulp of the logarithm is |v| ulps of the draw -- a few ulps at most
    for name, x0, x1 in (("r", r0, r1), ("b", b0, b1)):
        np.testing.assert_allclose(x1, x0, rtol=8e-15, atol=0, err_msg=name)
        assert (_ulp_diff(x0, x1) == 0).mean() > 0.5, name
    # c = ln(draw): a draw next to 1 gives a logarithm next to 0, where one ulp of the draw is many ulps of c
    np.testing.assert_allclose(c1, c0, rtol=1e-14, atol=4e-16)
    assert (c1 == c0).mean() > 0.5


def test_cws_device_draw_lets_the_host_decide_what_is_close(hb, monkeypatch):
    """With the band widened to 2e-3, about one attempt in a thousand goes through the host's re-decision: the tables
    must not change (the host decides as the device would have, except on true near-ties)."""
    k, s = 11, 16
    D = k ** 4
    r0, c0, b0 = hb.new_cws(s, D)
    monkeypatch.setenv("HULK_B200_CWS_TIE_EPS", "2e-3")
    with hb.HistoSketch(k, 9, s, 1.0) as hs:
        hs.generate_tables_device()
        r1, c1, b1 = hs.tables()
    np.testing.assert_allclose(r1, r0, rtol=8e-15, atol=0)
    np.testing.assert_allclose(c1, c0, rtol=1e-14, atol=4e-16)
    np.testing.assert_allclose(b1, b0, rtol=8e-15, atol=0)


def test_sketch_with_device_drawn_tables_matches_the_oracle(hb, oracle):
    """Whole path with the tables drawn on the device: mins equal, weights to 1e-12 (the tables differ from the host
    draw by at most a few ulp)."""
    k, w, s = 11, 9, 24
    D = k ** 4
    r, c, b = hb.new_cws(s, D)
    reads = hb.synthetic_reads(6000, 150, seed=21)
    offs = np.arange(6001, dtype=np.uint64) * np.uint64(150)
    ref = oracle.HistoSketch(k, s, D, 1.0, r, c, b)
    ref.run(w, reads.reshape(-1), offs, interval=2000, parallel=True)
    mins_ref, weights_ref = ref.get()
    with hb.HistoSketch(k, w, s, 1.0) as hs:
        hs.generate_tables_device()
        mins, weights, _ = hb.sketch_reads(hs, [(reads.reshape(-1), offs)], interval=2000)
    np.testing.assert_array_equal(mins, mins_ref)
    np.testing.assert_allclose(weights, weights_ref, rtol=W_RTOL, atol=0)


def test_flush_semantics(hb):
    k, s = 5, 8
    D = k ** 4
    tables = _tables(s, D, 3)
    with hb.HistoSketch(k, 4, s, tables=tables) as hs:
        hs.flush()                                           # empty spectrum: no-op (boss.go:117)
        mins, weights = hs.finish()
        assert (mins == 0).all() and (weights == 1.7976931348623157e308).all()
        assert hs.stats()["n_flushes"] == 0
        hs.add_seqs([b"ACGTACGTAC"])                         # 3 minimizers -> 3/625 bins < 1 %
        hs.flush()
        with pytest.raises(hb.HulkError) as e:
            hs.finish()
        assert e.value.code == -6 and "not used yet" in str(e.value)
    with hb.HistoSketch(k, 4, s) as hs:
        with pytest.raises(hb.HulkError) as e:
            hs.flush()                                       # tables never set
        assert e.value.code == -21


def test_context_reuse_snapshot_and_async_input(hb, oracle):
    # one context, several samples: reset() returns it to the freshly-created state; snapshot_async works into
    # pageable memory (copy engine) and pinned memory (kernel write); ASYNC_INPUT + sync_inputs
    import ctypes as C
    k, s = 9, 24
    D = k ** 4
    tables = _tables(s, D, 41)
    samples = [random_reads(2500, 90, seed=300 + i, ragged=30) for i in range(3)]
    want = []
    for reads in samples:
        ref = oracle.HistoSketch(k, s, D, 1.0, *tables)
        ref.run(9, *oracle.pack_reads(reads), interval=1000)
        want.append(ref.get())
    L = hb.load()
    pin = C.c_void_p()
    assert L.hulk_b200_alloc_pinned(C.byref(pin), 16 * s) == 0
    with hb.HistoSketch(k, 9, s, tables=tables, async_input=True) as hs:
        for reads, (mo, wo) in zip(samples, want):
            hs.reset()
            bases, offs = hb.pack_reads(reads)
            for a in range(0, len(reads), 1000):
                hs.add_reads(bases, offs[a:a + 1001])
                hs.flush()
            assert L.hulk_b200_sync_inputs(hs._ctx) == 0
            page_m, page_w = np.zeros(s, np.uint64), np.zeros(s)
            assert L.hulk_b200_snapshot_async(hs._ctx, page_m.ctypes.data, page_w.ctypes.data) == 0
            assert L.hulk_b200_snapshot_async(hs._ctx, pin.value, pin.value + 8 * s) == 0
            hs.sync()
            pin_m = np.ctypeslib.as_array(C.cast(pin, C.POINTER(C.c_uint64)), shape=(s,)).copy()
            pin_w = np.ctypeslib.as_array(C.cast(C.c_void_p(pin.value + 8 * s), C.POINTER(C.c_double)), shape=(s,)).copy()
            mins, weights = hs.finish()
            for m, w in ((mins, weights), (page_m, page_w), (pin_m, pin_w)):
                np.testing.assert_array_equal(m, mo)
                np.testing.assert_allclose(w, wo, rtol=W_RTOL, atol=0)
            st = hs.stats()
            assert st["n_reads"] == len(reads) and st["n_flushes"] == 3
    L.hulk_b200_free_pinned(pin)


def test_intervals_match_oracle_run(hb, oracle):
    k, w, s = 11, 9, 40
    D = hb.spectrum_size(k)
    tables = _tables(s, D, 9)
    reads = random_reads(2500, 120, seed=12)
    bases, offs = oracle.pack_reads(reads)
    for interval in (0, 1000, 500, 2500):
        ref = oracle.HistoSketch(k, s, D, 1.0, *tables)
        nm, nf = ref.run(w, bases, offs, interval=interval)
        with hb.HistoSketch(k, w, s, tables=tables) as hs:
            mins, weights, st = hb.sketch_reads(hs, [(bases, offs)], interval=interval)
        np.testing.assert_array_equal(mins, ref.get()[0])
        np.testing.assert_allclose(weights, ref.get()[1], rtol=W_RTOL)
        assert st["n_minimizers"] == nm


def test_slot_sharding_is_identical_to_single_context(hb):
    # multi-GPU layout: every rank counts a read shard, histograms are summed, each rank sweeps its slots
    k, w, s = 11, 9, 64
    D = hb.spectrum_size(k)
    tables = _tables(s, D, 31)
    reads = [random_reads(2000, 150, seed=40 + i) for i in range(2)]
    with hb.HistoSketch(k, w, s, tables=tables) as whole:
        for rd in reads:
            whole.add_seqs(rd)
        whole.flush()
        mins, weights = whole.finish()
    parts = []
    shards = [hb.HistoSketch(k, w, s, slots=(g * 16, (g + 1) * 16), tables=tuple(t[g * 16:(g + 1) * 16] for t in tables))
              for g in range(4)]
    try:
        shards[0].add_seqs(reads[0])
        shards[1].add_seqs(reads[1])
        h0, h1 = shards[0].histogram(), shards[1].histogram()
        # what the NCCL all-reduce leaves in every rank's histogram: the sum of all partial spectra
        shards[0].merge_histogram(h1)
        shards[1].merge_histogram(h0)
        shards[2].merge_histogram(h0 + h1)
        shards[3].merge_histogram(h0 + h1)
        for sh in shards:
            sh.flush()
            parts.append(sh.finish())
    finally:
        for sh in shards:
            sh.close()
    np.testing.assert_array_equal(np.concatenate([p[0] for p in parts]), mins)
    np.testing.assert_array_equal(np.concatenate([p[1] for p in parts]), weights)


def test_folded_table_is_the_rounding_of_the_float64_coefficient(hb, monkeypatch):
    # the screen's coefficient table K = c * exp(b - r): bfloat16 (default) or fp32, NaN padding up to 512 bins
    k, s = 7, 5
    D = k ** 4
    r, c, b = _tables(s, D, 2)
    want = c * np.exp(b - r)
    with hb.HistoSketch(k, 5, s, tables=(r, c, b)) as hs:
        K = hs.folded_table()
    assert K.shape[1] % 512 == 0 and np.isnan(K[:, D:]).all()
    np.testing.assert_allclose(K[:, :D], want, rtol=2.0 ** -8)                   # bf16: 8 significant bits, round to nearest
    assert ((np.ascontiguousarray(K[:, :D]).view(np.uint32) & 0xffff) == 0).all()
    monkeypatch.setenv("HULK_B200_K3_FP32", "1")
    with hb.HistoSketch(k, 5, s, tables=(r, c, b)) as hs:
        K = hs.folded_table()
    np.testing.assert_allclose(K[:, :D], want.astype(np.float32), rtol=2e-7)


@pytest.mark.parametrize("fp32", ["0", "1"])
def test_both_screens_give_the_same_sketch(hb, oracle, monkeypatch, fp32):
    # the bf16 and the fp32 screen only decide which chunks are re-evaluated in float64: same mins, same weights
    monkeypatch.setenv("HULK_B200_K3_FP32", fp32)
    k, s = 11, 96
    D = hb.spectrum_size(k)
    tables = _tables(s, D, 51)
    batches = [random_reads(4000, 120, seed=500 + i) for i in range(6)]
    for decay in (1.0, 0.4):
        _run_both(hb, oracle, k, 9, s, decay, batches, tables, rtol_f=(0.0 if decay == 1.0 else 1e-9))


# ---- size-independent properties at a larger size -------------------------------------------------
def test_properties_one_million_reads(hb):
    n, L, k, s = 1_000_000, 150, 21, 64
    D = hb.spectrum_size(k)
    tables = _tables(s, D, 55)
    reads = hb.synthetic_reads(n, L, seed=1).reshape(-1)
    with hb.HistoSketch(k, 9, s, tables=tables) as a:
        a.add_reads_fixed(reads, n, L)
        h = a.histogram()
        st = a.stats()
        assert int(h.astype(np.int64).sum()) == st["n_minimizers"]          # every set member lands in one bin
        assert 26.0 < st["n_minimizers"] / n < 28.0
        a.flush()
        mins_a, w_a = a.finish()
    # linearity: two half-batches into one context == one batch; order of batches is irrelevant
    with hb.HistoSketch(k, 9, s, tables=tables) as b:
        half = (n // 2) * L
        b.add_reads_fixed(reads[half:], n - n // 2, L)
        b.add_reads_fixed(reads[:half], n // 2, L)
        np.testing.assert_array_equal(b.histogram(), h)
        b.flush()
        mins_b, w_b = b.finish()
        # idempotence: flushing an empty spectrum changes nothing
        b.flush()
        mins_c, w_c = b.finish()
    np.testing.assert_array_equal(mins_a, mins_b)
    np.testing.assert_array_equal(w_a, w_b)
    np.testing.assert_array_equal(mins_b, mins_c)
    np.testing.assert_array_equal(w_b, w_c)
    assert (w_a < 0).all() and (mins_a < D).all()
