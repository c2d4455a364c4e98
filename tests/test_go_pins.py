"""Pins produced by the reference itself (go/golden/main.go run inside will-rowe/hulk v1.0.0) against the CPU oracle.

tests/golden/go_pins.json is NOT in the repository yet: the image this was built in has no Go toolchain, so the
program has never been run ("parity unpinned").  Whoever has Go follows the recipe in
go/golden/main.go's header, drops the file in, and these tests turn every restated function of the oracle --
minimizer sets, jump-hash binning, count-min estimates, leesper/go_rng's Gamma and uniform streams, the CWS update --
from "restated and cross-checked" into "equal to the reference's own output".  Bit-exact throughout (float64 values
travel as their bit patterns), except the sketch weights, which pass through exp/log of the platform's libm (1e-12).
"""
import gzip
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

PINS = os.path.join(GOLDEN, "go_pins.json")
pytestmark = pytest.mark.skipif(not os.path.exists(PINS),
                                reason="tests/golden/go_pins.json absent: run go/golden/main.go in a Go toolchain "
                                       "(see its header) to pin the oracle against the reference")


@pytest.fixture(scope="module")
def pins():
    with open(PINS) as fh:
        return json.load(fh)


def _f64(bits):
    return np.array(bits, dtype=np.uint64).view(np.float64)


def test_go_rng_streams(pins, oracle):
    """leesper/go_rng Gamma(2,1) and Float64Range(0,1), seed 1: the streams newCWS consumes (histosketch.go:103-116)."""
    g = oracle.GoRand(1)
    got = np.array([g.gamma(2.0, 1.0) for _ in range(len(pins["rng_gamma_2_1_bits"]))])
    np.testing.assert_array_equal(got.view(np.uint64), np.array(pins["rng_gamma_2_1_bits"], dtype=np.uint64))
    u = oracle.GoRand(1)
    got = np.array([u.float64() for _ in range(len(pins["rng_uniform_0_1_bits"]))])
    np.testing.assert_array_equal(got.view(np.uint64), np.array(pins["rng_uniform_0_1_bits"], dtype=np.uint64))
    # and the tables built from them
    n = len(pins["rng_gamma_2_1_bits"]) // 2
    r, c, b = oracle.new_cws(1, n)
    gam, uni = _f64(pins["rng_gamma_2_1_bits"]), _f64(pins["rng_uniform_0_1_bits"])
    np.testing.assert_array_equal(r[0], gam[0:2 * n:2])
    np.testing.assert_allclose(c[0], np.log(gam[1:2 * n:2]), rtol=1e-15)
    np.testing.assert_array_equal(b[0], uni[:n] * r[0])


def test_minimizer_sets_and_spectrum(pins, oracle, fixture_reads):
    k, w, D = pins["k"], pins["w"], pins["num_bins"]
    assert len(pins["minimizers"]) == len(fixture_reads)
    for seq, want in zip(fixture_reads, pins["minimizers"]):
        np.testing.assert_array_equal(np.sort(oracle.minimizers(k, w, seq)), np.array(want, dtype=np.uint64))
    bases, offs = oracle.pack_reads(fixture_reads)
    hist, _ = oracle.count_reads(k, w, D, bases, offs)
    assert int((hist != 0).sum()) == pins["used_bins"]
    assert hashlib.md5(hist.astype("<u4").tobytes()).hexdigest() == pins["spectrum_md5_u32le"]
    nz = np.nonzero(hist)[0]
    np.testing.assert_array_equal(nz, np.array(pins["spectrum_bins"]))
    np.testing.assert_array_equal(hist[nz], np.array(pins["spectrum_freq"]))


def test_countmin_estimates(pins, oracle):
    D = pins["num_bins"]
    for pin in pins["countmin"]:
        z = np.zeros((1, D))
        hs = oracle.HistoSketch(pins["k"], 1, D, pin["decay"], z + 1.0, z, z)
        got = np.array([hs.add_element(b, v) for b, v in zip(pin["bins"], pin["values"])])
        np.testing.assert_array_equal(got.view(np.uint64), np.array(pin["estimates_bits"], dtype=np.uint64))


def test_histosketch_of_the_fixture(pins, oracle):
    D = pins["num_bins"]
    hist = np.zeros(D)
    hist[np.array(pins["spectrum_bins"])] = np.array(pins["spectrum_freq"])
    for pin in pins["histosketch"]:
        r, c, b = oracle.new_cws(pin["s"], D)
        hs = oracle.HistoSketch(pin["k"], pin["s"], D, pin["decay"], r, c, b)
        hs.flush(hist.copy())
        mins, weights = hs.get()
        np.testing.assert_array_equal(mins, np.array(pin["mins"], dtype=np.uint64))
        np.testing.assert_allclose(weights, _f64(pin["weights_bits"]), rtol=1e-12, atol=0)


def test_minhash_sketches_of_the_fixture(pins, oracle, fixture_reads):
    """KMVsketch / KHFsketch fed with every minimizer of the fixture (src/minhash/kmv.go:40-71, khf.go:35-45): the feed
    hulk_b200_minhash_enable switches on, against the restatement in oracle/pyref.py."""
    if "minhash_kmv" not in pins:
        pytest.skip("go_pins.json was produced by an older go/golden/main.go")
    from oracle import pyref as P
    s = len(pins["minhash_khf"])
    kmv, khf = P.KMVsketch(pins["k"], s), P.KHFsketch(pins["k"], s)
    for seq in fixture_reads:
        for m in oracle.minimizers(pins["k"], pins["w"], seq):
            kmv.add_hash(int(m))
            khf.add_hash(int(m))
    assert kmv.get_sketch() == pins["minhash_kmv"]
    assert khf.get_sketch() == pins["minhash_khf"]
