"""CPU-only checks of the product's host side: the C ABI library loads and exports every symbol of
include/hulk_b200.h, fails loudly without a GPU, and its host-only pieces (CWS table generator,
md5, JSON writer, the device math header compiled for the host) agree with the oracle."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, has_gpu
from oracle import pyref as P

import hulk_b200
from hulk_b200 import _native as N


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "hulk_b200.h")).read()
    declared = sorted(set(re.findall(r"HULK_B200_API[^;(]*?(hulk_b200_\w+)\s*\(", hdr)))
    assert declared == sorted(N.EXPORTS)
    L = hulk_b200.load()
    for name in declared:
        assert hasattr(L, name), name
    assert L.hulk_b200_version() == b"1.0.0"
    assert L.hulk_b200_strerror(-4) == b"sequence length must be >= w + k - 1"
    assert L.hulk_b200_strerror(-6) == b"not used yet"


def test_product_does_not_touch_the_oracle():
    # the product tree must never import, link or execute anything under oracle/
    for dirpath, _, files in os.walk(os.path.join(ROOT, "hulk_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".inc")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in txt.lower(), os.path.join(dirpath, f)


@pytest.mark.skipif(has_gpu(), reason="only meaningful on a box without a GPU")
def test_create_fails_loudly_without_gpu():
    with pytest.raises(hulk_b200.HulkError) as e:
        hulk_b200.HistoSketch(21, 9, 50)
    assert e.value.code == N.ECUDA and "no CPU fallback" in str(e.value)


def test_parameter_validation_order():
    # argument checks happen before any CUDA call, so they are testable here
    L = hulk_b200.load()
    for kw, code in [(dict(k=32), N.EK), (dict(w=257), N.EW), (dict(decay_ratio=1.5), N.EDECAY),
                     (dict(decay_ratio=-0.1), N.EDECAY), (dict(k=1), N.EBINS), (dict(num_bins=-5), N.ENEGBINS)]:
        p = N.Params()
        p.k, p.w, p.sketch_size, p.decay_ratio = 21, 9, 50, 1.0
        for a, v in kw.items():
            setattr(p, a, v)
        ctx = C.c_void_p()
        assert L.hulk_b200_create(C.byref(p), C.byref(ctx)) == code, kw


def test_new_cws_equals_oracle(oracle):
    s, D = 5, 14641
    r, c, b = hulk_b200.new_cws(s, D)
    ro, co, bo = oracle.new_cws(s, D)
    np.testing.assert_array_equal(r, ro)
    np.testing.assert_array_equal(c, co)
    np.testing.assert_array_equal(b, bo)
    r2, c2, b2 = hulk_b200.new_cws(s, D, 2, 4)
    np.testing.assert_array_equal(r2, ro[2:4])
    np.testing.assert_array_equal(b2, bo[2:4])


@pytest.mark.parametrize("s,D,sb,se,T,chunk", [(3, 1000, 0, 3, 4, 64), (5, 777, 2, 5, 3, 100), (2, 50000, 0, 2, 7, 4096),
                                               (4, 14641, 1, 3, 8, 1 << 14), (6, 14641, 0, 6, 5, 1000)])
def test_parallel_cws_draw_is_bit_identical(oracle, s, D, sb, se, T, chunk):
    # jump-ahead of Go's ALFG + sequential resolution of the rejection sampler's phase: same tables as the
    # one-generator loop of histosketch.go:95-126, whatever the thread count and chunk length
    L = hulk_b200.load()
    rows = se - sb
    r, c, b = (np.full((rows, D), np.nan) for _ in range(3))
    rc = L.hulk_b200_new_cws_parallel(s, D, sb, se, r.ctypes.data, c.ctypes.data, b.ctypes.data, T, chunk)
    assert rc == 0
    ro, co, bo = oracle.new_cws(s, D)
    np.testing.assert_array_equal(r, ro[sb:se])
    np.testing.assert_array_equal(c, co[sb:se])
    np.testing.assert_array_equal(b, bo[sb:se])


def test_md5_and_json_equal_oracle_side():
    rng = np.random.default_rng(0)
    mins = rng.integers(0, 923521, 64).astype(np.uint64)
    weights = np.concatenate([-rng.random(40) * 10.0 ** rng.integers(-12, 3, 40), rng.random(20) * 1e-8,
                              [1.7976931348623157e308, 0.0, -0.0, 123456789012345678901234.0]])
    assert hulk_b200.md5_mins(mins) == P.md5_of_mins(mins)
    for fn, banner in (("a.fq,b.fq.gz,", "blank"), ("STDIN", "label <&> \"q\" \\ \t  é")):
        mine = hulk_b200.sketch_json(fn, 21, mins, weights, 194481, True, banner)
        assert mine == P.sketch_json(fn, 21, mins, weights, 194481, True, banner)
    import json
    doc = json.loads(hulk_b200.sketch_json("x,", 31, mins, weights, 923521, False))
    assert list(doc) == ["class", "filename", "hash_function", "license", "signatures", "version", "banner_label"]
    sk = doc["signatures"][0]["Sketch"]
    assert list(sk) == ["ksize", "md5sum", "mins", "weights", "num", "num_histogram_bins", "concept_drift"]
    assert sk["weights"] == [float(x) for x in weights] and sk["mins"] == mins.tolist()


def test_json_with_minhash_signatures(tmp_path):
    """`hulk sketch --kmv / --khf`: extra signatures in the reference's order, fields ksize, md5sum, mins, num
    (src/pipeline/sketch.go:227-234,289-294; src/minhash/kmv.go:12-21, khf.go:11-17); an empty sketch is refused
    with sketchio.Add's error (src/sketchio/sketchio.go:59-61)."""
    import json
    rng = np.random.default_rng(3)
    mins = rng.integers(0, 194481, 16).astype(np.uint64)
    weights = -rng.random(16)
    kmv = np.sort(rng.integers(0, 2 ** 63, 16).astype(np.uint64))
    khf = hulk_b200.sketch.khf_unfed(16)
    assert khf.tolist() == [2 ** 64 - 1] * 16
    for kw in ({"khf": khf}, {"kmv": kmv}, {"kmv": kmv, "khf": khf}):
        mine = hulk_b200.sketch_json("a.fq,", 21, mins, weights, 194481, False, "blank", **kw)
        assert mine == P.sketch_json("a.fq,", 21, mins, weights, 194481, False, "blank", **kw)
        doc = json.loads(mine)
        assert [g["Algorithm"] for g in doc["signatures"]] == ["histosketch"] + list(kw)
        for g in doc["signatures"][1:]:
            assert list(g["Sketch"]) == ["ksize", "md5sum", "mins", "num"]
            assert g["Sketch"]["md5sum"] == P.md5_of_mins(kw[g["Algorithm"]]) and g["Sketch"]["num"] == 16
    assert (hulk_b200.sketch_json("a.fq,", 21, mins, weights, 194481, False)
            == P.sketch_json("a.fq,", 21, mins, weights, 194481, False))
    for kw in ({"kmv": np.zeros(0, dtype=np.uint64)}, {"khf": []}):
        with pytest.raises(hulk_b200.HulkError) as e:
            hulk_b200.sketch_json("a.fq,", 21, mins, weights, 194481, False, **kw)
        assert e.value.code == N.ENOSKETCH
        with pytest.raises(ValueError):
            P.sketch_json("a.fq,", 21, mins, weights, 194481, False, **kw)
    # the read side finds each signature by algorithm (sketchio.go:197-260); only histosketches carry weights
    path = tmp_path / "multi.json"
    path.write_text(hulk_b200.sketch_json("a.fq,", 21, mins, weights, 194481, False, kmv=kmv, khf=khf))
    m, w, _ = hulk_b200.load_sketch(str(path), 21, "khf")
    assert m.tolist() == khf.tolist() and w.size == 0
    m, w, _ = hulk_b200.load_sketch(str(path), 21, "kmv")
    assert m.tolist() == kmv.tolist() and w.size == 0
    m, w, _ = hulk_b200.load_sketch(str(path), 21)
    assert m.tolist() == mins.tolist() and w.tolist() == weights.tolist()


def test_go_float_formatting_cases():
    cases = {1.7976931348623157e308: "1.7976931348623157e+308", 1e-7: "1e-7", 1.5e-9: "1.5e-9", 1e-6: "0.000001",
             9.999e-7: "9.999e-7", 1e21: "1e+21", 1e20: "100000000000000000000", -3.25: "-3.25", 100.0: "100",
             5e-324: "5e-324", -0.0123: "-0.0123", 2.5e-10: "2.5e-10", 0.1: "0.1"}
    mins = np.zeros(len(cases), dtype=np.uint64)
    doc = hulk_b200.sketch_json("f,", 21, mins, np.array(list(cases)), 16, False)
    got = [ln.strip().rstrip(",") for ln in doc.split('"weights": [')[1].split("]")[0].strip().split("\n")]
    assert got == list(cases.values())
    assert [P.go_float(v) for v in cases] == list(cases.values())


def test_fastq_reader_on_reference_fixture(fixture_reads):
    reads = hulk_b200.read_fastq(os.path.join(GOLDEN, "c1_reads.fq.gz"))
    assert reads == fixture_reads


def test_fastq_reader_quirks(tmp_path):
    # empty lines are nil in the reference and re-fill the same slot (src/pipeline/sketch.go:50,139-159)
    p = tmp_path / "q.fq"
    p.write_bytes(b"@r1\n\nACGT\n+\nIIII\n@r2\r\nGGCC\r\n+\r\nIIII")
    assert hulk_b200.read_fastq(str(p)) == [b"ACGT", b"GGCC"]
    bad = tmp_path / "bad.fq"
    bad.write_bytes(b"r1\nACGT\n+\nIIII\n")
    with pytest.raises(ValueError):
        hulk_b200.read_fastq(str(bad))
    fa = tmp_path / "x.fa"
    fa.write_bytes(b">c1 desc\nACGT\nTTAA\n>c2\nGG\n\n>c3\nAAAA\n")
    assert hulk_b200.read_fasta(str(fa)) == [b"ACGTTTAA", b"GG"]        # stops at the empty line


def test_synthetic_reads_definition():
    a = hulk_b200.synthetic_reads(64, 150, seed=1)
    b = hulk_b200.synthetic_reads(32, 150, seed=1, first_read=32)
    assert a.shape == (64, 150) and (a[32:] == b).all()
    assert set(np.unique(a).tolist()) <= set(b"ACGT")
    # base j of read i from the definition in BASELINE.md section 3
    def splitmix(x):
        x = (x + 0x9E3779B97F4A7C15) & (2 ** 64 - 1)
        z = x
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2 ** 64 - 1)
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2 ** 64 - 1)
        return z ^ (z >> 31)
    for i, j in ((0, 0), (3, 31), (3, 32), (63, 149)):
        word = splitmix(1 ^ (i * 5 + j // 32))
        assert a[i, j] == b"ACGT"[(word >> (2 * (j % 32))) & 3]


HD_TEST_SRC = r"""
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include "hd_math.h"
int main(int argc, char** argv) {
    // stdin: lines "h key mask" | "j key n" | "n byte"; stdout: results
    char op; unsigned long long a, b;
    while (scanf(" %c %llu %llu", &op, &a, &b) == 3) {
        if (op == 'h') printf("%llu\n", (unsigned long long)hulk::hash64(a, b));
        else if (op == 'j') printf("%d\n", hulk::jump_hash(a, (int32_t)b));
        else printf("%u\n", hulk::nt4((uint32_t)a));
    }
    return 0;
}
"""


def test_device_math_header_compiled_for_host_matches_oracle(oracle, tmp_path):
    # hd_math.h is the exact source the kernels use (multiply forms of hash64, arithmetic nt4)
    src = tmp_path / "hd_test.cpp"
    src.write_text(HD_TEST_SRC)
    exe = tmp_path / "hd_test"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O2", "-std=c++17", "-ffp-contract=off", "-I", os.path.join(ROOT, "hulk_b200", "csrc"),
                    "-o", str(exe), str(src)], check=True)
    rng = np.random.default_rng(1)
    lines, want = [], []
    for k in (2, 4, 11, 21, 31):
        mask = (1 << (2 * k)) - 1
        for key in [0, 1, mask] + [int(x) & mask for x in rng.integers(0, 2 ** 63, 300, dtype=np.uint64)]:
            lines.append(f"h {key} {mask}")
            want.append(oracle.hash64(key, mask))
    for n in (1, 2, 10, 2000, 14641, 194481, 923521, 2 ** 31 - 1):
        for key in [0, 1, 2 ** 64 - 1] + [int(x) for x in rng.integers(0, 2 ** 64, 200, dtype=np.uint64)]:
            lines.append(f"j {key} {n}")
            want.append(oracle.jump(key, n))
    for b in range(256):
        lines.append(f"n {b} 0")
        want.append(oracle.nt4(b))
    out = subprocess.run([str(exe)], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout
    got = [int(x) for x in out.split()]
    assert got == want


SCAN_TEST_SRC = r"""
#include <algorithm>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include "k1_scan.h"
#include "k1_scan2.h"
struct VH { uint64_t* b; uint64_t& operator()(int t) const { return b[t]; } };
// second-generation w = 9 scan (k1_scan2.h): the candidate list, converted back to the reference's X
template <int K>
static std::vector<uint64_t> scan_v2(const char* buf, int len, uint32_t cap) {
    std::vector<uint64_t> store(cap + 1);
    hulk::K1List<1> L{store.data(), cap, 0u, 0.0};
    hulk::k1_scan_read_w9_v2<K, 1>(hulk::ByteSrcW{(const uint8_t*)buf, len, 0}, len, L);
    std::vector<uint64_t> out;
    if (L.n > cap) { out.push_back(~0ull); return out; }                 // overflow: the kernel hands the read on
    for (uint32_t e = 0; e < L.n; e++) out.push_back(hulk::K1Repr<K>::to_x(store[e]));
    return out;
}
template <int K>
static std::vector<uint64_t> scan_v2_k(int k, const char* buf, int len, uint32_t cap) {
    if (k == K) return scan_v2<K>(buf, len, cap);
    if constexpr (K < 31) return scan_v2_k<K + 1>(k, buf, len, cap);
    return {};
}
// same role as the kernel's SmemWordSrc: aligned word loads + funnel shift
struct WordSrc {
    const uint32_t* words; uint32_t sh;
    uint32_t get4(int32_t i) const {
        const uint64_t two = ((uint64_t)words[(i >> 2) + 1] << 32) | words[i >> 2];
        return (uint32_t)(two >> sh);
    }
};
int main() {
    // stdin: "s k w seq" -> sorted distinct window minima (byte source and word source must agree)
    //        "J key n"   -> jump_hash_fast + number of ambiguous steps
    char op[4];
    while (scanf(" %3s", op) == 1) {
        if (op[0] == 's') {
            int k, w; static char buf[1 << 20];
            if (scanf("%d %d %1048575s", &k, &w, buf) != 3) return 1;
            const int len = (int)strlen(buf);
            for (int i = 0; i < len; i++) if (buf[i] >= '0' && buf[i] <= '3') buf[i] -= '0';   // raw 0..3 bytes
            std::vector<uint64_t> a, b, vh(w + 1);
            hulk::k1_scan_read<false>(hulk::ByteSrc{(const uint8_t*)buf, len}, len, k, w, VH{vh.data()},
                                      [&](uint64_t m, bool on) { if (on) a.push_back(m); });
            if (hulk::k1_fp_compare_ok(k, w)) {       // the sentinel of the FP-compare variant (host: integer compares)
                hulk::k1_scan_read<true>(hulk::ByteSrc{(const uint8_t*)buf, len}, len, k, w, VH{vh.data()},
                                         [&](uint64_t m, bool on) { if (on) b.push_back(m); });
                if (b != a) { printf("MISMATCH\n"); return 2; }
            }
            for (int mis = 0; mis < 4; mis++) {      // every misalignment of the word source
                std::vector<uint32_t> words((len + mis) / 4 + 4, 0xA5A5A5A5u);
                memcpy((uint8_t*)words.data() + mis, buf, len);
                std::vector<uint64_t> c;
                hulk::k1_scan_read<false>(WordSrc{words.data(), (uint32_t)mis * 8u}, len, k, w, VH{vh.data()},
                                          [&](uint64_t m, bool on) { if (on) c.push_back(m); });
                if (c != a) { printf("MISMATCH\n"); return 2; }
            }
            if (w == 9) {                             // the register-resident w = 9 scan emits the same sequence
                for (int fp = 0; fp < 2; fp++) {
                    if (fp && !hulk::k1_fp_compare_ok(k, w)) continue;
                    std::vector<uint64_t> d;
                    auto em = [&](uint64_t m, bool on) { if (on) d.push_back(m); };
                    if (fp) hulk::k1_scan_read_w9<true>(hulk::ByteSrc8{(const uint8_t*)buf, len, 0}, len, k, em);
                    else hulk::k1_scan_read_w9<false>(hulk::ByteSrc8{(const uint8_t*)buf, len, 0}, len, k, em);
                    if (d != a) { printf("MISMATCH_W9\n"); return 2; }
                }
                if (k >= 8) {                         // k1_scan2.h: same minima with adjacent repeats dropped
                    std::vector<uint64_t> want;
                    for (uint64_t x : a) if (want.empty() || want.back() != x) want.push_back(x);
                    if (scan_v2_k<8>(k, buf, len, 1u << 20) != want) { printf("MISMATCH_V2\n"); return 2; }
                    // a list that is too short reports the overflow and never writes past its capacity
                    if (want.size() > 12 && scan_v2_k<8>(k, buf, len, 12).at(0) != ~0ull) { printf("MISMATCH_V2_OVF\n"); return 2; }
                }
            }
            std::sort(a.begin(), a.end());
            a.erase(std::unique(a.begin(), a.end()), a.end());
            printf("%zu", a.size());
            for (uint64_t x : a) printf(" %llu", (unsigned long long)x);
            printf("\n");
        } else {
            unsigned long long key; long long n;
            if (scanf("%llu %lld", &key, &n) != 2) return 1;
            uint32_t amb = 0, amb_fx = 0;
            const int32_t got = hulk::jump_hash_fast(key, (int32_t)n, &amb);
            // the fixed-point step of the binning kernel (num_buckets <= 2^20) must land in the same bucket
            if (n <= (long long)hulk::JUMP_FX_MAX_BUCKETS && hulk::jump_hash_fx(key, (int32_t)n, &amb_fx) != got) { printf("MISMATCH_FX\n"); return 2; }
            printf("%d %u\n", got, amb + amb_fx);
        }
    }
    return 0;
}
"""


def test_kernel_scan_and_fast_jump_compiled_for_host_match_oracle(oracle, tmp_path):
    # k1_scan.h / hd_math.h are the exact sources the CUDA kernel compiles: the 4-bases-per-word scan,
    # the word-wise encoder with its fallback, and the bracketed fast jump step
    src = tmp_path / "scan_test.cpp"
    src.write_text(SCAN_TEST_SRC)
    exe = tmp_path / "scan_test"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O2", "-std=c++17", "-ffp-contract=off", "-frounding-math", "-I",
                    os.path.join(ROOT, "hulk_b200", "csrc"), "-o", str(exe), str(src)], check=True)
    from conftest import random_reads
    rng = np.random.default_rng(7)
    lines, want = [], []
    cases = [(21, 9, 150), (31, 9, 151), (11, 9, 149), (4, 4, 60), (21, 1, 80), (15, 32, 200), (5, 9, 40),
             (3, 7, 30), (7, 40, 120), (21, 200, 400), (2, 2, 5), (27, 9, 35), (28, 9, 100), (1, 9, 9), (8, 9, 16),
             (9, 9, 17), (21, 9, 29), (21, 9, 33), (17, 9, 150), (19, 9, 101), (23, 9, 150), (25, 9, 97), (29, 9, 150),
             (31, 9, 64), (12, 9, 150), (13, 9, 77), (16, 9, 150), (22, 9, 150), (30, 9, 99), (10, 9, 41)]
    for k, w, L in cases:
        reads = (random_reads(6, L, seed=k * 100 + w) + random_reads(4, L, seed=k + w, n_frac=0.05, lower_frac=0.3)
                 + [b"A" * L, (b"ACGTU" * L)[:L], (b"acgn0123RYKM" * L)[:L], (b"AC" * L)[:L]])
        for rd in reads:
            if len(rd) < k + w - 1:
                continue
            lines.append("s %d %d %s" % (k, w, rd.decode()))
            raw = bytes((c - 48) if 48 <= c <= 51 else c for c in rd)
            m = np.sort(oracle.minimizers(k, w, raw))
            want.append(" ".join([str(len(m))] + [str(int(x)) for x in m]))
    n_jump = 0
    for n in (2, 10, 2000, 14641, 194481, 923521, 2 ** 20, 2 ** 31 - 1):
        for key in [0, 1, 2 ** 64 - 1] + [int(x) for x in rng.integers(0, 2 ** 64, 400, dtype=np.uint64)]:
            lines.append("J %d %d" % (key, n))
            want.append(oracle.jump(key, n))
            n_jump += 1
    out = subprocess.run([str(exe)], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout
    got = out.strip().split("\n")
    assert len(got) == len(want)
    n_amb = 0
    for g, w_ in zip(got, want):
        if isinstance(w_, str):
            assert g == w_
        else:
            b, amb = g.split()
            assert int(b) == w_
            n_amb += int(amb)
    # the exact step is a rare fallback, not the common path (only buckets near 2^31 see it at all)
    assert n_amb < n_jump


LONG_TEST_SRC = r"""
#include <algorithm>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include "k1_scan.h"
#include "k1_long.h"
struct VH { uint64_t* b; uint64_t& operator()(int t) const { return b[t]; } };
int main() {
    // stdin: "k w seg seq" -> the slices of k1_long.h, put end to end, against one pass of k1_scan.h over the sequence;
    // prints the sorted distinct minima (or MISMATCH)
    int k, w, seg; static char buf[1 << 20];
    while (scanf("%d %d %d %1048575s", &k, &w, &seg, buf) == 4) {
        const int len = (int)strlen(buf);
        for (int i = 0; i < len; i++) if (buf[i] >= '0' && buf[i] <= '3') buf[i] -= '0';   // raw 0..3 bytes
        std::vector<uint64_t> whole, sliced, vh(w + 1);
        hulk::k1_scan_read<false>(hulk::ByteSrc{(const uint8_t*)buf, len}, len, k, w, VH{vh.data()},
                                  [&](uint64_t m, bool on) { if (on) whole.push_back(m); });
        for (int64_t b = 0; b < len; b += seg)
            hulk::k1_scan_range((const uint8_t*)buf, len, k, w, b, b + seg, VH{vh.data()},
                                [&](uint64_t m) { sliced.push_back(m); });
        if (sliced != whole) { printf("MISMATCH\n"); continue; }
        // a range that ends past the sequence, an empty range, and the table size rule
        std::vector<uint64_t> tail;
        hulk::k1_scan_range((const uint8_t*)buf, len, k, w, len - 1, (int64_t)len + 1000, VH{vh.data()},
                            [&](uint64_t m) { tail.push_back(m); });
        hulk::k1_scan_range((const uint8_t*)buf, len, k, w, len, (int64_t)len + 5, VH{vh.data()},
                            [&](uint64_t m) { tail.push_back(~0ull); });
        if (tail.size() > 1 || (tail.size() == 1 && (whole.empty() || tail[0] != whole.back()))) { printf("MISMATCH_TAIL\n"); continue; }
        const uint64_t e = hulk::k1_long_table_entries((uint64_t)len, k) - 8;
        if ((e & (e - 1)) || e < 2 * (uint64_t)(len - k + 1) || e < 64) { printf("MISMATCH_CAP\n"); continue; }
        std::sort(whole.begin(), whole.end());
        whole.erase(std::unique(whole.begin(), whole.end()), whole.end());
        printf("%zu", whole.size());
        for (uint64_t x : whole) printf(" %llu", (unsigned long long)x);
        printf("\n");
    }
    return 0;
}
"""


def test_long_sequence_slices_compiled_for_host_match_one_pass_and_oracle(oracle, tmp_path):
    # k1_long.h is the exact source k1_long_scan compiles: a slice of a long sequence restarts the rolling k-mers and
    # the window k + w positions early and must emit what the sequential pass (minimizer.go:96-204) emits there
    src = tmp_path / "long_test.cpp"
    src.write_text(LONG_TEST_SRC)
    exe = tmp_path / "long_test"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O2", "-std=c++17", "-ffp-contract=off", "-I", os.path.join(ROOT, "hulk_b200", "csrc"),
                    "-o", str(exe), str(src)], check=True)
    from conftest import random_reads
    lines, want = [], []
    cases = [(21, 9, 2000, 256), (21, 9, 1500, 1), (21, 9, 700, 7), (31, 9, 3000, 320), (11, 9, 1000, 160), (4, 4, 300, 64),
             (21, 1, 400, 50), (15, 32, 1200, 376), (5, 9, 500, 33), (7, 40, 900, 100), (21, 200, 2500, 1768),
             (2, 2, 60, 3), (28, 9, 800, 296), (1, 9, 200, 16), (30, 256, 4000, 2288), (3, 7, 90, 90), (21, 9, 5000, 4999)]
    for k, w, L, seg in cases:
        reads = (random_reads(2, L, seed=k * 1000 + w) + random_reads(3, L, seed=k + w + seg, n_frac=0.03, lower_frac=0.3)
                 + [b"A" * L, (b"ACGTU" * L)[:L], (b"acgn0123RYKM" * L)[:L], (b"AC" * L)[:L],
                    (b"N" * (L // 2)) + random_reads(1, L - L // 2, seed=seg)[0]])
        for rd in reads:
            lines.append("%d %d %d %s" % (k, w, seg, rd.decode()))
            raw = bytes((c - 48) if 48 <= c <= 51 else c for c in rd)
            m = np.sort(oracle.minimizers(k, w, raw))
            want.append(" ".join([str(len(m))] + [str(int(x)) for x in m]))
    # and shapes nobody picked: random k, w, slice length (down to 1, up to past the end), sequence length and N density
    rng = np.random.default_rng(2024)
    for case in range(80):
        k = int(rng.integers(1, 32))
        w = int(rng.choice([1, 2, 5, 9, 9, 9, 16, 33, 100, 256]))
        L = int(rng.integers(k + w - 1, k + w + 3000))
        seg = int(rng.choice([1, 2, 17, 64, 256, 8 * (k + w), L, L + 7, int(rng.integers(1, L + 1))]))
        rd = random_reads(1, L, seed=5000 + case, n_frac=float(rng.choice([0.0, 0.01, 0.2])), lower_frac=0.2)[0]
        lines.append("%d %d %d %s" % (k, w, seg, rd.decode()))
        m = np.sort(oracle.minimizers(k, w, rd))
        want.append(" ".join([str(len(m))] + [str(int(x)) for x in m]))
    out = subprocess.run([str(exe)], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout
    got = out.strip().split("\n")
    assert len(got) == len(want)
    for ln, g, w_ in zip(lines, got, want):
        assert g == w_, ln[:40]


def test_bench_generators_on_cpu(monkeypatch):
    """bench.py's device-side generators, run here on torch's CPU device: the i.i.d. reads equal the host definition
    (hulk_b200.synthetic_reads) and the realism variant draws every read
    from the genome, either strand, independent of how the reads are split over steps and ranks."""
    import torch
    sys.path.insert(0, ROOT)
    import bench
    got = bench.synthetic_reads_torch(torch, 300, 150, 1, 1000, "cpu").numpy()
    np.testing.assert_array_equal(got, hulk_b200.synthetic_reads(300, 150, seed=1, first_read=1000))
    monkeypatch.setattr(bench, "GENOME_BASES", 100_000)
    genome = bench.genome_torch(torch, "cpu")
    assert genome.numel() == 100_000 and set(bytes(genome.numpy())) == set(b"ACGT")
    reads = bench.genome_reads_torch(torch, genome, 400, 150, 1, 0, "cpu")
    text = bytes(genome.numpy())
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    fwd = sum(bytes(r) in text for r in reads.numpy())
    rev = sum(bytes(r).translate(comp)[::-1] in text for r in reads.numpy())
    assert fwd + rev >= 400 and fwd > 100 and rev > 100
    again = bench.genome_reads_torch(torch, genome, 50, 150, 1, 200, "cpu")
    assert (again == reads[200:250]).all()


# ---- packed read transport, host side (csrc/pack.cpp) ------------------------------------------------
def _nt4_table():
    t = np.full(256, 4, np.uint8)
    t[:4] = [0, 1, 2, 3]
    for ch, c in zip("ACGTU", [0, 1, 2, 3, 3]):                     # src/minimizer/minimizer.go:13-30
        t[ord(ch)] = c
        t[ord(ch.lower())] = c
    return t


@pytest.mark.parametrize("isa", ["scalar", "avx2", "avx512"])
def test_pack_bases_is_the_nt4_table_two_bits_per_base(isa, monkeypatch):
    """hulk_b200_pack_bases against seq_nt4_table: codes of every non-4 base, ascending positions of the code-4 bases,
    for every byte value, every vector path the host has, one and several threads, sizes around the vector and piece
    boundaries."""
    import hulk_b200 as hb
    monkeypatch.setenv("HULK_B200_PACK_ISA", isa)        # a path the CPU lacks falls back to the next one down
    nt4 = _nt4_table()
    rng = np.random.default_rng(11)
    alphabets = [np.arange(256, dtype=np.uint8), np.frombuffer(b"ACGT", np.uint8),
                 np.frombuffer(b"ACGTacgtUuN\x00\x01\x02\x03", np.uint8)]
    for n in [0, 1, 3, 4, 5, 63, 64, 65, 127, 1000, 65535, 65536, 65537, 200001, 1_000_003]:
        for alpha in alphabets:
            b = alpha[rng.integers(0, alpha.size, n)].copy()
            codes = nt4[b]
            want_exc = np.nonzero(codes == 4)[0].astype(np.uint32)
            for threads in (1, 3):
                packed, exc, n_exc = hb.pack_bases(b, threads)
                assert packed.size == (n + 3) // 4
                assert n_exc == want_exc.size
                np.testing.assert_array_equal(exc, want_exc)
                got = np.unpackbits(packed, bitorder="little").reshape(-1, 2)[:n]
                got = got[:, 0] | (got[:, 1] << 1)
                keep = codes != 4                                    # the code stored under an exception is a don't-care
                np.testing.assert_array_equal(got[keep], codes[keep])
    # a cap smaller than the exception count: the true count is still reported
    b = np.full(5000, ord("N"), np.uint8)
    _, exc, n_exc = hb.pack_bases(b, 2, exceptions_cap=100)
    assert n_exc == 5000 and exc.size == 100 and (exc == np.arange(100)).all()
