// k1_scan2.h -- second-generation per-read minimizer scan for w = 9 (the reference default,
// cmd/sketch.go:52), written against the pipe budget of an sm_100a sub-partition instead of against
// the 64-bit source form.  Host/device code: tests/test_host_logic.py compiles this very file with g++
// and checks it against the CPU restatement of the reference.
//
// Reference semantics (paths relative to the reference checkout):
//   src/minimizer/minimizer.go:96-204  findMinimizers (rolling 2-bit k-mer pair, canonical pick,
//       X = hash64(canon)<<8 | kmerSpan, monotone-deque window minimum, per-read set)
//
// What changed against k1_scan.h (same block decomposition: blocks of w-1 = 8 positions, window minimum =
// min(suffix of the previous block, prefix of this one)):
//   * every 64-bit quantity is carried as two 32-bit halves and the 2k-bit masks of hash64 are deferred:
//     the low 2k bits of a product only depend on the low 2k bits of its factors, so only the xor-shifts
//     need clean inputs, and there the mask folds into the xor (one LOP3);
//   * values that fit 52 bits (2k + 8 <= 52, i.e. k <= 22) travel as the DOUBLE 2^52 + X: differences of two
//     such numbers are exact, so min(a, b) = a + ((b-a) - |b-a|)/2 is three exact FP64-pipe operations
//     (DADD, DADD, DFMA) and costs the (busiest) integer ALU pipe nothing -- there is no 64-bit integer
//     min on this machine, the compare-and-select form is one DSETP plus two ALU selects;
//   * k = 23..31: the span never exceeds 31, so bits 5..7 of X are always zero and X' = hash << 5 | span
//     (hash cut to the 56 bits that survive the reference's << 8) is a lossless, order-preserving 61-bit
//     form of X; it travels as a double BIT PATTERN (positive and finite: IEEE order = integer order,
//     one DSETP per compare) and X is rebuilt when the lists are drained;
//   * k <= 12: X itself fits 32 bits: one VIMNMX per minimum;
//   * a FAST block (8 positions, all inside the read, all with the full span, no byte other than
//     ACGTUacgtu among the last k, odd k so a k-mer never equals its reverse complement) has no
//     per-position range tests, no sentinel selects and no "skip" test; everything else (first and last
//     blocks of a read, the k + 9 positions behind an N, even k) takes the GENERAL block, which is the
//     reference's loop literally.  Both work on the same state, block by block.
#pragma once
#include <stdint.h>

#include <cmath>

#include "hd_math.h"
#include "k1_scan.h"

namespace hulk {

HULK_HD uint32_t fsh_l(uint32_t lo, uint32_t hi, uint32_t s) {   // high word of (hi:lo) << s, 0 < s < 32
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, s);
#else
    return (hi << s) | (lo >> (32u - s));
#endif
}
HULK_HD uint32_t fsh_r(uint32_t lo, uint32_t hi, uint32_t s) {   // low word of (hi:lo) >> s, 0 < s < 32
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, s);
#else
    return (lo >> s) | (hi << (32u - s));
#endif
}
HULK_HD uint64_t dbl_bits(double x) { return ((uint64_t)dbl_hi(x) << 32) | dbl_lo(x); }
HULK_HD double bits_dbl(uint64_t u) { return dbl_make((uint32_t)(u >> 32), (uint32_t)u); }

// ---- word-wise base encoding, second form ---------------------------------------------------
// Four ASCII bases -> four 2-bit codes (one per byte) and a word that is non-zero iff some byte is not one of
// ACGTUacgtu.  The expected letter is rebuilt arithmetically from the code (A + 2 c + 2 (c >> 1) + 11 [c == 3],
// byte-wise, no carries: at most 0x41 + 6 + 2 + 11), three multiply-adds on the FMA pipe instead of the
// byte-permute lookup of nt4x4().
HULK_HD uint32_t nt4x4b(uint32_t w, uint32_t &bad) {
    const uint32_t v = ((w >> 1) ^ (w >> 2)) & 0x03030303u;
    const uint32_t v1 = (v >> 1) & 0x01010101u;
    const uint32_t is_t = v & v1;                                        // code 3: T or U
    uint32_t e = v * 2u + 0x41414141u;
    e = v1 * 2u + e;
    e = is_t * 11u + e;
    bad = ((w & 0xDFDFDFDFu) ^ e) & ~is_t;                               // U = T + 1
    return v;
}

// ---- value representations ------------------------------------------------------------------
// ARITH (2k + 8 <= 52): a value is the double 2^52 + X, SENT = 2^53.
// BITS  (otherwise)   : a value is the double whose bit pattern is X' = (hash & (2^56 - 1)) << 5 | span, SENT = +inf.
template <int K>
struct K1Repr {
    static constexpr int BITS = 2 * K;
    static constexpr bool ARITH = (BITS + 8 <= 52);
    static constexpr uint64_t PREFIX = ARITH ? 0x4330000000000000ull : 0ull;
    static constexpr uint64_t SENT = ARITH ? 0x4340000000000000ull : 0x7FF0000000000000ull;
    static constexpr uint64_t MASK = (1ull << BITS) - 1ull;
    // stored form -> the reference's X
    static constexpr uint64_t H56 = (1ull << 56) - 1ull;
    // the reference's X -> stored form and back
    static HULK_HD uint64_t from_x(uint64_t x) { return ARITH ? (PREFIX | x) : (((x >> 8) << 5) | (x & 31ull)); }
    static HULK_HD uint64_t to_x(uint64_t stored) {
        return ARITH ? (stored & 0x000FFFFFFFFFFFFFull) : (((stored >> 5) << 8) | (stored & 31ull));
    }
};

// min of two values in either representation
template <bool ARITH>
HULK_HD double k1_vmin(double a, double b) {
    if (ARITH) {                                   // exact: both are integers in [2^52, 2^53]
        const double d = b - a;
        const double e = d - fabs(d);
        return fma(0.5, e, a);
    }
    return (b < a) ? b : a;                        // DSETP + two selects
}

// This lane's candidate list: window minima that differ from their predecessor, in position order.
// Entry e lives at base[e * STRIDE]; entries past `cap` are dropped but still counted (overflow test).
template <int STRIDE>
struct K1List {
    uint64_t *base;
    uint32_t cap;
    uint32_t n;
    double last;                                   // the previous window minimum (stored form)
    HULK_HD void emit(double m, bool on) {                               // minimizer.go:186-199
        const bool fresh = on && (m != last);
        if (fresh) base[(size_t)(n < cap ? n : cap - 1u) * STRIDE] = dbl_bits(m);
        n += fresh ? 1u : 0u;
        last = fresh ? m : last;
    }
};

// ---- FAST block: 8 positions inside the read, clean bases, odd k (17 <= k <= 31) ---------------
// PH = 0: interior (every position holds a k-mer with the full span).  The head of a read is the same for
// every read, so its blocks are the same code with the position-dependent parts folded at compile time:
// PH = 3: in front of the first k-mer (roll only); PH = 1: the block that holds position k-1 (offset U0: no
// k-mer before it, spans k-8.. from there, minimizer.go:127-131); PH = 2: the block after it (the remaining
// partial spans, then k).  PH = 4: the LAST block of a read whose length is not a multiple of 8: an interior block of
// which only the first `rem` positions exist (the codes behind them are arbitrary; what is computed from them is dropped).
template <int K, int STRIDE, int PH>
HULK_HD void k1_fast_block(const uint32_t codes0, const uint32_t codes1, uint32_t &f_lo, uint32_t &f_hi, uint32_t &r_lo,
                           uint32_t &r_hi, double (&A)[8], K1List<STRIDE> &L, const int32_t rem = 8) {
    using R = K1Repr<K>;
    static_assert(K >= 17 && K <= 31 && (K & 1), "fast block: odd k, two-word k-mers");
    constexpr int U0 = (K - 1) & 7;                // offset of position k-1 in its block
    constexpr uint32_t HM = (uint32_t)(R::MASK >> 32);
    constexpr int RSH = 2 * (K - 1) - 32 + 2;      // where the incoming complement code sits in r_hi before the shift
    const uint32_t rcodes0 = codes0 ^ 0x03030303u, rcodes1 = codes1 ^ 0x03030303u;
    double X[8];
HULK_UNROLL
    for (int u = 0; u < 8; u++) {
        const uint32_t c = byte_perm(u < 4 ? codes0 : codes1, 0u, 0x4440u | (uint32_t)(u & 3));
        const uint32_t rc = byte_perm(u < 4 ? rcodes0 : rcodes1, 0u, 0x4440u | (uint32_t)(u & 3));
        const uint32_t nf_hi = fsh_l(f_lo, f_hi, 2) & HM;                 // minimizer.go:134
        f_lo = f_lo * 4u + c;
        f_hi = nf_hi;
        const uint32_t nr_lo = fsh_r(r_lo, r_hi, 2);                      // minimizer.go:137
        r_hi = (r_hi + (rc << RSH)) >> 2;
        r_lo = nr_lo;
        if (PH == 3 || (PH == 1 && u < U0)) continue;                     // minimizer.go:140-142
        const int span = (PH == 0 || PH == 4) ? K : PH == 1 ? K - 8 + (u - U0) : (u < U0 ? K - U0 + u : K);   // :127-131
        const bool lt = dbl_make(f_hi, f_lo) < dbl_make(r_hi, r_lo);      // k-mers as bit patterns: tiny positive doubles
        uint32_t lo = lt ? f_lo : r_lo, hi = lt ? f_hi : r_hi;            // minimizer.go:150-153
        // hash64 (minimizer.go:33-42), masks deferred into the xor-shifts
        uint64_t p = (uint64_t)lo * 2097151ull + 0xFFFFFFFFFFFFFFFFull;
        lo = (uint32_t)p;
        hi = hi * 2097151u + (uint32_t)(p >> 32);
        {
            constexpr uint64_t M = R::MASK >> 24;                         // bits of key >> 24
            lo ^= fsh_r(lo, hi, 24) & (uint32_t)M;
            if (M >> 32) hi ^= ((hi & HM) >> 24);
        }
        p = (uint64_t)lo * 265ull;
        lo = (uint32_t)p;
        hi = hi * 265u + (uint32_t)(p >> 32);
        {
            constexpr uint64_t M = R::MASK >> 14;
            lo ^= fsh_r(lo, hi, 14) & (uint32_t)M;
            if (M >> 32) hi ^= ((hi & HM) >> 14);
        }
        p = (uint64_t)lo * 21ull;
        lo = (uint32_t)p;
        hi = hi * 21u + (uint32_t)(p >> 32);
        {
            constexpr uint64_t M = R::MASK >> 28;
            lo ^= fsh_r(lo, hi, 28) & (uint32_t)M;
            if (M >> 32) hi ^= ((hi & HM) >> 28);
        }
        p = (uint64_t)lo * 2147483649ull;
        lo = (uint32_t)p;
        hi = hi * 2147483649u + (uint32_t)(p >> 32);
        if (R::ARITH) {                                                   // 2^52 + (hash << 8 | k)
            constexpr uint32_t XM = (uint32_t)(((R::MASK << 8) | 0xffull) >> 32);
            X[u] = dbl_make((fsh_l(lo, hi, 8) & XM) | 0x43300000u, lo * 256u + (uint32_t)span);
        } else {                                                          // bit pattern hash56 << 5 | k
            constexpr uint32_t YM = (uint32_t)((((R::MASK & R::H56) << 5) | 31ull) >> 32);
            X[u] = dbl_make(fsh_l(lo, hi, 5) & YM, lo * 32u + (uint32_t)span);
        }
    }
    if (PH == 3) return;
    constexpr int T0 = PH == 1 ? U0 : 0;
    double pref = X[T0];
HULK_UNROLL
    for (int t = 0; t < 8; t++) {
        if (t < T0) {                                                     // no k-mer here: "not in the deque"
            A[t] = bits_dbl(R::SENT);
            continue;
        }
        if (t > T0) pref = k1_vmin<R::ARITH>(pref, X[t]);
        const double m = k1_vmin<R::ARITH>(A[t], pref);                   // suffix of the previous block, prefix of this one
        A[t] = X[t];
        if (PH == 4) {
            const bool fresh = t < rem && m != L.last;                    // positions at or behind the read's end: nothing
            if (fresh) {
                L.base[(size_t)L.n * STRIDE] = dbl_bits(m);
                L.n++;
                L.last = m;
            }
            continue;
        }
        if (m != L.last) {                                                // minimizer.go:186-199 (room for 8 was checked)
            L.base[(size_t)L.n * STRIDE] = dbl_bits(m);
            L.n++;
        }
        L.last = m;
    }
    if (PH != 4) {                                                        // (nothing comes behind a read's last block)
HULK_UNROLL
        for (int x = 6; x >= 0; x--) A[x] = k1_vmin<R::ARITH>(A[x], A[x + 1]);   // suffix minima in place
    }
}

// ---- GENERAL block: the reference's loop, position by position, in the same representation -----
template <int K, int STRIDE>
HULK_HD void k1_general_block(const uint32_t codes0, const uint32_t codes1, const int32_t i0, const int32_t len,
                              uint64_t &fwd, uint64_t &rev, double (&A)[8], K1List<STRIDE> &L) {
    using R = K1Repr<K>;
    constexpr int32_t W = 9;
    constexpr uint64_t mask = R::MASK;                                    // minimizer.go:103
    constexpr int shift = 2 * (K - 1);                                    // minimizer.go:104
    const double SENT = bits_dbl(R::SENT);
    const bool warm = i0 + 7 < K - 1;                                     // whole block in front of the first k-mer
    double pref = SENT;
HULK_UNROLL
    for (int h = 0; h < 2; h++) {
        const uint32_t codes = h ? codes1 : codes0;
        uint64_t canon[4];
        bool skip[4];
HULK_UNROLL
        for (int u = 0; u < 4; u++) {
            const uint32_t c = (codes >> (8 * u)) & 0xffu;                // 0..4
            fwd = ((fwd << 2) | (uint64_t)c) & mask;                      // :134
            rev = (rev >> 2) | ((uint64_t)(3u ^ c) << shift);             // :137 (not masked)
            skip[u] = fwd == rev;                                         // :145-147
            canon[u] = fwd < rev ? fwd : rev;                             // :150-153
        }
        if (warm) continue;                                               // :140-142
HULK_UNROLL
        for (int u = 0; u < 4; u++) {
            const int t = 4 * h + u;
            const int32_t i = i0 + t;
            const int32_t wi = i - W + 1;                                 // windowIndex :112
            const int32_t span = (wi + 1 < K) ? (wi + 1) : K;             // :127-131 (>= 0 for k >= 8)
            const uint64_t X = (hash64(canon[u], mask) << 8) | (uint64_t)(int64_t)span;   // :156-159
            const double Xs = bits_dbl(R::from_x(X));
            const bool real = (i >= K - 1) && (i < len) && !skip[u];
            const double Xe = real ? Xs : SENT;
            pref = k1_vmin<R::ARITH>(pref, Xe);
            const double m = k1_vmin<R::ARITH>(A[t], pref);
            A[t] = Xe;
            L.emit(m, real && i >= W - 1);                                // :186-199
        }
    }
    if (warm) return;
HULK_UNROLL
    for (int x = 6; x >= 0; x--) A[x] = k1_vmin<R::ARITH>(A[x], A[x + 1]);
}

// bytes of a word -> codes through the reference's table (minimizer.go:13-30, :115)
HULK_HD uint32_t nt4_bytes(uint32_t w) {
    uint32_t codes = 0;
HULK_UNROLL
    for (int u = 0; u < 4; u++) codes |= nt4((w >> (8 * u)) & 0xffu) << (8 * u);
    return codes;
}

// Source of bases for this scan: next(w0, w1) returns bytes i..i+3 and i+4..i+7 of the read (i = 0, 8, 16, ...)
// as little-endian words; bytes at or beyond the read's end are unspecified (never interpreted).
struct ByteSrcW {
    const uint8_t *p;
    int32_t len;
    int32_t i;
    HULK_HD void next(uint32_t &w0, uint32_t &w1) {
        uint32_t w[2] = {0, 0};
HULK_UNROLL
        for (int u = 0; u < 8; u++)
            if (i + u < len) w[u >> 2] |= (uint32_t)p[i + u] << (8 * (u & 3));
        i += 8;
        w0 = w[0];
        w1 = w[1];
    }
};

// ---- k = 9, 11 (odd, 2k + 8 <= 32): everything fits one 32-bit word -----------------------------------------
// The k-mer pair is one word each, hash64 works on 2k <= 22 bits -- its `>> 24` and `>> 28` xor-shifts and its last
// multiplication (key + key << 31) are the identity there -- X = hash << 8 | span has at most 30 bits, and a minimum is one
// integer instruction.  Same block decomposition and the same phases as k1_fast_block; the state between blocks is the
// k-mer pair and eight suffix minima as 32-bit words.  The list still holds the 64-bit stored form (2^52 + X).
constexpr uint32_t K1_SENT32 = 0xFFFFFFFFu;                                // larger than any X
template <int K, int STRIDE, int PH>
HULK_HD void k1_fast_block32(const uint32_t codes0, const uint32_t codes1, uint32_t &f, uint32_t &r, uint32_t (&A)[8],
                             uint32_t &last, K1List<STRIDE> &L, const int32_t rem = 8) {
    static_assert((K == 9 || K == 11), "32-bit block: odd k with 2k + 8 <= 32 and k - 1 >= 8");
    constexpr int U0 = (K - 1) & 7;
    constexpr uint32_t M = (1u << (2 * K)) - 1u;
    constexpr int SH = 2 * (K - 1);
    uint32_t X[8];
HULK_UNROLL
    for (int u = 0; u < 8; u++) {
        const uint32_t c = ((u < 4 ? codes0 : codes1) >> (8 * (u & 3))) & 0xffu;
        f = ((f << 2) | c) & M;                                           // minimizer.go:134
        r = (r >> 2) | ((3u ^ c) << SH);                                  // :137
        if (PH == 3 || (PH == 1 && u < U0)) continue;                     // :140-142
        const int span = (PH == 0 || PH == 4) ? K : PH == 1 ? K - 8 + (u - U0) : (u < U0 ? K - U0 + u : K);   // :127-131
        uint32_t key = f < r ? f : r;                                     // :150-153 (odd k: never equal)
        key = key * 2097151u + 0xFFFFFFFFu;                               // hash64 (:33-42), masks deferred to the xor-shift
        key = key * 265u;
        key &= M;
        key ^= key >> 14;
        key = key * 21u;
        X[u] = ((key << 8) & (M << 8)) | (uint32_t)span;                  // :156-159
    }
    if (PH == 3) return;
    constexpr int T0 = PH == 1 ? U0 : 0;
    uint32_t pref = X[T0];
HULK_UNROLL
    for (int t = 0; t < 8; t++) {
        if (t < T0) {
            A[t] = K1_SENT32;
            continue;
        }
        if (t > T0) pref = pref < X[t] ? pref : X[t];
        const uint32_t m = A[t] < pref ? A[t] : pref;
        A[t] = X[t];
        const bool fresh = (PH != 4 || t < rem) && m != last;             // minimizer.go:186-199 (room for 8 was checked)
        if (fresh) {
            L.base[(size_t)L.n * STRIDE] = K1Repr<K>::PREFIX | (uint64_t)m;
            L.n++;
            last = m;
        }
    }
    if (PH != 4) {
HULK_UNROLL
        for (int x = 6; x >= 0; x--) A[x] = A[x] < A[x + 1] ? A[x] : A[x + 1];
    }
}

template <int K, int STRIDE, class Src>
HULK_HD void k1_scan_read_w9_32(Src src, const int32_t len, K1List<STRIDE> &L) {
    using R = K1Repr<K>;
    uint32_t f = 0, r = 0, last = K1_SENT32;
    uint32_t A[8];
HULK_UNROLL
    for (int x = 0; x < 8; x++) A[x] = K1_SENT32;
    L.last = bits_dbl(R::SENT);
    int32_t fast_from = K + 7;
    for (int32_t i0 = 0; i0 < len; i0 += 8) {
        uint32_t w0, w1, bad0, bad1;
        src.next(w0, w1);
        uint32_t c0 = nt4x4b(w0, bad0), c1 = nt4x4b(w1, bad1);
        const int32_t rem = len - i0;
        if (rem < 8) {
            const uint32_t keep0 = rem >= 4 ? 0xFFFFFFFFu : (1u << (8 * rem)) - 1u;
            const uint32_t keep1 = rem <= 4 ? 0u : (1u << (8 * (rem - 4))) - 1u;
            bad0 &= keep0;
            bad1 &= keep1;
        }
        if (bad0 | bad1) {
            c0 = nt4_bytes(w0);
            c1 = nt4_bytes(w1);
            fast_from = i0 + K + 9;
        }
        constexpr int32_t HEAD = 8 * ((K - 1) >> 3);
        const bool room = L.n + 8u <= L.cap;
        const bool quick = room && rem >= 8;
        if (room && rem < 8 && i0 >= fast_from) {
            k1_fast_block32<K, STRIDE, 4>(c0, c1, f, r, A, last, L, rem);
        } else if (quick && i0 >= fast_from) {
            k1_fast_block32<K, STRIDE, 0>(c0, c1, f, r, A, last, L);
        } else if (quick && fast_from == K + 7) {
            if (i0 < HEAD) k1_fast_block32<K, STRIDE, 3>(c0, c1, f, r, A, last, L);
            else if (i0 == HEAD) k1_fast_block32<K, STRIDE, 1>(c0, c1, f, r, A, last, L);
            else k1_fast_block32<K, STRIDE, 2>(c0, c1, f, r, A, last, L);
        } else {
            // the reference's loop on the 64-bit state (a code 4 makes rev outgrow 2k bits: minimizer.go:137 does not mask it)
            uint64_t fwd = f, rev = r;
            double Ad[8];
HULK_UNROLL
            for (int x = 0; x < 8; x++) Ad[x] = A[x] == K1_SENT32 ? bits_dbl(R::SENT) : bits_dbl(R::PREFIX | (uint64_t)A[x]);
            L.last = last == K1_SENT32 ? bits_dbl(R::SENT) : bits_dbl(R::PREFIX | (uint64_t)last);
            k1_general_block<K, STRIDE>(c0, c1, i0, len, fwd, rev, Ad, L);
            // back to words: a pair that no longer fits (a code 4 still inside it) keeps the general block until it has left
            if ((fwd | rev) >> 32) fast_from = i0 + K + 9 > fast_from ? i0 + K + 9 : fast_from;
            f = (uint32_t)fwd;
            r = (uint32_t)rev;
HULK_UNROLL
            for (int x = 0; x < 8; x++) {
                const uint64_t bits = dbl_bits(Ad[x]);
                A[x] = bits == R::SENT ? K1_SENT32 : (uint32_t)R::to_x(bits);
            }
            const uint64_t lb = dbl_bits(L.last);
            last = lb == R::SENT ? K1_SENT32 : (uint32_t)R::to_x(lb);
        }
    }
}

// Scan one read of at least 9 + k - 1 bases (the caller applied minimizer.go:62-76); the candidates end up in L
// (stored form of K1Repr<K>, position order, adjacent duplicates removed).  8 <= k <= 31.
template <int K, int STRIDE, class Src>
HULK_HD void k1_scan_read_w9_v2(Src src, const int32_t len, K1List<STRIDE> &L) {
    using R = K1Repr<K>;
    static_assert(K >= 8 && K <= 31, "w = 9 needs k >= 8 for a non-negative span");
    if constexpr (K == 9 || K == 11) {                                    // one-word k-mers and values
        k1_scan_read_w9_32<K, STRIDE>(src, len, L);
        return;
    }
    constexpr bool HAS_FAST = (K & 1) && K >= 17;
    const double SENT = bits_dbl(R::SENT);
    uint32_t f_lo = 0, f_hi = 0, r_lo = 0, r_hi = 0;
    double A[8];
HULK_UNROLL
    for (int x = 0; x < 8; x++) A[x] = SENT;
    L.last = SENT;                                                        // never a value
    int32_t fast_from = K + 7;                                            // first block start with the full span everywhere
    for (int32_t i0 = 0; i0 < len; i0 += 8) {
        uint32_t w0, w1, bad0, bad1;
        src.next(w0, w1);
        uint32_t c0 = nt4x4b(w0, bad0), c1 = nt4x4b(w1, bad1);
        const int32_t rem = len - i0;                                     // positions of this block inside the read
        if (rem < 8) {                                                    // bytes at or past the end say nothing about the read
            const uint32_t keep0 = rem >= 4 ? 0xFFFFFFFFu : (1u << (8 * rem)) - 1u;
            const uint32_t keep1 = rem <= 4 ? 0u : (1u << (8 * (rem - 4))) - 1u;
            bad0 &= keep0;
            bad1 &= keep1;
        }
        if (bad0 | bad1) {                                                // N, IUPAC, raw 0..3 bytes, ...
            c0 = nt4_bytes(w0);
            c1 = nt4_bytes(w1);
            fast_from = i0 + K + 9;                                       // a code 4 lives k steps in fwd, k + 1 in rev
        }
        constexpr int32_t HEAD = 8 * ((K - 1) >> 3);                       // start of the block that holds position k-1
        const bool room = HAS_FAST && L.n + 8u <= L.cap;
        const bool quick = room && rem >= 8;
        if (room && rem < 8 && i0 >= fast_from) {                         // the read's last, partial block
            if constexpr (HAS_FAST) k1_fast_block<K, STRIDE, 4>(c0, c1, f_lo, f_hi, r_lo, r_hi, A, L, rem);
        } else if (quick && i0 >= fast_from) {
            if constexpr (HAS_FAST) k1_fast_block<K, STRIDE, 0>(c0, c1, f_lo, f_hi, r_lo, r_hi, A, L);
        } else if (quick && fast_from == K + 7) {                         // head of a read without any foreign byte so far
            if constexpr (HAS_FAST) {
                if (i0 < HEAD) k1_fast_block<K, STRIDE, 3>(c0, c1, f_lo, f_hi, r_lo, r_hi, A, L);
                else if (i0 == HEAD) k1_fast_block<K, STRIDE, 1>(c0, c1, f_lo, f_hi, r_lo, r_hi, A, L);
                else k1_fast_block<K, STRIDE, 2>(c0, c1, f_lo, f_hi, r_lo, r_hi, A, L);
            }
        } else {
            uint64_t fwd = ((uint64_t)f_hi << 32) | f_lo, rev = ((uint64_t)r_hi << 32) | r_lo;
            k1_general_block<K, STRIDE>(c0, c1, i0, len, fwd, rev, A, L);
            f_lo = (uint32_t)fwd;
            f_hi = (uint32_t)(fwd >> 32);
            r_lo = (uint32_t)rev;
            r_hi = (uint32_t)(rev >> 32);
        }
    }
}

}  // namespace hulk
