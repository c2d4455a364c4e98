// k1_scan.h -- the per-read minimizer scan of stage 1, as host/device code so the exact source
// the kernel runs can also be compiled for the host and unit-tested there (tests/test_host_logic.py).
//
// Reference semantics (paths relative to the reference checkout):
//   src/minimizer/minimizer.go:96-204  findMinimizers (rolling 2-bit k-mer pair, canonical pick,
//       X = hash64(canon)<<8 | kmerSpan, monotone-deque window minimum, per-read set)
//
// What the deque computes, restated for a SIMT machine: position i (i >= k-1, fwd != rev) emits
//   m_i = min{ X_j : i-w < j <= i, j >= k-1, fwd_j != rev_j }   when i >= w-1,
// and the read contributes the SET {m_i}.  (The deque's tie rule only affects the stored
// position, which never leaves the function.)  The window minimum is evaluated with the
// two-pass block decomposition (prefix minima of the current block of w k-mers, suffix minima
// of the previous one) so every lane runs the same instruction stream regardless of data.
//
// Bases are consumed four at a time (one 32-bit word): word-wise 2-bit encoding with an
// "all ACGTU" check, then four k-mer hashes whose dependency chains are independent (ILP),
// then four window updates.
#pragma once
#include <stdint.h>

#include "hd_math.h"

namespace hulk {

// Source of bases: get4(i) returns bytes i..i+3 of the read (i % 4 == 0) as a little-endian
// word; bytes at or beyond the read's end are unspecified (never interpreted).
struct ByteSrc {              // any address space, byte loads only (never reads past `len`)
    const uint8_t *p;
    int32_t len;
    HULK_HD uint32_t get4(int32_t i) const {
        uint32_t w = 0;
HULK_UNROLL
        for (int u = 0; u < 4; u++)
            if (i + u < len) w |= (uint32_t)p[i + u] << (8 * u);
        return w;
    }
};

// 64-bit unsigned minimum / equality.  When every value is below 0x7FF0'0000'0000'0000 the bit
// patterns are non-negative finite doubles (or +inf for the sentinel), whose IEEE order equals the
// integer order, so one DSETP on the otherwise idle FP64 pipe replaces the two-instruction 64-bit integer
// compare on the (busiest) ALU pipe.  FP = false keeps the plain integer forms (k > 27, or w > k + 1 where the reference's
// kmerSpan goes negative and sign-extends into the top bits).
template <bool FP>
HULK_HD uint64_t umin64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    if (FP) return (__longlong_as_double((long long)a) < __longlong_as_double((long long)b)) ? a : b;   // DSETP + 2 SEL
#endif
    return (a < b) ? a : b;
}
template <bool FP>
HULK_HD bool ueq64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    if (FP) return __longlong_as_double((long long)a) == __longlong_as_double((long long)b);
#endif
    return a == b;
}
template <bool FP>
struct Sentinel {   // "no k-mer": larger than every minimizer value, never emitted
    static constexpr uint64_t value = FP ? 0x7FF0000000000000ull : ~0ull;
};
HULK_HD bool k1_fp_compare_ok(int32_t k, int32_t w) { return k <= 27 && w <= k + 1; }

// One window-minimum state machine (block decomposition).  vh(t): buffer of w + 1 entries, entry w
// stays at the sentinel so the look-ahead load needs no bounds test.  Written so that the only
// branch is the (warp-uniform) end-of-block test.
template <bool FP, class VH, class Emit>
struct WinMin {
    static constexpr uint64_t SENT = Sentinel<FP>::value;
    VH vh;
    Emit emit;
    int32_t w;
    int32_t t;
    uint64_t pref;
    HULK_HD WinMin(VH vh_, Emit emit_, int32_t w_) : vh(vh_), emit(emit_), w(w_), t(0), pref(SENT) {
        for (int x = 0; x <= w; x++) vh(x) = SENT;
    }
    // act: this position holds a k-mer (i >= k-1, inside the read); !real marks a skipped k-mer
    // (fwd == rev): it occupies its window position but never wins and emits nothing
    HULK_HD void step(uint64_t X, bool act, bool real, bool may_emit) {
        X = real ? X : SENT;
        pref = umin64<FP>(pref, X);
        const uint64_t suf = vh(t + 1);                                  // previous block, positions t+1..w-1
        if (act) vh(t) = X;
        emit(umin64<FP>(pref, suf), may_emit);                           // minimizer.go:186-199
        t += act ? 1 : 0;
        if (t == w) {                                                    // block complete: suffix minima in place
            uint64_t run = SENT;
            for (int x = w - 1; x >= 0; x--) {
                run = umin64<FP>(run, vh(x));
                vh(x) = run;
            }
            t = 0;
            pref = SENT;
        }
    }
};

// Scan one read.  `emit(m, on)` is called for every position; when `on` is true m is that position's
// window minimum (position order).  The caller has already applied the reference's length checks
// (minimizer.go:62-76) and provides a buffer of w + 1 entries.
template <bool FP, class Src, class VH, class Emit>
HULK_HD void k1_scan_read(const Src src, int32_t len, int32_t k, int32_t w, VH vh, Emit emit) {
    const uint64_t mask = (1ull << (2 * k)) - 1ull;       // minimizer.go:103  (k <= 31)
    const int shift = 2 * (k - 1);                        // minimizer.go:104
    uint64_t fwd = 0, rev = 0;
    WinMin<FP, VH, Emit> win(vh, emit, w);
    for (int32_t i0 = 0; i0 < len; i0 += 4) {
        const uint32_t word = src.get4(i0);
        bool fast;
        uint32_t codes = nt4x4(word, fast);
        if (!fast) {                                                      // N, IUPAC, raw 0..3 bytes, ...
            codes = 0;
HULK_UNROLL
            for (int u = 0; u < 4; u++) codes |= nt4((word >> (8 * u)) & 0xffu) << (8 * u);   // minimizer.go:115
        }
        uint64_t X[4], canon[4];
        bool skip[4];
HULK_UNROLL
        for (int u = 0; u < 4; u++) {
            const uint32_t c = (codes >> (8 * u)) & 0xffu;               // 0..4
            // (positions at or past the read's end roll garbage in; nothing after them is ever used)
            fwd = ((fwd << 2) | (uint64_t)c) & mask;                      // :134
            rev = (rev >> 2) | ((uint64_t)(3u ^ c) << shift);             // :137 (not masked)
            // (a k = 31 read with N's can set bit 62 of rev; the integer forms are used there)
            skip[u] = ueq64<FP>(fwd, rev);                                // :145-147
            canon[u] = umin64<FP>(fwd, rev);                              // :150-153
        }
        if (i0 + 3 < k - 1) continue;                                     // :140-142 (whole group before the first k-mer)
HULK_UNROLL
        for (int u = 0; u < 4; u++) {                                     // four independent hash chains
            const int32_t wi = i0 + u - w + 1;                            // windowIndex :112
            const int32_t span = (wi + 1 < k) ? (wi + 1) : k;             // :127-131
            X[u] = (hash64(canon[u], mask) << 8) | (uint64_t)(int64_t)span;   // :156-159
        }
HULK_UNROLL
        for (int u = 0; u < 4; u++) {
            const int32_t i = i0 + u;
            const bool act = (i >= k - 1) && (i < len);                   // :140-142
            win.step(X[u], act, act && !skip[u], act && !skip[u] && i >= w - 1);
        }
    }
}

// ---- w = 9 (the reference's default, cmd/sketch.go:52): the whole window state in registers -----
// With blocks of B = w - 1 = 8 positions a window of w positions always spans exactly two consecutive
// blocks: the suffix of the previous block from offset t and the prefix of the current block up to
// offset t.  So m_i = min(S_prev[t], P_cur[t]) with every index a compile-time constant once the block
// is unrolled, the suffix minima S live in eight registers, and no shared-memory window buffer (nor
// the loads/stores that maintain it) is needed.  Blocks are the absolute base ranges [8j, 8j + 8);
// positions that hold no k-mer (i < k-1, i >= len, fwd == rev) carry the sentinel, which is exactly
// "not in the deque" of the reference.

// Source of bases, eight at a time: next8() returns bytes i..i+7 of the read (i = 0, 8, 16, ...) as a
// little-endian 64-bit word; bytes at or beyond the read's end are unspecified (never interpreted).
struct ByteSrc8 {
    const uint8_t *p;
    int32_t len;
    int32_t i;
    HULK_HD uint64_t next8() {
        uint64_t v = 0;
HULK_UNROLL
        for (int u = 0; u < 8; u++)
            if (i + u < len) v |= (uint64_t)p[i + u] << (8 * u);
        i += 8;
        return v;
    }
};

// One group of four positions (base indices i0h .. i0h + 3, block offsets 4 H .. 4 H + 3).
// INTERIOR: every position of the block holds a k-mer with the full span and lies inside the read
// (i >= k + w - 2 and i < len), so the range tests and the span arithmetic drop out.
template <bool FP, bool INTERIOR, int H, class Emit>
HULK_HD void k1_w9_group(const uint32_t codes, const int32_t i0h, const int32_t len, const int32_t k,
                         const uint64_t mask, const int shift, const bool warm, uint64_t &fwd, uint64_t &rev,
                         uint64_t &pref, uint64_t (&A)[8], Emit &emit) {
    constexpr int32_t W = 9;
    constexpr uint64_t SENT = Sentinel<FP>::value;
    uint64_t X[4], canon[4];
    bool skip[4];
HULK_UNROLL
    for (int u = 0; u < 4; u++) {
        const uint32_t c = (codes >> (8 * u)) & 0xffu;                    // 0..4
        fwd = ((fwd << 2) | (uint64_t)c) & mask;                          // :134
        rev = (rev >> 2) | ((uint64_t)(3u ^ c) << shift);                 // :137 (not masked)
        skip[u] = ueq64<FP>(fwd, rev);                                    // :145-147
        canon[u] = umin64<FP>(fwd, rev);                                  // :150-153
    }
    if (warm) return;                                                     // :140-142
HULK_UNROLL
    for (int u = 0; u < 4; u++) {
        int32_t span = k;
        if (!INTERIOR) {
            const int32_t wi = i0h + u - W + 1;                           // windowIndex :112
            span = (wi + 1 < k) ? (wi + 1) : k;                           // :127-131
        }
        X[u] = (hash64(canon[u], mask) << 8) | (uint64_t)(int64_t)span;   // :156-159
    }
HULK_UNROLL
    for (int u = 0; u < 4; u++) {
        const int t = 4 * H + u;
        const int32_t i = i0h + u;
        const bool real = INTERIOR ? !skip[u] : ((i >= k - 1) && (i < len) && !skip[u]);
        const uint64_t Xe = real ? X[u] : SENT;
        pref = umin64<FP>(pref, Xe);
        const uint64_t m = umin64<FP>(pref, A[t]);                        // prefix of this block, suffix of the previous one
        A[t] = Xe;
        emit(m, INTERIOR ? real : (real && i >= W - 1));                  // minimizer.go:186-199
    }
}

// KC: compile-time k (0 = use the run-time argument): with k fixed the 2k-bit mask, the shifts and the
// xor-shift steps of hash64 on words that are known to be zero fold away.
template <bool FP, int KC = 0, class Src8, class Emit>
HULK_HD void k1_scan_read_w9(Src8 src, int32_t len, int32_t k_arg, Emit emit) {
    constexpr int32_t W = 9, B = 8;
    const int32_t k = KC ? KC : k_arg;
    constexpr uint64_t SENT = Sentinel<FP>::value;
    const uint64_t mask = (1ull << (2 * k)) - 1ull;       // minimizer.go:103  (k <= 31)
    const int shift = 2 * (k - 1);                        // minimizer.go:104
    uint64_t fwd = 0, rev = 0, pref = SENT;
    uint64_t A[B];                                        // A[t..7]: suffix minima of the previous block; A[0..t-1]: this block's values
HULK_UNROLL
    for (int x = 0; x < B; x++) A[x] = SENT;
    for (int32_t i0 = 0; i0 < len; i0 += B) {
        const uint64_t word = src.next8();
        uint32_t codes[2];
HULK_UNROLL
        for (int h = 0; h < 2; h++) {
            const uint32_t w32 = (uint32_t)(word >> (32 * h));
            bool fast;
            codes[h] = nt4x4(w32, fast);
            if (!fast) {                                                  // N, IUPAC, raw 0..3 bytes, ...
                codes[h] = 0;
HULK_UNROLL
                for (int u = 0; u < 4; u++) codes[h] |= nt4((w32 >> (8 * u)) & 0xffu) << (8 * u);   // minimizer.go:115
            }
        }
        const bool warm = i0 + (B - 1) < k - 1;                           // whole block in front of the first k-mer
        if (i0 >= k + W - 2 && i0 + B <= len) {                           // the bulk of every read
            k1_w9_group<FP, true, 0>(codes[0], i0, len, k, mask, shift, false, fwd, rev, pref, A, emit);
            k1_w9_group<FP, true, 1>(codes[1], i0 + 4, len, k, mask, shift, false, fwd, rev, pref, A, emit);
        } else {
            k1_w9_group<FP, false, 0>(codes[0], i0, len, k, mask, shift, warm, fwd, rev, pref, A, emit);
            k1_w9_group<FP, false, 1>(codes[1], i0 + 4, len, k, mask, shift, warm, fwd, rev, pref, A, emit);
            if (warm) continue;
        }
        uint64_t run = SENT;
HULK_UNROLL
        for (int x = B - 1; x >= 0; x--) {                                // suffix minima in place
            run = umin64<FP>(run, A[x]);
            A[x] = run;
        }
        pref = SENT;
    }
}

}  // namespace hulk
