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

// One window-minimum state machine (block decomposition).  vh(t): w-entry buffer.
template <class VH, class Emit>
struct WinMin {
    VH vh;
    Emit emit;
    int32_t w;
    int32_t t;
    uint64_t pref;
    HULK_HD WinMin(VH vh_, Emit emit_, int32_t w_) : vh(vh_), emit(emit_), w(w_), t(0), pref(~0ull) {
        for (int x = 0; x < w; x++) vh(x) = ~0ull;
    }
    // X == ~0 marks a skipped k-mer (fwd == rev): it never wins and emits nothing
    HULK_HD void step(uint64_t X, bool may_emit) {
        pref = (X < pref) ? X : pref;
        const uint64_t suf = (t + 1 < w) ? vh(t + 1) : ~0ull;            // previous block, positions t+1..w-1
        vh(t) = X;
        if (may_emit) emit((pref < suf) ? pref : suf);                   // minimizer.go:186-199
        if (++t == w) {                                                  // block complete: suffix minima in place
            uint64_t run = ~0ull;
            for (int x = w - 1; x >= 0; x--) {
                const uint64_t v = vh(x);
                run = (v < run) ? v : run;
                vh(x) = run;
            }
            t = 0;
            pref = ~0ull;
        }
    }
};

// Scan one read.  `emit(m)` receives every window minimum in position order.
// The caller has already applied the reference's length checks (minimizer.go:62-76).
template <class Src, class VH, class Emit>
HULK_HD void k1_scan_read(const Src src, int32_t len, int32_t k, int32_t w, VH vh, Emit emit) {
    const uint64_t mask = (1ull << (2 * k)) - 1ull;       // minimizer.go:103  (k <= 31)
    const int shift = 2 * (k - 1);                        // minimizer.go:104
    uint64_t fwd = 0, rev = 0;
    WinMin<VH, Emit> win(vh, emit, w);
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
            if (i0 + u < len) {
                fwd = ((fwd << 2) | (uint64_t)c) & mask;                  // :134
                rev = (rev >> 2) | ((uint64_t)(3u ^ c) << shift);         // :137 (not masked)
            }
            skip[u] = (fwd == rev);                                       // :145-147
            canon[u] = (fwd > rev) ? rev : fwd;                           // :150-153
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
            if (i >= k - 1 && i < len)                                    // :140-142
                win.step(skip[u] ? ~0ull : X[u], !skip[u] && i >= w - 1);
        }
    }
}

}  // namespace hulk
