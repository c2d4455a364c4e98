// k1_long.h -- a slice of one long sequence, scanned on its own.  Host/device code: tests/test_host_logic.py
// compiles this file with g++ and checks that the slices of a sequence, put end to end, emit exactly what
// one pass over the whole sequence emits.
//
// Reference semantics (paths relative to the reference checkout):
//   src/minimizer/minimizer.go:96-204  findMinimizers, one sequential pass per sequence.  `hulk sketch --fasta`
//   hands whole contigs / chromosomes to it (src/pipeline/sketch.go:99-135), so a "read" may be hundreds of
//   megabases long.
//
// Why a slice can be scanned alone.  What position i emits is min{ X_j : i-w < j <= i } over the positions that
// hold a k-mer (k1_scan.h), and X_j is a function of
//   * the forward k-mer: the last k base codes (the masked register forgets everything older; a code-4 byte
//     bleeds one bit into the slot of the base before it, which is still one of the last k),
//   * the reverse k-mer: the last k + 1 base codes (the register is not masked, minimizer.go:137: the third
//     bit of 3 ^ 4 = 7 enters at bit 2k and is shifted out k + 1 steps later),
//   * kmerSpan, a function of the absolute position (minimizer.go:127-131).
// So a scan that starts from zeroed registers k + w positions in front of `begin` has the reference's registers
// from its position k + 1 on and the reference's window from its position k + w on -- exactly at `begin`.
// Positions in front of `begin` roll and fill the window but emit nothing; the absolute position is kept for
// the span and for the two start-of-sequence rules (i < k-1: no k-mer yet, i < w-1: no full window yet).
#pragma once
#include <stdint.h>

#include "hd_math.h"
#include "k1_scan.h"

namespace hulk {

// number of positions a slice rolls in front of its first emitting position
HULK_HD int64_t k1_long_warmup(int32_t k, int32_t w) { return (int64_t)k + (int64_t)w; }

// Positions [begin, end) of the sequence seq[0, len): emit(m) is called, in position order, with the window
// minimum of every position of the range that emits one in the reference's loop.  `vh`: w + 1 entries.
template <class VH, class Emit>
HULK_HD void k1_scan_range(const uint8_t *seq, int64_t len, int32_t k, int32_t w, int64_t begin, int64_t end, VH vh,
                           Emit emit) {
    if (end > len) end = len;
    if (begin >= end) return;
    int64_t s0 = begin - k1_long_warmup(k, w);
    if (s0 < 0) s0 = 0;
    const uint64_t mask = (1ull << (2 * k)) - 1ull;                          // minimizer.go:103
    const int shift = 2 * (k - 1);                                           // minimizer.go:104
    uint64_t fwd = 0, rev = 0;
    auto gate = [&](uint64_t m, bool on) { if (on) emit(m); };
    WinMin<false, VH, decltype(gate)> win(vh, gate, w);
    for (int64_t i = s0; i < end; i++) {
        const uint64_t c = nt4(seq[i]);                                      // :115
        fwd = ((fwd << 2) | c) & mask;                                       // :134
        rev = (rev >> 2) | ((3ull ^ c) << shift);                            // :137
        if (i - s0 < k - 1) continue;                                        // :140-142 (s0 = 0), registers still filling (s0 > 0)
        const bool real = fwd != rev;                                        // :145-147
        const int64_t wi = i - w + 1;                                        // windowIndex :112
        const int64_t span = (wi + 1 < k) ? (wi + 1) : k;                    // :127-131
        const uint64_t X = (hash64(fwd < rev ? fwd : rev, mask) << 8) | (uint64_t)span;   // :150-159
        win.step(X, true, real, real && i >= w - 1 && i >= begin);           // :186-199
    }
}

// The per-sequence set (mapset, minimizer.go:189-198) when many slices of one sequence insert at once: open
// addressing over `cap` entries (a power of two, at least twice the number of k-mers), 0 = empty, the value 0
// itself is tracked in entry `cap`.  Returns true for the first insertion of a value.
HULK_HD uint64_t k1_long_table_entries(uint64_t len, int32_t k) {            // entries to reserve, flag entry included
    const uint64_t nk = len - (uint64_t)k + 1;
    uint64_t cap = 64;
    while (cap < 2 * nk && cap < (1ull << 62)) cap <<= 1;                    // (terminates on garbage lengths too)
    return cap + 8;                                                          // (keeps every table 64-byte aligned)
}
HULK_HD uint64_t k1_long_slot(uint64_t m) { return (m * 0x9E3779B97F4A7C15ull) >> 17; }

}  // namespace hulk
