// k1_minimizer.cuh -- stage 1+2 of the sketch hot path on sm_100a:
//   reads (ASCII, in HBM) -> per-read SET of window minimizers -> jump-hash bin -> histogram.
//
// Reference semantics (paths relative to the reference checkout):
//   src/minimizer/minimizer.go:96-204  findMinimizers (see k1_scan.h for the scan itself)
//   src/pipeline/minion.go:51-57 + boss.go:90-95 -> src/kmerspectrum/kmerspectrum.go:67-81 AddHash
//
// Work decomposition: one thread per read, in three phases:
//   scan    every lane walks its read (k1_scan.h); window minima that differ from the previous one go to
//           the lane's private list in shared memory;
//   de-dup  the list is reduced in place to the exact per-read set (minimizer.go:189-198);
//   jump    every set member is binned with jump.Hash.  In the production configuration the warp's lists
//           are compacted and appended to a per-batch queue in global memory, and a second kernel
//           (k1_jump_queue: no shared memory, one wave of warps, keys handed out inside each warp as walks
//           finish) bins the whole queue; two jump-hash walks per lane are in flight (ILP) and advance
//           K1_JUMP_BATCH steps between two refill points with the bracketed fast step of hd_math.h.
//           Bins are counted with global (L2) atomics -- the k^4-bin histogram does not fit shared memory
//           for k > 11.
// Two scan kernels: k1_minimizer_histogram_w9 (w = 9, the reference default: window state in registers,
// reads streamed from global memory, warps independent, 32-read tasks handed out dynamically) and
// k1_minimizer_histogram (any w <= 32: 128 reads per CTA tile staged by one 1-D bulk TMA copy, window
// buffers in shared memory).  k1_generic takes what neither can (w > 32, reads whose candidate list
// overflows).
#pragma once
#include <stdint.h>

#include "hd_math.h"
#include "k1_scan.h"
#include "k1_scan2.h"
#include "ptx_util.cuh"

namespace hulk {

constexpr int K1_TPB = 128;       // threads (= reads) per CTA tile
constexpr int K1_W_FAST = 32;     // largest w handled by the shared-memory path

// A long sequence (k1_long.cuh): its slices are scanned by many threads that share one set.
struct K1LongTask {
    uint64_t r, b0, len;           // read index in the batch, first byte, length
    uint64_t tab, cap;             // its open-addressing table: first arena entry, entries (power of two; flag at [cap])
    uint64_t seg_first, n_seg;     // its slices are the batch's slices seg_first .. seg_first + n_seg - 1
};
struct K1LongCtl {
    unsigned long long n_tasks, n_segs;
    unsigned long long zero_begin, zero_end;   // arena entries the tables of this launch occupy
};

struct K1Params {
    const uint8_t *bases;          // device
    uint64_t bases_bytes;          // bytes that may legally be read starting at `bases`
    const uint64_t *offsets;       // device, n_reads+1 entries, or nullptr when fixed_len != 0
    uint32_t fixed_len;
    uint64_t n_reads;
    uint64_t read_base;            // global index of read 0 of this batch (error reporting)
    uint64_t off_base;             // value of offsets[] that corresponds to bases[0]
    uint32_t list_cap;             // per-read candidate list capacity (window minima, adjacent-distinct)
    uint32_t k, w;
    int32_t D;
    uint32_t tile_cap;             // bytes per shared-memory tile buffer (multiple of 16)
    uint32_t *hist;                // D bins
    unsigned long long *n_minimizers;
    unsigned long long *err_word;  // (global read index << 8 | code) of the first offending read
    // reads the fast path cannot finish (list overflow) are queued for k1_generic
    unsigned int *ovf_count;
    unsigned long long *ovf_list;
    uint32_t ovf_cap;
    // parity tap: dump the per-read sets instead of counting them
    uint64_t *dump;
    uint32_t dump_cap;
    uint32_t *dump_counts;
    // the batch's minimizer queue: k1 scan kernels append their per-read sets, k1_jump_queue bins them
    uint64_t *queue;
    unsigned long long *queue_cursor;      // [0]: keys appended so far, [1]: next 32-read task (dynamic scheduling)
    uint64_t queue_cap;
    // generic path scratch
    // (arena entries [0, slab_entries) are the k1_generic threads' own tables, slab_size each; the rest is handed out
    // through arena_cursor)
    uint64_t *arena;
    unsigned long long *arena_cursor;
    uint64_t arena_entries;
    uint64_t slab_entries;
    uint32_t slab_size;
    // sequences of long_min bases or more are left to k1_long.cuh by every other kernel (~0: there are none)
    uint64_t long_min;
    uint32_t long_seg;             // positions per slice
    uint32_t long_cap;             // entries of long_tasks
    K1LongTask *long_tasks;
    K1LongCtl *long_ctl;
    // MinHash feed (k1_minhash.cuh): the kernels that bin their minimizers themselves also append them to the queue
    uint32_t feed_queue;
};

constexpr uint32_t K1_SLAB_MAX = 8192;   // most entries of a k1_generic thread's own table (sets of up to 4096 values)
constexpr uint32_t K1_ERR_EMPTY = 3;   // HULK_B200_EEMPTYSEQ
constexpr uint32_t K1_ERR_SHORT = 4;   // HULK_B200_ESHORTSEQ
constexpr uint32_t K1_ERR_OVF = 31;    // overflow queue / arena exhausted -> HULK_B200_ENOMEM

__device__ __forceinline__ uint64_t k1_read_off(const K1Params &p, uint64_t r) {
    return p.offsets ? (p.offsets[r] - p.off_base) : r * (uint64_t)p.fixed_len;
}
__device__ __forceinline__ void k1_report(const K1Params &p, uint64_t r, uint32_t code) {
    atomicMin(p.err_word, (unsigned long long)(((p.read_base + r) << 8) | code));
}
// one more member of read r's set for the batch queue (the queue is sized for them when feed_queue is set)
__device__ __forceinline__ void k1_feed_queue(const K1Params &p, uint64_t r, uint64_t m) {
    const unsigned long long at = atomicAdd(p.queue_cursor, 1ull);
    if (at < p.queue_cap) p.queue[at] = m;
    else k1_report(p, r, K1_ERR_OVF);
}

struct K1SmemVH {
    uint64_t *base;   // [w][K1_TPB], this thread's column
    __device__ __forceinline__ uint64_t &operator()(int t) const { return base[t * K1_TPB]; }
};

// bases of a read staged in shared memory: aligned 32-bit loads + funnel shift (the buffer has
// 16 bytes of slack behind the tile, so the word behind the last base may be read, never used)
struct SmemWordSrc {
    const uint32_t *words;    // aligned word that holds base 0
    uint32_t sh;              // (byte offset of base 0 inside that word) * 8
    __device__ __forceinline__ uint32_t get4(int32_t i) const {
        const uint32_t lo = words[i >> 2], hi = words[(i >> 2) + 1];
        return __funnelshift_r(lo, hi, sh);
    }
};

constexpr int K1_JUMP_BATCH = 4;      // jump steps between two refills of a lane's chains
constexpr int K1_WARPS = K1_TPB / 32;

// keeps a loop-invariant double in registers (the compiler would otherwise re-materialise the
// 64-bit immediate with two moves in front of every use)
// Loop-invariant doubles of the jump step live in the constant bank: DFMA/DMUL take a c[bank][offset]
// operand directly, so no instruction is spent re-materialising a 64-bit immediate in front of each use.
__constant__ double k1_jump_consts[2] = {1.0 - JUMP_EPS, 1.0 + JUMP_EPS};
__constant__ double k1_jump_fx_consts[3] = {JUMP_TWO52 - 1.0, JUMP_TWO83M, 1.0};

__device__ __forceinline__ double k1_pin(double x) {
    asm volatile("" : "+d"(x));
    return x;
}

// ---- jump.Hash walks over a key queue (kmerspectrum.go:67-81: bins[jump.Hash(kmer, numBins)]++) ----
// This thread bins keys g, g + stride, g + 2 stride, ... < total.  Two walks per thread are in flight
// (ILP); they advance K1_JUMP_BATCH steps between two refill points.  The step itself is hd_math.h
// jump_step_fast, split into evaluate / commit.  Must be entered by whole warps.
// DYN = false: static assignment, this thread's keys are g, g + stride, ...
// DYN = true : the warp owns keys [g, total) (g warp-uniform, stride unused); a walk that ends takes the
//              warp's next unclaimed key (ballot + popc, no atomics), so every lane stays busy until the
//              segment is exhausted, whatever the per-key step counts are.
template <bool DYN, int BATCH, class Load>
__device__ __forceinline__ void k1_jump_walk(Load load, uint32_t g, const uint32_t stride, const uint32_t total,
                                             uint32_t *const hist, const uint32_t nb) {
    const uint32_t lane_lt = (1u << (threadIdx.x & 31)) - 1u;
    uint64_t key[2] = {0, 0};
    double jd1[2] = {1.0, 1.0};
    uint32_t bkt[2] = {0, 0};
    bool busy[2] = {false, false}, loaded[2] = {false, false};
    const double c_lo = k1_jump_consts[0], c_hi = k1_jump_consts[1];
    const double two52 = k1_pin(JUMP_TWO52), two52m1 = k1_pin(JUMP_TWO52 - 1.0), one = k1_pin(1.0);
    for (;;) {
#pragma unroll
        for (int c = 0; c < 2; c++) {
            if (!busy[c]) {
                if (loaded[c]) {                                         // its walk ended in the last batch
                    atomicAdd(&hist[bkt[c]], 1u);
                    loaded[c] = false;
                }
            }
            uint32_t mine = g;
            if (DYN) {
                const uint32_t need = __ballot_sync(0xffffffffu, !busy[c]);
                mine = g + __popc(need & lane_lt);
                g = min(g + (uint32_t)__popc(need), total);
            }
            if (!busy[c] && mine < total) {
                key[c] = load(mine);
                if (!DYN) g += stride;
                bkt[c] = 0;                                              // first step of jump.Hash: b = 0
                jd1[c] = one;
                busy[c] = loaded[c] = true;
            }
        }
        if (!__any_sync(0xffffffffu, busy[0] | busy[1])) break;
#pragma unroll
        for (int it = 0; it < BATCH; it++) {
            double tlo[2];
            bool fin[2], amb[2];
#pragma unroll
            for (int c = 0; c < 2; c++) {                                // evaluate: no side effects, the two walks interleave
                key[c] = key[c] * 2862933555777941757ull + 1ull;
                const uint32_t q = (uint32_t)(key[c] >> 33) + 1u;                    // 1 .. 2^31
                const double qd = dbl_make(0x43300000u, q) - two52;                  // (double)q, exact
                const double Q = dbl_make(dbl_hi(qd) - (31u << 20), dbl_lo(qd));     // q * 2^-31, exact
                const double r0 = rcp_seed(Q);
                const double e = fma(-Q, r0, one);
                const double R = fma(r0, e, r0);                                     // ~ 2^31 / q
                const double x = jd1[c] * R;
                tlo[c] = __fma_rd(x, c_lo, two52);                                   // 2^52 + floor(x (1 - EPS))
                const double thi = __fma_rd(x, c_hi, two52);
                fin[c] = (dbl_hi(thi) != 0x43300000u) || (dbl_lo(tlo[c]) >= nb);
                amb[c] = !fin[c] && (dbl_lo(tlo[c]) != dbl_lo(thi));
            }
            if ((amb[0] && busy[0]) || (amb[1] && busy[1])) {            // ~2^-21 per step: the true division
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    if (amb[c] && busy[c]) {
                        uint32_t b = bkt[c];
                        double j1 = jd1[c];
                        fin[c] = jump_step_exact(key[c], b, j1, nb) != 0;
                        tlo[c] = JUMP_TWO52 + (double)b;                 // as the fast step reports it
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < 2; c++) {                                // commit
                const bool adv = busy[c] && !fin[c];
                bkt[c] = adv ? dbl_lo(tlo[c]) : bkt[c];
                jd1[c] = adv ? tlo[c] - two52m1 : jd1[c];                // (double)(bucket + 1), exact
                busy[c] = adv;
            }
        }
    }
}

// ---- the same walks with the fixed-point step (hd_math.h jump_step_fx), for num_buckets <= 2^20 ----------
// State per walk: the LCG key and the 32-bit bucket.  Per step: LCG (IMAD.WIDE + 2 IMAD), (double)q and
// (double)(b + 1) by the 2^52 trick (one DADD each, the constant high words stay in their registers), the
// reciprocal seed (MUFU.RCP64H), residual, product and folded Newton step (DFMA, DMUL, DFMA), y = 2^32 + x
// (DFMA), "finished" as one DSETP against 2^32 + n, the ambiguity band as one IMAD + ISETP on the fraction bits,
// floor(x) by one funnel shift.  About 20 instructions against 33 for the bracketed step, 7 of them on the FP64 pipe.
template <int BATCH, bool SMEM, class Load>
__device__ __forceinline__ void k1_jump_walk_fx(Load load, uint32_t *const next_key, uint32_t g, const uint32_t total,
                                                uint32_t *const hist, const uint32_t nb) {
    const uint32_t lane_lt = (1u << (threadIdx.x & 31)) - 1u;
    uint64_t key[2] = {0, 0}, spare[2] = {0, 0};
    constexpr uint32_t NO_BIN = 0xFFFFFFFFu;                             // the walk holds no key (nothing left to count)
    uint32_t bkt[2] = {NO_BIN, NO_BIN};
    // busy: the walk is under way (when it ends, bkt keeps its bin until the next refill point counts it); have: a
    // spare key is waiting (fetched one refill point ahead of its use, so no walk starts on a load still in flight)
    bool busy[2] = {false, false}, have[2] = {false, false};
    // loop-invariant operands come from the constant bank (an FP64 instruction takes one c[][] operand directly), so none
    // of them is re-materialised with moves inside the loop: [0] 2^52 - 1, [1] 2^83 - 2^31, [2] 1
    const double two52m1 = k1_jump_fx_consts[0], two83m = k1_jump_fx_consts[1], one = k1_jump_fx_consts[2];
    const double ynb = k1_pin(JUMP_TWO32 + (double)nb);
    // hand-out of the warp's next unclaimed keys to the empty spares.  SMEM: a counter in shared memory that only the
    // lanes in need touch (an atomic add each); otherwise ballot + popc over the warp, no memory at all
    auto top_up = [&](int c) {
        uint32_t mine;
        if (SMEM) {
            if (have[c]) return;
            mine = atomicAdd(next_key, 1u);
        } else {
            const uint32_t need = __ballot_sync(0xffffffffu, !have[c]);
            mine = g + __popc(need & lane_lt);
            g = min(g + (uint32_t)__popc(need), total);
            if (have[c]) return;
        }
        if (mine < total) {
            spare[c] = load(mine);
            have[c] = true;
        }
    };
    top_up(0);
    top_up(1);
    for (;;) {
#pragma unroll
        for (int c = 0; c < 2; c++) {
            if (!busy[c]) {
                if (bkt[c] != NO_BIN) atomicAdd(&hist[bkt[c]], 1u);      // its walk ended in the last batch
                key[c] = spare[c];
                bkt[c] = have[c] ? 0u : NO_BIN;                          // first step of jump.Hash: b = 0
                busy[c] = have[c];
                have[c] = false;
            }
            top_up(c);
        }
        if (!__any_sync(0xffffffffu, busy[0] | busy[1] | have[0] | have[1])) break;
#pragma unroll
        for (int it = 0; it < BATCH; it++) {
            uint32_t nbk[2];
            bool fin[2], amb[2];
#pragma unroll
            for (int c = 0; c < 2; c++) {                                // evaluate: no side effects, the two walks interleave
                key[c] = key[c] * 2862933555777941757ull + 1ull;
                const double qd = dbl_make(0x43300000u, (uint32_t)(key[c] >> 33)) - two52m1;   // (double)q
                const double jd1 = dbl_make(0x45200000u, bkt[c]) - two83m;                     // (b + 1) 2^31, exact
                const double r0 = rcp_seed(qd);
                const double e = fma(-qd, r0, one);
                const double jr = jd1 * r0;
                const double x = fma(jr, e, jr);                                               // ~ (b + 1) 2^31 / q
                const double y = x + JUMP_TWO32;                                               // 2^32 + x
                amb[c] = (uint32_t)(dbl_lo(y) * 4096u + 3u * 4096u) < 6u * 4096u;              // fraction in -3 .. 2 units of 2^-20
                fin[c] = y >= ynb;
                nbk[c] = __funnelshift_l(dbl_lo(y), dbl_hi(y), 12);                            // floor(x)
            }
            if ((amb[0] && busy[0]) || (amb[1] && busy[1])) {            // ~2^-17 per step: the true division
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    if (amb[c] && busy[c]) {
                        uint32_t b = bkt[c];
                        double j1;
                        fin[c] = jump_step_exact(key[c], b, j1, nb) != 0;
                        nbk[c] = b;
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < 2; c++) {                                // commit
                const bool adv = busy[c] && !fin[c];
                bkt[c] = adv ? nbk[c] : bkt[c];
                busy[c] = adv;
            }
        }
    }
}

// ---- the part of a warp's work behind the scan: exact per-read sets, then jump-hash binning -------
// wl: the warp's list block [list_cap][32] (entry e of lane l at wl[e * 32 + l]); n: this lane's number
// of adjacent-distinct window minima (0 when the read is invalid or was queued for k1_generic).
// KREP: 0 = the lists hold the reference's X; k = they hold the stored form of K1Repr<k> (k1_scan2.h), converted
// back when a value leaves the list block.
template <int KREP>
__device__ __forceinline__ uint64_t k1_list_value(uint64_t stored) {
    if constexpr (KREP != 0) return K1Repr<KREP>::to_x(stored);
    else return stored;
}

template <bool DUMP, bool FP, bool QUEUE, int KREP = 0>
__device__ __forceinline__ void k1_finish_lists(const K1Params &p, uint64_t *wl, uint64_t *my_list, const int lane,
                                                const uint32_t list_cap, uint32_t n, const bool valid,
                                                const bool overflow, const uint64_t r, const uint32_t nb,
                                                unsigned long long &local_minimizers) {
    // ---- exact per-read set: drop values already present earlier in the list (minimizer.go:189-198).
    // Four candidates at a time: every entry of the accepted prefix is loaded once and compared with all
    // four (independent compares, no load waits on a compare), then the four are compared with each
    // other; survivors are appended in order (in place: m_out never passes the read cursor).
    constexpr uint64_t NONE = Sentinel<FP>::value;           // never a list value
    // A value can only come back later in the list when the same k-mer occurs twice in the read with a smaller one
    // in between: rare.  So the lists are first only TESTED -- loads and compares, four independent accumulators,
    // nothing stored -- and the set construction below runs for a warp where the test fires (a false alarm from
    // stale entries behind a short list only costs that pass).
    // The test looks at the LOW WORDS only (one 32-bit load and one integer compare-and-accumulate per pair instead of a
    // 64-bit load and a 64-bit compare): two different values share a low word once in 2^24 pairs (the span byte is
    // the same for nearly all of them), which again only costs the set pass.
    bool maybe_dup = false;
    {
        const uint32_t n_max = __reduce_max_sync(0xffffffffu, n);
        const uint32_t *const lo = reinterpret_cast<const uint32_t *>(my_list);       // entry e: lo[e * 64] (little endian)
        for (uint32_t a = 0; a < n_max; a += 4) {
            uint32_t x[4];
#pragma unroll
            for (int u = 0; u < 4; u++) x[u] = (a + u < n) ? lo[(a + u) * 64] : 0xFFFFFFFFu - (uint32_t)u;   // distinct pads
            for (uint32_t b = 0; b < a; b += 4) {                        // a is a multiple of four
                uint32_t y[4];
#pragma unroll
                for (int v = 0; v < 4; v++) y[v] = lo[(b + v) * 64];
#pragma unroll
                for (int v = 0; v < 4; v++)
#pragma unroll
                    for (int u = 0; u < 4; u++) maybe_dup |= y[v] == x[u];
            }
            maybe_dup |= (x[1] == x[0]) | (x[2] == x[0]) | (x[2] == x[1]) | (x[3] == x[0]) | (x[3] == x[1]) | (x[3] == x[2]);
        }
    }
    const bool exact = __any_sync(0xffffffffu, maybe_dup);
    uint32_t m_out = exact ? 0u : n;
    for (uint32_t a = 0; exact && a < n; a += 4) {
        uint64_t x[4];
        bool dup[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const bool have = a + u < n;
            x[u] = have ? my_list[(a + u) * 32] : NONE;
            dup[u] = !have;
        }
        uint32_t b = 0;
        for (; b + 2 <= m_out; b += 2) {
            const uint64_t y0 = my_list[b * 32], y1 = my_list[(b + 1) * 32];
#pragma unroll
            for (int u = 0; u < 4; u++) dup[u] |= ueq64<FP>(y0, x[u]) | ueq64<FP>(y1, x[u]);
        }
        if (b < m_out) {
            const uint64_t y0 = my_list[b * 32];
#pragma unroll
            for (int u = 0; u < 4; u++) dup[u] |= ueq64<FP>(y0, x[u]);
        }
        dup[1] |= ueq64<FP>(x[1], x[0]);
        dup[2] |= ueq64<FP>(x[2], x[0]) | ueq64<FP>(x[2], x[1]);
        dup[3] |= ueq64<FP>(x[3], x[0]) | ueq64<FP>(x[3], x[1]) | ueq64<FP>(x[3], x[2]);
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (!dup[u]) {
                if (m_out != a + u) my_list[m_out * 32] = x[u];
                m_out++;
            }
        }
    }
    if (DUMP) {
        if (valid) {
            for (uint32_t e = 0; e < m_out && e < p.dump_cap; e++)
                p.dump[r * p.dump_cap + e] = k1_list_value<KREP>(my_list[e * 32]);
            p.dump_counts[r] = m_out;
        } else if (r < p.n_reads && !overflow) {
            p.dump_counts[r] = 0;
        }
    } else {
        // ---- jump: kmerspectrum.go:67-81, bins[jump.Hash(kmer, numBins)]++ for every set member.
        // Rows 0 .. m_min-1 of the warp's list block are full; the entries of the longer sets
        // behind them are packed row by row right after, so the block becomes one dense queue of
        // `total` keys and lane l simply walks keys l, l + 32, l + 64, ...: every lane gets the
        // same number of keys whatever its own read produced.  Two walks per lane are in flight
        // (ILP); they advance K1_JUMP_BATCH steps between two refill points.  The step itself is
        // hd_math.h jump_step_fast, split into evaluate / commit.
        local_minimizers += m_out;
        const uint32_t m_min = __reduce_min_sync(0xffffffffu, m_out);
        const uint32_t m_max = __reduce_max_sync(0xffffffffu, m_out);
        uint32_t total = m_min * 32;
        for (uint32_t e = m_min; e < m_max; e++) {
            const bool has = m_out > e;
            const uint64_t x = has ? wl[e * 32 + lane] : 0ull;
            const uint32_t mask = __ballot_sync(0xffffffffu, has);
            __syncwarp();                                    // row e is in registers before anyone overwrites it
            if (has) wl[total + __popc(mask & ((1u << lane) - 1u))] = x;     // total <= e * 32: never a later row
            total += __popc(mask);
        }
        __syncwarp();
        if (QUEUE) {
            // hand the warp's dense key block to the batch queue (one reservation per warp, coalesced copy);
            // k1_jump_queue bins the whole batch afterwards without any shared memory
            unsigned long long base = 0;
            if (lane == 0 && total) base = atomicAdd(p.queue_cursor, (unsigned long long)total);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base + total > p.queue_cap) {
                if (lane == 0) k1_report(p, r, K1_ERR_OVF);
            } else {
                for (uint32_t j = lane; j < total; j += 32) p.queue[base + j] = k1_list_value<KREP>(wl[j]);
            }
        } else {
            k1_jump_walk<false, K1_JUMP_BATCH>([&](uint32_t g) { return k1_list_value<KREP>(wl[g]); }, (uint32_t)lane, 32u,
                                               total, p.hist, nb);
        }
    }
}

template <bool DUMP, bool FP, bool QUEUE>
__global__ void __launch_bounds__(K1_TPB) k1_minimizer_histogram(const K1Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *tile = smem;                                                          // tile_cap + 16 bytes
    uint64_t *vh_all = reinterpret_cast<uint64_t *>(smem + (size_t)p.tile_cap + 16);   // [w + 1][K1_TPB]
    uint64_t *list_all = vh_all + (size_t)(p.w + 1) * K1_TPB;                      // [warps][list_cap][32]
    uint64_t *bar = list_all + (size_t)p.list_cap * K1_TPB;                        // mbarrier
    uint64_t *tile_src = bar + 1;                  // global address the tile was staged from (0 = not staged)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t ntiles = (p.n_reads + K1_TPB - 1) / K1_TPB;
    if (blockIdx.x >= ntiles) return;

    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();

    uint32_t phase = 0;
    uint64_t *my_vh = vh_all + tid;
    const uint32_t list_cap = p.list_cap;
    uint64_t *wl = list_all + (size_t)warp * list_cap * 32;    // this warp's lists: entry e of lane l at wl[e * 32 + l]
    uint64_t *my_list = wl + lane;                             // this lane's column (stride 32)
    unsigned long long local_minimizers = 0;
    const uint32_t nb = (uint32_t)p.D;

    for (uint64_t tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
        if (tid == 0) {
            // start the bulk copy of the tile's byte range (or mark it "read straight from global")
            const uint64_t r0 = tl * K1_TPB;
            const uint64_t r1 = (r0 + K1_TPB < p.n_reads) ? r0 + K1_TPB : p.n_reads;
            const uint64_t b0 = k1_read_off(p, r0), b1 = k1_read_off(p, r1);
            const uintptr_t lo = reinterpret_cast<uintptr_t>(p.bases), hi = lo + p.bases_bytes;
            const uintptr_t a0 = (lo + b0) & ~(uintptr_t)15, a1 = (lo + b1 + 15) & ~(uintptr_t)15;
            const bool ok = (a1 > a0) && (a1 - a0 <= p.tile_cap) && a0 >= lo && a1 <= hi;
            if (ok) {
                *tile_src = (uint64_t)a0;
                mbar_arrive_expect_tx(bar, (uint32_t)(a1 - a0));
                bulk_g2s(tile, reinterpret_cast<const void *>(a0), (uint32_t)(a1 - a0), bar);
            } else {
                *tile_src = 0;
                mbar_arrive(bar);
            }
        }
        mbar_wait(bar, phase);
        phase ^= 1;

        // ---- scan ----------------------------------------------------------------------------
        const uint64_t r = tl * K1_TPB + tid;
        uint32_t n = 0;            // window minima that differ from their predecessor (all of them, even past list_cap)
        bool valid = false;
        if (r < p.n_reads) {
            const uint64_t b0 = k1_read_off(p, r), b1 = k1_read_off(p, r + 1);
            const uint64_t len64 = b1 - b0;
            if (len64 < 1) {
                k1_report(p, r, K1_ERR_EMPTY);                                   // minimizer.go:71-73
            } else if (len64 < (uint64_t)(p.w + p.k - 1)) {
                k1_report(p, r, K1_ERR_SHORT);                                   // minimizer.go:74-76
            } else if (len64 >= p.long_min) {
                valid = true;                                                    // k1_long.cuh takes it: queued below
                n = list_cap + 1u;
            } else {
                valid = true;
                const uint64_t src = *tile_src;
                uint64_t last = Sentinel<FP>::value;                             // never a minimizer
                auto emit = [&](uint64_t m, bool on) {
                    const bool fresh = on && !ueq64<FP>(m, last);
                    if (fresh) my_list[min(n, list_cap - 1u) * 32u] = m;         // entries past the cap are dropped,
                    n += fresh ? 1u : 0u;                                        // n still counts them (overflow test)
                    last = fresh ? m : last;
                };
                if (src) {
                    const uint32_t off = (uint32_t)(reinterpret_cast<uintptr_t>(p.bases) + b0 - src);
                    const SmemWordSrc ws{reinterpret_cast<const uint32_t *>(tile + (off & ~3u)), (off & 3u) * 8u};
                    k1_scan_read<FP>(ws, (int32_t)len64, (int32_t)p.k, (int32_t)p.w, K1SmemVH{my_vh}, emit);
                } else {
                    const ByteSrc bs{p.bases + b0, (int32_t)len64};
                    k1_scan_read<FP>(bs, (int32_t)len64, (int32_t)p.k, (int32_t)p.w, K1SmemVH{my_vh}, emit);
                }
            }
        }
        const bool overflow = valid && n > list_cap;
        if (overflow) {
            // hand the read to the generic kernel (exact de-dup with an unbounded set)
            const unsigned int slot = atomicAdd(p.ovf_count, 1u);
            if (slot < p.ovf_cap) p.ovf_list[slot] = ((unsigned long long)n << 32) | r;   // (r < 2^32 per launch)
            else k1_report(p, r, K1_ERR_OVF);
            n = 0;
            valid = false;
        }
        k1_finish_lists<DUMP, FP, QUEUE>(p, wl, my_list, lane, list_cap, n, valid, overflow, r, nb, local_minimizers);
        __syncthreads();   // everyone is done with the tile before thread 0 refills it
    }
    if (!DUMP) {
        for (int o = 16; o > 0; o >>= 1) local_minimizers += __shfl_down_sync(0xffffffffu, local_minimizers, o);
        if (lane == 0 && local_minimizers) atomicAdd(p.n_minimizers, local_minimizers);
    }
}

// ------------------------------------------------------------------------------------------
// w = 9 path (the reference's default window, cmd/sketch.go:52): k1_scan_read_w9 keeps the whole
// window state in registers, so shared memory only holds the candidate lists and five CTAs
// (20 warps) fit one SM instead of three.  No tile is staged: every lane streams its own read
// from global memory with aligned 8-byte loads issued one block (8 bases, ~600 instructions)
// ahead of their use; the batch is read exactly once from HBM, sectors are shared through L2.
// Warps are independent (no CTA barrier anywhere): warp g takes reads [32 g', 32 g' + 32) for
// g' = g, g + #warps, ...
// ------------------------------------------------------------------------------------------
constexpr int K1_W9_CTAS_PER_SM = 5;

struct GlobalSrc8 {
    const uint64_t *p;      // aligned unit that holds the next unread base
    uint64_t cur, nxt;
    uint32_t sh;            // (byte offset of the read inside its first unit) * 8
    uintptr_t lim;          // no 8-byte load may touch this address or beyond
    uintptr_t end;          // end of this read: bytes at or past it are never interpreted
    __device__ __forceinline__ uint64_t ld(const uint64_t *q) const {
        const uintptr_t a = reinterpret_cast<uintptr_t>(q);
        if (a + 8 <= lim) return __ldg(reinterpret_cast<const unsigned long long *>(q));
        uint64_t v = 0;                                    // the last few bytes of the batch
        for (int j = 0; j < 8; j++)
            if (a + j < end) v |= (uint64_t)reinterpret_cast<const uint8_t *>(q)[j] << (8 * j);
        return v;
    }
    __device__ __forceinline__ GlobalSrc8(const uint8_t *read, uint64_t len, uintptr_t lim_)
        : lim(lim_), end(reinterpret_cast<uintptr_t>(read) + len) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(read);
        p = reinterpret_cast<const uint64_t *>(a & ~(uintptr_t)7);
        sh = (uint32_t)(a & 7) * 8u;
        cur = ld(p);
        nxt = ld(p + 1);
    }
    __device__ __forceinline__ uint64_t next8() {
        const uint64_t r = (cur >> sh) | ((nxt << 1) << (63u - sh));
        cur = nxt;
        p++;
        nxt = ld(p + 1);
        return r;
    }
};

template <bool DUMP, bool FP, bool QUEUE, int KC = 0>
__global__ void __launch_bounds__(K1_TPB, K1_W9_CTAS_PER_SM) k1_minimizer_histogram_w9(const K1Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *list_all = reinterpret_cast<uint64_t *>(smem);                       // [warps][list_cap][32]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t list_cap = p.list_cap;
    uint64_t *wl = list_all + (size_t)warp * list_cap * 32;
    uint64_t *my_list = wl + lane;
    unsigned long long local_minimizers = 0;
    const uint32_t nb = (uint32_t)p.D;
    const uint64_t ntasks = (p.n_reads + 31) / 32;
    const uintptr_t lim = reinterpret_cast<uintptr_t>(p.bases) + p.bases_bytes;

    // tasks (32 reads) are handed out dynamically when a counter exists: the grid may be smaller than the
    // task count without a straggler round, so it can be sized to leave SM room for the flush chain
    unsigned long long *const task_counter = QUEUE ? p.queue_cursor + 1 : nullptr;
    for (uint64_t task = (uint64_t)blockIdx.x * K1_WARPS + warp;;) {
        if (task_counter) {
            unsigned long long t = 0;
            if (lane == 0) t = atomicAdd(task_counter, 1ull);
            task = __shfl_sync(0xffffffffu, t, 0);
        }
        if (task >= ntasks) break;
        const uint64_t r = task * 32 + lane;
        uint32_t n = 0;
        bool valid = false;
        if (r < p.n_reads) {
            const uint64_t b0 = k1_read_off(p, r), b1 = k1_read_off(p, r + 1);
            const uint64_t len64 = b1 - b0;
            if (len64 < 1) {
                k1_report(p, r, K1_ERR_EMPTY);                                   // minimizer.go:71-73
            } else if (len64 < (uint64_t)(9 + p.k - 1)) {
                k1_report(p, r, K1_ERR_SHORT);                                   // minimizer.go:74-76
            } else if (len64 >= p.long_min) {
                valid = true;                                                    // k1_long.cuh takes it: queued below
                n = list_cap + 1u;
            } else {
                valid = true;
                uint64_t last = Sentinel<FP>::value;                             // never a minimizer
                auto emit = [&](uint64_t m, bool on) {
                    const bool fresh = on && !ueq64<FP>(m, last);
                    if (fresh) my_list[min(n, list_cap - 1u) * 32u] = m;         // entries past the cap are dropped,
                    n += fresh ? 1u : 0u;                                        // n still counts them (overflow test)
                    last = fresh ? m : last;
                };
                k1_scan_read_w9<FP, KC>(GlobalSrc8(p.bases + b0, len64, lim), (int32_t)len64, (int32_t)p.k, emit);
            }
        }
        const bool overflow = valid && n > list_cap;
        if (overflow) {
            const unsigned int slot = atomicAdd(p.ovf_count, 1u);
            if (slot < p.ovf_cap) p.ovf_list[slot] = ((unsigned long long)n << 32) | r;   // (r < 2^32 per launch)
            else k1_report(p, r, K1_ERR_OVF);
            n = 0;
            valid = false;
        }
        __syncwarp();
        k1_finish_lists<DUMP, FP, QUEUE>(p, wl, my_list, lane, list_cap, n, valid, overflow, r, nb, local_minimizers);
        __syncwarp();                                                            // the queue is drained before the lists refill
        if (!task_counter) task += (uint64_t)gridDim.x * K1_WARPS;
    }
    if (!DUMP) {
        for (int o = 16; o > 0; o >>= 1) local_minimizers += __shfl_down_sync(0xffffffffu, local_minimizers, o);
        if (lane == 0 && local_minimizers) atomicAdd(p.n_minimizers, local_minimizers);
    }
}

// ------------------------------------------------------------------------------------------
// Second-generation w = 9 scan (k1_scan2.h), k folded in at compile time.  Same work decomposition as
// k1_minimizer_histogram_w9 (one lane per read, independent warps, dynamic 32-read tasks, lists in shared
// memory); the per-position work is about 45 instructions instead of 77 and spread over the ALU, FMA and FP64
// pipes.
// ------------------------------------------------------------------------------------------
// Bases of one read, eight at a time, from aligned 8-byte global loads issued one block ahead.  The byte
// offset inside a word is undone by one PRMT per word (selector fixed per read), the word offset inside the
// 8-byte unit by three selects on a per-read predicate.
struct GlobalSrcW {
    const uint2 *p;          // aligned unit that holds the next unread base
    uint2 cur, nxt;
    uint32_t sel;            // 0x3210 + 0x1111 * (address & 3)
    bool hiw;                // the read starts in the high word of its first unit
    uintptr_t lim;           // no 8-byte load may touch this address or beyond
    uintptr_t end;           // end of this read: bytes at or past it are never interpreted
    __device__ __forceinline__ uint2 ld(const uint2 *q) const {
        const uintptr_t a = reinterpret_cast<uintptr_t>(q);
        if (a + 8 <= lim) return __ldg(q);
        uint2 v = make_uint2(0u, 0u);                      // the last few bytes of the batch
        for (int j = 0; j < 8; j++)
            if (a + j < end) (j < 4 ? v.x : v.y) |= (uint32_t) reinterpret_cast<const uint8_t *>(q)[j] << (8 * (j & 3));
        return v;
    }
    __device__ __forceinline__ GlobalSrcW(const uint8_t *read, uint64_t len, uintptr_t lim_)
        : lim(lim_), end(reinterpret_cast<uintptr_t>(read) + len) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(read);
        p = reinterpret_cast<const uint2 *>(a & ~(uintptr_t)7);
        sel = 0x3210u + 0x1111u * (uint32_t)(a & 3);
        hiw = (a & 4) != 0;
        cur = ld(p);
        nxt = ld(p + 1);
    }
    __device__ __forceinline__ void next(uint32_t &w0, uint32_t &w1) {
        const uint32_t t0 = hiw ? cur.y : cur.x, t1 = hiw ? nxt.x : cur.y, t2 = hiw ? nxt.y : nxt.x;
        w0 = __byte_perm(t0, t1, sel);
        w1 = __byte_perm(t1, t2, sel);
        cur = nxt;
        p++;
        nxt = ld(p + 1);
    }
};

template <bool DUMP, bool QUEUE, int K>
__global__ void __launch_bounds__(K1_TPB, K1_W9_CTAS_PER_SM) k1_scan_w9_v2(const K1Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *list_all = reinterpret_cast<uint64_t *>(smem);                       // [warps][list_cap][32]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t list_cap = p.list_cap;
    uint64_t *wl = list_all + (size_t)warp * list_cap * 32;
    uint64_t *my_list = wl + lane;
    unsigned long long local_minimizers = 0;
    const uint32_t nb = (uint32_t)p.D;
    const uint64_t ntasks = (p.n_reads + 31) / 32;
    const uintptr_t lim = reinterpret_cast<uintptr_t>(p.bases) + p.bases_bytes;

    // a grid with a warp for every task runs each task once (the host's default); a smaller grid takes them from a
    // counter (persistent CTAs)
    unsigned long long *const task_counter =
        (QUEUE && (uint64_t)gridDim.x * K1_WARPS < ntasks) ? p.queue_cursor + 1 : nullptr;
    for (uint64_t task = (uint64_t)blockIdx.x * K1_WARPS + warp;;) {
        if (task_counter) {
            unsigned long long t = 0;
            if (lane == 0) t = atomicAdd(task_counter, 1ull);
            task = __shfl_sync(0xffffffffu, t, 0);
        }
        if (task >= ntasks) break;
        const uint64_t r = task * 32 + lane;
        uint32_t n = 0;
        bool valid = false;
        if (r < p.n_reads) {
            const uint64_t b0 = k1_read_off(p, r), b1 = k1_read_off(p, r + 1);
            const uint64_t len64 = b1 - b0;
            if (len64 < 1) {
                k1_report(p, r, K1_ERR_EMPTY);                                   // minimizer.go:71-73
            } else if (len64 < (uint64_t)(9 + K - 1)) {
                k1_report(p, r, K1_ERR_SHORT);                                   // minimizer.go:74-76
            } else if (len64 >= p.long_min) {
                valid = true;                                                    // k1_long.cuh takes it: queued below
                n = list_cap + 1u;
            } else {
                valid = true;
                K1List<32> L{my_list, list_cap, 0u, 0.0};
                k1_scan_read_w9_v2<K, 32>(GlobalSrcW(p.bases + b0, len64, lim), (int32_t)len64, L);
                n = L.n;
            }
        }
        const bool overflow = valid && n > list_cap;
        if (overflow) {
            const unsigned int slot = atomicAdd(p.ovf_count, 1u);
            if (slot < p.ovf_cap) p.ovf_list[slot] = ((unsigned long long)n << 32) | r;   // (r < 2^32 per launch)
            else k1_report(p, r, K1_ERR_OVF);
            n = 0;
            valid = false;
        }
        __syncwarp();
        k1_finish_lists<DUMP, true, QUEUE, K>(p, wl, my_list, lane, list_cap, n, valid, overflow, r, nb, local_minimizers);
        __syncwarp();                                                            // the queue is drained before the lists refill
        if (!task_counter) task += (uint64_t)gridDim.x * K1_WARPS;
    }
    if (!DUMP) {
        for (int o = 16; o > 0; o >>= 1) local_minimizers += __shfl_down_sync(0xffffffffu, local_minimizers, o);
        if (lane == 0 && local_minimizers) atomicAdd(p.n_minimizers, local_minimizers);
    }
}

// ------------------------------------------------------------------------------------------
// Second half of stage 1+2 when the scan kernels run in QUEUE mode: bin every key of the batch queue.
// No shared memory and few registers, so it runs at full occupancy and shares an SM with the CWS
// filter's TMA ring or the next interval's scan CTAs; the queue is dense, so the walks are balanced
// across the whole grid whatever the per-read set sizes were.
// ------------------------------------------------------------------------------------------
constexpr int K1_JUMP_TPB = 256;
constexpr int K1_JUMP_CTAS_PER_SM = 3;      // 24 warps per SM, one wave: every warp walks one contiguous segment
template <int BATCH>
__global__ void __launch_bounds__(K1_JUMP_TPB) k1_jump_queue(const K1Params p) {
    const unsigned long long filled = *p.queue_cursor;
    const uint64_t total = filled < p.queue_cap ? filled : p.queue_cap;
    const uint64_t nwarps = (uint64_t)gridDim.x * (K1_JUMP_TPB / 32);
    const uint64_t gw = (uint64_t)blockIdx.x * (K1_JUMP_TPB / 32) + (threadIdx.x >> 5);
    const uint32_t seg_begin = (uint32_t)(total * gw / nwarps), seg_end = (uint32_t)(total * (gw + 1) / nwarps);
    const uint64_t *const q = p.queue;
    k1_jump_walk<true, BATCH>([&](uint32_t i) { return q[i]; }, seg_begin, 0u, seg_end, p.hist, (uint32_t)p.D);
}
// the same kernel with the fixed-point step (num_buckets <= 2^20)
template <int BATCH, bool SMEM>
__global__ void __launch_bounds__(K1_JUMP_TPB) k1_jump_queue_fx(const K1Params p) {
    const unsigned long long filled = *p.queue_cursor;
    const uint64_t total = filled < p.queue_cap ? filled : p.queue_cap;
    const uint64_t nwarps = (uint64_t)gridDim.x * (K1_JUMP_TPB / 32);
    const uint64_t gw = (uint64_t)blockIdx.x * (K1_JUMP_TPB / 32) + (threadIdx.x >> 5);
    const uint32_t seg_begin = (uint32_t)(total * gw / nwarps), seg_end = (uint32_t)(total * (gw + 1) / nwarps);
    const uint64_t *const q = p.queue;
    __shared__ uint32_t next_key[K1_JUMP_TPB / 32];                      // per warp: the next unclaimed key of its segment
    if (SMEM) {
        if ((threadIdx.x & 31) == 0) next_key[threadIdx.x >> 5] = seg_begin;
        __syncwarp();
    }
    k1_jump_walk_fx<BATCH, SMEM>([&](uint32_t i) { return q[i]; }, &next_key[threadIdx.x >> 5], seg_begin, seg_end, p.hist,
                                 (uint32_t)p.D);
}

// reciprocal self-test (parity tap): for q = q0 + i the seed's and the refined reciprocal's relative errors,
// as residuals 1 - q r evaluated with one FMA; out[0] = max |1 - q r0|, out[1] = max |1 - q R| over the grid's range
__global__ void k_rcp_selftest(uint32_t q0, uint32_t n, unsigned long long *out) {
    double m0 = 0.0, m1 = 0.0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const double qd = (double)(q0 + (uint32_t)i);
        const double r0 = rcp_seed(qd);
        const double e = fma(-qd, r0, 1.0);
        const double R = fma(r0, e, r0);
        const double e2 = fma(-qd, R, 1.0);
        m0 = fmax(m0, fabs(e));
        m1 = fmax(m1, fabs(e2));
    }
    atomicMax(&out[0], (unsigned long long)__double_as_longlong(m0));    // non-negative doubles order like integers
    atomicMax(&out[1], (unsigned long long)__double_as_longlong(m1));
}
// fixed-point jump hash tap (parity tests): out[i] = jump_hash_fx(keys[i]), n_amb += ambiguous steps
__global__ void k_jump_fx_tap(const uint64_t *keys, uint64_t n, int32_t buckets, int32_t *out, unsigned long long *n_amb) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t amb = 0;
    out[i] = jump_hash_fx(keys[i], buckets, &amb);
    if (amb) atomicAdd(n_amb, (unsigned long long)amb);
}

// ------------------------------------------------------------------------------------------
// Generic path: any w <= 256, any read length, exact set via an open-addressing table carved
// from a global arena.  One thread per queued read (or per read of the batch when use_queue
// is false, i.e. w > K1_W_FAST).  Slow by design; it only sees reads the fast path queued.
// ------------------------------------------------------------------------------------------
struct K1LocalVH {
    uint64_t *buf;
    __device__ __forceinline__ uint64_t &operator()(int t) const { return buf[t]; }
};

template <bool DUMP>
__global__ void __launch_bounds__(64) k1_generic(const K1Params p, const bool use_queue) {
    uint64_t vhbuf[257];
    const uint64_t total = use_queue ? (uint64_t)min(*p.ovf_count, p.ovf_cap) : p.n_reads;
    const uint64_t gt = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long local_minimizers = 0;
    for (uint64_t q = gt; q < total; q += (uint64_t)gridDim.x * blockDim.x) {
        // a queued read comes with the number of values its scan produced (an upper bound of its set's size)
        const unsigned long long entry = use_queue ? p.ovf_list[q] : 0ull;
        const uint64_t r = use_queue ? (entry & 0xffffffffull) : q;
        const uint64_t n_hint = entry >> 32;
        const uint64_t b0 = k1_read_off(p, r), b1 = k1_read_off(p, r + 1);
        const uint64_t len64 = b1 - b0;
        if (DUMP) p.dump_counts[r] = 0;
        if (len64 >= p.long_min) continue;                                       // k1_long.cuh takes it
        if (len64 < 1) { k1_report(p, r, K1_ERR_EMPTY); continue; }
        if (len64 < (uint64_t)(p.w + p.k - 1)) { k1_report(p, r, K1_ERR_SHORT); continue; }
        // table capacity: power of two >= 2 * (values that can be inserted: at most one per k-mer).  Tables of up to
        // slab_size entries live in this thread's own slab, read after read; larger ones come from the shared cursor
        const uint64_t nk = len64 - p.k + 1;
        const uint64_t n_max = (n_hint && n_hint < nk) ? n_hint : nk;
        uint64_t cap = 64;
        while (cap < 2 * n_max && cap < (1ull << 62)) cap <<= 1;
        uint64_t *tab;
        if (cap <= p.slab_size && (gt + 1) * p.slab_size <= p.slab_entries) {
            tab = p.arena + gt * p.slab_size;
        } else {
            const unsigned long long at = p.slab_entries + atomicAdd(p.arena_cursor, (unsigned long long)cap);
            if (at + cap > p.arena_entries) { k1_report(p, r, K1_ERR_OVF); continue; }
            tab = p.arena + at;
        }
        for (uint64_t x = 0; x < cap; x++) tab[x] = 0;        // 0 = empty; the value 0 itself is tracked apart
        bool seen_zero = false;
        uint32_t n_set = 0;
        uint64_t last = 0;
        bool have_last = false;
        k1_scan_read<false>(ByteSrc{p.bases + b0, (int32_t)len64}, (int32_t)len64, (int32_t)p.k, (int32_t)p.w, K1LocalVH{vhbuf}, [&](uint64_t m, bool on) {
            if (!on) return;
            if (have_last && m == last) return;
            last = m;
            have_last = true;
            bool is_new;
            if (m == 0) {
                is_new = !seen_zero;
                seen_zero = true;
            } else {
                uint64_t h = (m * 0x9E3779B97F4A7C15ull) >> 17;
                for (;;) {
                    h &= (cap - 1);
                    const uint64_t cur = tab[h];
                    if (cur == 0) { tab[h] = m; is_new = true; break; }
                    if (cur == m) { is_new = false; break; }
                    h++;
                }
            }
            if (!is_new) return;
            if (DUMP) {
                if (n_set < p.dump_cap) p.dump[r * p.dump_cap + n_set] = m;
            } else {
                atomicAdd(&p.hist[jump_hash(m, p.D)], 1u);
                if (p.feed_queue) k1_feed_queue(p, r, m);
            }
            n_set++;
        });
        if (DUMP) p.dump_counts[r] = n_set;
        local_minimizers += n_set;
    }
    if (!DUMP && local_minimizers) atomicAdd(p.n_minimizers, local_minimizers);
}

// device-side jump hash tap (parity tests)
__global__ void k_jump_tap(const uint64_t *keys, uint64_t n, int32_t buckets, int32_t *out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = jump_hash(keys[i], buckets);
}

}  // namespace hulk
