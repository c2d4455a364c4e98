// k1_minimizer.cuh -- stage 1+2 of the sketch hot path on sm_100a:
//   reads (ASCII, in HBM) -> per-read SET of window minimizers -> jump-hash bin -> histogram.
//
// Reference semantics (paths relative to the reference checkout):
//   src/minimizer/minimizer.go:96-204  findMinimizers (rolling 2-bit k-mer pair, canonical pick,
//       X = hash64(canon)<<8 | kmerSpan, monotone-deque window minimum, per-read set)
//   src/pipeline/minion.go:51-57 + boss.go:90-95 -> src/kmerspectrum/kmerspectrum.go:67-81 AddHash
//
// What the deque computes, restated for a SIMT machine: position i (i >= k-1, fwd != rev) emits
//   m_i = min{ X_j : i-w < j <= i, j >= k-1, fwd_j != rev_j }   when i >= w-1,
// and the read contributes the SET {m_i}.  (The deque's tie rule only affects the stored
// position, which never leaves the function.)  The window minimum is evaluated with the
// two-pass block decomposition (prefix minima of the current block of w k-mers, suffix minima
// of the previous one) so every lane runs the same instruction stream regardless of data.
//
// Work decomposition: one thread per read, 128 reads per CTA tile; the tile's bytes are one
// contiguous range of the batch and are staged into shared memory with a single 1-D bulk TMA
// copy (the other resident CTAs of the SM cover its latency).  Candidates (window minima that
// differ from the previous one) go to a per-thread shared-memory list; after the scan the list
// is de-duplicated exactly and every distinct minimizer is jump-hashed and counted with a
// global (L2) atomic -- the k^4-bin histogram does not fit shared memory for k > 11.
#pragma once
#include <stdint.h>

#include "hd_math.h"
#include "ptx_util.cuh"

namespace hulk {

constexpr int K1_TPB = 128;       // threads (= reads) per CTA tile
constexpr int K1_W_FAST = 32;     // largest w handled by the shared-memory path

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
    // generic path scratch
    uint64_t *arena;
    unsigned long long *arena_cursor;
    uint64_t arena_entries;
};

constexpr uint32_t K1_ERR_EMPTY = 3;   // HULK_B200_EEMPTYSEQ
constexpr uint32_t K1_ERR_SHORT = 4;   // HULK_B200_ESHORTSEQ
constexpr uint32_t K1_ERR_OVF = 31;    // overflow queue / arena exhausted -> HULK_B200_ENOMEM

__device__ __forceinline__ uint64_t k1_read_off(const K1Params &p, uint64_t r) {
    return p.offsets ? (p.offsets[r] - p.off_base) : r * (uint64_t)p.fixed_len;
}
__device__ __forceinline__ void k1_report(const K1Params &p, uint64_t r, uint32_t code) {
    atomicMin(p.err_word, (unsigned long long)(((p.read_base + r) << 8) | code));
}

// Scan one read.  `vh(t)` is the w-entry block buffer, `emit(m)` receives every window minimum
// (position order).  Returns false if the read fails the reference's length checks.
template <class VH, class Emit>
__device__ __forceinline__ void k1_scan_read(const uint8_t *__restrict__ seq, int32_t len, int32_t k, int32_t w,
                                             VH vh, Emit emit) {
    const uint64_t mask = (1ull << (2 * k)) - 1ull;       // minimizer.go:103  (k <= 31)
    const int shift = 2 * (k - 1);                        // minimizer.go:104
    uint64_t fwd = 0, rev = 0;
    uint64_t pref = ~0ull;                                // prefix minimum of the current block
    int t = 0;                                            // position inside the current block
    for (int x = 0; x < w; x++) vh(x) = ~0ull;
    for (int32_t i = 0; i < len; i++) {
        const uint32_t c = nt4(seq[i]);                                   // minimizer.go:115
        fwd = ((fwd << 2) | (uint64_t)c) & mask;                          // :134
        rev = (rev >> 2) | ((uint64_t)(3u ^ c) << shift);                 // :137 (not masked)
        if (i < k - 1) continue;                                          // :140-142
        const bool skip = (fwd == rev);                                   // :145-147
        uint64_t X = ~0ull;
        if (!skip) {
            const int32_t wi = i - w + 1;                                 // windowIndex :112
            const int32_t span = (wi + 1 < k) ? (wi + 1) : k;             // :127-131
            const uint64_t canon = (fwd > rev) ? rev : fwd;               // :150-153
            X = (hash64(canon, mask) << 8) | (uint64_t)(int64_t)span;     // :156-159
        }
        pref = (X < pref) ? X : pref;
        const uint64_t suf = (t + 1 < w) ? vh(t + 1) : ~0ull;             // previous block, positions t+1..w-1
        vh(t) = X;
        if (!skip && i >= w - 1) emit((pref < suf) ? pref : suf);         // :186-199
        if (++t == w) {                                                   // block complete: suffix minima in place
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
}

struct K1SmemVH {
    uint64_t *base;   // [w][K1_TPB], this thread's column
    __device__ __forceinline__ uint64_t &operator()(int t) const { return base[t * K1_TPB]; }
};

template <bool DUMP>
__global__ void __launch_bounds__(K1_TPB) k1_minimizer_histogram(const K1Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *tile = smem;                                                          // tile_cap bytes
    uint64_t *vh_all = reinterpret_cast<uint64_t *>(smem + (size_t)p.tile_cap);    // [w][K1_TPB]
    uint64_t *list_all = vh_all + (size_t)p.w * K1_TPB;                            // [list_cap][K1_TPB]
    uint64_t *bar = list_all + (size_t)p.list_cap * K1_TPB;                        // mbarrier
    uint64_t *tile_src = bar + 1;                  // global address the tile was staged from (0 = not staged)

    const int tid = threadIdx.x;
    const uint64_t ntiles = (p.n_reads + K1_TPB - 1) / K1_TPB;
    if (blockIdx.x >= ntiles) return;

    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();

    uint32_t phase = 0;
    uint64_t *my_vh = vh_all + tid;
    uint64_t *my_list = list_all + tid;
    const uint32_t list_cap = p.list_cap;
    unsigned long long local_minimizers = 0;

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

        const uint64_t r = tl * K1_TPB + tid;
        uint32_t n = 0;            // entries in my list
        bool overflow = false;
        bool valid = false;
        if (r < p.n_reads) {
            const uint64_t b0 = k1_read_off(p, r), b1 = k1_read_off(p, r + 1);
            const uint64_t len64 = b1 - b0;
            if (len64 < 1) {
                k1_report(p, r, K1_ERR_EMPTY);                                   // minimizer.go:71-73
            } else if (len64 < (uint64_t)(p.w + p.k - 1)) {
                k1_report(p, r, K1_ERR_SHORT);                                   // minimizer.go:74-76
            } else {
                valid = true;
                const uint64_t src = *tile_src;
                const uint8_t *seq = src ? tile + (reinterpret_cast<uintptr_t>(p.bases) + b0 - src)
                                         : p.bases + b0;
                uint64_t last = 0;
                k1_scan_read(seq, (int32_t)len64, (int32_t)p.k, (int32_t)p.w, K1SmemVH{my_vh},
                             [&](uint64_t m) {
                                 if (n == 0 || m != last) {
                                     if (n < list_cap) { my_list[(size_t)n * K1_TPB] = m; n++; }
                                     else overflow = true;
                                     last = m;
                                 }
                             });
            }
        }
        if (valid && overflow) {
            // hand the read to the generic kernel (exact de-dup with an unbounded set)
            const unsigned int slot = atomicAdd(p.ovf_count, 1u);
            if (slot < p.ovf_cap) p.ovf_list[slot] = r;
            else k1_report(p, r, K1_ERR_OVF);
            n = 0;
            valid = false;
        }
        // exact per-read set: drop values already present earlier in the list (minimizer.go:189-198)
        uint32_t m_out = 0;
        for (uint32_t a = 0; a < n; a++) {
            const uint64_t x = my_list[(size_t)a * K1_TPB];
            bool dup = false;
            for (uint32_t b = 0; b < m_out; b++) dup |= (my_list[(size_t)b * K1_TPB] == x);
            if (!dup) { my_list[(size_t)m_out * K1_TPB] = x; m_out++; }
        }
        if (DUMP) {
            if (valid) {
                for (uint32_t e = 0; e < m_out && e < p.dump_cap; e++)
                    p.dump[r * p.dump_cap + e] = my_list[(size_t)e * K1_TPB];
                p.dump_counts[r] = m_out;
            } else if (r < p.n_reads && !overflow) {
                p.dump_counts[r] = 0;
            }
        } else {
            // kmerspectrum.go:67-81: bins[jump.Hash(kmer, numBins)]++ for every set member.
            // Each lane walks its own list; a lane that finishes a key immediately starts the next
            // one, so the warp only idles for the difference in total jump steps between lanes.
            local_minimizers += m_out;
            if (m_out > 0) {
                uint32_t e = 0;
                uint64_t key = my_list[0];
                int64_t b = -1, j = 0;
                const int64_t nb = p.D;
                for (;;) {
                    if (j >= nb) {
                        atomicAdd(&p.hist[(int32_t)b], 1u);
                        if (++e >= m_out) break;
                        key = my_list[(size_t)e * K1_TPB];
                        b = -1;
                        j = 0;
                    }
                    b = j;
                    key = key * 2862933555777941757ull + 1ull;
                    j = (int64_t)((double)(b + 1) * (2147483648.0 / (double)((key >> 33) + 1)));
                }
            }
        }
        __syncthreads();   // everyone is done with the tile before thread 0 refills it
    }
    if (!DUMP) {
        for (int o = 16; o > 0; o >>= 1) local_minimizers += __shfl_down_sync(0xffffffffu, local_minimizers, o);
        if ((tid & 31) == 0 && local_minimizers) atomicAdd(p.n_minimizers, local_minimizers);
    }
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
    uint64_t vhbuf[256];
    const uint64_t total = use_queue ? (uint64_t)min(*p.ovf_count, p.ovf_cap) : p.n_reads;
    unsigned long long local_minimizers = 0;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total;
         q += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = use_queue ? p.ovf_list[q] : q;
        const uint64_t b0 = k1_read_off(p, r), b1 = k1_read_off(p, r + 1);
        const uint64_t len64 = b1 - b0;
        if (DUMP) p.dump_counts[r] = 0;
        if (len64 < 1) { k1_report(p, r, K1_ERR_EMPTY); continue; }
        if (len64 < (uint64_t)(p.w + p.k - 1)) { k1_report(p, r, K1_ERR_SHORT); continue; }
        // table capacity: power of two >= 2 * (number of k-mers)
        const uint64_t nk = len64 - p.k + 1;
        uint64_t cap = 64;
        while (cap < 2 * nk) cap <<= 1;
        const unsigned long long at = atomicAdd(p.arena_cursor, (unsigned long long)cap);
        if (at + cap > p.arena_entries) { k1_report(p, r, K1_ERR_OVF); continue; }
        uint64_t *tab = p.arena + at;
        for (uint64_t x = 0; x < cap; x++) tab[x] = 0;        // 0 = empty; the value 0 itself is tracked apart
        bool seen_zero = false;
        uint32_t n_set = 0;
        uint64_t last = 0;
        bool have_last = false;
        k1_scan_read(p.bases + b0, (int32_t)len64, (int32_t)p.k, (int32_t)p.w, K1LocalVH{vhbuf}, [&](uint64_t m) {
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
