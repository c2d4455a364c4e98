// k1_long.cuh -- stage 1+2 for LONG sequences (`hulk sketch --fasta`: contigs, chromosomes; long reads).
//
// Reference semantics (paths relative to the reference checkout):
//   src/minimizer/minimizer.go:96-204   findMinimizers: one sequential pass and one set per sequence
//   src/pipeline/sketch.go:99-135       FASTA records arrive as one sequence each, whatever their length
//   src/kmerspectrum/kmerspectrum.go:67-81  bins[jump.Hash(minimizer, numBins)]++ for every set member
//
// The read-per-lane kernels of k1_minimizer.cuh give one lane to a sequence, and k1_generic one thread: a 5 Mbp
// genome then takes seconds while the rest of the machine idles.  Here a sequence of `long_min` bases or more is
// cut into slices of `long_seg` positions, one thread per slice (k1_long.h: a slice restarts the rolling k-mers
// and the window k + w positions early and emits exactly what the sequential pass emits for its positions), and
// the slices of a sequence share one open-addressing set in the scratch arena, filled with 64-bit compare-and-swap:
// the thread that inserts a value first bins it.  The histogram is a sum, so neither the order of the slices nor
// the winner of a race shows in the result.
//
// Three launches, only made when the host knows (or cannot rule out) that the batch holds such a sequence:
//   k1_long_plan   one CTA: collects the long sequences the other kernels passed over, gives each its table
//                  (arena bump allocation behind k1_generic's) and its range of slice numbers (two prefix sums);
//   k1_long_zero   clears the tables;
//   k1_long_scan   one thread per slice, grid-stride.
#pragma once
#include <stdint.h>

#include "k1_long.h"
#include "k1_minimizer.cuh"

namespace hulk {

// Sequences this long go to the sliced scan.  Measured (profiles/r02x_long_reads.txt, 100-160 Mbases resident in HBM):
// reads of ~2000 bases 2.8 -> 11.0 Gbases/s, of ~8000 bases 1.7 -> 13.7 Gbases/s against one k1_generic thread per read.
constexpr uint32_t K1_LONG_MIN = 1u << 10;
constexpr int K1_LONG_PLAN_TPB = 1024;
constexpr int K1_LONG_TPB = 128;

// positions per slice: the warm-up (k + w) stays a small part of the work
__host__ __device__ inline uint32_t k1_long_seg(uint32_t k, uint32_t w) {
    const uint32_t s = 8u * (k + w);
    return s < 256u ? 256u : s;
}

// exclusive prefix sum over the CTA's 1024 threads (every thread calls it); `total` is the sum over the CTA
__device__ __forceinline__ unsigned long long k1_plan_scan(const unsigned long long v, unsigned long long *ws /* [33] */,
                                                           unsigned long long &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) ws[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const unsigned long long w = ws[lane];
        unsigned long long winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        ws[lane] = winc - w;
        if (lane == 31) ws[32] = winc;
    }
    __syncthreads();
    const unsigned long long r = ws[warp] + inc - v;
    total = ws[32];
    __syncthreads();                                             // ws is reused by the next call
    return r;
}

// use_list: the candidates are the reads queued by the fast kernels; otherwise every read of the batch (w > 32)
__global__ void __launch_bounds__(K1_LONG_PLAN_TPB) k1_long_plan(const K1Params p, const bool use_list) {
    __shared__ unsigned int s_n;
    __shared__ unsigned long long s_end;
    __shared__ unsigned long long ws[33];
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    const uint64_t total = use_list ? (uint64_t)min(*p.ovf_count, p.ovf_cap) : p.n_reads;
    for (uint64_t q = threadIdx.x; q < total; q += K1_LONG_PLAN_TPB) {
        const uint64_t r = use_list ? (p.ovf_list[q] & 0xffffffffull) : q;      // (the high word is k1_generic's size hint)
        const uint64_t b0 = k1_read_off(p, r), len = k1_read_off(p, r + 1) - b0;
        if (len < p.long_min) continue;
        const unsigned int slot = atomicAdd(&s_n, 1u);
        if (slot < p.long_cap) {
            K1LongTask t{};
            t.r = r;
            t.b0 = b0;
            t.len = len;
            p.long_tasks[slot] = t;
        } else {
            k1_report(p, r, K1_ERR_OVF);
        }
    }
    __syncthreads();
    const uint32_t n = s_n < p.long_cap ? s_n : p.long_cap;
    // tables behind k1_generic's allocations (it is done: nobody else allocates now), slices numbered in task order
    const unsigned long long base = (p.slab_entries + *p.arena_cursor + 7ull) & ~7ull;
    if (threadIdx.x == 0) s_end = base;
    __syncthreads();
    unsigned long long carry_e = 0, carry_s = 0;
    for (uint32_t c0 = 0; c0 < n; c0 += K1_LONG_PLAN_TPB) {
        const uint32_t i = c0 + threadIdx.x;
        const bool have = i < n;
        K1LongTask t{};
        if (have) t = p.long_tasks[i];
        const unsigned long long entries = have ? k1_long_table_entries(t.len, (int32_t)p.k) : 0ull;
        unsigned long long tot_e, tot_s;
        const unsigned long long tab = base + carry_e + k1_plan_scan(entries, ws, tot_e);
        const bool fits = have && tab + entries <= p.arena_entries;       // (once one does not fit, none behind it does)
        const unsigned long long nseg = fits ? (t.len + p.long_seg - 1) / p.long_seg : 0ull;
        const unsigned long long first = carry_s + k1_plan_scan(nseg, ws, tot_s);
        if (have) {
            t.tab = tab;
            t.cap = entries - 8;
            t.seg_first = first;
            t.n_seg = nseg;
            p.long_tasks[i] = t;
            if (fits) atomicMax(&s_end, tab + entries);
            else k1_report(p, t.r, K1_ERR_OVF);
        }
        carry_e += tot_e;
        carry_s += tot_s;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    *p.arena_cursor = s_end - p.slab_entries;
    K1LongCtl c;
    c.n_tasks = n;
    c.n_segs = carry_s;
    c.zero_begin = base;
    c.zero_end = s_end;
    *p.long_ctl = c;
}

// What the host wants to know about a batch whose offsets only exist on the device: out[0] = the longest read,
// out[1] = the arena entries the sets of its long sequences take (both start at 0)
__global__ void __launch_bounds__(256) k1_length_stats(const uint64_t *offsets, const uint64_t n_reads, const uint64_t long_min,
                                                       const int32_t k, unsigned long long *out) {
    unsigned long long mx = 0, ent = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_reads; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t len = offsets[i + 1] - offsets[i];
        mx = len > mx ? len : mx;
        if (len >= long_min) ent += k1_long_table_entries(len, k);
    }
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long m2 = __shfl_down_sync(0xffffffffu, mx, o);
        mx = m2 > mx ? m2 : mx;
        ent += __shfl_down_sync(0xffffffffu, ent, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&out[0], mx);
        if (ent) atomicAdd(&out[1], ent);
    }
}

__global__ void __launch_bounds__(256) k1_long_zero(const K1Params p) {
    const K1LongCtl c = *p.long_ctl;
    // (both ends are multiples of 8 entries, the arena itself is 256-byte aligned)
    uint4 *const base = reinterpret_cast<uint4 *>(p.arena + c.zero_begin);
    const uint64_t n16 = (c.zero_end - c.zero_begin) / 2;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x)
        base[i] = make_uint4(0u, 0u, 0u, 0u);
}

// first insertion of m into the sequence's set?
__device__ __forceinline__ bool k1_long_insert(uint64_t *tab, const uint64_t cap, const uint64_t m) {
    if (m == 0) return atomicExch(reinterpret_cast<unsigned long long *>(tab + cap), 1ull) == 0ull;
    uint64_t h = k1_long_slot(m);
    for (;;) {
        h &= cap - 1;
        unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(tab + h);
        if (cur == 0ull) cur = atomicCAS(reinterpret_cast<unsigned long long *>(tab + h), 0ull, (unsigned long long)m);
        if (cur == 0ull) return true;
        if (cur == (unsigned long long)m) return false;
        h++;
    }
}

template <bool DUMP>
__global__ void __launch_bounds__(K1_LONG_TPB) k1_long_scan(const K1Params p) {
    uint64_t vhbuf[257];
    const K1LongCtl c = *p.long_ctl;
    unsigned long long local_minimizers = 0;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < c.n_segs; g += (uint64_t)gridDim.x * blockDim.x) {
        // the task whose slice range holds g: the last one that starts at or before g (tasks without slices share
        // their start with the task behind them)
        uint32_t lo = 0, hi = (uint32_t)c.n_tasks;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (p.long_tasks[mid].seg_first <= g) lo = mid;
            else hi = mid;
        }
        const K1LongTask t = p.long_tasks[lo];
        const int64_t begin = (int64_t)((g - t.seg_first) * p.long_seg);
        uint64_t *const tab = p.arena + t.tab;
        uint64_t last = 0;
        bool have_last = false;
        uint32_t n_new = 0;
        k1_scan_range(p.bases + t.b0, (int64_t)t.len, (int32_t)p.k, (int32_t)p.w, begin, begin + (int64_t)p.long_seg,
                      K1LocalVH{vhbuf}, [&](uint64_t m) {
            if (have_last && m == last) return;                  // a minimum usually holds for several positions
            last = m;
            have_last = true;
            if (!k1_long_insert(tab, t.cap, m)) return;
            if (DUMP) {
                const uint32_t e = atomicAdd(&p.dump_counts[t.r], 1u);
                if (e < p.dump_cap) p.dump[t.r * p.dump_cap + e] = m;
            } else {
                atomicAdd(&p.hist[jump_hash(m, p.D)], 1u);
                if (p.feed_queue) k1_feed_queue(p, t.r, m);
            }
            n_new++;
        });
        local_minimizers += n_new;
    }
    if (!DUMP) {
        for (int o = 16; o > 0; o >>= 1) local_minimizers += __shfl_down_sync(0xffffffffu, local_minimizers, o);
        if ((threadIdx.x & 31) == 0 && local_minimizers) atomicAdd(p.n_minimizers, local_minimizers);
    }
}

}  // namespace hulk
