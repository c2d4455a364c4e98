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
//                  (arena bump allocation behind k1_generic's) and its range of slice numbers;
//   k1_long_zero   clears the tables;
//   k1_long_scan   one thread per slice, grid-stride.
#pragma once
#include <stdint.h>

#include "k1_long.h"
#include "k1_minimizer.cuh"

namespace hulk {

constexpr uint32_t K1_LONG_MIN = 1u << 14;      // sequences this long go to the sliced scan
constexpr uint32_t K1_LONG_TASKS = 1u << 16;    // long sequences per launch
constexpr int K1_LONG_PLAN_TPB = 1024;
constexpr int K1_LONG_TPB = 128;

// positions per slice: the warm-up (k + w) stays a small part of the work
__host__ __device__ inline uint32_t k1_long_seg(uint32_t k, uint32_t w) {
    const uint32_t s = 8u * (k + w);
    return s < 256u ? 256u : s;
}

// use_list: the candidates are the reads queued by the fast kernels; otherwise every read of the batch (w > 32)
__global__ void __launch_bounds__(K1_LONG_PLAN_TPB) k1_long_plan(const K1Params p, const bool use_list) {
    __shared__ unsigned int s_n;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    const uint64_t total = use_list ? (uint64_t)min(*p.ovf_count, p.ovf_cap) : p.n_reads;
    for (uint64_t q = threadIdx.x; q < total; q += K1_LONG_PLAN_TPB) {
        const uint64_t r = use_list ? p.ovf_list[q] : q;
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
    if (threadIdx.x != 0) return;
    const uint32_t n = s_n < p.long_cap ? s_n : p.long_cap;
    unsigned long long cursor = *p.arena_cursor;                 // k1_generic is done: nobody else allocates now
    cursor = (cursor + 7ull) & ~7ull;
    const unsigned long long zero_begin = cursor;
    unsigned long long segs = 0;
    for (uint32_t i = 0; i < n; i++) {
        K1LongTask t = p.long_tasks[i];
        const uint64_t entries = k1_long_table_entries(t.len, (int32_t)p.k);
        t.seg_first = segs;
        if (cursor + entries > p.arena_entries) {
            k1_report(p, t.r, K1_ERR_OVF);
            t.n_seg = 0;
        } else {
            t.tab = cursor;
            t.cap = entries - 8;
            t.n_seg = (t.len + p.long_seg - 1) / p.long_seg;
            cursor += entries;
            segs += t.n_seg;
        }
        p.long_tasks[i] = t;
    }
    *p.arena_cursor = cursor;
    K1LongCtl c;
    c.n_tasks = n;
    c.n_segs = segs;
    c.zero_begin = zero_begin;
    c.zero_end = cursor;
    *p.long_ctl = c;
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
