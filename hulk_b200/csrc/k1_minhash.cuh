// k1_minhash.cuh -- the two MinHash side sketches of `hulk sketch --kmv / --khf`, fed from stage 1's minimizer queue.
//
// Reference semantics (paths relative to the reference checkout):
//   src/minhash/khf.go:35-45   KHFsketch.AddHash: for every slot i, val = hv + i * hv (uint64 wrap); keep the minimum
//   src/minhash/kmv.go:40-71   KMVsketch.AddHash: a max-heap of at most s values; a full heap takes hv only when
//                              hv < heap[0] (strict).  No duplicate check: what it ends up holding is the s smallest
//                              values of the MULTISET of everything added (equal values are interchangeable, so the
//                              order of arrival does not show)
//   src/minhash/kmv.go:160-176 SetSketch/GetSketch: the heap's content sorted low -> high
//   src/pipeline/boss.go:18-19,70-71,90-95  the boss builds both sketches next to the k-mer spectrum but its collector
//                              only calls kmerSpectrum.AddHash ("not used yet"); with the feed switched on
//                              (hulk_b200_minhash_enable) every minimizer the collector receives also goes to the
//                              two AddHash methods above -- the wiring SURVEY section 8(f) row 4 describes.  Off (the
//                              default) the library behaves like the reference does today: both sketches stay unfed.
//
// Both are order-independent summaries of the multiset of minimizers, so they are computed per launch from the batch's
// minimizer queue (k1_minimizer.cuh; with the feed on, k1_generic and k1_long_scan append what they bin to the same
// queue) and merged: KHF with a 64-bit atomic minimum per slot, KMV as bottom-s(pool + candidates).
//   k1_khf_queue    one pass over the queue per 256 slots; lane l of a warp owns slots l, l + 32, ...: running
//                   minima in registers, (slot + 1) * hv as one 64 x 32-bit product, one atomicMin per slot at the end
//   k1_kmv_filter   queue keys below the pool's largest value (all keys while the pool is not full) -> candidates
//   k1_kmv_select   one CTA: the s-th smallest of pool + candidates by an 8-bit radix select (eight counting passes over
//                   the candidates, which are few once the pool is full), then the new pool
#pragma once
#include <stdint.h>

#include "k1_minimizer.cuh"

namespace hulk {

constexpr int K1_KHF_TPB = 256;
constexpr int K1_KHF_R = 8;                       // slots per lane and pass (32 * 8 = 256 slots per pass)
constexpr int K1_KMV_SELECT_TPB = 1024;

struct K1KmvState {
    unsigned long long n_pool;                    // values in the pool (<= s)
    unsigned long long thr;                       // the pool's largest value once it holds s values
    unsigned long long n_cand;                    // candidates of the launch in flight
    unsigned int cur;                             // which of the two pool buffers is current
    unsigned int pad;
};

struct K1MinhashParams {
    const uint64_t *queue;
    const unsigned long long *queue_cursor;
    uint64_t queue_cap;
    uint32_t s;
    unsigned long long *khf;                      // [s], starts at MaxUint64 (khf.go:20-32); nullptr: not wanted
    K1KmvState *kmv;                              // nullptr: not wanted
    uint64_t *kmv_pool;                           // [2][s]
    uint64_t *kmv_cand;                           // [queue_cap]
};

__device__ __forceinline__ uint64_t k1_mh_filled(const K1MinhashParams &p) {
    const unsigned long long f = *p.queue_cursor;
    return f < p.queue_cap ? f : p.queue_cap;
}

__global__ void __launch_bounds__(K1_KHF_TPB) k1_khf_queue(const K1MinhashParams p) {
    const uint64_t total = k1_mh_filled(p);
    const uint64_t nwarps = (uint64_t)gridDim.x * (K1_KHF_TPB / 32);
    const uint64_t gw = (uint64_t)blockIdx.x * (K1_KHF_TPB / 32) + (threadIdx.x >> 5);
    const uint64_t i0 = total * gw / nwarps, i1 = total * (gw + 1) / nwarps;
    if (i0 >= i1) return;
    const uint32_t lane = threadIdx.x & 31;
    for (uint32_t base = 0; base < p.s; base += 32 * K1_KHF_R) {
        unsigned long long mn[K1_KHF_R];
#pragma unroll
        for (int r = 0; r < K1_KHF_R; r++) mn[r] = ~0ull;
        for (uint64_t i = i0; i < i1; i++) {
            const unsigned long long hv = p.queue[i];                            // one address per warp: a broadcast load
#pragma unroll
            for (int r = 0; r < K1_KHF_R; r++) {
                const unsigned long long val = hv * (unsigned long long)(base + lane + 32u * r + 1u);   // khf.go:39
                mn[r] = val < mn[r] ? val : mn[r];                               // khf.go:41-43
            }
        }
#pragma unroll
        for (int r = 0; r < K1_KHF_R; r++) {
            const uint32_t slot = base + lane + 32u * r;
            if (slot < p.s && mn[r] != ~0ull) atomicMin(&p.khf[slot], mn[r]);
        }
    }
}

__global__ void __launch_bounds__(256) k1_kmv_filter(const K1MinhashParams p) {
    const uint64_t total = k1_mh_filled(p);
    const bool full = p.kmv->n_pool >= p.s;
    const unsigned long long thr = p.kmv->thr;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long hv = p.queue[i];
        if (full && !(hv < thr)) continue;                                       // kmv.go:62 (strict)
        const unsigned long long at = atomicAdd(&p.kmv->n_cand, 1ull);
        p.kmv_cand[at] = hv;                                                     // (at < total <= queue_cap)
    }
}

__global__ void __launch_bounds__(K1_KMV_SELECT_TPB) k1_kmv_select(const K1MinhashParams p) {
    __shared__ unsigned int hist[256];
    __shared__ unsigned long long s_prefix, s_rank;
    __shared__ unsigned int s_out, s_eq;
    K1KmvState *const st = p.kmv;
    const uint64_t n_c = st->n_cand, n_p = st->n_pool, total = n_c + n_p;
    const uint64_t *const pool_in = p.kmv_pool + (size_t)st->cur * p.s;
    uint64_t *const pool_out = p.kmv_pool + (size_t)(st->cur ^ 1u) * p.s;
    const uint64_t *const cand = p.kmv_cand;
    const int tid = threadIdx.x;
    if (n_c == 0) return;                                                        // nothing new: the pool stands
    auto key = [&](uint64_t i) { return i < n_c ? cand[i] : pool_in[i - n_c]; };
    if (total <= p.s) {                                                          // kmv.go:57-59: the heap is not full yet
        for (uint64_t i = tid; i < total; i += K1_KMV_SELECT_TPB) pool_out[i] = key(i);
        __syncthreads();
        if (total == p.s) {                                                      // full from now on: its largest value
            unsigned long long mx = 0;
            for (uint64_t i = tid; i < total; i += K1_KMV_SELECT_TPB) { const unsigned long long v = key(i); mx = v > mx ? v : mx; }
            if (tid == 0) s_prefix = 0;
            __syncthreads();
            atomicMax(&s_prefix, mx);
            __syncthreads();
        }
        if (tid == 0) {
            st->thr = (total == p.s) ? s_prefix : ~0ull;
            st->n_pool = total;
            st->n_cand = 0;
            st->cur ^= 1u;
        }
        return;
    }
    // ---- the s-th smallest value v of the multiset (1-based rank s), most significant byte first
    if (tid == 0) { s_prefix = 0; s_rank = p.s; }
    for (int pass = 0; pass < 8; pass++) {
        const int shift = 56 - 8 * pass;
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        const unsigned long long prefix = s_prefix;
        for (uint64_t i = tid; i < total; i += K1_KMV_SELECT_TPB) {
            const unsigned long long v = key(i);
            if (pass == 0 || (v >> (shift + 8)) == prefix) atomicAdd(&hist[(v >> shift) & 255ull], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned long long rank = s_rank, cum = 0;
            unsigned int d = 0;
            for (; d < 256; d++) {
                if (cum + hist[d] >= rank) break;
                cum += hist[d];
            }
            s_rank = rank - cum;                                                 // rank among the values that share the new prefix
            s_prefix = (prefix << 8) | d;
        }
        __syncthreads();
    }
    const unsigned long long v = s_prefix, take_eq = s_rank;                     // take_eq copies of v belong to the bottom s
    if (tid == 0) { s_out = 0; s_eq = 0; }
    __syncthreads();
    for (uint64_t i = tid; i < total; i += K1_KMV_SELECT_TPB) {
        const unsigned long long x = key(i);
        bool take = x < v;
        if (x == v) take = atomicAdd(&s_eq, 1u) < take_eq;
        if (take) pool_out[atomicAdd(&s_out, 1u)] = x;
    }
    __syncthreads();
    if (tid == 0) {
        st->thr = v;
        st->n_pool = p.s;
        st->n_cand = 0;
        st->cur ^= 1u;
    }
}

}  // namespace hulk
