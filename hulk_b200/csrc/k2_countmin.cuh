// k2_countmin.cuh -- stage 3a of the sketch hot path: the persistent count-min sketch that every
// non-zero histogram bin passes through before the CWS update.
//
// Reference semantics (paths relative to the reference checkout):
//   src/pipeline/boss.go:112-128      flush: skip if no bin used, Dump, Wipe
//   src/kmerspectrum/kmerspectrum.go:84-112  Dump: "not used yet" if < 1% of bins used; bins
//                                     are delivered in ascending order
//   src/histosketch/histosketch.go:132       estiFreq = cmSketch.Add(bin, value)
//   src/countmin/countmin.go:28-57,103-147   7 x 2000 float64 counters; key = bin*(d+1);
//                                     column = jump.Hash(key, 2000); decay scales ALL counters
//                                     by exp(-ratio) before EVERY Add.
//
// The sequential loop "for bin ascending: Q[d][col_d(bin)] += v; f = min_d Q[d][col_d(bin)]" is a
// segmented inclusive scan: counter (d, col) sees exactly the bins of its static list
// {bin : jump(bin*(d+1), 2000) == col} in ascending order.  The lists (CSR) are built once per
// context; per flush one warp per counter scans its list.  Without decay the sums are integers
// in float64 (exact, order independent).  With decay, counter value after the hit at global add
// index t is Q(t) = Q(t_prev) * w^(t - t_prev) + v, scanned with the associative operator
// (t1,B1) o (t2,B2) = (t2, B1 * w^(t2 - t1) + B2); w^n is exp(n ln w) instead of n roundings of a
// repeated product (differs from the reference by <= ~n * 2^-53 relative).
#pragma once
#include <cuda_bf16.h>
#include <math.h>
#include <stdint.h>

#include "hd_math.h"

namespace hulk {

constexpr uint32_t CMS_DEPTH = 7;      // ceil(ln(1-0.99)/ln(0.5))   countmin.go:32
constexpr uint32_t CMS_WIDTH = 2000;   // ceil(2/0.001)              countmin.go:31
constexpr uint32_t CMS_CELLS = CMS_DEPTH * CMS_WIDTH;
constexpr unsigned long long F_EMPTY_BITS = 0x7FF0000000000000ull;   // +inf: bin not in this flush

// per-flush control block in device memory
struct FlushCtl {
    // two flushes can be in flight (count-min of flush i+1 while the CWS sweep of flush i runs): the
    // per-flush words are indexed by the flush's parity `fi`
    unsigned int nnz[2];           // used bins of the histogram being flushed
    unsigned int go[2];            // 1: this flush runs; 0: histogram empty (no-op) or error
    int err;                       // sticky: HULK_B200_ESPARSE once a flush was < 1% used
    unsigned int pad;
    unsigned long long t0;         // AddElement calls before this flush
    unsigned long long n_adds;     // AddElement calls so far
    unsigned long long n_flushes;  // non-empty flushes so far
    unsigned long long n_rescans;  // fp64 chunk re-evaluations (k3_resolve)
};

// ---- context creation: static column map and its CSR ----
__global__ void k2_build_cols(int32_t D, uint16_t *cols /*[7][D]*/, unsigned int *counts /*[14000]*/) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)D * CMS_DEPTH) return;
    const uint32_t d = (uint32_t)(i / D);
    const uint64_t bin = (uint64_t)(i % D);
    const uint64_t hash = bin + (uint64_t)d * bin;                         // countmin.go:122
    const int32_t g = jump_hash(hash, (int32_t)CMS_WIDTH);                 // countmin.go:125
    cols[i] = (uint16_t)g;
    atomicAdd(&counts[d * CMS_WIDTH + g], 1u);
}
// single block: exclusive scan of the 14000 list lengths
__global__ void k2_scan_counts(const unsigned int *counts, unsigned int *start /*[14001]*/) {
    __shared__ unsigned int warp_tot[32];
    __shared__ unsigned int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < CMS_CELLS; base += blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        const unsigned int v = (i < CMS_CELLS) ? counts[i] : 0;
        unsigned int s = v;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int u = __shfl_up_sync(0xffffffffu, s, o);
            if ((threadIdx.x & 31) >= o) s += u;
        }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned int t = (threadIdx.x < (blockDim.x >> 5)) ? warp_tot[threadIdx.x] : 0;
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int u = __shfl_up_sync(0xffffffffu, t, o);
                if (threadIdx.x >= o) t += u;
            }
            warp_tot[threadIdx.x] = t;   // inclusive
        }
        __syncthreads();
        const unsigned int before = carry + ((threadIdx.x >> 5) ? warp_tot[(threadIdx.x >> 5) - 1] : 0);
        if (i < CMS_CELLS) start[i] = before + s - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += warp_tot[(blockDim.x >> 5) - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0) start[CMS_CELLS] = carry;
}
__global__ void k2_fill_csr(int32_t D, const uint16_t *cols, const unsigned int *start, unsigned int *cursor,
                            int32_t *csr_bins) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)D * CMS_DEPTH) return;
    const uint32_t d = (uint32_t)(i / D);
    const int32_t bin = (int32_t)(i % D);
    const uint32_t cell = d * CMS_WIDTH + cols[i];
    const unsigned int at = atomicAdd(&cursor[cell], 1u);
    csr_bins[start[cell] + at] = bin;
}
// one thread per counter: put its list in ascending bin order (insertion sort; lists are short)
__global__ void k2_sort_csr(const unsigned int *start, int32_t *csr_bins) {
    const uint32_t cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= CMS_CELLS) return;
    const unsigned int b = start[cell], e = start[cell + 1];
    for (unsigned int i = b + 1; i < e; i++) {
        const int32_t v = csr_bins[i];
        unsigned int j = i;
        while (j > b && csr_bins[j - 1] > v) { csr_bins[j] = csr_bins[j - 1]; j--; }
        csr_bins[j] = v;
    }
}

// ---- per flush ----
// (b) exclusive scan of the per-block counts + the flush decision.  Run by the LAST block of k2_mask_count[_peers] to
// finish (a ticket counter decides who that is), so the bitmap, the ranks and the decision are one launch.
__device__ __forceinline__ void k2_decide_tail(const uint32_t *block_count, uint32_t nblocks, uint32_t *__restrict__ block_prefix,
                                               int32_t D, FlushCtl *ctl, const int fi) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nblocks; base += blockDim.x) {
        if (threadIdx.x < 32) warp_tot[threadIdx.x] = 0;                  // blocks of fewer than 32 warps
        __syncthreads();
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = (i < nblocks) ? __ldcg(block_count + i) : 0;   // written by other blocks: read at the L2
        uint32_t s = v;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, s, o);
            if ((threadIdx.x & 31) >= o) s += u;
        }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t t = warp_tot[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, t, o);
                if (threadIdx.x >= o) t += u;
            }
            warp_tot[threadIdx.x] = t;
        }
        __syncthreads();
        const uint32_t before = carry + ((threadIdx.x >> 5) ? warp_tot[(threadIdx.x >> 5) - 1] : 0);
        if (i < nblocks) block_prefix[i] = before + s - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += warp_tot[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const unsigned int nnz = carry;                                   // used bins of this flush
        ctl->nnz[fi] = nnz;
        unsigned int go = 0;
        if (nnz != 0) {                                                   // boss.go:117
            const double prop = (double)nnz / (double)D;                  // kmerspectrum.go:92
            if (prop < 0.01) {                                            // kmerspectrum.go:94-96
                if (ctl->err == 0) ctl->err = -6;                         // HULK_B200_ESPARSE
            } else {
                go = 1;
            }
        }
        ctl->go[fi] = go;
        ctl->t0 = ctl->n_adds;
        if (go) {
            ctl->n_adds += nnz;
            ctl->n_flushes += 1;
        }
    }
}
// true in every thread of the block that finishes last (ticket returns to 0 for the next launch)
__device__ __forceinline__ bool k2_last_block(unsigned int *ticket) {
    __shared__ unsigned int last;
    __threadfence();                                                      // this thread's results first ...
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;   // ... then the ticket
    __syncthreads();
    if (last && threadIdx.x == 0) *ticket = 0;
    return last != 0u;
}

// (a) used-bin bitmap, per-32-bin and per-1024-bin used counts, reset of the estimate vector.
// A block covers 1024 bins with 256 threads (every warp four 32-bin words): small blocks find room next to the counting
// kernels of the next interval, which a 1024-thread block does not.
constexpr int K2_MASK_TPB = 256;
__global__ void __launch_bounds__(K2_MASK_TPB) k2_mask_count(const uint32_t *__restrict__ hist, int32_t D,
                                                             uint32_t *__restrict__ words, uint32_t *__restrict__ word_prefix,
                                                             uint32_t *__restrict__ block_count,
                                                             unsigned long long *__restrict__ fbits, FlushCtl *ctl,
                                                             const int fi, uint32_t *__restrict__ block_prefix,
                                                             unsigned int *ticket) {
    __shared__ uint32_t pc[32];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int wd = wid * 4 + j;                                     // word of the block
        const int64_t i = (int64_t)blockIdx.x * 1024 + wd * 32 + lane;
        const bool nz = (i < D) && (hist[i] != 0u);
        if (i < D) fbits[i] = F_EMPTY_BITS;
        const uint32_t word = __ballot_sync(0xffffffffu, nz);
        if (lane == 0) {
            words[(size_t)blockIdx.x * 32 + wd] = word;
            pc[wd] = __popc(word);
        }
    }
    __syncthreads();
    if (wid == 0) {
        const uint32_t v = pc[lane];
        uint32_t s = v;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += u;
        }
        word_prefix[(size_t)blockIdx.x * 32 + lane] = s - v;   // used bins before this word, inside the block
        if (lane == 31) block_count[blockIdx.x] = s;
    }
    if (k2_last_block(ticket)) k2_decide_tail(block_count, gridDim.x, block_prefix, D, ctl, fi);
}
// ---- multi-GPU: the spectrum of a flush is the sum of every GPU's counting buffer --------
// The GPUs of a node reach each other's memory over NVLink (peer access inside one process, CUDA IPC across processes),
// so the all-reduce is not a collective call: every GPU reads the other buffers straight from their owners while it
// builds its used-bin bitmap (k2_mask_count_peers).  0.78 MB per peer at k = 21: latency, not bandwidth.
// Ordering is by sequence numbers in device memory: the owner of a buffer PUSHES "counted up to use n" into every
// reader's flag array when its counting kernels are done (k_peer_signal), a reader waits on its LOCAL copy (k_peer_wait,
// one warp), and the reader's last block pushes "gathered use n" back to the owner, who waits for all of them before
// the buffer is wiped and counted into again.  Integer sums: the result is bit-identical to one GPU whatever the
// arrival order is.
constexpr uint32_t PEER_MAX = 16;                   // GPUs of one node
struct PeerTargets {                                // by value: up to PEER_MAX remote words to set to `seq`
    uint32_t *flag[PEER_MAX];
    uint32_t n;
    uint32_t seq;
};
struct PeerSources {                                // by value: the same spectrum buffer on every GPU
    const uint32_t *hist[PEER_MAX];
    uint32_t n;
};
__global__ void k_peer_signal(const PeerTargets t) {
    __threadfence_system();                         // everything this stream did before is visible system-wide first
    if (threadIdx.x < t.n) *reinterpret_cast<volatile uint32_t *>(t.flag[threadIdx.x]) = t.seq;
    __threadfence_system();
}
// one warp; flags: this GPU's array [n] for the buffer in question; gives up after timeout_ns (a peer died) with ctl->err set
__global__ void k_peer_wait(const uint32_t *flags, const uint32_t n, const uint32_t seq, FlushCtl *ctl,
                            const unsigned long long timeout_ns) {
    if (threadIdx.x < n) {
        const volatile uint32_t *f = flags + threadIdx.x;
        unsigned long long t0 = 0, now = 0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while ((int32_t)(*f - seq) < 0) {
            __nanosleep(200);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (now - t0 > timeout_ns) {
                if (ctl->err == 0) ctl->err = -30;  // HULK_B200_ECUDA: peer timeout
                break;
            }
        }
    }
    __threadfence_system();
}
// k2_mask_count over the SUM of the peers' buffers; the sum is kept (hist_sum) for the kernels behind it
__global__ void __launch_bounds__(K2_MASK_TPB) k2_mask_count_peers(const PeerSources src, int32_t D,
                                                                   uint32_t *__restrict__ hist_sum,
                                                                   uint32_t *__restrict__ words,
                                                                   uint32_t *__restrict__ word_prefix,
                                                                   uint32_t *__restrict__ block_count,
                                                                   unsigned long long *__restrict__ fbits, FlushCtl *ctl,
                                                                   const int fi, const PeerTargets done, unsigned int *ticket,
                                                                   uint32_t *__restrict__ block_prefix) {
    __shared__ uint32_t pc[32];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int wd = wid * 4 + j;
        const int64_t i = (int64_t)blockIdx.x * 1024 + wd * 32 + lane;
        uint32_t sum = 0;
        if (i < D) {
            for (uint32_t p = 0; p < src.n; p++) sum += __ldcv(src.hist[p] + i);   // never a stale cached copy
            hist_sum[i] = sum;
            fbits[i] = F_EMPTY_BITS;
        }
        const uint32_t word = __ballot_sync(0xffffffffu, sum != 0u);
        if (lane == 0) {
            words[(size_t)blockIdx.x * 32 + wd] = word;
            pc[wd] = __popc(word);
        }
    }
    __syncthreads();
    if (wid == 0) {
        const uint32_t v = pc[lane];
        uint32_t s = v;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += u;
        }
        word_prefix[(size_t)blockIdx.x * 32 + lane] = s - v;
        if (lane == 31) block_count[blockIdx.x] = s;
    }
    // the last block to finish tells every owner that this GPU is done with its buffer, then takes the flush decision
    if (k2_last_block(ticket)) {
        __threadfence_system();
        if (threadIdx.x < done.n) *reinterpret_cast<volatile uint32_t *>(done.flag[threadIdx.x]) = done.seq;
        k2_decide_tail(block_count, gridDim.x, block_prefix, D, ctl, fi);
    }
}

__device__ __forceinline__ uint32_t k2_rank(int32_t bin, const uint32_t *words, const uint32_t *word_prefix,
                                            const uint32_t *block_prefix) {
    // number of used bins <= bin (1-based position of `bin` among this flush's AddElement calls)
    const uint32_t w = words[bin >> 5];
    return block_prefix[bin >> 10] + word_prefix[bin >> 5] + __popc(w & (0xffffffffu >> (31 - (bin & 31))));
}

// (c) one warp per counter
__global__ void __launch_bounds__(256) k2_cms_update(const uint32_t *__restrict__ hist,
                                                     const unsigned int *__restrict__ csr_start,
                                                     const int32_t *__restrict__ csr_bins,
                                                     const uint32_t *__restrict__ words,
                                                     const uint32_t *__restrict__ word_prefix,
                                                     const uint32_t *__restrict__ block_prefix,
                                                     double *__restrict__ q, unsigned long long *__restrict__ fbits,
                                                     const FlushCtl *__restrict__ ctl, const int fi,
                                                     const int apply_scaling, const double decay_weight) {
    if (!ctl->go[fi]) return;
    // w^n as exp(n ln w): ~5x fewer instructions than pow(); n |ln w| 2^-53 relative, the same order as the
    // rounding drift of the reference's n repeated multiplications (tolerance 1e-9)
    const double lnw = apply_scaling ? log(decay_weight) : 0.0;
    const uint32_t cell = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (cell >= CMS_CELLS) return;
    const int lane = threadIdx.x & 31;
    const unsigned int beg = csr_start[cell], end = csr_start[cell + 1];
    if (!apply_scaling) {
        double carry = q[cell];
        for (unsigned int base = beg; base < end; base += 32) {
            const unsigned int e = base + lane;
            int32_t bin = -1;
            double v = 0.0;
            if (e < end) { bin = csr_bins[e]; v = (double)hist[bin]; }
            double s = v;
            for (int o = 1; o < 32; o <<= 1) {
                const double u = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += u;
            }
            if (v != 0.0) atomicMin(&fbits[bin], (unsigned long long)__double_as_longlong(carry + s));
            carry += __shfl_sync(0xffffffffu, s, 31);
        }
        if (lane == 0) q[cell] = carry;
    } else {
        const unsigned long long t0 = ctl->t0;
        double cval = q[cell];                 // counter value as of add index ct
        unsigned long long ct = t0;
        for (unsigned int base = beg; base < end; base += 32) {
            const unsigned int e = base + lane;
            int32_t bin = -1;
            double B = 0.0;
            unsigned long long t = 0;
            bool valid = false;
            if (e < end) {
                bin = csr_bins[e];
                const uint32_t c = hist[bin];
                if (c) {
                    valid = true;
                    B = (double)c;
                    t = t0 + k2_rank(bin, words, word_prefix, block_prefix);
                }
            }
            for (int o = 1; o < 32; o <<= 1) {
                const double Bl = __shfl_up_sync(0xffffffffu, B, o);
                const unsigned long long tl = __shfl_up_sync(0xffffffffu, t, o);
                const int vl = __shfl_up_sync(0xffffffffu, (int)valid, o);
                if (lane >= o && vl) {
                    if (valid) B = Bl * exp(lnw * (double)(t - tl)) + B;
                    else { B = Bl; t = tl; valid = true; }
                }
            }
            // B: contribution of this chunk's hits up to and including this lane, as of time t
            double est = 0.0;
            if (valid) est = cval * exp(lnw * (double)(t - ct)) + B;
            if (e < end && bin >= 0 && hist[bin] != 0u) atomicMin(&fbits[bin], (unsigned long long)__double_as_longlong(est));
            const int lastv = __shfl_sync(0xffffffffu, (int)valid, 31);
            if (lastv) {
                cval = __shfl_sync(0xffffffffu, est, 31);
                ct = __shfl_sync(0xffffffffu, t, 31);
            }
        }
        const unsigned long long t1 = t0 + ctl->nnz[fi];
        if (lane == 0) q[cell] = cval * exp(lnw * (double)(t1 - ct));
    }
}

// (d) estimate -> fp32 reciprocal for the streaming filter; wipe the histogram (kmerspectrum.go:58-64)
__global__ void k2_finalize(uint32_t *__restrict__ hist, int32_t D, const unsigned long long *__restrict__ fbits,
                            float *__restrict__ invf, __nv_bfloat16 *__restrict__ invf16,
                            const FlushCtl *__restrict__ ctl, const int fi) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= D) return;
    const unsigned int go = ctl->go[fi];
    if (go) {
        const bool used = hist[i] != 0u;
        const double f = __longlong_as_double((long long)fbits[i]);
        if (invf16) invf16[i] = used ? __double2bfloat16(1.0 / f) : __float2bfloat16(__int_as_float(0x7fc00000));
        else invf[i] = used ? (float)(1.0 / f) : __int_as_float(0x7fc00000);
    }
    // boss.go:117-128: the spectrum is wiped only when a dump happened (cardinality != 0 and no error)
    if (go) hist[i] = 0u;
}

}  // namespace hulk
