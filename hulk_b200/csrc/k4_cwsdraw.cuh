// k4_cwsdraw.cuh -- HistoSketch.newCWS on the device (src/histosketch/histosketch.go:95-126).
//
// The reference draws r, c = ln(.), b = U * r for every (slot, bin) from two generators with the same seed: Go's
// math/rand source (an additive lagged Fibonacci generator, S[n] = S[n-607] + S[n-273] mod 2^64) feeding
// leesper/go_rng's Gamma(2,1) -- R.C.H. Cheng's rejection sampler, two uniforms per attempt, one when the first falls
// outside (1e-7, 0.9999999) -- and Float64Range(0,1).  csrc/host_io.cpp spreads that over the host's cores; at k = 31,
// s = 1024 (0.95 G elements, 4.3 G raw outputs) it still takes ~11 s against 0.13 s of sketching, and the tables then
// cross PCIe (22.7 GB).  Here the same algorithm runs where the tables live:
//
//   k4_apply_poly   jump-ahead: a block's generator window advanced by n outputs, n given as x^n mod (x^607 - x^334 - 1)
//                   (exact integer arithmetic mod 2^64 -- the recurrence is linear)
//   k4_raw          every block regenerates its chunk of the RAW stream from its window (273 outputs per step: the
//                   shortest lag) into a round buffer
//   k4_scan         positions whose uniform is outside (1e-7, 0.9999999) (they shift the pairing of everything behind
//                   them; ~2e-7 of all, resolved sequentially on the host), outputs that convert to exactly 1.0 (the
//                   reference redraws: the device draw gives up), and the uniforms of b (fixed position per element)
//   k4_sample       one thread per ATTEMPT of a segment with known pairing: Cheng's test in float64; attempts whose test
//                   lies within 1e-9 of the boundary are listed for the host to re-decide with its own libm, so the
//                   accept/reject sequence -- the only thing that can shift every later entry -- is the host generator's
//   k4_count/k4_scatter  ranks of the accepted draws; draw m goes to element m / 2: r when m is even, c = ln when odd
//
// Values: x = 2 exp(v), v = ln(u1 / (1 - u1)) / sqrt(3) and c = ln x are CUDA's exp/log (<= 1 ulp each, a few ulp through the
// chain); the host generator uses glibc's, Go its own -- the three agree to ~1e-15, far inside the 1e-12 weight tolerance.
#pragma once
#include <stdint.h>

namespace hulk {

constexpr int ALFG_LEN = 607, ALFG_TAP = 273, ALFG_SHIFT = ALFG_LEN - ALFG_TAP;   // 334
constexpr uint32_t K4_CHUNK = 273u * 240u;        // raw outputs per block and round (a multiple of the shortest lag)
constexpr uint32_t K4_LOOKAHEAD = 2;              // the last attempt of a round may take its second uniform from the next

// window W[i] = S[pos - 607 + i]  ->  S[pos + n - 607 + i] = sum_j poly[j] * Wext[i + j], poly = x^n mod P
// bit >= 0: only blocks whose index has that bit set (initial positions by binary decomposition)
__global__ void __launch_bounds__(512) k4_apply_poly(uint64_t *__restrict__ states, const uint64_t *__restrict__ poly, const int bit) {
    if (bit >= 0 && !((blockIdx.x >> bit) & 1u)) return;
    __shared__ uint64_t wext[2 * ALFG_LEN - 1];
    __shared__ uint64_t p[ALFG_LEN];
    uint64_t *const w = states + (size_t)blockIdx.x * ALFG_LEN;
    for (int i = threadIdx.x; i < ALFG_LEN; i += blockDim.x) {
        wext[i] = w[i];
        p[i] = poly[i];
    }
    __syncthreads();
    for (int base = 0; base < ALFG_LEN - 1; base += ALFG_TAP) {          // extend by 606 outputs, 273 at a time
        const int u = base + (int)threadIdx.x;
        if ((int)threadIdx.x < ALFG_TAP && u < ALFG_LEN - 1) wext[ALFG_LEN + u] = wext[u] + wext[ALFG_SHIFT + u];
        __syncthreads();
    }
    uint64_t acc[2] = {0, 0};
    for (int j = 0; j < ALFG_LEN; j++) {
        const uint64_t c = p[j];
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int i = (int)threadIdx.x + q * (int)blockDim.x;
            if (i < ALFG_LEN) acc[q] += c * wext[i + j];
        }
    }
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const int i = (int)threadIdx.x + q * (int)blockDim.x;
        if (i < ALFG_LEN) w[i] = acc[q];
    }
}

// Block c writes raw outputs [c * K4_CHUNK, (c + 1) * K4_CHUNK) of the round to out and leaves its window at the end of
// its chunk; the last block writes `lookahead` more without moving its window further.
__global__ void __launch_bounds__(288) k4_raw(uint64_t *__restrict__ states, uint64_t *__restrict__ out, const uint32_t lookahead) {
    __shared__ uint64_t ring[ALFG_LEN];
    uint64_t *const w = states + (size_t)blockIdx.x * ALFG_LEN;
    for (int i = threadIdx.x; i < ALFG_LEN; i += blockDim.x) ring[i] = w[i];
    __syncthreads();
    uint64_t *const o = out + (size_t)blockIdx.x * K4_CHUNK;
    uint32_t head = 0;                                                   // physical index of logical 0
    const uint32_t u = threadIdx.x;
    for (uint32_t done = 0; done < K4_CHUNK; done += ALFG_TAP) {
        if (u < (uint32_t)ALFG_TAP) {
            uint32_t a = head + u, b = head + ALFG_SHIFT + u;
            a -= a >= (uint32_t)ALFG_LEN ? ALFG_LEN : 0;
            b -= b >= (uint32_t)ALFG_LEN ? ALFG_LEN : 0;
            const uint64_t v = ring[a] + ring[b];                         // S[n] = S[n - 607] + S[n - 273]
            ring[a] = v;                                                  // logical 334 + u of the window 273 further on
            o[done + u] = v;
        }
        head += ALFG_TAP;
        head -= head >= (uint32_t)ALFG_LEN ? ALFG_LEN : 0;
        __syncthreads();
    }
    for (int i = threadIdx.x; i < ALFG_LEN; i += blockDim.x) {
        uint32_t a = head + (uint32_t)i;
        a -= a >= (uint32_t)ALFG_LEN ? ALFG_LEN : 0;
        w[i] = ring[a];
    }
    if (blockIdx.x == gridDim.x - 1 && u < lookahead) {                   // lookahead <= 273: one more (partial) step
        uint32_t a = head + u, b = head + ALFG_SHIFT + u;
        a -= a >= (uint32_t)ALFG_LEN ? ALFG_LEN : 0;
        b -= b >= (uint32_t)ALFG_LEN ? ALFG_LEN : 0;
        o[K4_CHUNK + u] = ring[a] + ring[b];
    }
}

// rand.Float64() of one raw output: float64(Int63()) / (1 << 63)  (Go math/rand)
__device__ __forceinline__ double k4_uniform(uint64_t raw) {
    return __ll2double_rn((long long)(raw & 0x7fffffffffffffffull)) * (1.0 / 9223372036854775808.0);
}

struct K4ScanOut {
    unsigned int n_extreme;        // entries written to the extremes list (may exceed its capacity: then the round is redone smaller)
    unsigned int saw_one;          // some output converts to exactly 1.0
};
// raw: the round's outputs, n of them, global position of raw[0] = pos0.  b (may be null): unscaled uniforms of elements
// [skip, E) at their fixed positions.
__global__ void k4_scan(const uint64_t *__restrict__ raw, const uint64_t n, const uint64_t pos0, uint64_t *__restrict__ extremes,
                        const uint32_t extremes_cap, K4ScanOut *__restrict__ out, double *__restrict__ b, const uint64_t skip,
                        const uint64_t E) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const double u = k4_uniform(raw[i]);
        if (u == 1.0) out->saw_one = 1u;
        if (!(1e-7 < u && u < .9999999)) {
            const unsigned int at = atomicAdd(&out->n_extreme, 1u);
            if (at < extremes_cap) extremes[at] = pos0 + i;
        }
        const uint64_t e = pos0 + i;
        if (b && e >= skip && e < E) b[e - skip] = u;                     // histosketch.go:116, times r later
    }
}

// a run of attempts with known pairing: attempt j takes the uniforms at raw positions start + 2 j, start + 2 j + 1
struct K4Segment {
    uint64_t start;                // relative to the round's first raw output
    uint64_t first_attempt;        // index of its first attempt among the round's attempts
    uint64_t n_attempts;
};
constexpr int K4_MAX_SEGMENTS = 96;
struct K4Segments {
    K4Segment seg[K4_MAX_SEGMENTS];
    uint32_t n;
    uint32_t pad;
    uint64_t total_attempts;
};
struct K4Tie {                     // an attempt the host decides
    uint64_t attempt;
    uint64_t raw1, raw2;
};
struct K4SampleOut {
    unsigned int n_ties;
    unsigned int pad;
};
// Cheng's sampler for alpha = 2, beta = 1 as go_rng ports it from CPython's random.gammavariate
__global__ void k4_sample(const uint64_t *__restrict__ raw, const K4Segments segs, double *__restrict__ xs,
                          uint8_t *__restrict__ accept, K4Tie *__restrict__ ties, const uint32_t ties_cap,
                          K4SampleOut *__restrict__ out, const double tie_eps) {
    const double alpha = 2.0;
    const double ainv = sqrt(2.0 * alpha - 1.0), bbb = alpha - log(4.0), ccc = alpha + ainv, magic = 1.0 + log(4.5);
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < segs.total_attempts;
         g += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t k = 0;
        while (k + 1 < segs.n && g >= segs.seg[k + 1].first_attempt) k++;
        const uint64_t p = segs.seg[k].start + 2 * (g - segs.seg[k].first_attempt);
        const uint64_t raw1 = raw[p], raw2 = raw[p + 1];
        const double u1 = k4_uniform(raw1), u2 = 1.0 - k4_uniform(raw2);
        const double v = log(u1 / (1.0 - u1)) / ainv;
        const double x = alpha * exp(v);
        const double z = u1 * u1 * u2;
        const double r = bbb + ccc * v - x;
        const double t1 = r + magic - 4.5 * z, t2 = r - log(z);
        const bool acc = t1 >= 0.0 || t2 >= 0.0;
        // a test this close to its boundary could come out the other way with another libm: the host decides
        const bool tie = (fabs(t1) < tie_eps && !(t2 >= tie_eps)) || (fabs(t2) < tie_eps && !(t1 >= tie_eps));
        if (tie) {
            const unsigned int at = atomicAdd(&out->n_ties, 1u);
            if (at < ties_cap) ties[at] = K4Tie{g, raw1, raw2};
        }
        xs[g] = x;
        accept[g] = acc ? 1u : 0u;
    }
}

constexpr int K4_COUNT_TPB = 1024;
// accepted draws per block of 1024 attempts
__global__ void __launch_bounds__(K4_COUNT_TPB) k4_count(const uint8_t *__restrict__ accept, const uint64_t n, uint32_t *__restrict__ block_count) {
    __shared__ uint32_t warp_tot[32];
    const uint64_t g = (uint64_t)blockIdx.x * K4_COUNT_TPB + threadIdx.x;
    const bool a = g < n && accept[g];
    const uint32_t m = __popc(__ballot_sync(0xffffffffu, a));
    if ((threadIdx.x & 31) == 0) warp_tot[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t t = warp_tot[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) block_count[blockIdx.x] = t;
    }
}
// exclusive scan of the block counts (one block), total accepted draws of the round
__global__ void __launch_bounds__(1024) k4_scan_counts(const uint32_t *__restrict__ block_count, const uint32_t nblocks,
                                                        unsigned long long *__restrict__ block_prefix,
                                                        unsigned long long *__restrict__ total) {
    __shared__ unsigned long long warp_tot[32];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nblocks; base += blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        const unsigned long long v = i < nblocks ? block_count[i] : 0;
        unsigned long long s = v;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, s, o);
            if ((threadIdx.x & 31) >= o) s += t;
        }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned long long t = warp_tot[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long q = __shfl_up_sync(0xffffffffu, t, o);
                if (threadIdx.x >= o) t += q;
            }
            warp_tot[threadIdx.x] = t;
        }
        __syncthreads();
        const unsigned long long before = carry + ((threadIdx.x >> 5) ? warp_tot[(threadIdx.x >> 5) - 1] : 0);
        if (i < nblocks) block_prefix[i] = before + s - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += warp_tot[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
// draw m (over the whole job) -> element m / 2: r when m is even, c = ln(draw) when odd (histosketch.go:112-113)
__global__ void __launch_bounds__(K4_COUNT_TPB) k4_scatter(const uint8_t *__restrict__ accept, const double *__restrict__ xs, const uint64_t n,
                                                          const unsigned long long *__restrict__ block_prefix, const uint64_t produced,
                                                          const uint64_t need, const uint64_t skip_elems, double *__restrict__ r,
                                                          double *__restrict__ c) {
    __shared__ uint32_t warp_tot[32];
    const uint64_t g = (uint64_t)blockIdx.x * K4_COUNT_TPB + threadIdx.x;
    const bool a = g < n && accept[g];
    const uint32_t bal = __ballot_sync(0xffffffffu, a);
    const uint32_t lane = threadIdx.x & 31;
    if (lane == 0) warp_tot[threadIdx.x >> 5] = __popc(bal);
    __syncthreads();
    if (threadIdx.x < 32) {
        const uint32_t v = warp_tot[threadIdx.x];
        uint32_t s = v;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, s, o);
            if (threadIdx.x >= (unsigned)o) s += t;
        }
        warp_tot[threadIdx.x] = s - v;
    }
    __syncthreads();
    if (!a) return;
    const uint64_t m = produced + block_prefix[blockIdx.x] + warp_tot[threadIdx.x >> 5] + __popc(bal & ((1u << lane) - 1u));
    if (m >= need) return;
    const uint64_t elem = m >> 1;
    if (elem < skip_elems) return;
    const double x = xs[g];
    if (m & 1) c[elem - skip_elems] = log(x);
    else r[elem - skip_elems] = x;
}
__global__ void k4_scale_b(double *__restrict__ b, const double *__restrict__ r, const uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = (0.0 + b[i] * (1.0 - 0.0)) * r[i];                 // Float64Range(0, 1) * r  (histosketch.go:116)
}

}  // namespace hulk
