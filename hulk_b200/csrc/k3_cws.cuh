// k3_cws.cuh -- stage 3b of the sketch hot path: the HistoSketch consistent-weighted-sampling
// update of every sketch slot by every used histogram bin of a flush.
//
// Reference semantics (paths relative to the reference checkout):
//   src/histosketch/histosketch.go:30-33    getSample: A = c / (exp(ln f - b) * exp(r))
//   src/histosketch/histosketch.go:129-155  AddElement: for every slot j: curMin = W[j] (or
//       W[j]/decayWeight under concept drift); if A < curMin { Sketch[j] = bin; W[j] = A }
//   src/pipeline/sketch.go:281-285          called once per used bin, ascending bin order
//
// Design.  A = K / f with K = c * exp(b - r) fixed per (slot, bin).  K is folded once into a low-precision
// table (bfloat16 K16[rows][Dp] by default, fp32 K32 as a switch; Dp = bins padded to 512); a flush
// streams it exactly once:
//   k3_filter16 / k3_filter (HBM-bound, "the screen"): m32[slot][chunk] = min over the chunk's 512 bins of
//              K * (1/f) in that precision, row segments arriving in shared memory through a ring of 1-D
//              bulk TMA copies; raises cand[slot] when a chunk minimum is below the slot's bound;
//   k3_resolve (tiny): per flagged slot walks the chunks in bin order carrying W exactly like the
//              reference's loop; a chunk whose screen minimum could possibly pass the test
//              (m32 < thr + eps*|thr|, eps covering the roundings of K, 1/f and their product) is
//              re-evaluated bin by bin in float64 with the reference's own formula from the
//              float64 r, c, b tables.  Every value that reaches Sketch/Weights therefore comes
//              from the float64 formula; the screen only decides which chunks cannot matter.
#pragma once
#include <cuda_bf16.h>
#include <math.h>
#include <stdint.h>

#include "k2_countmin.cuh"
#include "ptx_util.cuh"

namespace hulk {

constexpr int K3_SUB = 512;                         // bins per chunk (one warp-reduction)
constexpr int K3_SEG = 4096;                        // bins per TMA stage (16 KB)
constexpr int K3_SUBS_PER_SEG = K3_SEG / K3_SUB;    // 8 = consumer warps
constexpr int K3_STAGES_MAX = 8;                    // ring depth (template parameter): 4 x 16 KB leaves room for k1 CTAs
constexpr int K3_CONSUMER_WARPS = K3_SUBS_PER_SEG;
constexpr int K3_THREADS = (K3_CONSUMER_WARPS + 1) * 32;
constexpr double K3_EPS = 1e-6;                     // >> 3 * 2^-24 (K32, (1/f)32 and product roundings)
// bf16 screen (default): K and 1/f are stored as bfloat16 and multiplied/compared as packed pairs, so a
// flush streams 2 bytes per (slot, bin) instead of 4 and issues half the instructions.  The guard band
// widens accordingly: bfloat16 keeps 8 significant bits, so each of the three roundings (K, 1/f, product) is
// off by at most 2^-8 relative; (1 + 2^-8)^3 - 1 = 0.011765 < 0.0125.
constexpr int K3_SEG16 = 8192;                      // bins per TMA stage of the bf16 table (16 KB)
constexpr int K3_SUBS_PER_SEG16 = K3_SEG16 / K3_SUB;
constexpr double K3_EPS16 = 0.0125;

// ---- context creation: fold the float64 tables into K32 (fp32 screen) / K16 (bf16 screen) ----
__global__ void k3_fold(const double *__restrict__ r, const double *__restrict__ c, const double *__restrict__ b,
                        uint32_t rows, int32_t D, uint64_t Dp, float *__restrict__ K32) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t total = (uint64_t)rows * Dp;
    if (i >= total) return;
    const uint64_t row = i / Dp, col = i % Dp;
    float v = __int_as_float(0x7fc00000);
    if (col < (uint64_t)D) {
        const uint64_t at = row * (uint64_t)D + col;
        v = (float)(c[at] * exp(b[at] - r[at]));
    }
    K32[i] = v;
}

__global__ void k3_fold16(const double *__restrict__ r, const double *__restrict__ c, const double *__restrict__ b,
                          uint32_t rows, int32_t D, uint64_t Dp, __nv_bfloat16 *__restrict__ K16) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t total = (uint64_t)rows * Dp;
    if (i >= total) return;
    const uint64_t row = i / Dp, col = i % Dp;
    __nv_bfloat16 v = __float2bfloat16(__int_as_float(0x7fc00000));
    if (col < (uint64_t)D) {
        const uint64_t at = row * (uint64_t)D + col;
        v = __double2bfloat16(c[at] * exp(b[at] - r[at]));       // round to nearest
    }
    K16[i] = v;
}

// Blackwell packed-pair multiply (FMUL2) and three-input minimum (FMNMX3; like fminf it returns the
// non-NaN operand): 4 instructions per 4 bins instead of 8.
__device__ __forceinline__ float k3_min3(float a, float b, float c) {
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ float2 k3_mul2(float ax, float ay, float bx, float by) {
    unsigned long long ra, rb, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(ax), "f"(ay));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(bx), "f"(by));
    asm("mul.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(rd));
    return r;
}

// ---- per flush: streaming filter ----
// tile = (segment, slot); tiles are enumerated segment-major so a CTA's consecutive tiles share
// the segment's (1/f) values (L1-resident); CTA c owns the contiguous tile range [c*T/G, (c+1)*T/G).
// The screen.  thr32[slot] is the slot's test bound as it stands at the start of the flush, rounded UP to
// float: a chunk whose fp32 minimum is not below it cannot pass k3_resolve's own test
// (m < thr + eps |thr|, thr = W or W / decayWeight), so cand[slot] stays down and k3_resolve skips the slot.
// This is exact because the bound only decreases during a flush: without concept drift W only decreases;
// with drift a replacement gives W' = A < W / decayWeight < W when W < 0, and for W >= 0 (or NaN), where the
// bound may grow, thr32 is +inf (every non-empty chunk is a candidate).  k3_resolve maintains thr32.
__device__ __forceinline__ float k3_thr32(const double W, const int drift, const double thr_scale, const double eps) {
    if (drift && !(W < 0.0)) return __int_as_float(0x7f800000);
    const double thr = W * thr_scale;            // thr_scale = 1 / decayWeight under drift (a reciprocal is fine
                                                 // for a screen with a 1e-6 guard band), else 1
    const double lim = thr + eps * fabs(thr) + 1e-37;
    if (lim != lim) return __int_as_float(0xff800000);       // -inf + inf: nothing can pass
    return __double2float_ru(lim);
}
__global__ void k3_fill_f32(float *p, uint32_t n, float v) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// tile = (segment, slot); tiles are enumerated segment-major so a CTA's consecutive tiles share the
// segment's (1/f) values (L1-resident); CTA c owns the contiguous tile range [c*T/G, (c+1)*T/G).
template <int K3_STAGES>
__global__ void __launch_bounds__(K3_THREADS, 2)
k3_filter(const float *__restrict__ K32, const uint64_t Dp, const float *__restrict__ invf, float *__restrict__ m32,
          const uint32_t rows, const uint32_t nseg, const float *__restrict__ thr32, unsigned int *__restrict__ cand,
          const FlushCtl *__restrict__ ctl, const int fi) {
    if (!ctl->go[fi]) return;
    extern __shared__ __align__(128) uint8_t smem[];
    float *stage = reinterpret_cast<float *>(smem);                                    // [STAGES][SEG]
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)K3_STAGES * K3_SEG * 4);
    uint64_t *empty = full + K3_STAGES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < K3_STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], K3_CONSUMER_WARPS);
        }
        fence_mbar_init();
    }
    __syncthreads();

    const uint64_t T = (uint64_t)rows * nseg;
    const uint64_t t_begin = T * blockIdx.x / gridDim.x;
    const uint32_t n_tiles = (uint32_t)(T * (blockIdx.x + 1) / gridDim.x - t_begin);
    const uint32_t nsub_row = (uint32_t)(Dp / K3_SUB);
    uint32_t seg = (uint32_t)(t_begin / rows), slot = (uint32_t)(t_begin % rows);       // one division per CTA

    if (warp == K3_CONSUMER_WARPS) {
        // producer warp: one lane feeds the ring
        if (lane == 0) {
            uint32_t s = 0, round = 0;
            for (uint32_t it = 0; it < n_tiles; it++) {
                if (round > 0) mbar_wait(&empty[s], (round - 1) & 1);
                const uint64_t col0 = (uint64_t)seg * K3_SEG;
                const uint32_t ncols = (uint32_t)((Dp - col0 < (uint64_t)K3_SEG) ? (Dp - col0) : K3_SEG);
                mbar_arrive_expect_tx(&full[s], ncols * 4u);
                bulk_g2s_evict_first(stage + (size_t)s * K3_SEG, K32 + (uint64_t)slot * Dp + col0, ncols * 4u,
                                     &full[s]);
                if (++slot == rows) { slot = 0; seg++; }
                if (++s == K3_STAGES) { s = 0; round++; }
            }
        }
    } else {
        // consumer warps: everything that depends on the segment only is recomputed when the segment changes
        const float *my_stage = stage + warp * K3_SUB;
        uint32_t s = 0, parity = 0;
        uint64_t col0 = (uint64_t)seg * K3_SEG + (uint64_t)warp * K3_SUB;
        bool live = col0 < Dp;
        const float4 *fs = reinterpret_cast<const float4 *>(invf + (live ? col0 : 0));
        float *mout = m32 + (uint64_t)slot * nsub_row + (uint64_t)seg * K3_SUBS_PER_SEG + warp;
        for (uint32_t it = 0; it < n_tiles; it++) {
            float thr = 0.f;
            if (lane == 0) thr = thr32[slot];       // issued ahead of the wait: its latency hides behind the stage
            mbar_wait(&full[s], parity);
            float m = __int_as_float(0x7f800000);
            if (live) {
                const float4 *ks = reinterpret_cast<const float4 *>(my_stage + (size_t)s * K3_SEG);
#pragma unroll
                for (int u = 0; u < K3_SUB / 128; u++) {
                    const float4 kv = ks[u * 32 + lane];
                    const float4 fv = __ldg(&fs[u * 32 + lane]);
                    const float2 p0 = k3_mul2(kv.x, kv.y, fv.x, fv.y), p1 = k3_mul2(kv.z, kv.w, fv.z, fv.w);
                    m = k3_min3(m, p0.x, p0.y);    // NaN (empty bin / padding) is ignored, as by fminf
                    m = k3_min3(m, p1.x, p1.y);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);  // the stage is free before the (global) epilogue
            if (live) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
                if (lane == 0) {
                    *mout = m;
                    if (m < thr) atomicOr(&cand[slot], 1u);
                }
            }
            mout += nsub_row;
            if (++slot == rows) {
                slot = 0;
                seg++;
                col0 = (uint64_t)seg * K3_SEG + (uint64_t)warp * K3_SUB;
                live = col0 < Dp;
                fs = reinterpret_cast<const float4 *>(invf + (live ? col0 : 0));
                mout = m32 + (uint64_t)seg * K3_SUBS_PER_SEG + warp;
            }
            if (++s == K3_STAGES) { s = 0; parity ^= 1; }
        }
    }
}

// The same screen over the bf16 table: a stage is 8192 bins, every consumer warp owns two 512-bin chunks of
// it; K16 and (1/f)16 are multiplied and minimised as packed bf16 pairs (HMUL2.BF16 / HMNMX2.BF16, NaN = empty
// bin or padding is ignored by the minimum), the two chunk minima are reduced and written as fp32.
template <int K3_STAGES>
__global__ void __launch_bounds__(K3_THREADS, 2)
k3_filter16(const __nv_bfloat16 *__restrict__ K16, const uint64_t Dp, const __nv_bfloat16 *__restrict__ invf16,
            float *__restrict__ m32, const uint32_t rows, const uint32_t nseg, const float *__restrict__ thr32,
            unsigned int *__restrict__ cand, const FlushCtl *__restrict__ ctl, const int fi) {
    if (!ctl->go[fi]) return;
    extern __shared__ __align__(128) uint8_t smem[];
    __nv_bfloat16 *stage = reinterpret_cast<__nv_bfloat16 *>(smem);                    // [STAGES][SEG16]
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)K3_STAGES * K3_SEG16 * 2);
    uint64_t *empty = full + K3_STAGES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < K3_STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], K3_CONSUMER_WARPS);
        }
        fence_mbar_init();
    }
    __syncthreads();

    const uint64_t T = (uint64_t)rows * nseg;
    const uint64_t t_begin = T * blockIdx.x / gridDim.x;
    const uint32_t n_tiles = (uint32_t)(T * (blockIdx.x + 1) / gridDim.x - t_begin);
    const uint32_t nsub_row = (uint32_t)(Dp / K3_SUB);
    uint32_t seg = (uint32_t)(t_begin / rows), slot = (uint32_t)(t_begin % rows);

    if (warp == K3_CONSUMER_WARPS) {
        if (lane == 0) {
            uint32_t s = 0, round = 0;
            for (uint32_t it = 0; it < n_tiles; it++) {
                if (round > 0) mbar_wait(&empty[s], (round - 1) & 1);
                const uint64_t col0 = (uint64_t)seg * K3_SEG16;
                const uint32_t ncols = (uint32_t)((Dp - col0 < (uint64_t)K3_SEG16) ? (Dp - col0) : K3_SEG16);
                mbar_arrive_expect_tx(&full[s], ncols * 2u);
                bulk_g2s_evict_first(stage + (size_t)s * K3_SEG16, K16 + (uint64_t)slot * Dp + col0, ncols * 2u,
                                     &full[s]);
                if (++slot == rows) { slot = 0; seg++; }
                if (++s == K3_STAGES) { s = 0; round++; }
            }
        }
    } else {
        constexpr int PER_WARP = K3_SEG16 / K3_CONSUMER_WARPS;                         // 1024 bins = 2 chunks
        const __nv_bfloat16 *my_stage = stage + warp * PER_WARP;
        uint32_t s = 0, parity = 0;
        uint64_t col0 = (uint64_t)seg * K3_SEG16 + (uint64_t)warp * PER_WARP;
        bool live0 = col0 < Dp, live1 = col0 + K3_SUB < Dp;
        const uint4 *fs = reinterpret_cast<const uint4 *>(invf16 + (live0 ? col0 : 0));
        float *mout = m32 + (uint64_t)slot * nsub_row + (uint64_t)seg * K3_SUBS_PER_SEG16 + 2 * warp;
        const __nv_bfloat162 inf2 = __floats2bfloat162_rn(INFINITY, INFINITY);
        for (uint32_t it = 0; it < n_tiles; it++) {
            float thr = 0.f;
            if (lane == 0) thr = thr32[slot];
            mbar_wait(&full[s], parity);
            __nv_bfloat162 acc[2] = {inf2, inf2};
            if (live0) {
                const uint4 *ks = reinterpret_cast<const uint4 *>(my_stage + (size_t)s * K3_SEG16);
#pragma unroll
                for (int u = 0; u < 4; u++) {                                          // u = 0,1: chunk 0; u = 2,3: chunk 1
                    if (u >= 2 && !live1) break;
                    const uint4 kv = ks[u * 32 + lane];                                // 8 bins
                    const uint4 fv = __ldg(&fs[u * 32 + lane]);
                    const uint32_t kw[4] = {kv.x, kv.y, kv.z, kv.w}, fw[4] = {fv.x, fv.y, fv.z, fv.w};
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const __nv_bfloat162 p = __hmul2(*reinterpret_cast<const __nv_bfloat162 *>(&kw[e]),
                                                         *reinterpret_cast<const __nv_bfloat162 *>(&fw[e]));
                        acc[u >> 1] = __hmin2(acc[u >> 1], p);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (live0) {
                float m0 = fminf(__low2float(acc[0]), __high2float(acc[0]));
                float m1 = fminf(__low2float(acc[1]), __high2float(acc[1]));
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    m0 = fminf(m0, __shfl_xor_sync(0xffffffffu, m0, o));
                    m1 = fminf(m1, __shfl_xor_sync(0xffffffffu, m1, o));
                }
                if (lane == 0) {
                    mout[0] = m0;
                    if (live1) mout[1] = m1;
                    if (fminf(m0, m1) < thr) atomicOr(&cand[slot], 1u);
                }
            }
            mout += nsub_row;
            if (++slot == rows) {
                slot = 0;
                seg++;
                col0 = (uint64_t)seg * K3_SEG16 + (uint64_t)warp * PER_WARP;
                live0 = col0 < Dp;
                live1 = col0 + K3_SUB < Dp;
                fs = reinterpret_cast<const uint4 *>(invf16 + (live0 ? col0 : 0));
                mout = m32 + (uint64_t)seg * K3_SUBS_PER_SEG16 + 2 * warp;
            }
            if (++s == K3_STAGES) { s = 0; parity ^= 1; }
        }
    }
}

// reference formula, float64 (histosketch.go:30-33)
__device__ __forceinline__ double k3_sample(double c, double b, double r, double f) {
    const double Yka = exp(log(f) - b);
    return c / (Yka * exp(r));
}

// ---- per flush: exact resolve, one warp per slot ----
__global__ void __launch_bounds__(128)
k3_resolve(const float *__restrict__ m32, const uint32_t nsub_row, const double *__restrict__ r,
           const double *__restrict__ c, const double *__restrict__ b, const int32_t D,
           const unsigned long long *__restrict__ fbits, const uint32_t rows, unsigned long long *__restrict__ sketch,
           double *__restrict__ weights, const int drift, const double decay_weight, unsigned int *__restrict__ cand,
           float *__restrict__ thr32, const double thr_scale, const double eps, FlushCtl *ctl, const int fi) {
    if (!ctl->go[fi]) return;
    const uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (slot >= rows) return;
    const int lane = threadIdx.x & 31;
    if (cand[slot] == 0u) return;                   // no chunk of this flush can change the slot
    __syncwarp();
    if (lane == 0) cand[slot] = 0u;                 // lowered for the next flush
    double W = weights[slot];
    unsigned long long S = sketch[slot];
    unsigned int rescans = 0;
    const float *mrow = m32 + (uint64_t)slot * nsub_row;
    const uint64_t rowoff = (uint64_t)slot * (uint64_t)D;

    constexpr int PRE = 8;                          // chunk minima fetched per lane before they are walked
    for (uint32_t cbase0 = 0; cbase0 < nsub_row; cbase0 += 32 * PRE) {
      float mv[PRE];
#pragma unroll
      for (int j = 0; j < PRE; j++) {
          const uint32_t ci = cbase0 + 32 * j + lane;
          mv[j] = (ci < nsub_row) ? mrow[ci] : INFINITY;
      }
#pragma unroll
      for (int j = 0; j < PRE; j++) {
        const uint32_t cbase = cbase0 + 32 * j;
        if (cbase >= nsub_row) break;
        const double m = (double)mv[j];
        uint32_t pending = 0xffffffffu;
        for (;;) {
            const double thr = drift ? W / decay_weight : W;             // histosketch.go:141-146
            // conservative: could any bin of this chunk satisfy A < thr ?  (false for NaN thr)
            const bool cand = (m < INFINITY) && (m < thr + eps * fabs(thr) + 1e-37);
            const uint32_t mask = __ballot_sync(0xffffffffu, cand) & pending;
            if (mask == 0) break;
            const int first = __ffs(mask) - 1;
            pending = (first == 31) ? 0u : (0xffffffffu << (first + 1));
            // float64 re-evaluation of chunk (cbase + first), bins ascending
            const uint32_t chunk = cbase + first;
            rescans++;
            const int32_t bin_end = min((int32_t)((chunk + 1) * K3_SUB), D);
            for (int32_t b0 = (int32_t)(chunk * K3_SUB); b0 < bin_end; b0 += 32) {
                const int32_t bin = b0 + lane;
                double A = INFINITY;
                bool used = false;
                if (bin < bin_end) {
                    const unsigned long long fb = fbits[bin];
                    if (fb != F_EMPTY_BITS) {
                        used = true;
                        const double f = __longlong_as_double((long long)fb);
                        A = k3_sample(c[rowoff + bin], b[rowoff + bin], r[rowoff + bin], f);
                    }
                }
                uint32_t todo = 0xffffffffu;
                for (;;) {
                    const double th = drift ? W / decay_weight : W;
                    const bool trig = used && (A < th);                   // histosketch.go:149
                    const uint32_t tm = __ballot_sync(0xffffffffu, trig) & todo;
                    if (tm == 0) break;
                    const int fl = __ffs(tm) - 1;
                    W = __shfl_sync(0xffffffffu, A, fl);                  // :150-151
                    S = (unsigned long long)(b0 + fl);
                    todo = (fl == 31) ? 0u : (0xffffffffu << (fl + 1));
                }
            }
        }
      }
    }
    if (lane == 0) {
        weights[slot] = W;
        sketch[slot] = S;
        thr32[slot] = k3_thr32(W, drift, thr_scale, eps);    // the next flush's screen
        if (rescans) atomicAdd(&ctl->n_rescans, (unsigned long long)rescans);
    }
}

}  // namespace hulk
