// smash.cu -- `hulk smash`: all-pairs similarity of sketches.
//
// Reference semantics (paths relative to the reference checkout):
//   cmd/smash.go:183-226            makeMatrix: similarity = 100 - 100 * distance, "%.2f"
//   src/sketchio/sketchio.go:262-306  HULKdata.GetDistance: sketches as float64 sets; for "weightedjaccard" BOTH
//                                   weight vectors are taken from the SUBJECT (hsB is asserted from
//                                   subjectSketchObj, sketchio.go:296) -- mirrored, not fixed
//   src/distances/distances.go:12-72  jaccard: 1 - matches / len;  GetWJD: 1 - sum_match min(wA,wB) / sum max(wA,wB)
//                                   with w = max(max(w,0), max(-w,0)), accumulated in slot order
//
// One thread per (subject, query) pair walks the s slots in order with IEEE double operations issued
// one by one (no fused multiply-add), so the matrix equals the reference's float64 arithmetic bit for bit.
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "../../include/hulk_b200.h"

namespace {

__global__ void k_smash(const unsigned long long *__restrict__ mins, const double *__restrict__ weights, uint32_t n,
                        uint32_t s, int weighted, double *__restrict__ sim, uint32_t i0) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;     // query
    const uint32_t i = i0 + blockIdx.y;                           // subject (grid.y holds at most 65535 rows per launch)
    if (j >= n) return;
    const unsigned long long *a = mins + (size_t)i * s, *b = mins + (size_t)j * s;
    double dist;
    if (!weighted) {
        double intersect = 0.0;
        for (uint32_t t = 0; t < s; t++)
            // the reference compares float64(min): equal uint64 values above 2^53 that round together count too
            if ((double)a[t] == (double)b[t]) intersect = __dadd_rn(intersect, 1.0);
        dist = __dsub_rn(1.0, __ddiv_rn(intersect, (double)s));
    } else {
        const double *wa = weights + (size_t)i * s;               // subject's weights stand in for both sides
        double intersect = 0.0, uni = 0.0;
        for (uint32_t t = 0; t < s; t++) {
            const double w = wa[t];
            const double wA = fmax(fmax(w, 0.0), fmax(-w, 0.0)), wB = wA;
            if ((double)a[t] == (double)b[t]) {
                if (wA < wB) { intersect = __dadd_rn(intersect, wA); uni = __dadd_rn(uni, wB); }
                else { intersect = __dadd_rn(intersect, wB); uni = __dadd_rn(uni, wA); }
            } else {
                uni = __dadd_rn(uni, (wA > wB) ? wA : wB);
            }
        }
        dist = __dsub_rn(1.0, __ddiv_rn(intersect, uni));
    }
    sim[(size_t)i * n + j] = __dsub_rn(100.0, __dmul_rn(dist, 100.0));     // cmd/smash.go:216
}

}  // namespace

extern "C" int hulk_b200_smash(const uint64_t *mins, const double *weights, uint32_t n, uint32_t s, int weighted,
                               int32_t device, double *similarity) {
    if (!mins || !similarity || (weighted && !weights) || n == 0 || s == 0) return HULK_B200_EARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return HULK_B200_ECUDA;
    if (cudaSetDevice(device) != cudaSuccess) return HULK_B200_ECUDA;
    unsigned long long *d_m = nullptr;
    double *d_w = nullptr, *d_s = nullptr;
    const size_t ne = (size_t)n * s;
    int rc = HULK_B200_OK;
    if (cudaMalloc(&d_m, ne * 8) != cudaSuccess || cudaMalloc(&d_s, (size_t)n * n * 8) != cudaSuccess ||
        (weighted && cudaMalloc(&d_w, ne * 8) != cudaSuccess)) {
        rc = HULK_B200_ENOMEM;
    } else {
        cudaMemcpy(d_m, mins, ne * 8, cudaMemcpyHostToDevice);
        if (weighted) cudaMemcpy(d_w, weights, ne * 8, cudaMemcpyHostToDevice);
        for (uint32_t i0 = 0; i0 < n; i0 += 65535u) {             // the reference has no limit on the number of sketches
            const dim3 grid((n + 127) / 128, n - i0 < 65535u ? n - i0 : 65535u);
            k_smash<<<grid, 128>>>(d_m, d_w, n, s, weighted, d_s, i0);
        }
        if (cudaGetLastError() != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) rc = HULK_B200_ECUDA;
        else cudaMemcpy(similarity, d_s, (size_t)n * n * 8, cudaMemcpyDeviceToHost);
    }
    if (d_m) cudaFree(d_m);
    if (d_w) cudaFree(d_w);
    if (d_s) cudaFree(d_s);
    return rc;
}
