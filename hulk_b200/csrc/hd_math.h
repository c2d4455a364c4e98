// hd_math.h -- pure integer/IEEE functions of the sketch hot path, usable from CUDA device code
// and (for CPU-side unit tests of the exact same source) from a plain C++ compiler.
//
// Reference semantics restated here (paths relative to the reference checkout):
//   nt4()       src/minimizer/minimizer.go:13-30   seq_nt4_table
//   hash64()    src/minimizer/minimizer.go:33-42   minimap2 invertible mix
//   jump_hash() github.com/dgryski/go-jump Hash(), called at src/kmerspectrum/kmerspectrum.go:70
//               and src/countmin/countmin.go:125
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define HULK_HD __host__ __device__ __forceinline__
#else
#define HULK_HD inline
#endif

namespace hulk {

// A/a->0 C/c->1 G/g->2 T/t/U/u->3, bytes 0..3 -> themselves, everything else -> 4.
// Branch-free arithmetic form (a 256-entry table indexed per lane would serialise in the
// constant cache and cost a dependent shared-memory load otherwise).
HULK_HD uint32_t nt4(uint32_t b) {
    const uint32_t idx = (b & 0xDFu) - 0x41u;              // letter index after upper-casing
    // 2-bit codes for A(0) C(2) G(6) T(19) U(20); validity mask for the same letters
    const uint64_t codes = (1ull << (2 * 2)) | (2ull << (2 * 6)) | (3ull << (2 * 19)) | (3ull << (2 * 20));
    const uint32_t valid = (1u << 0) | (1u << 2) | (1u << 6) | (1u << 19) | (1u << 20);
    uint32_t r = 4u;
    if (idx < 26u && ((valid >> idx) & 1u)) r = (uint32_t)(codes >> (2 * idx)) & 3u;
    if (b < 4u) r = b;
    return r;
}

// The shift-and-add lines of the reference are multiplications by constants modulo 2^64:
//   ~key + (key << 21)            == key * (2^21 - 1) - 1
//   key + (key << 3) + (key << 8) == key * 265
//   key + (key << 2) + (key << 4) == key * 21
//   key + (key << 31)             == key * (2^31 + 1)
// which map to IMAD on the FMA pipe and leave the ALU pipe to the xor-shifts.
HULK_HD uint64_t hash64(uint64_t key, uint64_t mask) {
    key = (key * 2097151ull - 1ull) & mask;
    key = key ^ (key >> 24);
    key = (key * 265ull) & mask;
    key = key ^ (key >> 14);
    key = (key * 21ull) & mask;
    key = key ^ (key >> 28);
    key = (key * 2147483649ull) & mask;
    return key;
}

// One step of the Lamping-Veach loop.  fl(2^31 / q) for an integer 1 <= q <= 2^31 is computed
// as an IEEE double division; the product has no addend, so FMA contraction cannot alter it.
HULK_HD int32_t jump_hash(uint64_t key, int32_t num_buckets) {
    int64_t b = -1, j = 0;
    while (j < (int64_t)num_buckets) {
        b = j;
        key = key * 2862933555777941757ull + 1ull;
        j = (int64_t)((double)(b + 1) * (2147483648.0 / (double)((key >> 33) + 1)));
    }
    return (int32_t)b;
}

}  // namespace hulk
