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
#if defined(__CUDA_ARCH__)
#define HULK_UNROLL _Pragma("unroll")
#else
#define HULK_UNROLL
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

// ---- fast path of the Lamping-Veach step -------------------------------------------------
// One step is j' = trunc(fl(fl(2^31 / q) * (j + 1))) with q = (key >> 33) + 1.  The exact IEEE
// quotient is only needed when the product lies within a few ulps of an integer, so the hot
// loop evaluates x ~= (j + 1) * 2^31 / q from the hardware reciprocal seed plus one Newton
// step (relative error <= JUMP_RCP_ERR = 2^-39.88: the seed is good to 2^-19.94, measured over every
// q in [1, 2^31] by hulk_b200_rcp_selftest, tests/test_gpu_parity.py) and
// brackets the reference value: with EPS >= JUMP_RCP_ERR + 2^-50 the reference product lies
// in [x(1-EPS), x(1+EPS)], so when both ends truncate to the same integer that integer IS the
// reference's result.  Otherwise (probability ~ 2 EPS x per step) the caller recomputes the
// step with the true division.  No rounding-mode or fast-math assumptions leak out: the
// result is bit-identical to jump_hash() above by construction.
constexpr double JUMP_EPS = 1.8189894035458565e-12;       // 2^-39
constexpr double JUMP_TWO52 = 4503599627370496.0;         // 2^52

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ double u32_to_double(uint32_t v) {   // exact, no conversion unit
    return __hiloint2double(0x43300000, (int)v) - JUMP_TWO52;
}
__device__ __forceinline__ double rcp_seed(double x) {          // MUFU.RCP64H, ~2^-23 relative
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
}
__device__ __forceinline__ double fma_down(double a, double b, double c) { return __fma_rd(a, b, c); }
__device__ __forceinline__ uint32_t dbl_lo(double x) { return (uint32_t)__double2loint(x); }
__device__ __forceinline__ uint32_t dbl_hi(double x) { return (uint32_t)__double2hiint(x); }
__device__ __forceinline__ double dbl_make(uint32_t hi, uint32_t lo) { return __hiloint2double((int)hi, (int)lo); }
#else
}  // namespace hulk
#include <cfenv>
#include <cmath>
#include <cstring>
namespace hulk {
inline uint32_t dbl_lo(double x) { uint64_t u; std::memcpy(&u, &x, 8); return (uint32_t)u; }
inline uint32_t dbl_hi(double x) { uint64_t u; std::memcpy(&u, &x, 8); return (uint32_t)(u >> 32); }
inline double dbl_make(uint32_t hi, uint32_t lo) { uint64_t u = ((uint64_t)hi << 32) | lo; double x; std::memcpy(&x, &u, 8); return x; }
inline double u32_to_double(uint32_t v) { return dbl_make(0x43300000u, v) - JUMP_TWO52; }
// host stand-in for the hardware seed: 1/x with the mantissa cut to 24 bits (same error class)
inline double rcp_seed(double x) { const double r = 1.0 / x; return dbl_make(dbl_hi(r), dbl_lo(r) & 0xF0000000u); }
inline double fma_down(double a, double b, double c) {
    const int old = std::fegetround();
    std::fesetround(FE_DOWNWARD);
    volatile double va = a, vb = b, vc = c;
    const double r = std::fma(va, vb, vc);
    std::fesetround(old);
    return r;
}
#endif

// state of one chain: key, current bucket b, jd1 = (double)(b + 1)
// returns 0: stepped (b, jd1 updated), 1: finished (b is the answer), 2: ambiguous -> use jump_step_exact
HULK_HD int jump_step_fast(uint64_t &key, uint32_t &b, double &jd1, uint32_t num_buckets) {
    key = key * 2862933555777941757ull + 1ull;
    const uint32_t q = (uint32_t)(key >> 33) + 1u;                       // 1 .. 2^31
    const double qd = u32_to_double(q);
    const double Q = dbl_make(dbl_hi(qd) - (31u << 20), dbl_lo(qd));     // q * 2^-31, exact
    const double r0 = rcp_seed(Q);
    const double e = fma(-Q, r0, 1.0);
    const double R = fma(r0, e, r0);                                     // ~ 2^31 / q
    const double x = jd1 * R;
    const double tl = fma_down(x, 1.0 - JUMP_EPS, JUMP_TWO52);           // 2^52 + floor(x (1-EPS))
    const double th = fma_down(x, 1.0 + JUMP_EPS, JUMP_TWO52);
    const uint32_t nl = dbl_lo(tl), nh = dbl_lo(th);
    if (dbl_hi(th) != 0x43300000u || nl >= num_buckets) return 1;        // x(1-EPS) >= 2^32 - 1 or >= n
    if (nl != nh) return 2;
    b = nl;
    jd1 = tl - (JUMP_TWO52 - 1.0);                                       // (double)(nl + 1), exact
    return 0;
}
// the same step with the reference's arithmetic; `key` has already been advanced by jump_step_fast
HULK_HD int jump_step_exact(uint64_t key, uint32_t &b, double &jd1, uint32_t num_buckets) {
    const int64_t j = (int64_t)((double)((int64_t)b + 1) * (2147483648.0 / (double)((key >> 33) + 1)));
    if (j >= (int64_t)num_buckets) return 1;
    b = (uint32_t)j;
    jd1 = (double)(j + 1);
    return 0;
}
// jump_hash() evaluated through the fast step (what the kernel's chains compute)
HULK_HD int32_t jump_hash_fast(uint64_t key, int32_t num_buckets, uint32_t *n_ambiguous = nullptr) {
    uint32_t b = 0;
    double jd1 = 1.0;
    for (;;) {
        int rc = jump_step_fast(key, b, jd1, (uint32_t)num_buckets);
        if (rc == 2) {
            if (n_ambiguous) ++*n_ambiguous;
            rc = jump_step_exact(key, b, jd1, (uint32_t)num_buckets);
        }
        if (rc) return (int32_t)b;
    }
}

// ---- fixed-point form of the fast step, for num_buckets <= 2^20 (every k^4-bin spectrum: 31^4 < 2^20) ----
// x = (b + 1) 2^31 / q is evaluated as above (seed, one Newton step folded into the product) and added to
// 2^32: y = 2^32 + x carries 20 fraction bits in the low word of its mantissa, floor(x) in the 32
// bits above them.  While x < 2^20 the value of y is within 1.6 units of 2^-20 of the real quotient
// (x 2^-39.88 from the reciprocal, 2^-21 from rounding y; the reference's own double rounding moves its product
// by less than 2^-30), so unless the fraction lies within 3 units of an integer (0xFFFFD .. 0xFFFFF, 0 .. 2) the
// integer part IS the reference's trunc(); otherwise (2^-17 per step) the caller recomputes the step with the
// true division.
// x >= 2^20 means "finished" whatever the error is (num_buckets <= 2^20).  One FMA replaces the two directed
// roundings of the bracket, and the state between steps is the 32-bit bucket alone.
constexpr double JUMP_TWO32 = 4294967296.0;               // 2^32
constexpr double JUMP_TWO83M = 9671406556917033397649408.0 - 2147483648.0;   // 2^83 - 2^31 = (2^52 - 1) 2^31, exact
constexpr uint32_t JUMP_FX_MAX_BUCKETS = 1u << 20;
// returns 0: stepped (b updated), 1: finished (b is the answer), 2: ambiguous -> use jump_step_exact
HULK_HD int jump_step_fx(uint64_t &key, uint32_t &b, const uint32_t num_buckets) {
    key = key * 2862933555777941757ull + 1ull;
    const double qd = dbl_make(0x43300000u, (uint32_t)(key >> 33)) - (JUMP_TWO52 - 1.0);   // (double)q, q = (key >> 33) + 1
    const double jd1 = dbl_make(0x45200000u, b) - JUMP_TWO83M;           // (2^83 + b 2^31) - (2^83 - 2^31) = (b + 1) 2^31
    const double r0 = rcp_seed(qd);
    const double e = fma(-qd, r0, 1.0);
    const double jr = jd1 * r0;
    const double x = fma(jr, e, jr);                                                       // ~ (b + 1) 2^31 / q
    const double y = x + JUMP_TWO32;
    if ((uint32_t)((dbl_lo(y) + 3u) << 12) < (6u << 12)) return 2;
    if (y >= JUMP_TWO32 + (double)num_buckets) return 1;
    b = (dbl_hi(y) << 12) | (dbl_lo(y) >> 20);                                             // floor(x): the exponent bits shift out
    return 0;
}
HULK_HD int32_t jump_hash_fx(uint64_t key, int32_t num_buckets, uint32_t *n_ambiguous = nullptr) {
    uint32_t b = 0;
    for (;;) {
        int rc = jump_step_fx(key, b, (uint32_t)num_buckets);
        if (rc == 2) {
            if (n_ambiguous) ++*n_ambiguous;
            double jd1;
            rc = jump_step_exact(key, b, jd1, (uint32_t)num_buckets);
        }
        if (rc) return (int32_t)b;
    }
}

// ---- word-wise base encoding ---------------------------------------------------------------
// Four ASCII bases in one little-endian word -> four 2-bit codes (byte j of the result = code of
// base j) and a flag telling whether every byte was one of ACGTUacgtu.  (b >> 1 ^ b >> 2) & 3
// maps A,C,G,T/U -> 0,1,2,3 for either case; the check rebuilds the expected letters from the
// codes with a byte permute and compares.  Words that fail fall back to nt4() per byte.
HULK_HD uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t sel) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    const uint64_t pool = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((pool >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
#endif
}
HULK_HD uint32_t nt4x4(uint32_t w, bool &all_acgtu) {
    const uint32_t v = ((w >> 1) ^ (w >> 2)) & 0x03030303u;
    const uint32_t s1 = v | (v >> 4);
    const uint32_t sel = byte_perm(s1, 0u, 0x4420u) & 0xffffu;           // nibble j = code of base j
    const uint32_t expect = byte_perm(0x54474341u /* "ACGT" */, 0u, sel);
    const uint32_t is_t = v & (v >> 1) & 0x01010101u;                    // code 3: accept T and U
    all_acgtu = (((w & 0xDFDFDFDFu) ^ expect) & ~is_t) == 0u;
    return v;
}

}  // namespace hulk
