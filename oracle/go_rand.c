/*
 * go_rand.c -- CPU restatement of the random streams behind HistoSketch.newCWS
 * (src/histosketch/histosketch.go:95-126).  TEST INFRASTRUCTURE ONLY (see hulk_oracle.c).
 *
 * Two third-party pieces are on that path and neither is vendored under the reference:
 *
 *  (1) Go's math/rand (stdlib; value stream frozen by the Go 1 compatibility promise):
 *      additive lagged Fibonacci generator, length 607, tap 273, seeded through the
 *      Lehmer LCG seedrand() and XOR-ed with the 607-entry rngCooked table.  The table is
 *      "the state of the generator after 780e10 iterations" of the same ALFG started from
 *      srand(1) with 20/10-bit seeding shifts (math/rand/gen_cooked.go).  It is not copied
 *      here: cooked_init() REGENERATES it by polynomial jump-ahead (x^N mod
 *      x^607 - x^334 - 1 over Z/2^64) and the result is accepted only because it
 *      reproduces the published seed-1 outputs (tests/test_oracle_gorand.py):
 *      Int63() = 5577006791947779410, 8674665223082153551, 6129484611666145821,
 *      4037200794235010051, 3916589616287113937; Float64() = 0.6046602879796196,
 *      0.9405090880450124, 0.6645600532184904; Intn(100) = 81, 87, 47, 59, 81, 18, 25, 40, 56, 0.
 *
 *  (2) github.com/leesper/go_rng v0.0.0-20190531154944-a612b043e353 (go.mod:9):
 *      UniformGenerator{rand.New(rand.NewSource(seed))}, Float64Range(a,b) = a + Float64()*(b-a);
 *      GammaGenerator owns its own UniformGenerator(seed); Gamma(alpha>1, beta) is a port of
 *      CPython's random.gammavariate (R.C.H. Cheng 1977).  Restated from that published
 *      algorithm; tests check this restatement against CPython's own implementation fed
 *      the same uniform stream.  PARITY UNPINNED against go_rng itself (no vectors exist).
 *      Note the quick-accept constant does not influence the output: "r + C - 4.5 z >= 0"
 *      implies "r >= ln z" for any C <= 1 + ln 4.5, so both tests accept the same draws.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

#define RNG_LEN 607
#define RNG_TAP 273
#define INT32MAX 2147483647LL

static uint64_t g_cooked[RNG_LEN];
static int g_cooked_ready = 0;

/* math/rand/rng.go seedrand(): x[n+1] = 48271 * x[n] mod (2**31 - 1) */
static int32_t seedrand(int32_t x) {
    const int32_t A = 48271, Q = 44488, R = 3399;
    int32_t hi = x / Q, lo = x % Q;
    x = A * lo - R * hi;
    if (x < 0) x += (int32_t)INT32MAX;
    return x;
}

/* c = a*b mod (x^607 - x^334 - 1), coefficients mod 2^64 */
static void polymul(const uint64_t *a, const uint64_t *b, uint64_t *out) {
    static uint64_t res[2 * RNG_LEN];
    memset(res, 0, sizeof(res));
    for (int i = 0; i < RNG_LEN; i++) {
        uint64_t ai = a[i];
        if (!ai) continue;
        for (int j = 0; j < RNG_LEN; j++) res[i + j] += ai * b[j];
    }
    for (int d = 2 * RNG_LEN - 2; d >= RNG_LEN; d--) {     /* x^607 = x^334 + 1 */
        uint64_t c = res[d];
        if (!c) continue;
        res[d] = 0;
        res[d - RNG_LEN + (RNG_LEN - RNG_TAP)] += c;
        res[d - RNG_LEN] += c;
    }
    memcpy(out, res, sizeof(uint64_t) * RNG_LEN);
}

static void cooked_init(void) {
    if (g_cooked_ready) return;
    /* gen_cooked.go srand(1): 20/10-bit shifts, no cooked XOR */
    uint64_t vec[RNG_LEN];
    int32_t x = 1;
    for (int i = -20; i < RNG_LEN; i++) {
        x = seedrand(x);
        if (i >= 0) {
            uint64_t u = (uint64_t)x << 20;
            x = seedrand(x); u ^= (uint64_t)x << 10;
            x = seedrand(x); u ^= (uint64_t)x;
            vec[i] = u;
        }
    }
    /* ALFG as a linear recurrence s_n = s_{n-607} + s_{n-273}; the output of step m is
     * written to slot (334 - m) mod 607, so s_m (m = -606..0) = vec[(334 - m) mod 607]. */
    uint64_t init[RNG_LEN];
    for (int i = 0; i < RNG_LEN; i++) {
        long m = -606 + i;
        init[i] = vec[(int)(((334 - m) % RNG_LEN + RNG_LEN) % RNG_LEN)];
    }
    const uint64_t N = 7800000000000ULL;                   /* 780e10 */
    uint64_t result[RNG_LEN] = {0}, base[RNG_LEN] = {0};
    result[0] = 1; base[1] = 1;
    for (uint64_t n = N; n; n >>= 1) {
        if (n & 1) polymul(result, base, result);
        polymul(base, base, base);
    }
    /* the last 607 outputs s_m, m = N-606..N, land in slot (334 - m) mod 607 */
    for (uint64_t t = 0; t < RNG_LEN; t++) {
        uint64_t val = 0;
        for (int i = 0; i < RNG_LEN; i++) val += result[i] * init[i];
        uint64_t m_mod = (N - 606 + t) % RNG_LEN;
        int slot = (int)((334 + RNG_LEN - m_mod) % RNG_LEN);
        g_cooked[slot] = val;
        /* result *= x */
        uint64_t c = result[RNG_LEN - 1];
        memmove(result + 1, result, sizeof(uint64_t) * (RNG_LEN - 1));
        result[0] = c;
        result[RNG_LEN - RNG_TAP] += c;
    }
    g_cooked_ready = 1;
}

ORACLE_API void go_rand_cooked(uint64_t *out /* 607 */) {
    cooked_init();
    memcpy(out, g_cooked, sizeof(g_cooked));
}

/* math/rand/rng.go rngSource */
typedef struct {
    int tap, feed;
    uint64_t vec[RNG_LEN];
} go_source_t;

static void go_source_seed(go_source_t *r, int64_t seed) {
    cooked_init();
    r->tap = 0;
    r->feed = RNG_LEN - RNG_TAP;
    seed = seed % INT32MAX;
    if (seed < 0) seed += INT32MAX;
    if (seed == 0) seed = 89482311;
    int32_t x = (int32_t)seed;
    for (int i = -20; i < RNG_LEN; i++) {
        x = seedrand(x);
        if (i >= 0) {
            uint64_t u = (uint64_t)x << 40;
            x = seedrand(x); u ^= (uint64_t)x << 20;
            x = seedrand(x); u ^= (uint64_t)x;
            u ^= g_cooked[i];
            r->vec[i] = u;
        }
    }
}
static inline uint64_t go_source_uint64(go_source_t *r) {
    r->tap--;  if (r->tap < 0) r->tap += RNG_LEN;
    r->feed--; if (r->feed < 0) r->feed += RNG_LEN;
    uint64_t x = r->vec[r->feed] + r->vec[r->tap];
    r->vec[r->feed] = x;
    return x;
}
static inline int64_t go_int63(go_source_t *r) { return (int64_t)(go_source_uint64(r) & 0x7fffffffffffffffULL); }
/* math/rand/rand.go Float64(): float64(Int63()) / (1<<63), resample on 1.0 */
static inline double go_float64(go_source_t *r) {
    for (;;) {
        double f = (double)go_int63(r) / 9223372036854775808.0;
        if (f != 1.0) return f;
    }
}

ORACLE_API void *go_rand_new(int64_t seed) {
    go_source_t *r = (go_source_t *)malloc(sizeof(go_source_t));
    go_source_seed(r, seed);
    return r;
}
ORACLE_API void go_rand_free(void *p) { free(p); }
ORACLE_API int64_t go_rand_int63(void *p) { return go_int63((go_source_t *)p); }
ORACLE_API double go_rand_float64(void *p) { return go_float64((go_source_t *)p); }
/* math/rand/rand.go Int31n (n not a power of two, n <= 1<<31 - 1) via Intn */
ORACLE_API int32_t go_rand_intn(void *p, int32_t n) {
    go_source_t *r = (go_source_t *)p;
    if ((n & (n - 1)) == 0) return (int32_t)(go_int63(r) >> 32) & (n - 1);
    int32_t max = (int32_t)((1U << 31) - 1 - (1U << 31) % (uint32_t)n);
    int32_t v = (int32_t)(go_int63(r) >> 32);
    while (v > max) v = (int32_t)(go_int63(r) >> 32);
    return v % n;
}

/* go_rng GammaGenerator.Gamma(alpha > 1, beta) == CPython random.gammavariate */
static double go_rng_gamma(go_source_t *u, double alpha, double beta) {
    const double SG_MAGICCONST = 1.0 + log(4.5);
    double ainv = sqrt(2.0 * alpha - 1.0);
    double bbb = alpha - log(4.0);
    double ccc = alpha + ainv;
    for (;;) {
        double u1 = go_float64(u);
        if (!(1e-7 < u1 && u1 < .9999999)) continue;
        double u2 = 1.0 - go_float64(u);
        double v = log(u1 / (1.0 - u1)) / ainv;
        double x = alpha * exp(v);
        double z = u1 * u1 * u2;
        double r = bbb + ccc * v - x;
        if (r + SG_MAGICCONST - 4.5 * z >= 0.0 || r >= log(z)) return x * beta;
    }
}
ORACLE_API double go_rng_gamma_draw(void *p, double alpha, double beta) {
    return go_rng_gamma((go_source_t *)p, alpha, beta);
}

/* src/histosketch/histosketch.go:95-126 newCWS: r, c, b are s x D row-major.
 * Rows [slot_begin, slot_end) are written (relative to r/c/b = start of slot_begin), but the
 * streams are always advanced from slot 0 so any row range equals the full table's rows. */
ORACLE_API void hulk_oracle_new_cws(uint32_t s, int32_t D, uint32_t slot_begin, uint32_t slot_end,
                                    double *r, double *c, double *b) {
    go_source_t *gamma_src = (go_source_t *)go_rand_new(1);          /* DISTRIBUTION_SEED :20 */
    go_source_t *unif_src = (go_source_t *)go_rand_new(1);
    if (slot_end > s) slot_end = s;
    for (uint32_t i = 0; i < slot_end; i++) {
        for (int32_t j = 0; j < D; j++) {
            double rv = go_rng_gamma(gamma_src, 2, 1);
            double cv = log(go_rng_gamma(gamma_src, 2, 1));
            double bv = (0.0 + go_float64(unif_src) * (1.0 - 0.0)) * rv;   /* Float64Range(0,1)*r */
            if (i >= slot_begin) {
                size_t at = (size_t)(i - slot_begin) * (size_t)D + (size_t)j;
                r[at] = rv; c[at] = cv; b[at] = bv;
            }
        }
    }
    go_rand_free(gamma_src);
    go_rand_free(unif_src);
}
