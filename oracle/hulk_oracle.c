/*
 * hulk_oracle.c -- CPU restatement of the `hulk sketch` hot path (will-rowe/hulk v1.0.0).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (hulk_b200/) never links, imports or executes anything in oracle/.
 *
 * Every function cites the reference file:line it restates (paths relative to the
 * reference checkout).  Arithmetic is float64 / uint64 exactly as the Go code has it.
 * The reference's flush is racy (src/pipeline/boss.go:114 "TODO: pause the minimizer chan");
 * this oracle implements the race-free intent: every read up to an interval boundary
 * is fully counted before the flush is taken.
 *
 * Parity status:
 *   - minimizer / hash64 / jump hash / histogram / count-min / CWS update: pinned by the
 *     reference's own fixture (testing/test-reads-small.fq.gz), the published jump-hash
 *     and math/rand known answers in tests/, and a second independent restatement
 *     (oracle/pyref.py).  The reference has no golden values of its own for this path.
 *   - CWS tables: Go math/rand stream pinned by known answers (oracle/go_rand.c);
 *     leesper/go_rng's gamma sampler is restated from its documented origin (CPython
 *     random.gammavariate) -- parity unpinned for that one function (source not vendored).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* ---------------------------------------------------------------------------------- */
/* src/minimizer/minimizer.go:13-30  seq_nt4_table                                     */
/* ---------------------------------------------------------------------------------- */
static uint8_t NT4[256];
static int nt4_ready = 0;
static void nt4_init(void) {
    if (nt4_ready) return;
    for (int i = 0; i < 256; i++) NT4[i] = 4;
    NT4[0] = 0; NT4[1] = 1; NT4[2] = 2; NT4[3] = 3;          /* row 0 of the table */
    NT4['A'] = NT4['a'] = 0;
    NT4['C'] = NT4['c'] = 1;
    NT4['G'] = NT4['g'] = 2;
    NT4['T'] = NT4['t'] = 3;
    NT4['U'] = NT4['u'] = 3;
    nt4_ready = 1;
}
ORACLE_API uint8_t hulk_oracle_nt4(uint8_t b) { nt4_init(); return NT4[b]; }

/* src/minimizer/minimizer.go:33-42  hash64 (minimap2 invertible mix) */
ORACLE_API uint64_t hulk_oracle_hash64(uint64_t key, uint64_t mask) {
    key = (~key + (key << 21)) & mask;
    key = key ^ key >> 24;
    key = ((key + (key << 3)) + (key << 8)) & mask;
    key = key ^ key >> 14;
    key = ((key + (key << 2)) + (key << 4)) & mask;
    key = key ^ key >> 28;
    key = (key + (key << 31)) & mask;
    return key;
}

/* github.com/dgryski/go-jump v0.0.0-20170409065014-e1f439676b57 (go.mod:8), Hash().
 * Not vendored; this is the published Lamping-Veach loop ("A Fast, Minimal Memory,
 * Consistent Hash Algorithm", 2014), call sites src/kmerspectrum/kmerspectrum.go:70
 * and src/countmin/countmin.go:125.  Pinned by the paper's/library's vectors in tests. */
ORACLE_API int32_t hulk_oracle_jump(uint64_t key, int32_t num_buckets) {
    int64_t b = -1, j = 0;
    while (j < (int64_t)num_buckets) {
        b = j;
        key = key * 2862933555777941757ULL + 1;
        j = (int64_t)((double)(b + 1) * ((double)(1LL << 31) / (double)((key >> 33) + 1)));
    }
    return (int32_t)b;
}

/* ---------------------------------------------------------------------------------- */
/* src/minimizer/minimizer.go:59-93 NewMinimizerSketch + :96-204 findMinimizers        */
/* Returns 0 or a negative error mirroring the reference's four checks.               */
/* out[] receives the per-read SET in first-insertion order; *n_out its cardinality.  */
/* ---------------------------------------------------------------------------------- */
#define ERR_W      (-1)  /* "w must be: 0 < w < 257"                 minimizer.go:62-64 */
#define ERR_K      (-2)  /* "k size must be: 0 < k < 32"             minimizer.go:65-67 */
#define ERR_EMPTY  (-3)  /* "sequence length must be > 0"            minimizer.go:71-73 */
#define ERR_SHORT  (-4)  /* "sequence length must be >= w + k - 1"   minimizer.go:74-76 */
#define ERR_CAP    (-5)  /* caller's buffer too small (oracle only)                     */
#define ERR_SPARSE (-6)  /* "not used yet"                           kmerspectrum.go:94-96 */

typedef struct { uint64_t X; int32_t Y; } pair_t;   /* src/queue/queue.go:6-9 */

ORACLE_API int hulk_oracle_minimizers(uint32_t k_, uint32_t w_, const uint8_t *seq, int64_t len64,
                                      uint64_t *out, int64_t cap, int64_t *n_out) {
    nt4_init();
    *n_out = 0;
    if (w_ > 256) return ERR_W;
    if (k_ > 31) return ERR_K;
    int32_t len = (int32_t)len64;
    if (len < 1) return ERR_EMPTY;
    if (len < (int32_t)(w_ + k_ - 1)) return ERR_SHORT;
    const int32_t k = (int32_t)k_, w = (int32_t)w_;

    uint64_t kmers[2] = {0, 0};
    int32_t kmerSpan = 0;
    const uint64_t bitmask = (k_ >= 32) ? ~0ULL : ((1ULL << (2 * k_)) - 1ULL);
    const uint64_t bitshift = (uint64_t)(int64_t)(2 * (k - 1));     /* Go: uint64(2*(k-1)) */

    /* monotone deque (src/queue/queue.go), at most w+1 live entries */
    pair_t *q = (pair_t *)malloc(sizeof(pair_t) * (size_t)(len + 1));
    int64_t qh = 0, qt = 0;                                         /* [qh, qt) */
    int64_t n = 0;
    /* membership test of the per-read set (mapset, minimizer.go:189-198): a linear scan for read-sized
     * inputs, an open-addressing table for chromosome-sized ones -- same set, first-insertion order kept */
    uint64_t *tab = NULL, tab_mask = 0;
    uint8_t *used = NULL;
    if (len > 8192) {
        uint64_t tcap = 64;
        while (tcap < (uint64_t)len) tcap <<= 1;
        tab = (uint64_t *)malloc(sizeof(uint64_t) * tcap);
        used = (uint8_t *)calloc(tcap, 1);
        tab_mask = tcap - 1;
    }

    for (int32_t i = 0; i < len; i++) {
        int32_t windowIndex = i - w + 1;
        uint8_t c = NT4[seq[i]];
        /* c > 3: empty TODO in the reference (minimizer.go:118-122) -- not skipped */
        if ((windowIndex + 1) < k) kmerSpan = windowIndex + 1; else kmerSpan = k;
        kmers[0] = (kmers[0] << 2 | (uint64_t)c) & bitmask;
        /* Go shifts >= 64 yield 0 */
        uint64_t hi = (bitshift >= 64) ? 0 : (((uint64_t)3 ^ (uint64_t)c) << bitshift);
        kmers[1] = (kmers[1] >> 2) | hi;
        if (i < k - 1) continue;
        if (kmers[0] == kmers[1]) continue;
        unsigned strand = 0;
        if (kmers[0] > kmers[1]) strand = 1;
        pair_t cur;
        cur.X = hulk_oracle_hash64(kmers[strand], bitmask) << 8 | (uint64_t)(int64_t)kmerSpan;
        cur.Y = i;
        if (qt > qh) {
            while (qt > qh && !(q[qh].Y > (i - w))) qh++;            /* minimizer.go:165-170 */
            while (qt > qh && !(q[qt - 1].X < cur.X)) qt--;          /* minimizer.go:173-178 */
        }
        q[qt++] = cur;
        if (windowIndex >= 0) {
            uint64_t m = q[qh].X;
            int found = 0;
            if (tab) {                                               /* long read: the same set, hashed */
                uint64_t h = (m * 0x9E3779B97F4A7C15ULL) >> 20;
                for (;; h++) {
                    h &= tab_mask;
                    if (!used[h]) { used[h] = 1; tab[h] = m; break; }
                    if (tab[h] == m) { found = 1; break; }
                }
            } else {
                for (int64_t t = 0; t < n; t++) if (out[t] == m) { found = 1; break; }
            }
            if (!found) {
                if (n >= cap) { free(q); free(tab); free(used); return ERR_CAP; }
                out[n++] = m;
            }
        }
    }
    free(q);
    free(tab);
    free(used);
    *n_out = n;
    return 0;
}

/* ---------------------------------------------------------------------------------- */
/* src/kmerspectrum/kmerspectrum.go:67-81 AddHash, applied to every member of every    */
/* read's set (src/pipeline/minion.go:51-57 -> boss.go:90-95).  hist is the float64    */
/* bins[] of the reference; counts are exact integers.                                 */
/* ---------------------------------------------------------------------------------- */
ORACLE_API int hulk_oracle_count_reads(uint32_t k, uint32_t w, int32_t D, const uint8_t *bases,
                                       const uint64_t *offsets, int64_t n_reads, double *hist,
                                       uint64_t *n_minimizers) {
    int rc_all = 0;
#pragma omp parallel
    {
        uint64_t *buf = NULL; int64_t bufcap = 0;
        uint64_t local_min = 0;
#pragma omp for schedule(dynamic, 256)
        for (int64_t r = 0; r < n_reads; r++) {
            int64_t len = (int64_t)(offsets[r + 1] - offsets[r]);
            if (len + 1 > bufcap) { bufcap = 2 * (len + 1); buf = (uint64_t *)realloc(buf, sizeof(uint64_t) * (size_t)bufcap); }
            int64_t n = 0;
            int rc = hulk_oracle_minimizers(k, w, bases + offsets[r], len, buf, bufcap, &n);
            if (rc != 0) {
#pragma omp critical
                { if (rc_all == 0) rc_all = rc; }
                continue;
            }
            for (int64_t t = 0; t < n; t++) {
                int32_t bin = hulk_oracle_jump(buf[t], D);
#pragma omp atomic
                hist[bin] += 1.0;
            }
            local_min += (uint64_t)n;
        }
#pragma omp atomic
        *n_minimizers += local_min;
        free(buf);
    }
    return rc_all;
}

/* ---------------------------------------------------------------------------------- */
/* src/countmin/countmin.go                                                            */
/* ---------------------------------------------------------------------------------- */
typedef struct {
    uint32_t depth, width;      /* countmin.go:31-32 */
    double *q;                  /* depth x width */
    int apply_scaling;          /* countmin.go:50-55 */
    double decay_weight;
} cms_t;

static cms_t *cms_new(double epsilon, double delta, double decay_ratio) {
    cms_t *c = (cms_t *)calloc(1, sizeof(cms_t));
    c->width = (uint32_t)ceil(2 / epsilon);                          /* :31 */
    c->depth = (uint32_t)ceil(log(1 - delta) / log(0.5));            /* :32 */
    c->q = (double *)calloc((size_t)c->width * c->depth, sizeof(double));
    if (decay_ratio > 0.0 && decay_ratio < 1.0) {                    /* :50-55 */
        c->decay_weight = exp(-decay_ratio);
        c->apply_scaling = 1;
    } else {
        c->apply_scaling = 0;                                        /* decayWeight stays 0.0 */
        c->decay_weight = 0.0;
    }
    return c;
}
static void cms_free(cms_t *c) { if (c) { free(c->q); free(c); } }

/* countmin.go:103-147 Add -> scale + traverse */
static double cms_add(cms_t *c, uint64_t element, double increment) {
    if (c->apply_scaling) {
        size_t n = (size_t)c->width * c->depth;
        for (size_t i = 0; i < n; i++) c->q[i] = c->q[i] * c->decay_weight;
    }
    double cur = DBL_MAX;
    for (uint32_t d = 0; d < c->depth; d++) {
        uint64_t hash = element + ((uint64_t)d * element);
        int32_t g = hulk_oracle_jump(hash, (int32_t)c->width);
        double *cell = &c->q[(size_t)d * c->width + (size_t)g];
        if (increment != 0.0) *cell += increment;
        if (*cell < cur) cur = *cell;
    }
    return cur;
}

/* ---------------------------------------------------------------------------------- */
/* src/histosketch/histosketch.go                                                      */
/* ---------------------------------------------------------------------------------- */
typedef struct {
    uint32_t k, s;
    int32_t D;
    int drift;                  /* ApplyConceptDrift, histosketch.go:79-81 */
    const double *r, *c, *b;    /* s x D row-major, borrowed */
    uint64_t *sketch;           /* "mins"    */
    double *weights;            /* "weights" */
    cms_t *cms;
    uint64_t n_adds;            /* number of AddElement calls so far */
} hs_t;

#define ERR_HS_K      (-10)  /* "histosketching only supports k <= 31"        :53-55 */
#define ERR_HS_DECAY  (-11)  /* "decay ratio must be between 0.0 and 1.0"     :56-64 */
#define ERR_HS_BINS   (-12)  /* "histogram must have at least 2 bins"         :65-67 */

ORACLE_API int hulk_oracle_hs_new(uint32_t k, uint32_t s, int32_t D, double decay, const double *r,
                                  const double *c, const double *b, void **out) {
    *out = NULL;
    if (k > 31) return ERR_HS_K;
    if (decay < 0.0 || decay > 1.0 || decay != decay) return ERR_HS_DECAY;
    if (D < 2) return ERR_HS_BINS;
    hs_t *h = (hs_t *)calloc(1, sizeof(hs_t));
    h->k = k; h->s = s; h->D = D;
    h->r = r; h->c = c; h->b = b;
    h->sketch = (uint64_t *)calloc(s ? s : 1, sizeof(uint64_t));
    h->weights = (double *)calloc(s ? s : 1, sizeof(double));
    h->cms = cms_new(0.001, 0.99, decay);                            /* countmin.go:11,14 */
    h->drift = (decay != 1.0);
    for (uint32_t i = 0; i < s; i++) { h->sketch[i] = 0; h->weights[i] = DBL_MAX; }
    *out = h;
    return 0;
}
ORACLE_API void hulk_oracle_hs_free(void *p) {
    hs_t *h = (hs_t *)p;
    if (!h) return;
    free(h->sketch); free(h->weights); cms_free(h->cms); free(h);
}

/* histosketch.go:30-33 getSample */
static inline double get_sample(const hs_t *h, uint64_t i, uint32_t j, double freq) {
    size_t at = (size_t)j * (size_t)h->D + (size_t)i;
    double Yka = exp(log(freq) - h->b[at]);
    return h->c[at] / (Yka * exp(h->r[at]));
}

/* histosketch.go:129-155 AddElement; returns the count-min estimate (for test taps) */
ORACLE_API double hulk_oracle_hs_add_element(void *p, uint64_t bin, double value) {
    hs_t *h = (hs_t *)p;
    double f = cms_add(h->cms, bin, value);
    h->n_adds++;
    const double dw = h->cms->decay_weight;
    for (uint32_t j = 0; j < h->s; j++) {
        double A = get_sample(h, bin, j, f);
        double curMin = h->drift ? h->weights[j] / dw : h->weights[j];
        if (A < curMin) { h->sketch[j] = bin; h->weights[j] = A; }
    }
    return f;
}

/* src/pipeline/boss.go:112-128 flush + kmerspectrum.go:84-112 Dump + :58-64 Wipe.
 * f_out (optional, D doubles) receives the count-min estimate of every non-zero bin.
 * parallel == 0: the literal loop "for each non-zero bin ascending: AddElement".
 * parallel != 0: the same (slot, bin) operations in the same per-slot order, but with the
 * count-min estimates of the whole flush computed first and the slot loop outermost and
 * spread over OpenMP threads (slots are independent chains: histosketch.go:135-153 touches
 * only Sketch[j], SketchWeights[j]).  Used for the all-cores CPU baseline; tests check it is
 * bit-identical to the literal loop. */
ORACLE_API int hulk_oracle_hs_flush(void *p, double *hist, double *f_out, int parallel) {
    hs_t *h = (hs_t *)p;
    int64_t used = 0;
    for (int32_t i = 0; i < h->D; i++) if (hist[i] != 0.0) used++;
    if (used == 0) return 0;                                         /* boss.go:117 */
    double prop = (double)used / (double)h->D;
    if (prop < 0.01) return ERR_SPARSE;                              /* kmerspectrum.go:94-96 */
    if (!parallel) {
        for (int32_t i = 0; i < h->D; i++) {
            if (hist[i] != 0.0) {
                double f = hulk_oracle_hs_add_element(h, (uint64_t)i, hist[i]);
                if (f_out) f_out[i] = f;
            }
        }
    } else {
        int32_t *bins = (int32_t *)malloc(sizeof(int32_t) * (size_t)used);
        double *fs = (double *)malloc(sizeof(double) * (size_t)used);
        int64_t n = 0;
        for (int32_t i = 0; i < h->D; i++) {
            if (hist[i] != 0.0) {
                bins[n] = i;
                fs[n] = cms_add(h->cms, (uint64_t)i, hist[i]);
                if (f_out) f_out[i] = fs[n];
                n++;
            }
        }
        h->n_adds += (uint64_t)n;
        const double dw = h->cms->decay_weight;
#pragma omp parallel for schedule(static)
        for (uint32_t j = 0; j < h->s; j++) {
            uint64_t S = h->sketch[j];
            double W = h->weights[j];
            for (int64_t t = 0; t < n; t++) {
                double A = get_sample(h, (uint64_t)bins[t], j, fs[t]);
                double curMin = h->drift ? W / dw : W;
                if (A < curMin) { S = (uint64_t)bins[t]; W = A; }
            }
            h->sketch[j] = S;
            h->weights[j] = W;
        }
        free(bins); free(fs);
    }
    memset(hist, 0, sizeof(double) * (size_t)h->D);                  /* Wipe */
    return 0;
}

ORACLE_API void hulk_oracle_hs_get(void *p, uint64_t *mins, double *weights) {
    hs_t *h = (hs_t *)p;
    memcpy(mins, h->sketch, sizeof(uint64_t) * h->s);
    memcpy(weights, h->weights, sizeof(double) * h->s);
}
ORACLE_API void hulk_oracle_hs_get_cms(void *p, double *q /* 7*2000 */) {
    hs_t *h = (hs_t *)p;
    memcpy(q, h->cms->q, sizeof(double) * (size_t)h->cms->depth * h->cms->width);
}

/* ---------------------------------------------------------------------------------- */
/* src/pipeline/sketch.go:182-224 SeqMinimizer.Run + :271-285 Sketcher.Run (race-free) */
/* Drives reads -> per-interval flush -> final flush.  hist_scratch: D doubles.        */
/* ---------------------------------------------------------------------------------- */
ORACLE_API int hulk_oracle_run(void *p, uint32_t w, const uint8_t *bases, const uint64_t *offsets,
                               int64_t n_reads, uint64_t interval, int parallel,
                               uint64_t *n_minimizers, uint64_t *n_flushes) {
    hs_t *h = (hs_t *)p;
    double *hist = (double *)calloc((size_t)h->D, sizeof(double));
    int rc = 0;
    *n_minimizers = 0; *n_flushes = 0;
    int64_t done = 0;
    while (done < n_reads && rc == 0) {
        int64_t chunk = n_reads - done;
        if (interval != 0 && (uint64_t)chunk > interval) chunk = (int64_t)interval;
        rc = hulk_oracle_count_reads(h->k, w, h->D, bases, offsets + done, chunk, hist, n_minimizers);
        if (rc) break;
        done += chunk;
        if (interval != 0 && (uint64_t)chunk == interval) {          /* seqCount % Interval == 0 */
            rc = hulk_oracle_hs_flush(h, hist, NULL, parallel);
            (*n_flushes)++;
        }
    }
    if (rc == 0) { rc = hulk_oracle_hs_flush(h, hist, NULL, parallel); (*n_flushes)++; }   /* sketch.go:221 */
    free(hist);
    return rc;
}

ORACLE_API int hulk_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
ORACLE_API void hulk_oracle_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
