// host_io.cpp -- host-side pieces of the sketch path that sit either side of the GPU kernels:
//   * the CWS sample-table generator of HistoSketch.newCWS (src/histosketch/histosketch.go:95-126):
//     Go math/rand (seed 1) driving leesper/go_rng's Gamma(2,1) and Float64Range(0,1);
//   * helpers.MD5sum over the mins (src/helpers/helpers.go:156-166);
//   * the HULKdata JSON exactly as encoding/json.MarshalIndent(..., "", "    ") emits it
//     (src/sketchio/sketchio.go:20-34,78-97; src/histosketch/histosketch.go:36-47).
// No CUDA in this file.
#include <algorithm>
#include <atomic>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/hulk_b200.h"

namespace {

// ------------------------------------------------------------------------------------------
// Go math/rand rngSource (ALFG 607/273) -- value stream frozen by the Go 1 compatibility promise
// ------------------------------------------------------------------------------------------
constexpr int kRngLen = 607, kRngTap = 273;
const uint64_t kRngCooked[kRngLen] = {
#include "go_rng_cooked.inc"
};

struct GoSource {
    int tap = 0, feed = kRngLen - kRngTap;
    uint64_t vec[kRngLen];

    static int32_t seedrand(int32_t x) {          // x[n+1] = 48271 * x[n] mod (2**31 - 1)
        const int32_t A = 48271, Q = 44488, R = 3399;
        const int32_t hi = x / Q, lo = x % Q;
        x = A * lo - R * hi;
        if (x < 0) x += 2147483647;
        return x;
    }
    explicit GoSource(int64_t seed) {
        seed %= 2147483647LL;
        if (seed < 0) seed += 2147483647LL;
        if (seed == 0) seed = 89482311;
        int32_t x = (int32_t)seed;
        for (int i = -20; i < kRngLen; i++) {
            x = seedrand(x);
            if (i >= 0) {
                uint64_t u = (uint64_t)x << 40;
                x = seedrand(x);
                u ^= (uint64_t)x << 20;
                x = seedrand(x);
                u ^= (uint64_t)x;
                vec[i] = u ^ kRngCooked[i];
            }
        }
    }
    inline uint64_t next() {
        if (--tap < 0) tap += kRngLen;
        if (--feed < 0) feed += kRngLen;
        const uint64_t x = vec[feed] + vec[tap];
        vec[feed] = x;
        return x;
    }
    // ---- jump-ahead.  The recurrence s[n] = s[n-607] + s[n-273] (mod 2^64) is linear, so the state N
    // outputs further on is x^N mod (x^607 - x^334 - 1) applied to the current window (the same algebra
    // tools/gen_rng_cooked.py uses to rebuild rngCooked).  vec[(334 - m) mod 607] holds s[m].
    static int pmod(long v) { v %= kRngLen; return (int)(v < 0 ? v + kRngLen : v); }
    // out = a * b mod (x^607 - x^334 - 1), coefficients mod 2^64
    static void polymul(const uint64_t *a, const uint64_t *b, uint64_t *out) {
        std::vector<uint64_t> res(2 * kRngLen - 1, 0);
        for (int i = 0; i < kRngLen; i++) {
            const uint64_t ai = a[i];
            if (!ai) continue;
            for (int j = 0; j < kRngLen; j++) res[i + j] += ai * b[j];
        }
        for (int d = 2 * kRngLen - 2; d >= kRngLen; d--) {       // x^607 = x^334 + 1
            const uint64_t c = res[d];
            res[d - kRngLen + (kRngLen - kRngTap)] += c;
            res[d - kRngLen] += c;
        }
        std::copy(res.begin(), res.begin() + kRngLen, out);
    }
    static void xpow(uint64_t n, uint64_t *out) {                 // x^n mod P
        std::vector<uint64_t> result(kRngLen, 0), base(kRngLen, 0);
        result[0] = 1;
        base[1] = 1;
        while (n) {
            if (n & 1) polymul(result.data(), base.data(), result.data());
            n >>= 1;
            if (n) polymul(base.data(), base.data(), base.data());
        }
        std::copy(result.begin(), result.end(), out);
    }
    // advance by n outputs given xn = x^n mod P
    void jump(const uint64_t *xn, uint64_t n) {
        const int ph = pmod(-(long)tap);                          // outputs so far, mod 607
        uint64_t init[kRngLen], res[kRngLen], outv[kRngLen];
        for (int j = 0; j < kRngLen; j++) init[j] = vec[pmod(334L - ph + 606 - j)];   // s[n0 - 606 + j]
        std::copy(xn, xn + kRngLen, res);
        for (int t = 0; t < kRngLen; t++) {                       // s[n0 + n - 606 + t]
            uint64_t acc = 0;
            for (int j = 0; j < kRngLen; j++) acc += res[j] * init[j];
            outv[t] = acc;
            const uint64_t c = res[kRngLen - 1];                  // res *= x
            for (int j = kRngLen - 1; j > 0; j--) res[j] = res[j - 1];
            res[0] = c;
            res[kRngLen - kRngTap] += c;
        }
        const int ph2 = (int)((ph + n % kRngLen) % kRngLen);
        for (int t = 0; t < kRngLen; t++) vec[pmod(334L - ph2 + 606 - t)] = outv[t];
        tap = pmod(-(long)ph2);
        feed = pmod(334L - ph2);
    }
    void skip(uint64_t n) {
        if (n < 4096) { for (uint64_t i = 0; i < n; i++) next(); return; }
        std::vector<uint64_t> xn(kRngLen);
        xpow(n, xn.data());
        jump(xn.data(), n);
    }
    // Float64() of one raw output; *is_one: it rounded to 1.0 (rand.Float64 draws again)
    static inline double to_float64(uint64_t raw, bool *is_one) {
        const double f = (double)(int64_t)(raw & 0x7fffffffffffffffULL) / 9223372036854775808.0;
        *is_one = (f == 1.0);
        return f;
    }
    // rand.Float64(): float64(Int63()) / (1 << 63), resampled when it rounds to 1.0
    inline double float64() {
        for (;;) {
            const double f = (double)(int64_t)(next() & 0x7fffffffffffffffULL) / 9223372036854775808.0;
            if (f != 1.0) return f;
        }
    }
};

// go_rng GammaGenerator.Gamma(alpha > 1, beta): R.C.H. Cheng's rejection sampler as ported from
// CPython's random.gammavariate.
inline double go_rng_gamma(GoSource &u, double alpha, double beta) {
    static const double kMagic = 1.0 + std::log(4.5);
    const double ainv = std::sqrt(2.0 * alpha - 1.0);
    const double bbb = alpha - std::log(4.0);
    const double ccc = alpha + ainv;
    for (;;) {
        const double u1 = u.float64();
        if (!(1e-7 < u1 && u1 < .9999999)) continue;
        const double u2 = 1.0 - u.float64();
        const double v = std::log(u1 / (1.0 - u1)) / ainv;
        const double x = alpha * std::exp(v);
        const double z = u1 * u1 * u2;
        const double r = bbb + ccc * v - x;
        if (r + kMagic - 4.5 * z >= 0.0 || r >= std::log(z)) return x * beta;
    }
}

// ------------------------------------------------------------------------------------------
// MD5 (RFC 1321)
// ------------------------------------------------------------------------------------------
struct Md5 {
    uint32_t a = 0x67452301u, b = 0xefcdab89u, c = 0x98badcfeu, d = 0x10325476u;
    uint64_t total = 0;
    uint8_t block[64];
    size_t fill = 0;

    static uint32_t rol(uint32_t x, int s) { return (x << s) | (x >> (32 - s)); }
    void compress(const uint8_t *p) {
        static const uint32_t K[64] = {
            0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501,
            0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821,
            0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8,
            0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a,
            0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
            0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
            0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1,
            0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
        static const int S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9,  14, 20, 5, 9,
                                  14, 20, 5, 9,  14, 20, 5, 9,  14, 20, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23,
                                  4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};
        uint32_t M[16];
        for (int i = 0; i < 16; i++)
            M[i] = (uint32_t)p[4 * i] | ((uint32_t)p[4 * i + 1] << 8) | ((uint32_t)p[4 * i + 2] << 16) |
                   ((uint32_t)p[4 * i + 3] << 24);
        uint32_t A = a, B = b, C = c, D = d;
        for (int i = 0; i < 64; i++) {
            uint32_t F;
            int g;
            if (i < 16) { F = (B & C) | (~B & D); g = i; }
            else if (i < 32) { F = (D & B) | (~D & C); g = (5 * i + 1) & 15; }
            else if (i < 48) { F = B ^ C ^ D; g = (3 * i + 5) & 15; }
            else { F = C ^ (B | ~D); g = (7 * i) & 15; }
            F = F + A + K[i] + M[g];
            A = D; D = C; C = B;
            B = B + rol(F, S[i]);
        }
        a += A; b += B; c += C; d += D;
    }
    void update(const uint8_t *p, size_t n) {
        total += n;
        while (n) {
            const size_t take = (64 - fill < n) ? 64 - fill : n;
            memcpy(block + fill, p, take);
            fill += take; p += take; n -= take;
            if (fill == 64) { compress(block); fill = 0; }
        }
    }
    void final(uint8_t out[16]) {
        const uint64_t bits = total * 8;
        const uint8_t pad = 0x80;
        update(&pad, 1);
        const uint8_t zero = 0;
        while (fill != 56) update(&zero, 1);
        uint8_t len[8];
        for (int i = 0; i < 8; i++) len[i] = (uint8_t)(bits >> (8 * i));
        update(len, 8);
        const uint32_t v[4] = {a, b, c, d};
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) out[4 * i + j] = (uint8_t)(v[i] >> (8 * j));
    }
};

// ------------------------------------------------------------------------------------------
// encoding/json formatting
// ------------------------------------------------------------------------------------------
// floatEncoder: strconv.AppendFloat(b, f, fmt, -1, 64) with fmt 'e' iff |f| < 1e-6 || |f| >= 1e21,
// then "e-0X" cleaned up to "e-X".
std::string go_float(double f) {
    if (f == 0) return std::signbit(f) ? "-0" : "0";
    char tmp[64];
    const double a = std::fabs(f);
    auto res = std::to_chars(tmp, tmp + sizeof(tmp), a, std::chars_format::scientific);   // shortest round-trip
    std::string sci(tmp, res.ptr);
    const size_t epos = sci.find('e');
    std::string mant = sci.substr(0, epos);
    const int exp10 = std::stoi(sci.substr(epos + 1));
    std::string digits;
    for (char ch : mant)
        if (ch != '.') digits.push_back(ch);
    while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
    std::string out;
    if (a < 1e-6 || a >= 1e21) {
        out = digits.substr(0, 1);
        if (digits.size() > 1) out += "." + digits.substr(1);
        out += 'e';
        out += (exp10 < 0) ? '-' : '+';
        const int ae = exp10 < 0 ? -exp10 : exp10;
        char eb[16];
        snprintf(eb, sizeof(eb), "%02d", ae);
        out += eb;
        const size_t n = out.size();
        if (n >= 4 && out[n - 4] == 'e' && out[n - 3] == '-' && out[n - 2] == '0') {
            out[n - 2] = out[n - 1];
            out.pop_back();
        }
    } else {
        const int point = exp10 + 1;   // digits before the decimal point
        if (point <= 0) {
            out = "0." + std::string((size_t)(-point), '0') + digits;
        } else if ((size_t)point >= digits.size()) {
            out = digits + std::string((size_t)point - digits.size(), '0');
        } else {
            out = digits.substr(0, (size_t)point) + "." + digits.substr((size_t)point);
        }
    }
    return (f < 0 ? "-" : "") + out;
}

// encodeState.string with escapeHTML = true
std::string go_string(const char *s) {
    std::string out = "\"";
    const unsigned char *p = reinterpret_cast<const unsigned char *>(s ? s : "");
    while (*p) {
        const unsigned char ch = *p;
        if (ch < 0x80) {
            if (ch == '"') out += "\\\"";
            else if (ch == '\\') out += "\\\\";
            else if (ch == '\n') out += "\\n";
            else if (ch == '\r') out += "\\r";
            else if (ch == '\t') out += "\\t";
            else if (ch < 0x20 || ch == '<' || ch == '>' || ch == '&') {
                char b[8];
                snprintf(b, sizeof(b), "\\u%04x", ch);
                out += b;
            } else out.push_back((char)ch);
            p++;
            continue;
        }
        // multi-byte UTF-8: U+2028 / U+2029 are escaped, invalid bytes become U+FFFD
        int n = (ch >= 0xF0) ? 4 : (ch >= 0xE0) ? 3 : (ch >= 0xC0) ? 2 : 0;
        bool ok = n != 0;
        for (int i = 1; ok && i < n; i++) ok = (p[i] & 0xC0) == 0x80;
        if (!ok) { out += "\\ufffd"; p++; continue; }
        if (n == 3 && p[0] == 0xE2 && p[1] == 0x80 && (p[2] == 0xA8 || p[2] == 0xA9)) {
            out += (p[2] == 0xA8) ? "\\u2028" : "\\u2029";
        } else {
            out.append(reinterpret_cast<const char *>(p), (size_t)n);
        }
        p += n;
    }
    out += "\"";
    return out;
}

// ------------------------------------------------------------------------------------------
// newCWS on all host cores, bit-identical to the sequential draw.
//
// The gamma stream is one generator consumed in order by a rejection sampler, so draw #m sits at a
// data-dependent position.  What makes it splittable: (1) the raw generator can jump ahead; (2) an
// attempt consumes two uniforms except when u1 falls outside (1e-7, 0.9999999), which consumes one --
// those rare positions (~2e-7 of all) are found by a scan of the raw stream and resolved sequentially,
// which fixes the attempt phase at every chunk boundary; (3) every thread then runs the sampler over its
// chunk of the RAW stream into a local buffer, a prefix sum over the accepted counts gives each chunk
// its first draw index, and the draws are scattered to r (even draws) and c = ln (odd draws).
// The uniform stream of b has a fixed position per element.  A raw output that converts to exactly 1.0
// (Float64 redraws; probability 2^-53 per output) shifts every later position: the parallel draw gives up
// and the caller falls back to the sequential loop.
// ------------------------------------------------------------------------------------------
struct GammaChunk {
    uint64_t begin = 0, end = 0;              // raw positions [begin, end)
    uint64_t start = 0;                       // first attempt start >= begin
    std::vector<uint64_t> extremes;           // raw positions in [begin, end) whose uniform is outside (1e-7, .9999999)
    std::vector<double> draws;                // accepted x, in order
    bool saw_one = false;
};

bool new_cws_parallel(uint32_t slot_begin, uint32_t slot_end, int32_t num_bins, double *r, double *c, double *b,
                      unsigned n_threads, uint64_t chunk_raw) {
    const uint64_t D = (uint64_t)num_bins;
    const uint64_t E = (uint64_t)slot_end * D;                    // elements drawn (rows below slot_begin are discarded)
    const uint64_t skip_elems = (uint64_t)slot_begin * D;
    const uint64_t need = 2 * E;                                  // gamma draws: r, c, r, c, ...
    if (E == 0) return true;
    const unsigned T = std::max(1u, n_threads);
    static const double kMagic = 1.0 + std::log(4.5);
    const double alpha = 2.0, ainv = std::sqrt(2.0 * alpha - 1.0), bbb = alpha - std::log(4.0), ccc = alpha + ainv;

    // per-thread generator positioned at its first chunk; the stride between a thread's chunks is fixed
    std::vector<GoSource> src(T, GoSource(1));
    std::vector<uint64_t> stride_poly(kRngLen);
    {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < T; t++)
            th.emplace_back([&, t] {
                if (t == 0) GoSource::xpow((uint64_t)T * chunk_raw, stride_poly.data());
                else src[t].skip((uint64_t)t * chunk_raw);
            });
        for (auto &x : th) x.join();
    }
    std::vector<GammaChunk> ch(T);
    std::vector<std::vector<uint64_t>> raw(T);
    uint64_t produced = 0;                                        // gamma draws placed so far
    uint64_t next_start = 0;                                      // first attempt start of the next chunk
    std::atomic<bool> failed(false);
    for (uint64_t round = 0; produced < need && !failed; round++) {
        // pass 1: raw outputs of every chunk (+2 look-ahead) and their extreme positions
        {
            std::vector<std::thread> th;
            for (unsigned t = 0; t < T; t++)
                th.emplace_back([&, t] {
                    GammaChunk &k = ch[t];
                    k.begin = (round * T + t) * chunk_raw;
                    k.end = k.begin + chunk_raw;
                    k.extremes.clear();
                    k.saw_one = false;
                    GoSource g = src[t];                          // copy: src[t] stays at the chunk start for the jump
                    std::vector<uint64_t> &v = raw[t];
                    v.resize(chunk_raw + 2);
                    for (uint64_t i = 0; i < chunk_raw + 2; i++) {
                        const uint64_t x = g.next();
                        v[i] = x;
                        bool one;
                        const double u = GoSource::to_float64(x, &one);
                        if (one) k.saw_one = true;
                        if (i < chunk_raw && !(1e-7 < u && u < .9999999)) k.extremes.push_back(k.begin + i);
                    }
                    src[t].jump(stride_poly.data(), (uint64_t)T * chunk_raw);
                });
            for (auto &x : th) x.join();
        }
        // resolve the attempt phase at every chunk boundary (sequential, a handful of events)
        for (unsigned t = 0; t < T; t++) {
            GammaChunk &k = ch[t];
            if (k.saw_one) { failed = true; break; }
            uint64_t pos = next_start;                            // an attempt starts here (begin or begin + 1)
            k.start = pos;
            for (uint64_t e : k.extremes) {
                if (e < pos) continue;
                if (((e - pos) & 1) == 0) pos = e + 1;            // it is a u1: the attempt consumed one uniform
            }
            next_start = ((k.end > pos) && ((k.end - pos) & 1)) ? k.end + 1 : std::max(k.end, pos);
        }
        if (failed) break;
        // pass 2: the sampler over every chunk
        {
            std::vector<std::thread> th;
            for (unsigned t = 0; t < T; t++)
                th.emplace_back([&, t] {
                    GammaChunk &k = ch[t];
                    const std::vector<uint64_t> &v = raw[t];
                    k.draws.clear();
                    k.draws.reserve(chunk_raw / 2 + 16);
                    bool one;
                    uint64_t p = k.start;
                    while (p < k.end) {
                        const double u1 = GoSource::to_float64(v[p - k.begin], &one);
                        if (!(1e-7 < u1 && u1 < .9999999)) { p += 1; continue; }
                        const double u2 = 1.0 - GoSource::to_float64(v[p + 1 - k.begin], &one);
                        p += 2;
                        const double vv = std::log(u1 / (1.0 - u1)) / ainv;
                        const double x = alpha * std::exp(vv);
                        const double z = u1 * u1 * u2;
                        const double rr = bbb + ccc * vv - x;
                        if (rr + kMagic - 4.5 * z >= 0.0 || rr >= std::log(z)) k.draws.push_back(x * 1.0);
                    }
                });
            for (auto &x : th) x.join();
        }
        // scatter: draw m -> element m / 2, r when m is even, c = ln when odd
        {
            std::vector<uint64_t> first(T);
            for (unsigned t = 0; t < T; t++) {
                first[t] = produced;
                produced += ch[t].draws.size();
            }
            std::vector<std::thread> th;
            for (unsigned t = 0; t < T; t++)
                th.emplace_back([&, t] {
                    const GammaChunk &k = ch[t];
                    uint64_t m = first[t];
                    for (size_t i = 0; i < k.draws.size() && m < need; i++, m++) {
                        const uint64_t elem = m >> 1;
                        if (elem < skip_elems) continue;
                        const uint64_t at = elem - skip_elems;
                        if (m & 1) c[at] = std::log(k.draws[i]);
                        else r[at] = k.draws[i];
                    }
                });
            for (auto &x : th) x.join();
        }
    }
    if (failed) return false;
    // b = U * r, element e uses uniform #e of its own generator (histosketch.go:104,116)
    {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < T; t++)
            th.emplace_back([&, t] {
                const uint64_t e0 = E * t / T, e1 = E * (t + 1) / T;
                GoSource g(1);
                g.skip(e0);
                for (uint64_t e = e0; e < e1; e++) {
                    bool one;
                    const double u = GoSource::to_float64(g.next(), &one);
                    if (one) { failed = true; return; }
                    if (e >= skip_elems) b[e - skip_elems] = (0.0 + u * (1.0 - 0.0)) * r[e - skip_elems];
                }
            });
        for (auto &x : th) x.join();
    }
    return !failed;
}

}  // namespace

// ---- internal (not in include/hulk_b200.h): the pieces of Go's generator the device-side draw needs --------------
// (hidden visibility: linked from api.cu inside the same library)
extern "C" void hulk_b200_internal_alfg_window(int64_t seed, uint64_t *w) {
    // W[i] = S[i - 607]: the 607 values in front of the first output, so that S[n] = S[n - 607] + S[n - 273].
    // The first outputs are vec[333 - n] + vec[606 - n] (tap starts at 0, feed at 334, both step down).
    GoSource g(seed);
    for (int i = 0; i < kRngLen - kRngTap; i++) w[i] = g.vec[kRngLen - kRngTap - 1 - i];
    for (int i = 0; i < kRngTap; i++) w[kRngLen - kRngTap + i] = g.vec[kRngLen - 1 - i];
}
extern "C" void hulk_b200_internal_alfg_xpow(uint64_t n, uint64_t *poly) { GoSource::xpow(n, poly); }
extern "C" void hulk_b200_internal_alfg_polymul(const uint64_t *a, const uint64_t *b, uint64_t *out) {   // out may alias a or b
    GoSource::polymul(a, b, out);
}
// Cheng's acceptance test for one attempt with the host's libm (the arbiter of near-ties found on the device)
extern "C" int hulk_b200_internal_gamma_accepts(uint64_t raw1, uint64_t raw2) {
    static const double kMagic = 1.0 + std::log(4.5);
    const double alpha = 2.0, ainv = std::sqrt(2.0 * alpha - 1.0), bbb = alpha - std::log(4.0), ccc = alpha + ainv;
    bool one;
    const double u1 = GoSource::to_float64(raw1, &one), u2 = 1.0 - GoSource::to_float64(raw2, &one);
    const double vv = std::log(u1 / (1.0 - u1)) / ainv;
    const double x = alpha * std::exp(vv);
    const double z = u1 * u1 * u2;
    const double rr = bbb + ccc * vv - x;
    return (rr + kMagic - 4.5 * z >= 0.0 || rr >= std::log(z)) ? 1 : 0;
}

extern "C" {

// test hook: the parallel draw with an explicit thread count and raw-chunk length; returns 1 when it had to give up
int hulk_b200_new_cws_parallel(uint32_t s, int32_t num_bins, uint32_t slot_begin, uint32_t slot_end, double *r,
                               double *c, double *b, uint32_t n_threads, uint64_t chunk_raw) {
    if (!r || !c || !b || num_bins < 0 || slot_begin > slot_end || chunk_raw < 16) return HULK_B200_EARG;
    if (slot_end > s) slot_end = s;
    return new_cws_parallel(slot_begin, slot_end, num_bins, r, c, b, n_threads, chunk_raw) ? 0 : 1;
}

int hulk_b200_new_cws(uint32_t s, int32_t num_bins, uint32_t slot_begin, uint32_t slot_end, double *r, double *c,
                      double *b) {
    if (!r || !c || !b || num_bins < 0 || slot_begin > slot_end) return HULK_B200_EARG;
    if (slot_end > s) slot_end = s;
    // large tables: all host cores (bit-identical to the loop below, see new_cws_parallel)
    const uint64_t elems = (uint64_t)slot_end * (uint64_t)num_bins;
    const unsigned hw = std::thread::hardware_concurrency();
    if (elems >= (1u << 20) && hw > 1) {
        const unsigned T = std::min(hw, 32u);
        if (new_cws_parallel(slot_begin, slot_end, num_bins, r, c, b, T, 1u << 21)) return HULK_B200_OK;
    }
    GoSource gamma_src(1), unif_src(1);                       // DISTRIBUTION_SEED  histosketch.go:20,103-104
    for (uint32_t i = 0; i < slot_end; i++) {
        for (int32_t j = 0; j < num_bins; j++) {
            const double rv = go_rng_gamma(gamma_src, 2, 1);                  // :112
            const double cv = std::log(go_rng_gamma(gamma_src, 2, 1));        // :113
            const double bv = (0.0 + unif_src.float64() * (1.0 - 0.0)) * rv;  // :116
            if (i >= slot_begin) {
                const size_t at = (size_t)(i - slot_begin) * (size_t)num_bins + (size_t)j;
                r[at] = rv; c[at] = cv; b[at] = bv;
            }
        }
    }
    return HULK_B200_OK;
}

void hulk_b200_md5_mins(const uint64_t *mins, uint32_t n, char out_hex[33]) {
    Md5 md;
    for (uint32_t i = 0; i < n; i++) {
        uint8_t le[8];
        for (int j = 0; j < 8; j++) le[j] = (uint8_t)(mins[i] >> (8 * j));     // binary.LittleEndian.PutUint64
        md.update(le, 8);
    }
    uint8_t dig[16];
    md.final(dig);
    for (int i = 0; i < 16; i++) snprintf(out_hex + 2 * i, 3, "%02x", dig[i]);
    out_hex[32] = 0;
}

// one MinHash signature (minhash.KMVsketch / minhash.KHFsketch: ksize, md5sum, mins, num -- kmv.go:12-21, khf.go:11-17)
static void minhash_signature(std::string &o, const char *algo, uint32_t k, const uint64_t *mins, uint32_t n) {
    const std::string I = "    ", I4 = I + I + I + I, I5 = I4 + I;
    char md5[33];
    hulk_b200_md5_mins(mins, n, md5);
    o += I + I + "{\n";
    o += I + I + I + "\"Algorithm\": \"" + algo + "\",\n";
    o += I + I + I + "\"Sketch\": {\n";
    o += I4 + "\"ksize\": " + std::to_string(k) + ",\n";
    o += I4 + "\"md5sum\": \"" + md5 + "\",\n";
    o += I4 + "\"mins\": [\n";
    for (uint32_t i = 0; i < n; i++) o += I5 + std::to_string(mins[i]) + (i + 1 < n ? ",\n" : "\n");
    o += I4 + "],\n";
    o += I4 + "\"num\": " + std::to_string(n) + "\n";
    o += I + I + I + "}\n";
    o += I + I + "}";
}

int64_t hulk_b200_sketch_json(char *buf, uint64_t cap, const char *filename, const char *banner_label, uint32_t k,
                              const uint64_t *mins, const double *weights, uint32_t s, int32_t num_bins,
                              int concept_drift) {
    return hulk_b200_sketch_json_minhash(buf, cap, filename, banner_label, k, mins, weights, s, num_bins,
                                         concept_drift, nullptr, 0, nullptr, 0);
}

int64_t hulk_b200_sketch_json_minhash(char *buf, uint64_t cap, const char *filename, const char *banner_label,
                                      uint32_t k, const uint64_t *mins, const double *weights, uint32_t s,
                                      int32_t num_bins, int concept_drift, const uint64_t *kmv_mins, uint32_t kmv_n,
                                      const uint64_t *khf_mins, uint32_t khf_n) {
    // HULKdata.Add refuses a sketch without mins (sketchio.go:59-61); signatures are added in the order
    // histosketch, kmv, khf (src/pipeline/sketch.go:227-234,289-294)
    if ((kmv_mins && kmv_n == 0) || (khf_mins && khf_n == 0)) return HULK_B200_ENOSKETCH;
    for (uint32_t i = 0; i < s; i++)
        if (!std::isfinite(weights[i])) return HULK_B200_EARG;   // json: unsupported value
    char md5[33];
    hulk_b200_md5_mins(mins, s, md5);
    const std::string I = "    ";
    std::string o;
    o.reserve(256 + (size_t)s * 64);
    o += "{\n";
    o += I + "\"class\": \"hulk_sketch\",\n";
    o += I + "\"filename\": " + go_string(filename) + ",\n";
    o += I + "\"hash_function\": \"ntHash\",\n";                 // sketchio.go:48 (stale constant, kept)
    o += I + "\"license\": \"CC0\",\n";
    o += I + "\"signatures\": [\n";
    o += I + I + "{\n";
    o += I + I + I + "\"Algorithm\": \"histosketch\",\n";
    o += I + I + I + "\"Sketch\": {\n";
    const std::string I4 = I + I + I + I, I5 = I4 + I;
    o += I4 + "\"ksize\": " + std::to_string(k) + ",\n";
    o += I4 + "\"md5sum\": \"" + md5 + "\",\n";
    if (s) {
        o += I4 + "\"mins\": [\n";
        for (uint32_t i = 0; i < s; i++) o += I5 + std::to_string(mins[i]) + (i + 1 < s ? ",\n" : "\n");
        o += I4 + "],\n";
        o += I4 + "\"weights\": [\n";
        for (uint32_t i = 0; i < s; i++) o += I5 + go_float(weights[i]) + (i + 1 < s ? ",\n" : "\n");
        o += I4 + "],\n";
    } else {
        o += I4 + "\"mins\": [],\n";
        o += I4 + "\"weights\": [],\n";
    }
    o += I4 + "\"num\": " + std::to_string(s) + ",\n";
    o += I4 + "\"num_histogram_bins\": " + std::to_string(num_bins) + ",\n";
    o += I4 + "\"concept_drift\": " + (concept_drift ? "true" : "false") + "\n";
    o += I + I + I + "}\n";
    o += I + I + "}";
    if (kmv_mins) { o += ",\n"; minhash_signature(o, "kmv", k, kmv_mins, kmv_n); }
    if (khf_mins) { o += ",\n"; minhash_signature(o, "khf", k, khf_mins, khf_n); }
    o += "\n";
    o += I + "],\n";
    o += I + "\"version\": \"" HULK_B200_VERSION "\",\n";
    o += I + "\"banner_label\": " + go_string(banner_label) + "\n";
    o += "}";
    if (buf && cap) {
        const size_t n = (o.size() < cap - 1) ? o.size() : cap - 1;
        memcpy(buf, o.data(), n);
        buf[n] = 0;
    }
    return (int64_t)o.size();
}

int hulk_b200_write_json(const char *path, const char *filename, const char *banner_label, uint32_t k,
                         const uint64_t *mins, const double *weights, uint32_t s, int32_t num_bins,
                         int concept_drift) {
    return hulk_b200_write_json_minhash(path, filename, banner_label, k, mins, weights, s, num_bins, concept_drift,
                                        nullptr, 0, nullptr, 0);
}

int hulk_b200_write_json_minhash(const char *path, const char *filename, const char *banner_label, uint32_t k,
                                 const uint64_t *mins, const double *weights, uint32_t s, int32_t num_bins,
                                 int concept_drift, const uint64_t *kmv_mins, uint32_t kmv_n,
                                 const uint64_t *khf_mins, uint32_t khf_n) {
    if (s == 0) return HULK_B200_ENOSKETCH;                      // sketchio.go:59-61
    const int64_t need = hulk_b200_sketch_json_minhash(nullptr, 0, filename, banner_label, k, mins, weights, s,
                                                       num_bins, concept_drift, kmv_mins, kmv_n, khf_mins, khf_n);
    if (need < 0) return (int)need;
    std::vector<char> buf((size_t)need + 1);
    hulk_b200_sketch_json_minhash(buf.data(), buf.size(), filename, banner_label, k, mins, weights, s, num_bins,
                                  concept_drift, kmv_mins, kmv_n, khf_mins, khf_n);
    FILE *fh = fopen(path, "wb");                                // ioutil.WriteFile(..., 0644)
    if (!fh) return HULK_B200_EIO;
    const size_t wr = fwrite(buf.data(), 1, (size_t)need, fh);
    const int rc = fclose(fh);
    return (wr == (size_t)need && rc == 0) ? HULK_B200_OK : HULK_B200_EIO;
}

}  // extern "C"
