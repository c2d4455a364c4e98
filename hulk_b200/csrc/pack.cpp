// pack.cpp -- host side of the packed read transport (include/hulk_b200.h, "packed reads").
//
// The minimizer scan only ever sees a base through seq_nt4_table (src/minimizer/minimizer.go:13-30, :115):
// A/a C/c G/g T/t U/u and the raw bytes 0..3 give the codes 0..3, every other byte gives 4.  So a batch of
// reads travels to the GPU losslessly FOR THIS PATH as 2 bits per base plus the (sorted, usually empty) list
// of positions whose code is 4 -- a quarter of the bytes of the ASCII form over PCIe, which is what bounds the
// host-fed pipeline.  The device side (k0_unpack / k0_patch in api.cu) turns the
// stream back into letters the scan kernels read: ACGT for the codes, N for the listed positions.
//
// Layout: base i of the batch (reads back to back, exactly the ASCII layout) sits in bits 2 (i & 3) .. +1 of
// packed[i >> 2].  No per-read alignment: packing is a pure streaming transform of the ASCII buffer.
//
// Speed: 64 bases per step with AVX-512BW (32 with AVX2, a table otherwise), and the batch is cut into
// 64 KiB pieces handed to a small pool of spinning worker threads -- a 15 MB interval has to be packed in
// about 0.1 ms to keep up with the GPU.
#include <immintrin.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/hulk_b200.h"

namespace {

// ---- seq_nt4_table (minimizer.go:13-30) ----------------------------------------------------------
struct Nt4Table {
    uint8_t t[256];
    Nt4Table() {
        memset(t, 4, sizeof t);
        t[0] = 0; t[1] = 1; t[2] = 2; t[3] = 3;
        t[(int)'A'] = t[(int)'a'] = 0;
        t[(int)'C'] = t[(int)'c'] = 1;
        t[(int)'G'] = t[(int)'g'] = 2;
        t[(int)'T'] = t[(int)'t'] = 3;
        t[(int)'U'] = t[(int)'u'] = 3;
    }
};
const Nt4Table kNt4;

// One piece of the batch: bases [b0, b1) -> packed bytes [b0/4, ceil(b1/4)); b0 is a multiple of 64.
// Exceptions (code 4) are appended to `exc` as positions relative to the batch.
// The input streams from DRAM once; the hardware prefetchers stop at every 4 KiB page, so the loops ask for the line
// this far ahead themselves (a prefetch never faults: running past the end of the batch is harmless).
static const size_t kPrefetchAhead = []{ const char *e = getenv("HULK_B200_PACK_PREFETCH"); return e ? (size_t)atol(e) : (size_t)2048; }();

typedef void (*PieceFn)(const uint8_t *bases, uint64_t b0, uint64_t b1, uint8_t *packed, std::vector<uint32_t> &exc);

void piece_scalar(const uint8_t *bases, uint64_t b0, uint64_t b1, uint8_t *packed, std::vector<uint32_t> &exc) {
    uint64_t i = b0;
    for (; i + 4 <= b1; i += 4) {
        const uint32_t c0 = kNt4.t[bases[i]], c1 = kNt4.t[bases[i + 1]], c2 = kNt4.t[bases[i + 2]], c3 = kNt4.t[bases[i + 3]];
        if ((c0 | c1 | c2 | c3) & 4u) {
            if (c0 & 4u) exc.push_back((uint32_t)i);
            if (c1 & 4u) exc.push_back((uint32_t)(i + 1));
            if (c2 & 4u) exc.push_back((uint32_t)(i + 2));
            if (c3 & 4u) exc.push_back((uint32_t)(i + 3));
        }
        packed[i >> 2] = (uint8_t)((c0 & 3u) | ((c1 & 3u) << 2) | ((c2 & 3u) << 4) | ((c3 & 3u) << 6));
    }
    if (i < b1) {
        uint32_t v = 0;
        for (uint64_t j = i; j < b1; j++) {
            const uint32_t c = kNt4.t[bases[j]];
            if (c & 4u) exc.push_back((uint32_t)j);
            v |= (c & 3u) << (2 * (j - i));
        }
        packed[i >> 2] = (uint8_t)v;
    }
}

__attribute__((target("avx2"))) void piece_avx2(const uint8_t *bases, uint64_t b0, uint64_t b1, uint8_t *packed,
                                               std::vector<uint32_t> &exc) {
    const __m256i m03 = _mm256_set1_epi8(0x03), mDF = _mm256_set1_epi8((char)0xDF), mFC = _mm256_set1_epi8((char)0xFC);
    const __m256i lut = _mm256_setr_epi8('A', 'C', 'G', 'T', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                         'A', 'C', 'G', 'T', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i chU = _mm256_set1_epi8('U'), zero = _mm256_setzero_si256();
    const __m256i w14 = _mm256_set1_epi16(0x0401), w116 = _mm256_set1_epi32(0x00100001);
    const __m256i gather = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                            0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    uint64_t i = b0;
    for (; i + 32 <= b1; i += 32) {
        if ((i & 32) == 0) _mm_prefetch(reinterpret_cast<const char *>(bases + i + kPrefetchAhead), _MM_HINT_T0);
        const __m256i w = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(bases + i));
        const __m256i v = _mm256_and_si256(_mm256_xor_si256(_mm256_srli_epi16(w, 1), _mm256_srli_epi16(w, 2)), m03);
        const __m256i up = _mm256_and_si256(w, mDF);
        const __m256i lt4 = _mm256_cmpeq_epi8(_mm256_and_si256(w, mFC), zero);              // raw bytes 0..3: themselves
        const __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(up, _mm256_shuffle_epi8(lut, v)),
                                                           _mm256_cmpeq_epi8(up, chU)), lt4);
        const __m256i code = _mm256_blendv_epi8(v, w, lt4);
        uint32_t bad = ~(uint32_t)_mm256_movemask_epi8(ok);
        while (bad) {
            exc.push_back((uint32_t)(i + (uint64_t)__builtin_ctz(bad)));
            bad &= bad - 1;
        }
        const __m256i p16 = _mm256_maddubs_epi16(code, w14);           // c0 + 4 c1 per 16-bit lane
        const __m256i p32 = _mm256_madd_epi16(p16, w116);              // + 16 (c2 + 4 c3) per 32-bit lane: one byte
        const __m256i g = _mm256_shuffle_epi8(p32, gather);
        uint32_t lo = (uint32_t)_mm256_cvtsi256_si32(g);
        uint32_t hi = (uint32_t)_mm256_extract_epi32(g, 4);
        memcpy(packed + (i >> 2), &lo, 4);
        memcpy(packed + (i >> 2) + 4, &hi, 4);
    }
    if (i < b1) piece_scalar(bases, i, b1, packed, exc);
}

__attribute__((target("avx512f,avx512bw,avx512vl"))) void piece_avx512(const uint8_t *bases, uint64_t b0, uint64_t b1,
                                                                      uint8_t *packed, std::vector<uint32_t> &exc) {
    const __m512i m03 = _mm512_set1_epi8(0x03), mDF = _mm512_set1_epi8((char)0xDF);
    const __m512i lut = _mm512_broadcast_i32x4(_mm_setr_epi8('A', 'C', 'G', 'T', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0));
    const __m512i chU = _mm512_set1_epi8('U'), four = _mm512_set1_epi8(4);
    const __m512i w14 = _mm512_set1_epi16(0x0401), w116 = _mm512_set1_epi32(0x00100001);
    uint64_t i = b0;
    for (; i + 64 <= b1; i += 64) {
        _mm_prefetch(reinterpret_cast<const char *>(bases + i + kPrefetchAhead), _MM_HINT_T0);
        const __m512i w = _mm512_loadu_si512(bases + i);
        const __m512i v = _mm512_and_si512(_mm512_xor_si512(_mm512_srli_epi16(w, 1), _mm512_srli_epi16(w, 2)), m03);
        const __m512i up = _mm512_and_si512(w, mDF);
        const __mmask64 lt4 = _mm512_cmplt_epu8_mask(w, four);                               // raw bytes 0..3: themselves
        const __mmask64 ok = _mm512_cmpeq_epi8_mask(up, _mm512_shuffle_epi8(lut, v)) | _mm512_cmpeq_epi8_mask(up, chU) | lt4;
        const __m512i code = _mm512_mask_blend_epi8(lt4, v, w);
        uint64_t bad = ~(uint64_t)ok;
        while (bad) {
            exc.push_back((uint32_t)(i + (uint64_t)__builtin_ctzll(bad)));
            bad &= bad - 1;
        }
        const __m512i p16 = _mm512_maddubs_epi16(code, w14);
        const __m512i p32 = _mm512_madd_epi16(p16, w116);
        _mm_storeu_si128(reinterpret_cast<__m128i *>(packed + (i >> 2)), _mm512_cvtepi32_epi8(p32));
    }
    if (i < b1) piece_scalar(bases, i, b1, packed, exc);
}

PieceFn pick_piece_fn() {
    const char *e = getenv("HULK_B200_PACK_ISA");          // tests: "scalar", "avx2", "avx512"
    __builtin_cpu_init();
    const bool has512 = __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl");
    const bool has2 = __builtin_cpu_supports("avx2");
    if (e && !strcmp(e, "scalar")) return piece_scalar;
    if (e && !strcmp(e, "avx2") && has2) return piece_avx2;
    if (has512) return piece_avx512;
    if (has2) return piece_avx2;
    return piece_scalar;
}

// ---- worker pool -----------------------------------------------------------------------------------
// Workers spin on a generation counter for a while (a pipelined caller comes back every ~0.1 ms), then sleep on a
// condition variable.  One job at a time (guarded by job_mu); the caller works on the job too.
constexpr uint64_t kPiece = 64 * 1024;     // bases per piece (a multiple of 64)

struct Job {
    const uint8_t *bases = nullptr;
    uint64_t n_bases = 0;
    uint8_t *packed = nullptr;
    PieceFn fn = nullptr;
    std::atomic<uint64_t> next{0};          // next piece to take
    std::atomic<uint64_t> done{0};          // pieces finished
    uint64_t n_pieces = 0;
    std::vector<std::vector<uint32_t>> exc; // per piece
};

struct Pool {
    std::mutex job_mu;                      // one pack call at a time
    std::mutex mu;                          // guards `job` and the sleepers
    std::condition_variable cv;
    std::vector<std::thread> workers;
    std::atomic<uint64_t> gen{0};           // bumped when a job is posted
    std::atomic<int> want{0};               // workers that should take part in the current job
    std::atomic<bool> stop{false};
    // The job is shared: a worker that wakes up late (or is descheduled between waking and looking) still holds a valid
    // object, finds no piece left and goes back to waiting -- the caller never waits for anyone but the threads that
    // actually took a piece.
    std::shared_ptr<Job> job;

    static void run_pieces(Job *j) {
        for (;;) {
            const uint64_t p = j->next.fetch_add(1, std::memory_order_relaxed);
            if (p >= j->n_pieces) break;
            const uint64_t b0 = p * kPiece, b1 = std::min(j->n_bases, b0 + kPiece);
            j->fn(j->bases, b0, b1, j->packed, j->exc[p]);
            j->done.fetch_add(1, std::memory_order_release);
        }
    }
    void worker(int id) {
        uint64_t seen = 0;
        for (;;) {
            // wait for a new generation: spin first (a pipelined caller is back within ~0.1 ms), then sleep
            int spins = 0;
            while (gen.load(std::memory_order_acquire) == seen && !stop.load(std::memory_order_relaxed)) {
                if (++spins < 20000) {
                    _mm_pause();
                } else {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return gen.load(std::memory_order_acquire) != seen || stop.load(); });
                }
            }
            if (stop.load(std::memory_order_relaxed)) return;
            std::shared_ptr<Job> j;
            {
                std::lock_guard<std::mutex> lk(mu);
                seen = gen.load(std::memory_order_acquire);
                j = job;
            }
            if (j && id < want.load(std::memory_order_relaxed)) run_pieces(j.get());
        }
    }
    void ensure(int n) {                    // at least n workers (called under job_mu)
        while ((int)workers.size() < n) {
            const int id = (int)workers.size();
            workers.emplace_back([this, id] { worker(id); });
        }
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop.store(true);
        }
        cv.notify_all();
        for (auto &t : workers) t.join();
    }
};

Pool &pool() {
    static Pool *p = new Pool();            // leaked on purpose: no static-destruction order games with the threads
    return *p;
}

int default_threads() {
    const char *e = getenv("HULK_B200_PACK_THREADS");
    if (e && atoi(e) > 0) return std::min(atoi(e), 256);
    unsigned hc = 0;
#if defined(__linux__)
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0) hc = (unsigned)CPU_COUNT(&set);
#endif
    if (!hc) hc = std::thread::hardware_concurrency();
    if (!hc) hc = 4;
    // two CPUs stay free for the thread that issues the GPU work and whatever else the process runs: a packer that is
    // descheduled in the middle of a piece holds the whole batch up for a scheduler quantum
    if (hc > 3) hc -= 2;
    return (int)std::min(hc, 32u);
}

}  // namespace

extern "C" {

uint64_t hulk_b200_packed_bytes(uint64_t n_bases) { return (n_bases + 3) / 4; }

int hulk_b200_pack_bases(const uint8_t *bases, uint64_t n_bases, uint8_t *packed, uint32_t *exceptions,
                         uint64_t exceptions_cap, uint64_t *n_exceptions, int32_t n_threads) {
    if ((!bases && n_bases) || (!packed && n_bases) || !n_exceptions || (exceptions_cap && !exceptions)) return HULK_B200_EARG;
    if (n_bases >= (1ull << 32)) return HULK_B200_EARG;      // positions are 32 bits wide: pack batch by batch
    *n_exceptions = 0;
    if (n_bases == 0) return HULK_B200_OK;
    const PieceFn fn = pick_piece_fn();
    Pool &P = pool();
    std::lock_guard<std::mutex> guard(P.job_mu);
    std::shared_ptr<Job> jp = std::make_shared<Job>();
    Job &j = *jp;
    j.bases = bases;
    j.n_bases = n_bases;
    j.packed = packed;
    j.fn = fn;
    j.n_pieces = (n_bases + kPiece - 1) / kPiece;
    j.exc.resize(j.n_pieces);
    int threads = n_threads > 0 ? n_threads : default_threads();
    threads = (int)std::min<uint64_t>((uint64_t)threads, j.n_pieces);
    if (threads > 1) {
        P.ensure(threads - 1);
        P.want.store(threads - 1, std::memory_order_relaxed);
        {
            std::lock_guard<std::mutex> lk(P.mu);
            P.job = jp;
            P.gen.fetch_add(1, std::memory_order_release);
        }
        P.cv.notify_all();
    }
    Pool::run_pieces(&j);
    while (j.done.load(std::memory_order_acquire) < j.n_pieces) _mm_pause();
    if (threads > 1) {
        std::lock_guard<std::mutex> lk(P.mu);
        P.job.reset();
    }
    uint64_t n = 0;
    for (auto &v : j.exc) {
        for (uint32_t pos : v) {
            if (n < exceptions_cap) exceptions[n] = pos;
            n++;
        }
    }
    *n_exceptions = n;
    return HULK_B200_OK;
}

}  // extern "C"
