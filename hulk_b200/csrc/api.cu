// api.cu -- context management and the C ABI (include/hulk_b200.h) over the sm_100a kernels.
//
// HBM layout of one context (D bins, Dp = D padded to 512, rows = owned sketch slots):
//   reads      4 x staging buffer (raw ASCII, + offsets), a ring between the H2D copy stream and the k1 streams
//   hist       4 x uint32[D]          k^4-bin spectrum, one buffer per interval in flight (L2-resident atomics)
//   queue      4 x uint64[reads*cap]  per-batch minimizer queue: k1 scan -> k1_jump_queue
//   cms        double[7*2000] + static CSR of bins per counter (int32[7*D]) + uint16 cols[7*D]
//   f          2 x uint64[D] (float64 bits; +inf = bin unused in this flush), 2 x (1/f) as bfloat16[Dp] (or float)
//   r, c, b    double[rows*D] each    reference CWS tables (exact float64 re-evaluation)
//   K16        bfloat16[rows*Dp]      folded CWS coefficient c*exp(b-r) of the screen, streamed once per flush
//                                     (HULK_B200_K3_FP32=1: K32 float[rows*Dp])
//   m32        float[rows*Dp/512]     per-chunk screen minima of a flush; cand uint32[rows], thr32 float[rows]
//   sketch     uint64[rows], weights double[rows]
//
// Streams: `stream` (CWS sweep, caller-visible ordering point), a count-min stream, one k1 stream per
// spectrum buffer, a copy stream.  Events order "interval counted" -> count-min -> CWS sweep, and
// "buffer wiped" -> next use, so interval i+1.. is counted while interval i is flushed.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/hulk_b200.h"
#include "k1_minimizer.cuh"
#include "k1_long.cuh"
#include "k1_minhash.cuh"
#include "k2_countmin.cuh"
#include "k3_cws.cuh"
#include "k4_cwsdraw.cuh"

using namespace hulk;

// k values served by the second-generation w = 9 scan (k1_scan2.h: odd k with two-word k-mers; k = 9, 11 in one word)
#define K1_V2_FOR_EACH_K(X) X(9) X(11) X(17) X(19) X(21) X(23) X(25) X(27) X(29) X(31)

constexpr int NBUF = 4;     // spectrum buffers allocated; ctx->nbuf of them are cycled = intervals in flight at once
constexpr int NSTAGE = 6;   // host-input staging buffers (ring): deep enough that the feeder never waits for the copy that last used one

struct hulk_b200_ctx {
    hulk_b200_params P{};
    int32_t D = 0;
    uint32_t s = 0, rows = 0;
    uint64_t Dp = 0;
    uint32_t nsub_row = 0, nseg = 0, nblk = 0;
    bool drift = false, apply_scaling = false;
    bool overlap = true;            // k1 on its own streams (false: everything in order on the main stream)
    bool force_tile_path = false;   // HULK_B200_K1_TILE=1: use the staged-tile kernel even for w = 9 (A/B measurements)
    double decay_weight = 0.0;
    int sm_count = 148;

    cudaStream_t stream = nullptr, copy_stream = nullptr;
    bool own_stream = false;

    // stage 1+2
    // The spectrum is multi-buffered: the reads of interval i+1.. are counted into the next buffers (on
    // their own stream) while interval i is still being flushed on the main stream.
    uint32_t *d_hist[NBUF] = {};
    int cur_hist = 0;                          // buffer (and k1 stream) of the interval being counted
    int nbuf = 4;
    cudaStream_t k1_stream[NBUF] = {};
    cudaEvent_t ev_k1_last[NBUF] = {};         // last k1 launch into buffer b
    cudaEvent_t ev_hist_free[NBUF] = {};       // buffer b consumed and wiped by its flush
    cudaEvent_t ev_main = nullptr;
    bool k1_pending[NBUF] = {};                // k1 launches into buffer b since its last flush
    unsigned long long *d_nmin = nullptr, *d_errword = nullptr;
    uint8_t *d_stage[NSTAGE] = {};
    uint64_t stage_cap[NSTAGE] = {};
    uint64_t *d_off[NSTAGE] = {};
    uint64_t off_cap[NSTAGE] = {};
    cudaEvent_t ev_copy[NSTAGE] = {};
    // "the kernels that consumed stage buffer b are done": two events per buffer, used alternately, so that the event of
    // the PREVIOUS use can still be waited on (by the feeder thread, later) after the current use has been recorded
    cudaEvent_t ev_k1x[NSTAGE][2] = {};
    uint32_t k1_use[NSTAGE] = {};              // uses of stage buffer b so far
    cudaEvent_t ev_k1_prev(int b) const { return ev_k1x[b][(k1_use[b] + 1u) & 1u]; }   // last recorded (never recorded: a no-op to wait on)
    cudaError_t ev_k1_record(int b, cudaStream_t st) { const cudaError_t e = cudaEventRecord(ev_k1x[b][k1_use[b] & 1u], st); k1_use[b]++; return e; }
    int cur_buf = 0;
    // packed transport (HULK_B200_F_PACK_INPUT / push_reads_packed): 2 bits per base + positions of the code-4 bases
    int pack_threads = 0;                      // 0: off; > 0: host threads that pack a pushed ASCII batch; < 0: all CPUs
    uint8_t *h_pack[NSTAGE] = {};              // pinned: the packed form of the batch in flight through stage buffer b
    uint32_t *h_exc[NSTAGE] = {};
    uint64_t h_pack_cap[NSTAGE] = {}, h_exc_cap[NSTAGE] = {};
    uint8_t *d_pack[NSTAGE] = {};
    uint32_t *d_exc[NSTAGE] = {};
    uint64_t d_pack_cap[NSTAGE] = {}, d_exc_cap[NSTAGE] = {};
    // The feeder: with HULK_B200_F_ASYNC_INPUT a batch is packed and copied by this thread while the calling thread goes on
    // enqueueing the kernels that consume it (they wait, on the device, for the batch's sequence number to appear in
    // d_feed).  Packing a C2 interval takes about as long as counting it, so the two must not share a host thread.
    struct FeedReq {
        const uint8_t *src;         // ASCII bases of the batch
        uint64_t nb;
        const uint64_t *offsets;    // or nullptr
        uint64_t n_off;
        int buf;
        uint32_t seq;
        cudaEvent_t stage_free;     // the previous consumers of this stage buffer are done (waited for on the copy stream)
    };
    bool feeder_allowed = false;               // every stream has a hardware queue of its own (see hulk_b200_create)
    std::thread feeder;
    std::mutex feed_mu;
    std::condition_variable feed_cv;
    std::deque<FeedReq> feed_q;
    uint64_t feed_posted = 0, feed_done = 0;   // under feed_mu
    bool feed_stop = false;
    int feed_rc = 0;                           // first error of the feeder (under feed_mu)
    std::string feed_err;
    uint32_t feed_seq[NSTAGE] = {};            // uses of stage buffer b by the feeder path
    uint32_t *d_feed = nullptr;                // [NSTAGE][4]: sequence flag, mode (1 = packed), exceptions, pad
    uint32_t *h_feed = nullptr;                // pinned mirror of mode / exceptions, [NSTAGE][4]
    cudaEvent_t ev_tail0[NSTAGE] = {}, ev_tail[NSTAGE] = {};   // around the copy of the letters part of the batch in stage buffer b
    int feed_last_tail = -1;                   // stage buffer whose letters copy has not been timed yet
    uint64_t feed_last_tail_bytes = 0;
    double feed_bp = 0.0, feed_bl = 0.0;       // measured rates, bytes of bases per second: packing, link (0: not measured yet)
    double feed_frac = 0.75;                   // share of a batch that travels packed (feeder thread only)
    double feed_step = 1.0 / 16.0;             // how far the split moves per batch
    int feed_dir = 0;
    double feed_frac_sum = 0.0;
    bool feed_adapt = true;                    // HULK_B200_PACK_FRACTION=<0..1> pins the share
    std::atomic<uint64_t> feed_h2d{0}, feed_pack_ns{0}, feed_batches{0};
    std::atomic<uint64_t> feed_t_idle{0}, feed_t_evsync{0}, feed_t_enq{0}, feed_t_backpressure{0};   // ns (HULK_B200_FEED_DEBUG)
    CUresult (*cu_wait32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
    CUresult (*cu_write32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
    unsigned int *d_ovf_count[NBUF] = {};      // one overflow queue + scratch arena per k1 stream
    unsigned long long *d_ovf_list[NBUF] = {};
    uint32_t ovf_cap = 0;
    uint64_t *d_queue[NBUF] = {};              // per-batch minimizer queue (k1 scan -> k1_jump_queue)
    unsigned long long *d_queue_cursor[NBUF] = {};
    uint64_t queue_cap[NBUF] = {};
    int k1_ctas_per_sm = 4;                    // scan CTAs per SM in queue mode (tasks are handed out dynamically):
                                               // one short of what fits, so the flush chain always finds SM room
    uint64_t batch_max_len = 0;                // longest read of the batch being pushed (0: unknown)
    uint64_t batch_long_entries = 0;           // arena entries the sets of its long sequences take (k1_long.cuh)
    uint64_t max_launch_reads = 1ull << 22;    // reads per k1 launch (HULK_B200_MAX_LAUNCH_READS)
    int jump_ctas_per_sm = K1_JUMP_CTAS_PER_SM;
    int jump_batch = 4;                        // jump steps between two refill points of k1_jump_queue
    bool fused_jump = false;                   // HULK_B200_K1_FUSED=1: bin inside the scan kernel (A/B measurements)
    bool k1_persistent = false;                // HULK_B200_K1_PERSISTENT=1: scan CTAs loop over tasks handed out by a counter (A/B)
    bool jump_smem = false;                    // HULK_B200_JUMP_SMEM=1: keys handed out through a shared-memory counter (A/B measurements)
    bool jump_fx = true;                       // HULK_B200_JUMP_FX=0: keep the bracketed jump step for every D (A/B measurements)
    bool k1_v2 = true;                         // HULK_B200_K1_V2=0: keep the first-generation w = 9 scan (A/B measurements)
    uint64_t *d_arena[NBUF] = {};
    K1LongTask *d_long_tasks[NBUF] = {};       // long sequences of the launch in flight on each k1 stream (k1_long.cuh)
    K1LongCtl *d_long_ctl[NBUF] = {};
    unsigned long long *d_lenstat = nullptr;   // k1_length_stats' two words
    uint64_t long_cap[NBUF] = {};              // entries of d_long_tasks
    // MinHash side sketches fed from the minimizer queue (k1_minhash.cuh; hulk_b200_minhash_enable)
    bool mh_kmv = false, mh_khf = false;
    unsigned long long *d_khf = nullptr;       // [s]
    K1KmvState *d_kmv_state[NBUF] = {};        // one bottom-s pool per k1 stream, merged when read
    uint64_t *d_kmv_pool[NBUF] = {};           // [2][s]
    uint64_t *d_kmv_cand[NBUF] = {};           // [kmv_cand_cap]
    uint64_t kmv_cand_cap[NBUF] = {};
    bool long_path = true;                     // HULK_B200_LONG=0: long sequences stay with k1_generic (A/B)
    uint32_t list_cap_w9 = 192;                // longest candidate list of the w = 9 kernels (HULK_B200_LIST_CAP, 16..192)
    uint64_t long_min = K1_LONG_MIN;           // sequences this long take the sliced scan (HULK_B200_LONG_MIN, >= 1024)
    unsigned long long *d_arena_cursor[NBUF] = {};
    uint64_t arena_entries[NBUF] = {};

    // stage 3a
    FlushCtl *d_ctl = nullptr;
    uint16_t *d_cols = nullptr;
    unsigned int *d_csr_start = nullptr;
    int32_t *d_csr_bins = nullptr;
    uint32_t *d_words = nullptr, *d_word_prefix = nullptr, *d_block_count = nullptr, *d_block_prefix = nullptr;
    double *d_q = nullptr;
    // estimate vectors, double-buffered: the count-min update of flush i+1 (k2 stream) overlaps the CWS
    // sweep of flush i (main stream)
    unsigned long long *d_fbits[2] = {nullptr, nullptr};
    float *d_invf[2] = {nullptr, nullptr};
    int flush_idx = 0, last_flush_idx = 0;     // parity of the next / the latest flush
    cudaStream_t k2_stream = nullptr;
    cudaEvent_t ev_k2_done[2] = {nullptr, nullptr}, ev_k3_done[2] = {nullptr, nullptr};
    bool k3_pending[2] = {false, false};

    // stage 3b
    double *d_r = nullptr, *d_c = nullptr, *d_b = nullptr;
    float *d_K32 = nullptr, *d_m32 = nullptr;
    // bf16 screen (default; HULK_B200_K3_FP32=1 keeps the fp32 one): K16 replaces K32, (1/f)16 replaces (1/f)32
    bool filter16 = true;
    __nv_bfloat16 *d_K16 = nullptr;
    __nv_bfloat16 *d_invf16[2] = {nullptr, nullptr};
    double k3_eps = K3_EPS16;
    unsigned int *d_cand = nullptr;            // per slot: some chunk of the current flush may change it
    float *d_thr32 = nullptr;                  // per slot: the fp32 screen bound (k3_thr32), kept by k3_resolve
    int k3_stages = 4, k3_ctas_per_sm = 2;     // 2 x (4 x 16 KB) per SM: 5.7 TB/s alone, and k1 CTAs still fit next to it
    unsigned long long *d_sketch = nullptr;
    double *d_weights = nullptr;
    bool tables_set = false;
    uint64_t cws_ties = 0;                     // attempts of the device-side table draw that the host's libm decided
    // generate_cws_tables_async: the host draw runs on its own thread; the first flush (or set/finish) joins it
    std::thread gen_thread;
    std::vector<double> gen_r, gen_c, gen_b;
    int gen_rc = 0;
    bool gen_pending = false;

    // multi-GPU (k2_countmin.cuh, "the spectrum of a flush is the sum of every GPU's counting buffer")
    uint32_t world = 1, rank = 0;
    uint8_t *arena = nullptr;                  // the NBUF spectrum buffers + the two flag arrays: one allocation, one IPC handle
    size_t arena_bytes = 0, hist_stride = 0, off_counted = 0, off_gathered = 0;
    uint8_t *peer_arena[PEER_MAX] = {};        // every rank's arena as this device reaches it (own at [rank])
    bool peer_ipc[PEER_MAX] = {};              // opened with cudaIpcOpenMemHandle (another process)
    uint32_t use_seq[NBUF] = {};               // how many flushes buffer b has been through (never reset: the flags only grow)
    uint32_t peer_dirty[NBUF] = {};            // use of buffer b that every peer must have gathered before it is wiped
    unsigned long long peer_timeout_ns = 30000000000ull;   // a waiting GPU gives up after this long (HULK_B200_PEER_TIMEOUT_MS)
    uint32_t *d_hist_sum = nullptr;            // the summed spectrum of the flush under way
    unsigned int *d_ticket = nullptr;

    hulk_b200_stats st{};
    uint64_t extra_minimizers = 0;
    std::string err;

    // optional per-kernel-class timing
    bool profiling = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events[4];
    std::vector<cudaEvent_t> prof_pool;
    hulk_b200_profile prof{};
};

// RAII bracket: records an event pair around a kernel class when profiling is on
struct ProfScope {
    hulk_b200_ctx *ctx;
    int cls;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    static cudaEvent_t get(hulk_b200_ctx *ctx) {
        if (!ctx->prof_pool.empty()) {
            cudaEvent_t e = ctx->prof_pool.back();
            ctx->prof_pool.pop_back();
            return e;
        }
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        return e;
    }
    cudaStream_t st;
    ProfScope(hulk_b200_ctx *c, int k, cudaStream_t s = nullptr) : ctx(c), cls(k), st(s ? s : c->stream) {
        if (!ctx->profiling) return;
        e0 = get(ctx);
        e1 = get(ctx);
        cudaEventRecord(e0, st);
    }
    ~ProfScope() {
        if (!e0) return;
        cudaEventRecord(e1, st);
        ctx->prof_events[cls].emplace_back(e0, e1);
    }
};

static thread_local std::string g_create_err;

// host-side stopwatch for HULK_B200_FEED_STATS=1 runs: where the calling thread's time goes, by label
static const bool g_host_stats = [] { const char *e = getenv("HULK_B200_FEED_STATS"); return e && *e == '1'; }();
struct HostStat { const char *name; uint64_t ns = 0, n = 0; };
static HostStat g_hstat[12] = {{"push: backpressure"}, {"push: sizing"}, {"push: post"}, {"push: wait32"}, {"push: first_k1"},
                               {"push: unpack launches"}, {"push: k1 launches"}, {"push: records"}, {"flush"}, {"snapshot"},
                               {"push (whole call)"}, {"-"}};
struct HostLap {
    std::chrono::steady_clock::time_point t;
    HostLap() { if (g_host_stats) t = std::chrono::steady_clock::now(); }
    void lap(int i) {
        if (!g_host_stats) return;
        const auto now = std::chrono::steady_clock::now();
        g_hstat[i].ns += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(now - t).count();
        g_hstat[i].n++;
        t = now;
    }
};
static void host_stats_print() {
    if (!g_host_stats) return;
    for (auto &h : g_hstat)
        if (h.n) fprintf(stderr, "[host] %-24s %8llu calls  %9.3f ms  %7.2f us each\n", h.name, (unsigned long long)h.n, h.ns * 1e-6,
                         h.ns * 1e-3 / (double)h.n);
}

const char *hulk_b200_version(void) { return HULK_B200_VERSION; }

const char *hulk_b200_strerror(int code) {
    switch (code) {
        case HULK_B200_OK: return "ok";
        case HULK_B200_EW: return "w must be: 0 < w < 257";
        case HULK_B200_EK: return "k size must be: 0 < k < 32";
        case HULK_B200_EEMPTYSEQ: return "sequence length must be > 0";
        case HULK_B200_ESHORTSEQ: return "sequence length must be >= w + k - 1";
        case HULK_B200_ESPARSE: return "not used yet";
        case HULK_B200_EHSK: return "histosketching only supports k <= 31";
        case HULK_B200_EDECAY: return "decay ratio must be between 0.0 and 1.0";
        case HULK_B200_EBINS: return "histogram must have at least 2 bins";
        case HULK_B200_ENEGBINS: return "negative value used for number of k-mer spectrum bins";
        case HULK_B200_ENOSKETCH: return "no sketch was generated by the histosketch algorithm";
        case HULK_B200_EARG: return "invalid argument";
        case HULK_B200_ESTATE: return "call out of order";
        case HULK_B200_ECUDA: return "CUDA error";
        case HULK_B200_ENOMEM: return "out of memory";
        case HULK_B200_EIO: return "I/O error";
        case HULK_B200_EFASTQ: return "read ID in fastq file does not begin with @";
        case HULK_B200_ETOOLONG: return "bufio.Scanner: token too long";
        case HULK_B200_ENOSEQ: return "no sequences received";
        default: return "unknown error";
    }
}

const char *hulk_b200_last_error(const hulk_b200_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

static int fail(hulk_b200_ctx *ctx, int code, const std::string &detail = std::string()) {
    std::string msg = hulk_b200_strerror(code);
    if (!detail.empty()) msg += ": " + detail;
    if (ctx) ctx->err = msg;
    else g_create_err = msg;
    return code;
}

#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return fail(ctx, e__ == cudaErrorMemoryAllocation ? HULK_B200_ENOMEM : HULK_B200_ECUDA,  \
                        std::string(#call) + " -> " + cudaGetErrorString(e__));                      \
    } while (0)

#define LAUNCH_CHECK(name)                                                                           \
    do {                                                                                             \
        ctx->st.n_kernel_launches++;                                                                 \
        cudaError_t e__ = cudaGetLastError();                                                        \
        if (e__ != cudaSuccess)                                                                      \
            return fail(ctx, HULK_B200_ECUDA, std::string("launch ") + name + " -> " + cudaGetErrorString(e__)); \
    } while (0)

static uint32_t pow4(uint32_t a) {   // helpers.Pow(k, 4) in uint arithmetic, then int32()  cmd/sketch.go:118
    uint64_t p = (uint64_t)a * a;
    p = p * p;
    return (uint32_t)p;
}

template <class T>
static cudaError_t dmalloc(T **p, uint64_t n) {
    return cudaMalloc(reinterpret_cast<void **>(p), (size_t)(n ? n : 1) * sizeof(T));
}

void hulk_b200_destroy(hulk_b200_ctx *ctx) {
    if (!ctx) return;
    if (ctx->gen_thread.joinable()) ctx->gen_thread.join();
    if (ctx->feeder.joinable()) {
        {
            std::lock_guard<std::mutex> lk(ctx->feed_mu);
            ctx->feed_stop = true;
        }
        ctx->feed_cv.notify_all();
        ctx->feeder.join();
        host_stats_print();
        if (getenv("HULK_B200_FEED_STATS"))
            fprintf(stderr, "[feed] packing %.1f GB/s, link %.1f GB/s\n", ctx->feed_bp * 1e-9, ctx->feed_bl * 1e-9);
            fprintf(stderr, "[feed] requests %llu, %.3f of the bases packed (last split %.3f): feeder idle %.3f ms, waiting for its pinned "
                            "buffer %.3f, packing %.3f, enqueueing copies %.3f; caller held back %.3f ms\n",
                    (unsigned long long)ctx->feed_done, ctx->feed_frac_sum / (double)std::max<uint64_t>(1, ctx->feed_done), ctx->feed_frac, ctx->feed_t_idle.load() * 1e-6, ctx->feed_t_evsync.load() * 1e-6,
                    ctx->feed_pack_ns.load() * 1e-6, ctx->feed_t_enq.load() * 1e-6, ctx->feed_t_backpressure.load() * 1e-6);
    }
    cudaSetDevice(ctx->P.device);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    for (int i = 0; i < NBUF; i++)
        if (ctx->k1_stream[i]) cudaStreamSynchronize(ctx->k1_stream[i]);
    if (ctx->k2_stream) cudaStreamSynchronize(ctx->k2_stream);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < NBUF; i++) {
        void *per[] = {ctx->d_ovf_count[i], ctx->d_ovf_list[i], ctx->d_arena[i], ctx->d_arena_cursor[i],
                       ctx->d_queue[i], ctx->d_queue_cursor[i], ctx->d_long_tasks[i], ctx->d_long_ctl[i],
                       ctx->d_kmv_state[i], ctx->d_kmv_pool[i], ctx->d_kmv_cand[i]};
        for (void *p : per)
            if (p) cudaFree(p);
    }
    for (int i = 0; i < NSTAGE; i++) {
        if (ctx->d_stage[i]) cudaFree(ctx->d_stage[i]);
        if (ctx->d_off[i]) cudaFree(ctx->d_off[i]);
        if (ctx->d_pack[i]) cudaFree(ctx->d_pack[i]);
        if (ctx->d_exc[i]) cudaFree(ctx->d_exc[i]);
        if (ctx->h_pack[i]) cudaFreeHost(ctx->h_pack[i]);
        if (ctx->h_exc[i]) cudaFreeHost(ctx->h_exc[i]);
    }
    for (uint32_t p = 0; p < PEER_MAX; p++)
        if (ctx->peer_ipc[p] && ctx->peer_arena[p]) cudaIpcCloseMemHandle(ctx->peer_arena[p]);
    if (ctx->h_feed) cudaFreeHost(ctx->h_feed);
    if (ctx->d_khf) cudaFree(ctx->d_khf);
    if (ctx->d_lenstat) cudaFree(ctx->d_lenstat);
    void *ptrs[] = {ctx->d_feed, ctx->arena, ctx->d_hist_sum, ctx->d_ticket, ctx->d_nmin, ctx->d_errword, ctx->d_ctl,
                    ctx->d_cols, ctx->d_csr_start, ctx->d_csr_bins, ctx->d_words, ctx->d_word_prefix,
                    ctx->d_block_count, ctx->d_block_prefix, ctx->d_q, ctx->d_fbits[0], ctx->d_fbits[1], ctx->d_invf[0],
                    ctx->d_invf[1], ctx->d_r, ctx->d_c,
                    ctx->d_b, ctx->d_K32, ctx->d_m32, ctx->d_sketch, ctx->d_weights, ctx->d_cand, ctx->d_thr32,
                    ctx->d_K16, ctx->d_invf16[0], ctx->d_invf16[1]};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    for (int i = 0; i < NSTAGE; i++) {
        if (ctx->ev_copy[i]) cudaEventDestroy(ctx->ev_copy[i]);
        for (int q = 0; q < 2; q++) if (ctx->ev_k1x[i][q]) cudaEventDestroy(ctx->ev_k1x[i][q]);
        if (ctx->ev_tail[i]) cudaEventDestroy(ctx->ev_tail[i]);
        if (ctx->ev_tail0[i]) cudaEventDestroy(ctx->ev_tail0[i]);
    }
    for (int i = 0; i < NBUF; i++) {
        if (ctx->ev_k1_last[i]) cudaEventDestroy(ctx->ev_k1_last[i]);
        if (ctx->ev_hist_free[i]) cudaEventDestroy(ctx->ev_hist_free[i]);
        if (ctx->k1_stream[i]) cudaStreamDestroy(ctx->k1_stream[i]);
    }
    if (ctx->ev_main) cudaEventDestroy(ctx->ev_main);
    for (int i = 0; i < 2; i++) {
        if (ctx->ev_k2_done[i]) cudaEventDestroy(ctx->ev_k2_done[i]);
        if (ctx->ev_k3_done[i]) cudaEventDestroy(ctx->ev_k3_done[i]);
    }
    if (ctx->k2_stream) cudaStreamDestroy(ctx->k2_stream);
    for (int c = 0; c < 4; c++)
        for (auto &pr : ctx->prof_events[c]) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    for (cudaEvent_t e : ctx->prof_pool) cudaEventDestroy(e);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

// every stream of the context idle (inputs copied, all k1 launches and flushes done)
static int feed_drain(hulk_b200_ctx *ctx) {
    if (!ctx->feeder.joinable()) return HULK_B200_OK;
    std::unique_lock<std::mutex> lk(ctx->feed_mu);
    ctx->feed_cv.wait(lk, [&] { return ctx->feed_done == ctx->feed_posted; });
    if (ctx->feed_rc) {
        const int rc = ctx->feed_rc;
        ctx->err = ctx->feed_err;
        ctx->feed_rc = 0;
        return rc;
    }
    return HULK_B200_OK;
}
static int sync_all(hulk_b200_ctx *ctx) {
    { const int rc = feed_drain(ctx); if (rc) return rc; }
    CU(cudaStreamSynchronize(ctx->copy_stream));
    for (int i = 0; i < NBUF; i++) CU(cudaStreamSynchronize(ctx->k1_stream[i]));
    CU(cudaStreamSynchronize(ctx->k2_stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return HULK_B200_OK;
}

__global__ void k_snapshot(const unsigned long long *__restrict__ sketch, const double *__restrict__ weights,
                           unsigned long long *__restrict__ h_mins, double *__restrict__ h_weights, uint32_t rows);
__global__ void k_merge_hist(uint32_t *__restrict__ dst, const uint32_t *__restrict__ src, int32_t D);
__global__ void k0_unpack(const uint32_t *__restrict__ packed, uint64_t n_words, uint4 *__restrict__ ascii);
__global__ void k0_patch(const uint32_t *__restrict__ exc, uint64_t n, uint32_t shift, uint8_t *__restrict__ ascii);
__global__ void k0_unpack_fed(const uint32_t *__restrict__ packed, uint64_t n_words, uint4 *__restrict__ ascii,
                              const uint32_t *__restrict__ meta);
__global__ void k0_patch_fed(const uint32_t *__restrict__ exc, const uint32_t *__restrict__ meta, uint8_t *__restrict__ ascii);

static int create_impl(hulk_b200_ctx *ctx) {
    const hulk_b200_params &P = ctx->P;
    CU(cudaSetDevice(P.device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, P.device));
    ctx->sm_count = prop.multiProcessorCount;
    // the flush chain (k2, k3) is the critical path of a pipelined run: it gets the SMs first, the k1 streams
    // (counting the NEXT interval) fill what is left
    int prio_least = 0, prio_greatest = 0;
    CU(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    if (P.stream) {
        ctx->stream = reinterpret_cast<cudaStream_t>(P.stream);
    } else {
        CU(cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_greatest));
        ctx->own_stream = true;
    }
    CU(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < NSTAGE; i++) {
        CU(cudaEventCreateWithFlags(&ctx->ev_copy[i], cudaEventDisableTiming));
        for (int q = 0; q < 2; q++) CU(cudaEventCreateWithFlags(&ctx->ev_k1x[i][q], cudaEventDisableTiming));
        CU(cudaEventCreate(&ctx->ev_tail0[i]));
        CU(cudaEventCreate(&ctx->ev_tail[i]));
    }
    for (int i = 0; i < NBUF; i++) {
        CU(cudaEventCreateWithFlags(&ctx->ev_k1_last[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_hist_free[i], cudaEventDisableTiming));
        CU(cudaStreamCreateWithPriority(&ctx->k1_stream[i], cudaStreamNonBlocking, prio_least));
    }
    CU(cudaEventCreateWithFlags(&ctx->ev_main, cudaEventDisableTiming));
    CU(cudaStreamCreateWithPriority(&ctx->k2_stream, cudaStreamNonBlocking, prio_greatest));
    for (int i = 0; i < 2; i++) {
        CU(cudaEventCreateWithFlags(&ctx->ev_k2_done[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_k3_done[i], cudaEventDisableTiming));
    }
    const int32_t D = ctx->D;
    const uint32_t rows = ctx->rows;
    ctx->Dp = ((uint64_t)D + K3_SUB - 1) / K3_SUB * K3_SUB;
    ctx->nsub_row = (uint32_t)(ctx->Dp / K3_SUB);
    {
        const char *e = getenv("HULK_B200_K3_FP32");
        if (e && *e == '1') ctx->filter16 = false;
    }
    ctx->k3_eps = ctx->filter16 ? K3_EPS16 : K3_EPS;
    const uint64_t seg_bins = ctx->filter16 ? K3_SEG16 : K3_SEG;
    ctx->nseg = (uint32_t)((ctx->Dp + seg_bins - 1) / seg_bins);
    ctx->nblk = (uint32_t)(((uint64_t)D + 1023) / 1024);

    ctx->hist_stride = ((size_t)D * 4 + 255) & ~(size_t)255;
    ctx->off_counted = (size_t)NBUF * ctx->hist_stride;
    ctx->off_gathered = ctx->off_counted + sizeof(uint32_t) * NBUF * PEER_MAX;
    ctx->arena_bytes = ctx->off_gathered + sizeof(uint32_t) * NBUF * PEER_MAX;
    CU(dmalloc(&ctx->arena, ctx->arena_bytes));
    for (int i = 0; i < NBUF; i++) ctx->d_hist[i] = reinterpret_cast<uint32_t *>(ctx->arena + (size_t)i * ctx->hist_stride);
    ctx->peer_arena[0] = ctx->arena;
    CU(dmalloc(&ctx->d_nmin, 1));
    CU(dmalloc(&ctx->d_errword, 1));
    CU(dmalloc(&ctx->d_ticket, 1));
    CU(cudaMemset(ctx->d_ticket, 0, sizeof(unsigned int)));
    CU(dmalloc(&ctx->d_feed, NSTAGE * 4));
    CU(cudaMemset(ctx->d_feed, 0, sizeof(uint32_t) * NSTAGE * 4));
    CU(cudaHostAlloc(reinterpret_cast<void **>(&ctx->h_feed), sizeof(uint32_t) * NSTAGE * 4, cudaHostAllocDefault));
    {
        // stream memory operations of the driver API, reached through the runtime (no link-time dependency on libcuda):
        // "wait until the word at this device address is >= v" / "write v there" as stream-ordered operations
        cudaDriverEntryPointQueryResult q1, q2;
        void *f1 = nullptr, *f2 = nullptr;
        const char *off = getenv("HULK_B200_FEEDER");
        if (!(off && *off == '0') && ctx->feeder_allowed &&
            cudaGetDriverEntryPoint("cuStreamWaitValue32", &f1, cudaEnableDefault, &q1) == cudaSuccess &&
            cudaGetDriverEntryPoint("cuStreamWriteValue32", &f2, cudaEnableDefault, &q2) == cudaSuccess &&
            q1 == cudaDriverEntryPointSuccess && q2 == cudaDriverEntryPointSuccess && f1 && f2) {
            ctx->cu_wait32 = reinterpret_cast<decltype(ctx->cu_wait32)>(f1);
            ctx->cu_write32 = reinterpret_cast<decltype(ctx->cu_write32)>(f2);
        }
        cudaGetLastError();
    }
    ctx->ovf_cap = 1u << 20;
    for (int i = 0; i < NBUF; i++) {
        CU(dmalloc(&ctx->d_ovf_count[i], 1));
        CU(dmalloc(&ctx->d_ovf_list[i], ctx->ovf_cap));
        CU(dmalloc(&ctx->d_arena_cursor[i], 1));
        CU(dmalloc(&ctx->d_queue_cursor[i], 2));
        CU(dmalloc(&ctx->d_long_ctl[i], 1));
    }
    CU(dmalloc(&ctx->d_lenstat, 2));
    CU(dmalloc(&ctx->d_ctl, 1));
    CU(dmalloc(&ctx->d_cols, (uint64_t)D * CMS_DEPTH));
    CU(dmalloc(&ctx->d_csr_start, CMS_CELLS + 1));
    CU(dmalloc(&ctx->d_csr_bins, (uint64_t)D * CMS_DEPTH));
    CU(dmalloc(&ctx->d_words, (uint64_t)ctx->nblk * 32));
    CU(dmalloc(&ctx->d_word_prefix, (uint64_t)ctx->nblk * 32));
    CU(dmalloc(&ctx->d_block_count, ctx->nblk));
    CU(dmalloc(&ctx->d_block_prefix, ctx->nblk));
    CU(dmalloc(&ctx->d_q, CMS_CELLS));
    for (int i = 0; i < 2; i++) {
        CU(dmalloc(&ctx->d_fbits[i], D));
        CU(dmalloc(&ctx->d_invf[i], ctx->Dp));
        if (ctx->filter16) CU(dmalloc(&ctx->d_invf16[i], ctx->Dp));
    }
    CU(dmalloc(&ctx->d_m32, (uint64_t)rows * ctx->nsub_row));
    CU(dmalloc(&ctx->d_sketch, rows));
    CU(dmalloc(&ctx->d_cand, rows));
    CU(dmalloc(&ctx->d_thr32, rows));
    CU(dmalloc(&ctx->d_weights, rows));

    cudaStream_t st = ctx->stream;
    CU(cudaMemsetAsync(ctx->arena, 0, ctx->arena_bytes, st));
    for (int i = 0; i < NBUF; i++) {
        CU(cudaMemsetAsync(ctx->d_ovf_count[i], 0, 4, st));
        CU(cudaMemsetAsync(ctx->d_arena_cursor[i], 0, 8, st));
    }
    CU(cudaMemsetAsync(ctx->d_nmin, 0, 8, st));
    CU(cudaMemsetAsync(ctx->d_errword, 0xff, 8, st));
    CU(cudaMemsetAsync(ctx->d_ctl, 0, sizeof(FlushCtl), st));
    CU(cudaMemsetAsync(ctx->d_q, 0, sizeof(double) * CMS_CELLS, st));
    for (int i = 0; i < 2; i++) {
        CU(cudaMemsetAsync(ctx->d_invf[i], 0xff, sizeof(float) * ctx->Dp, st));                 // NaN padding
        if (ctx->filter16) CU(cudaMemsetAsync(ctx->d_invf16[i], 0xff, 2 * ctx->Dp, st));        // bf16 0xffff = NaN
    }
    CU(cudaMemsetAsync(ctx->d_sketch, 0, sizeof(unsigned long long) * (rows ? rows : 1), st));   // histosketch.go:84-87
    CU(cudaMemsetAsync(ctx->d_cand, 0, sizeof(unsigned int) * (rows ? rows : 1), st));
    if (rows) k3_fill_f32<<<(rows + 255) / 256, 256, 0, st>>>(ctx->d_thr32, rows, INFINITY);   // W = MaxFloat64
    {
        std::vector<double> w(rows ? rows : 1, 1.7976931348623157e308);             // math.MaxFloat64
        CU(cudaMemcpyAsync(ctx->d_weights, w.data(), sizeof(double) * rows, cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));
    }
    // static count-min column map + CSR
    {
        unsigned int *d_counts = nullptr, *d_cursor = nullptr;
        CU(dmalloc(&d_counts, CMS_CELLS));
        CU(dmalloc(&d_cursor, CMS_CELLS));
        CU(cudaMemsetAsync(d_counts, 0, sizeof(unsigned int) * CMS_CELLS, st));
        CU(cudaMemsetAsync(d_cursor, 0, sizeof(unsigned int) * CMS_CELLS, st));
        const uint64_t n = (uint64_t)D * CMS_DEPTH;
        k2_build_cols<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(D, ctx->d_cols, d_counts);
        LAUNCH_CHECK("k2_build_cols");
        k2_scan_counts<<<1, 1024, 0, st>>>(d_counts, ctx->d_csr_start);
        LAUNCH_CHECK("k2_scan_counts");
        k2_fill_csr<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(D, ctx->d_cols, ctx->d_csr_start, d_cursor,
                                                                 ctx->d_csr_bins);
        LAUNCH_CHECK("k2_fill_csr");
        k2_sort_csr<<<(CMS_CELLS + 127) / 128, 128, 0, st>>>(ctx->d_csr_start, ctx->d_csr_bins);
        LAUNCH_CHECK("k2_sort_csr");
        CU(cudaStreamSynchronize(st));
        cudaFree(d_counts);
        cudaFree(d_cursor);
    }
    CU(cudaFuncSetAttribute(k3_filter<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * K3_SEG * 4 + 2 * 4 * 8));
    CU(cudaFuncSetAttribute(k3_filter<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * K3_SEG * 4 + 2 * 6 * 8));
    CU(cudaFuncSetAttribute(k3_filter<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * K3_SEG * 4 + 2 * 8 * 8));
    CU(cudaFuncSetAttribute(k3_filter16<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * K3_SEG16 * 2 + 2 * 4 * 8));
    CU(cudaFuncSetAttribute(k3_filter16<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * K3_SEG16 * 2 + 2 * 6 * 8));
    CU(cudaFuncSetAttribute(k3_filter16<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * K3_SEG16 * 2 + 2 * 8 * 8));
    {
        const char *e = getenv("HULK_B200_K3_STAGES");
        if (e && (*e == '4' || *e == '6' || *e == '8')) ctx->k3_stages = *e - '0';
        e = getenv("HULK_B200_NBUF");
        if (e && *e >= '2' && *e <= '0' + NBUF) ctx->nbuf = *e - '0';
        e = getenv("HULK_B200_K3_CTAS");
        if (e && (*e == '1' || *e == '2' || *e == '3')) ctx->k3_ctas_per_sm = *e - '0';
    }
    {
        const int big = 200 * 1024;
        const auto attr = cudaFuncAttributeMaxDynamicSharedMemorySize;
        CU(cudaFuncSetAttribute(k1_minimizer_histogram<false, false, false>, attr, big));
        CU(cudaFuncSetAttribute(k1_minimizer_histogram<false, true, false>, attr, big));
        CU(cudaFuncSetAttribute(k1_minimizer_histogram<false, false, true>, attr, big));
        CU(cudaFuncSetAttribute(k1_minimizer_histogram<false, true, true>, attr, big));
        CU(cudaFuncSetAttribute(k1_minimizer_histogram<true, false, false>, attr, big));
        CU(cudaFuncSetAttribute(k1_minimizer_histogram<true, true, false>, attr, big));
        CU(cudaFuncSetAttribute(k1_minimizer_histogram_w9<false, false, false>, attr, big));
        CU(cudaFuncSetAttribute(k1_minimizer_histogram_w9<false, true, false>, attr, big));
        CU(cudaFuncSetAttribute(k1_minimizer_histogram_w9<false, false, true>, attr, big));
        CU(cudaFuncSetAttribute(k1_minimizer_histogram_w9<false, true, true>, attr, big));
        CU(cudaFuncSetAttribute(k1_minimizer_histogram_w9<true, false, false>, attr, big));
        CU(cudaFuncSetAttribute(k1_minimizer_histogram_w9<true, true, false>, attr, big));
        CU(cudaFuncSetAttribute(k1_minimizer_histogram_w9<false, true, true, 21>, attr, big));
        CU(cudaFuncSetAttribute(k1_minimizer_histogram_w9<false, true, true, 11>, attr, big));
        CU(cudaFuncSetAttribute(k1_minimizer_histogram_w9<false, false, true, 31>, attr, big));
#define K1_V2_ATTR(KK)                                                                   \
        CU(cudaFuncSetAttribute(k1_scan_w9_v2<false, true, KK>, attr, big));             \
        CU(cudaFuncSetAttribute(k1_scan_w9_v2<true, false, KK>, attr, big));
        K1_V2_FOR_EACH_K(K1_V2_ATTR)
#undef K1_V2_ATTR
    }
    // CUDA loads a kernel's code the first time it is launched, and that load can wait for the device to go idle.
    // A multi-GPU flush parks a (one-warp) kernel on the device until its peers have signalled; a first launch queued
    // behind it from the same host thread would then wait for a signal that thread has not sent yet.  So every kernel a
    // push or a flush can launch is loaded here, while nothing is running.
    {
        cudaFuncAttributes fa;
#define HULK_PRELOAD(F) CU(cudaFuncGetAttributes(&fa, F))
        HULK_PRELOAD(k_peer_signal);
        HULK_PRELOAD(k_peer_wait);
        HULK_PRELOAD(k2_mask_count_peers);
        HULK_PRELOAD(k2_mask_count);
        HULK_PRELOAD(k2_cms_update);
        HULK_PRELOAD(k2_finalize);
        HULK_PRELOAD(k3_resolve);
        HULK_PRELOAD(k3_fill_f32);
        HULK_PRELOAD(k_snapshot);
        HULK_PRELOAD(k_merge_hist);
        HULK_PRELOAD(k0_unpack);
        HULK_PRELOAD(k0_patch);
        HULK_PRELOAD(k0_unpack_fed);
        HULK_PRELOAD(k0_patch_fed);
        HULK_PRELOAD(k1_generic<false>);
        HULK_PRELOAD(k1_generic<true>);
        HULK_PRELOAD(k1_long_plan);
        HULK_PRELOAD(k1_length_stats);
        HULK_PRELOAD(k1_khf_queue);
        HULK_PRELOAD(k1_kmv_filter);
        HULK_PRELOAD(k1_kmv_select);
        HULK_PRELOAD(k1_long_zero);
        HULK_PRELOAD(k1_long_scan<false>);
        HULK_PRELOAD(k1_long_scan<true>);
        HULK_PRELOAD((k1_jump_queue<2>));
        HULK_PRELOAD((k1_jump_queue<4>));
        HULK_PRELOAD((k1_jump_queue_fx<2, false>));
        HULK_PRELOAD((k1_jump_queue_fx<3, false>));
        HULK_PRELOAD((k1_jump_queue_fx<4, false>));
        HULK_PRELOAD((k1_jump_queue_fx<3, true>));
        HULK_PRELOAD((k1_jump_queue_fx<4, true>));
#undef HULK_PRELOAD
    }
    {
        const char *e = getenv("HULK_B200_K1_TILE");
        ctx->force_tile_path = e && *e == '1';
        e = getenv("HULK_B200_K1_CTAS");
        if (e && *e >= '1' && *e <= '0' + K1_W9_CTAS_PER_SM) ctx->k1_ctas_per_sm = *e - '0';
        e = getenv("HULK_B200_LONG");
        ctx->long_path = !(e && *e == '0');
        e = getenv("HULK_B200_LIST_CAP");
        if (e && atoi(e) >= 16 && atoi(e) <= 192) ctx->list_cap_w9 = (uint32_t)atoi(e);
        e = getenv("HULK_B200_LONG_MIN");
        if (e && atoll(e) >= 1024) ctx->long_min = (uint64_t)atoll(e);
        e = getenv("HULK_B200_MAX_LAUNCH_READS");
        if (e && atoll(e) >= 32) ctx->max_launch_reads = std::min<uint64_t>((uint64_t)atoll(e), 1ull << 30);
        e = getenv("HULK_B200_JUMP_CTAS");
        if (e && *e >= '1' && *e <= '8') ctx->jump_ctas_per_sm = *e - '0';
        e = getenv("HULK_B200_JUMP_BATCH");
        if (e && (*e == '2' || *e == '3' || *e == '4')) ctx->jump_batch = *e - '0';
        e = getenv("HULK_B200_K1_PERSISTENT");
        if (e && *e == '1') ctx->k1_persistent = true;
        e = getenv("HULK_B200_PEER_TIMEOUT_MS");
        if (e && atoll(e) > 0) ctx->peer_timeout_ns = (unsigned long long)atoll(e) * 1000000ull;
        e = getenv("HULK_B200_JUMP_SMEM");
        if (e && *e == '1') ctx->jump_smem = true;
        e = getenv("HULK_B200_JUMP_FX");
        if (e && *e == '0') ctx->jump_fx = false;
        e = getenv("HULK_B200_K1_FUSED");
        ctx->fused_jump = e && *e == '1';
        e = getenv("HULK_B200_K1_V2");
        if (e && *e == '0') ctx->k1_v2 = false;
        e = getenv("HULK_B200_SERIAL");
        if (e && *e == '1') ctx->overlap = false;
        if (P.flags & HULK_B200_F_PACK_INPUT) ctx->pack_threads = -1;
        e = getenv("HULK_B200_PACK_FRACTION");
        if (e && *e) {
            ctx->feed_frac = std::min(1.0, std::max(0.0, atof(e)));
            ctx->feed_adapt = false;
        }
        e = getenv("HULK_B200_PACK_INPUT");                  // A/B runs and the CLI: 1 = on, 0 = off whatever the flag says
        if (e && *e == '1') ctx->pack_threads = -1;
        if (e && *e == '0') ctx->pack_threads = 0;
    }
    return HULK_B200_OK;
}

int hulk_b200_create(const hulk_b200_params *params, hulk_b200_ctx **out) {
    hulk_b200_ctx *ctx = nullptr;
    if (!params || !out) return fail(ctx, HULK_B200_EARG, "params/out is NULL");
    *out = nullptr;
    // validation in the order the reference reaches it
    if (params->num_bins < 0) return fail(ctx, HULK_B200_ENEGBINS);                       // kmerspectrum.go:33-35
    if (params->w > 256) return fail(ctx, HULK_B200_EW);                                  // minimizer.go:62-64
    if (params->k > 31) return fail(ctx, HULK_B200_EK);                                   // minimizer.go:65-67 / histosketch.go:53
    if (!(params->decay_ratio >= 0.0 && params->decay_ratio <= 1.0)) return fail(ctx, HULK_B200_EDECAY);
    const int32_t D = params->num_bins ? params->num_bins : (int32_t)pow4(params->k);
    if (D < 2) return fail(ctx, HULK_B200_EBINS);                                         // histosketch.go:65-67
    if (params->w == 0) return fail(ctx, HULK_B200_EW, "w = 0 never emits a window");     // reference would index q[-1]
    uint32_t sb = params->slot_begin, se = params->slot_end;
    if (sb == 0 && se == 0) se = params->sketch_size;
    if (sb > se || se > params->sketch_size) return fail(ctx, HULK_B200_EARG, "slot range");

    // The feeder path orders a counting stream behind a copy with stream memory operations, an order CUDA's scheduler does
    // not see: if the two streams shared a hardware queue, the wait could end up in front of the copy that satisfies it.
    // Every stream must therefore have its own queue: the context's ~8 streams (plus the caller's) need more than the
    // default 8 connections.  The setting is read when the device's context is created: it is set here if that has not
    // happened yet; a context created earlier without it keeps the safe path (packing on the calling thread).
    bool queues_ok;
    {
        const char *mc = getenv("CUDA_DEVICE_MAX_CONNECTIONS");
        if (mc) {
            queues_ok = atoi(mc) >= 16;
        } else {
            bool context_exists = true;               // unknown means: assume the worst
            cudaDriverEntryPointQueryResult q;
            void *f = nullptr;
            if (cudaGetDriverEntryPoint("cuDevicePrimaryCtxGetState", &f, cudaEnableDefault, &q) == cudaSuccess &&
                q == cudaDriverEntryPointSuccess && f) {
                unsigned int flags = 0;
                int active = 1;
                auto get_state = reinterpret_cast<CUresult (*)(CUdevice, unsigned int *, int *)>(f);
                if (get_state((CUdevice)params->device, &flags, &active) == CUDA_SUCCESS) context_exists = active != 0;
            }
            cudaGetLastError();
            if (!context_exists) setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 1);
            queues_ok = !context_exists;
        }
        const char *force = getenv("HULK_B200_FEEDER");
        if (force && *force == '1') queues_ok = true;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(ctx, HULK_B200_ECUDA, std::string("no CUDA device (no CPU fallback exists): ") +
                                              (e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e)));
    if (params->device < 0 || params->device >= ndev) return fail(ctx, HULK_B200_EARG, "device ordinal");

    ctx = new (std::nothrow) hulk_b200_ctx();
    if (!ctx) return fail(nullptr, HULK_B200_ENOMEM);
    ctx->P = *params;
    ctx->P.slot_begin = sb;
    ctx->P.slot_end = se;
    ctx->D = D;
    ctx->s = params->sketch_size;
    ctx->rows = se - sb;
    ctx->feeder_allowed = queues_ok;
    ctx->drift = (params->decay_ratio != 1.0);                                            // histosketch.go:79-81
    if (params->decay_ratio > 0.0 && params->decay_ratio < 1.0) {                         // countmin.go:50-55
        ctx->apply_scaling = true;
        ctx->decay_weight = std::exp(-params->decay_ratio);
    }
    const int rc = create_impl(ctx);
    if (rc != HULK_B200_OK) {
        g_create_err = ctx->err;
        hulk_b200_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return HULK_B200_OK;
}

// ------------------------------------------------------------------------------------------
// CWS tables
// ------------------------------------------------------------------------------------------
static int upload_tables(hulk_b200_ctx *ctx, const double *r, const double *c, const double *b) {
    const uint64_t n = (uint64_t)ctx->rows * (uint64_t)ctx->D;
    if (!ctx->d_r) {
        CU(dmalloc(&ctx->d_r, n));
        CU(dmalloc(&ctx->d_c, n));
        CU(dmalloc(&ctx->d_b, n));
        if (ctx->filter16) CU(dmalloc(&ctx->d_K16, (uint64_t)ctx->rows * ctx->Dp));
        else CU(dmalloc(&ctx->d_K32, (uint64_t)ctx->rows * ctx->Dp));
    }
    cudaStream_t st = ctx->stream;
    CU(cudaMemcpyAsync(ctx->d_r, r, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->d_c, c, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->d_b, b, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    ctx->st.h2d_bytes += 3 * sizeof(double) * n;
    const uint64_t total = (uint64_t)ctx->rows * ctx->Dp;
    if (total) {
        if (ctx->filter16)
            k3_fold16<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ctx->d_r, ctx->d_c, ctx->d_b, ctx->rows, ctx->D,
                                                                       ctx->Dp, ctx->d_K16);
        else
            k3_fold<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ctx->d_r, ctx->d_c, ctx->d_b, ctx->rows, ctx->D,
                                                                     ctx->Dp, ctx->d_K32);
        LAUNCH_CHECK("k3_fold");
    }
    CU(cudaStreamSynchronize(st));
    ctx->tables_set = true;
    return HULK_B200_OK;
}

int hulk_b200_set_cws_tables(hulk_b200_ctx *ctx, const double *r, const double *c, const double *b) {
    if (!ctx) return HULK_B200_EARG;
    if (ctx->rows && (!r || !c || !b)) return fail(ctx, HULK_B200_EARG, "table pointer is NULL");
    CU(cudaSetDevice(ctx->P.device));
    return upload_tables(ctx, r, c, b);
}

int hulk_b200_set_cws_tables_device(hulk_b200_ctx *ctx, const double *d_r, const double *d_c, const double *d_b) {
    if (!ctx) return HULK_B200_EARG;
    if (ctx->rows && (!d_r || !d_c || !d_b)) return fail(ctx, HULK_B200_EARG, "table pointer is NULL");
    CU(cudaSetDevice(ctx->P.device));
    const uint64_t n = (uint64_t)ctx->rows * (uint64_t)ctx->D;
    if (!ctx->d_r) {
        CU(dmalloc(&ctx->d_r, n));
        CU(dmalloc(&ctx->d_c, n));
        CU(dmalloc(&ctx->d_b, n));
        if (ctx->filter16) CU(dmalloc(&ctx->d_K16, (uint64_t)ctx->rows * ctx->Dp));
        else CU(dmalloc(&ctx->d_K32, (uint64_t)ctx->rows * ctx->Dp));
    }
    cudaStream_t st = ctx->stream;
    CU(cudaMemcpyAsync(ctx->d_r, d_r, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(ctx->d_c, d_c, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(ctx->d_b, d_b, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    const uint64_t total = (uint64_t)ctx->rows * ctx->Dp;
    if (total) {
        if (ctx->filter16)
            k3_fold16<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ctx->d_r, ctx->d_c, ctx->d_b, ctx->rows, ctx->D,
                                                                       ctx->Dp, ctx->d_K16);
        else
            k3_fold<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ctx->d_r, ctx->d_c, ctx->d_b, ctx->rows, ctx->D,
                                                                     ctx->Dp, ctx->d_K32);
        LAUNCH_CHECK("k3_fold");
    }
    CU(cudaStreamSynchronize(st));
    ctx->tables_set = true;
    return HULK_B200_OK;
}

// both side sketches as their constructors leave them: KHF slots at MaxUint64 (khf.go:20-32), an empty KMV heap
// (kmv.go:24-38)
static int minhash_clear(hulk_b200_ctx *ctx) {
    cudaStream_t st = ctx->stream;                           // (callers have drained every stream of the context)
    if (ctx->d_khf) CU(cudaMemsetAsync(ctx->d_khf, 0xff, sizeof(unsigned long long) * ctx->s, st));
    for (int i = 0; i < NBUF; i++)
        if (ctx->d_kmv_state[i]) CU(cudaMemsetAsync(ctx->d_kmv_state[i], 0, sizeof(K1KmvState), st));
    CU(cudaStreamSynchronize(st));
    return HULK_B200_OK;
}

int hulk_b200_minhash_enable(hulk_b200_ctx *ctx, int kmv, int khf) {
    if (!ctx) return HULK_B200_EARG;
    CU(cudaSetDevice(ctx->P.device));
    { const int rc = sync_all(ctx); if (rc) return rc; }
    if (ctx->st.n_reads) return fail(ctx, HULK_B200_ESTATE, "minhash_enable after reads were pushed");
    if (khf && !ctx->d_khf) CU(dmalloc(&ctx->d_khf, ctx->s));
    if (kmv) {
        for (int i = 0; i < NBUF; i++) {
            if (!ctx->d_kmv_state[i]) CU(dmalloc(&ctx->d_kmv_state[i], 1));
            if (!ctx->d_kmv_pool[i]) CU(dmalloc(&ctx->d_kmv_pool[i], 2 * (uint64_t)ctx->s));
        }
    }
    ctx->mh_kmv = kmv != 0;
    ctx->mh_khf = khf != 0;
    return minhash_clear(ctx);
}

int hulk_b200_get_khf(hulk_b200_ctx *ctx, uint64_t *mins) {
    if (!ctx || !mins) return HULK_B200_EARG;
    CU(cudaSetDevice(ctx->P.device));
    { const int rc = sync_all(ctx); if (rc) return rc; }
    if (!ctx->mh_khf || !ctx->d_khf) {                       // never fed: the constructor's state (khf.go:20-32)
        for (uint32_t i = 0; i < ctx->s; i++) mins[i] = ~0ull;
        return HULK_B200_OK;
    }
    CU(cudaMemcpy(mins, ctx->d_khf, sizeof(uint64_t) * ctx->s, cudaMemcpyDeviceToHost));
    return HULK_B200_OK;
}

int hulk_b200_get_kmv(hulk_b200_ctx *ctx, uint64_t *mins, uint32_t *n) {
    if (!ctx || !mins || !n) return HULK_B200_EARG;
    CU(cudaSetDevice(ctx->P.device));
    { const int rc = sync_all(ctx); if (rc) return rc; }
    *n = 0;
    if (!ctx->mh_kmv) return HULK_B200_OK;                   // never fed: an empty heap (kmv.go:24-38)
    // bottom-s of a union is the bottom-s of the parts' bottom-s: the pools of the k1 streams are merged here
    std::vector<uint64_t> all;
    for (int i = 0; i < NBUF; i++) {
        if (!ctx->d_kmv_state[i]) continue;
        K1KmvState st;
        CU(cudaMemcpy(&st, ctx->d_kmv_state[i], sizeof(st), cudaMemcpyDeviceToHost));
        const size_t have = all.size(), np = (size_t)std::min<unsigned long long>(st.n_pool, ctx->s);
        all.resize(have + np);
        if (np)
            CU(cudaMemcpy(all.data() + have, ctx->d_kmv_pool[i] + (size_t)st.cur * ctx->s, sizeof(uint64_t) * np,
                          cudaMemcpyDeviceToHost));
    }
    std::sort(all.begin(), all.end());                       // SetSketch: low -> high (kmv.go:160-168)
    if (all.size() > ctx->s) all.resize(ctx->s);
    std::copy(all.begin(), all.end(), mins);
    *n = (uint32_t)all.size();
    return HULK_B200_OK;
}

int hulk_b200_reset(hulk_b200_ctx *ctx) {
    if (!ctx) return HULK_B200_EARG;
    CU(cudaSetDevice(ctx->P.device));
    cudaStream_t st = ctx->stream;
    { const int rc = sync_all(ctx); if (rc) return rc; }
    for (int i = 0; i < NBUF; i++) {
        if (ctx->world > 1 && ctx->peer_dirty[i]) {      // a peer may still be reading this buffer
            const uint32_t *flags = reinterpret_cast<const uint32_t *>(ctx->arena + ctx->off_gathered) + (size_t)i * PEER_MAX;
            k_peer_wait<<<1, 32, 0, st>>>(flags, ctx->world, ctx->peer_dirty[i], ctx->d_ctl, ctx->peer_timeout_ns);
            ctx->peer_dirty[i] = 0;
        }
        CU(cudaMemsetAsync(ctx->d_hist[i], 0, sizeof(uint32_t) * (size_t)ctx->D, st));
        ctx->k1_pending[i] = false;
    }
    ctx->cur_hist = 0;
    ctx->flush_idx = ctx->last_flush_idx = 0;
    ctx->k3_pending[0] = ctx->k3_pending[1] = false;
    CU(cudaMemsetAsync(ctx->d_nmin, 0, 8, st));
    CU(cudaMemsetAsync(ctx->d_errword, 0xff, 8, st));
    CU(cudaMemsetAsync(ctx->d_ctl, 0, sizeof(FlushCtl), st));
    CU(cudaMemsetAsync(ctx->d_q, 0, sizeof(double) * CMS_CELLS, st));
    CU(cudaMemsetAsync(ctx->d_sketch, 0, sizeof(unsigned long long) * (ctx->rows ? ctx->rows : 1), st));
    CU(cudaMemsetAsync(ctx->d_cand, 0, sizeof(unsigned int) * (ctx->rows ? ctx->rows : 1), st));
    if (ctx->rows) k3_fill_f32<<<(ctx->rows + 255) / 256, 256, 0, st>>>(ctx->d_thr32, ctx->rows, INFINITY);
    std::vector<double> w(ctx->rows ? ctx->rows : 1, 1.7976931348623157e308);
    CU(cudaMemcpyAsync(ctx->d_weights, w.data(), sizeof(double) * ctx->rows, cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
    const uint64_t launches = ctx->st.n_kernel_launches;
    ctx->st = hulk_b200_stats{};
    ctx->st.n_kernel_launches = launches;
    ctx->feed_h2d = 0;
    ctx->feed_pack_ns = 0;
    ctx->feed_batches = 0;
    ctx->extra_minimizers = 0;
    ctx->err.clear();
    return minhash_clear(ctx);
}

int hulk_b200_set_overlap(hulk_b200_ctx *ctx, int enable) {
    if (!ctx) return HULK_B200_EARG;
    CU(cudaSetDevice(ctx->P.device));
    { const int rc = sync_all(ctx); if (rc) return rc; }
    ctx->overlap = enable != 0;
    return HULK_B200_OK;
}

int hulk_b200_profile_enable(hulk_b200_ctx *ctx, int enable) {
    if (!ctx) return HULK_B200_EARG;
    ctx->profiling = enable != 0;
    return HULK_B200_OK;
}
int hulk_b200_profile_read(hulk_b200_ctx *ctx, hulk_b200_profile *out) {
    if (!ctx || !out) return HULK_B200_EARG;
    CU(cudaSetDevice(ctx->P.device));
    { const int rc = sync_all(ctx); if (rc) return rc; }
    for (int c = 0; c < 4; c++) {
        for (auto &pr : ctx->prof_events[c]) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) {
                ctx->prof.ms[c] += ms;
                ctx->prof.launches[c] += 1;
            }
            ctx->prof_pool.push_back(pr.first);
            ctx->prof_pool.push_back(pr.second);
        }
        ctx->prof_events[c].clear();
    }
    *out = ctx->prof;
    ctx->prof = hulk_b200_profile{};
    return HULK_B200_OK;
}

// raw timeline of the recorded scopes (debug tap): rows of (class, start ms, end ms) relative to the
// earliest recorded start; consumes the records like profile_read
int hulk_b200_profile_timeline(hulk_b200_ctx *ctx, double *rows, uint64_t cap, uint64_t *n_out) {
    if (!ctx || !rows || !n_out) return HULK_B200_EARG;
    CU(cudaSetDevice(ctx->P.device));
    { const int rc = sync_all(ctx); if (rc) return rc; }
    cudaEvent_t base = nullptr;
    for (int c = 0; c < 4 && !base; c++)
        if (!ctx->prof_events[c].empty()) base = ctx->prof_events[c].front().first;
    uint64_t n = 0;
    for (int c = 0; c < 4; c++) {
        for (auto &pr : ctx->prof_events[c]) {
            float a = 0.f, b = 0.f;
            if (base && n < cap && cudaEventElapsedTime(&a, base, pr.first) == cudaSuccess &&
                cudaEventElapsedTime(&b, base, pr.second) == cudaSuccess) {
                rows[3 * n] = c; rows[3 * n + 1] = a; rows[3 * n + 2] = b;
                n++;
            }
            ctx->prof_pool.push_back(pr.first);
            ctx->prof_pool.push_back(pr.second);
        }
        ctx->prof_events[c].clear();
    }
    *n_out = n;
    return HULK_B200_OK;
}

int hulk_b200_generate_cws_tables_async(hulk_b200_ctx *ctx) {
    if (!ctx) return HULK_B200_EARG;
    if (ctx->gen_pending) return fail(ctx, HULK_B200_ESTATE, "table generation already running");
    const uint64_t n = (uint64_t)ctx->rows * (uint64_t)ctx->D;
    try {
        ctx->gen_r.resize(n ? n : 1); ctx->gen_c.resize(n ? n : 1); ctx->gen_b.resize(n ? n : 1);
    } catch (const std::bad_alloc &) {
        return fail(ctx, HULK_B200_ENOMEM, "host memory for the CWS tables");
    }
    ctx->gen_pending = true;
    ctx->gen_thread = std::thread([ctx] {
        ctx->gen_rc = hulk_b200_new_cws(ctx->s, ctx->D, ctx->P.slot_begin, ctx->P.slot_end, ctx->gen_r.data(),
                                        ctx->gen_c.data(), ctx->gen_b.data());
    });
    return HULK_B200_OK;
}
// join the generator thread (if any) and upload its tables
static int finish_async_tables(hulk_b200_ctx *ctx) {
    if (!ctx->gen_pending) return HULK_B200_OK;
    if (ctx->gen_thread.joinable()) ctx->gen_thread.join();
    ctx->gen_pending = false;
    int rc = ctx->gen_rc;
    if (rc == HULK_B200_OK) rc = upload_tables(ctx, ctx->gen_r.data(), ctx->gen_c.data(), ctx->gen_b.data());
    else fail(ctx, rc);
    std::vector<double>().swap(ctx->gen_r);
    std::vector<double>().swap(ctx->gen_c);
    std::vector<double>().swap(ctx->gen_b);
    return rc;
}

// ------------------------------------------------------------------------------------------
// newCWS drawn on the device (k4_cwsdraw.cuh)
// ------------------------------------------------------------------------------------------
extern "C" void hulk_b200_internal_alfg_window(int64_t seed, uint64_t *w);
extern "C" void hulk_b200_internal_alfg_xpow(uint64_t n, uint64_t *poly);
extern "C" void hulk_b200_internal_alfg_polymul(const uint64_t *a, const uint64_t *b, uint64_t *out);
extern "C" int hulk_b200_internal_gamma_accepts(uint64_t raw1, uint64_t raw2);

namespace {
struct DevBuf {                       // frees what the draw allocated, whichever way it leaves
    std::vector<void *> dev, host;
    ~DevBuf() {
        for (void *p : dev) cudaFree(p);
        for (void *p : host) cudaFreeHost(p);
    }
};
}  // namespace

// returns 1 when the draw has to be left to the host generator (an output converted to exactly 1.0, too many events)
static int draw_tables_on_device(hulk_b200_ctx *ctx) {
    const uint64_t D = (uint64_t)ctx->D;
    const uint64_t E = (uint64_t)ctx->P.slot_end * D, skip = (uint64_t)ctx->P.slot_begin * D;
    const uint64_t need = 2 * E;                                          // gamma draws: r, c, r, c, ...
    const uint64_t n_own = E - skip;
    cudaStream_t st = ctx->stream;
    if (E == 0) return HULK_B200_OK;
    const uint64_t L = K4_CHUNK;
    // ~4.55 raw outputs per element (1.14 attempts of two uniforms per draw, two draws); any shortfall is one more round
    uint64_t C = (uint64_t)(4.7 * (double)E / (double)L) + 1;
    C = std::min<uint64_t>(C, 1024);
    const uint64_t R = C * L, A = R / 2 + 2;                              // raw outputs and attempts (at most) per round
    const uint32_t ext_cap = 1u << 16, tie_cap = 1u << 12;
    // how close to its boundary an acceptance test has to be for the host to decide it (tests widen the band to drive
    // attempts through that path: HULK_B200_CWS_TIE_EPS)
    double tie_eps = 1e-9;
    {
        const char *e = getenv("HULK_B200_CWS_TIE_EPS");
        if (e && atof(e) > 0.0) tie_eps = atof(e);
    }
    const uint32_t nblk = (uint32_t)((A + K4_COUNT_TPB - 1) / K4_COUNT_TPB);
    DevBuf keep;
    uint64_t *d_states = nullptr, *d_poly = nullptr, *d_raw = nullptr, *d_ext = nullptr;
    double *d_xs = nullptr;
    uint8_t *d_acc = nullptr;
    K4Tie *d_ties = nullptr;
    uint32_t *d_bcount = nullptr;
    unsigned long long *d_bprefix = nullptr, *d_total = nullptr;
    K4ScanOut *d_scan = nullptr;
    K4SampleOut *d_samp = nullptr;
#define K4_ALLOC(ptr, count)                                   \
    do {                                                       \
        CU(dmalloc(&ptr, count));                              \
        keep.dev.push_back(ptr);                               \
    } while (0)
    K4_ALLOC(d_states, C * ALFG_LEN);
    K4_ALLOC(d_poly, ALFG_LEN);
    K4_ALLOC(d_raw, R + K4_LOOKAHEAD);
    K4_ALLOC(d_ext, ext_cap);
    K4_ALLOC(d_xs, A);
    K4_ALLOC(d_acc, A);
    K4_ALLOC(d_ties, tie_cap);
    K4_ALLOC(d_bcount, nblk);
    K4_ALLOC(d_bprefix, nblk);
    K4_ALLOC(d_total, 1);
    K4_ALLOC(d_scan, 1);
    K4_ALLOC(d_samp, 1);
#undef K4_ALLOC
    // generator windows: every block starts from the seeded window and jumps to block * L by binary decomposition
    {
        std::vector<uint64_t> w0(ALFG_LEN), all(C * ALFG_LEN), poly(ALFG_LEN);
        hulk_b200_internal_alfg_window(1, w0.data());                    // DISTRIBUTION_SEED  histosketch.go:20
        for (uint64_t c = 0; c < C; c++) std::copy(w0.begin(), w0.end(), all.begin() + c * ALFG_LEN);
        CU(cudaMemcpyAsync(d_states, all.data(), sizeof(uint64_t) * all.size(), cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));
        // Q_i = x^(L 2^i): block c applies the Q_i of its set bits; the jump between two rounds, x^((C - 1) L), is the
        // product of the Q_i of the set bits of C - 1
        std::vector<uint64_t> jump(ALFG_LEN, 0);
        jump[0] = 1;
        hulk_b200_internal_alfg_xpow(L, poly.data());
        for (int bit = 0; (1ull << bit) < C; bit++) {
            if (bit) hulk_b200_internal_alfg_polymul(poly.data(), poly.data(), poly.data());
            if (((C - 1) >> bit) & 1) hulk_b200_internal_alfg_polymul(jump.data(), poly.data(), jump.data());
            CU(cudaMemcpyAsync(d_poly, poly.data(), sizeof(uint64_t) * ALFG_LEN, cudaMemcpyHostToDevice, st));
            k4_apply_poly<<<(unsigned)C, 512, 0, st>>>(d_states, d_poly, bit);
            LAUNCH_CHECK("k4_apply_poly");
            CU(cudaStreamSynchronize(st));                                // poly is reused by the next bit
        }
        if (C > 1) {                                                      // between rounds: from the end of a chunk to its next one
            CU(cudaMemcpyAsync(d_poly, jump.data(), sizeof(uint64_t) * ALFG_LEN, cudaMemcpyHostToDevice, st));
            CU(cudaStreamSynchronize(st));
        }
    }
    std::vector<uint64_t> ext;
    std::vector<K4Tie> ties(tie_cap);
    uint64_t produced = 0, pos0 = 0, next_start = 0;
    while (produced < need) {
        k4_raw<<<(unsigned)C, 288, 0, st>>>(d_states, d_raw, K4_LOOKAHEAD);
        LAUNCH_CHECK("k4_raw");
        CU(cudaMemsetAsync(d_scan, 0, sizeof(K4ScanOut), st));
        k4_scan<<<ctx->sm_count * 8, 256, 0, st>>>(d_raw, R, pos0, d_ext, ext_cap, d_scan, ctx->d_b, skip, E);
        LAUNCH_CHECK("k4_scan");
        K4ScanOut so;
        CU(cudaMemcpyAsync(&so, d_scan, sizeof so, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (so.saw_one || so.n_extreme > ext_cap) return 1;
        ext.resize(so.n_extreme);
        if (so.n_extreme) CU(cudaMemcpy(ext.data(), d_ext, sizeof(uint64_t) * so.n_extreme, cudaMemcpyDeviceToHost));
        std::sort(ext.begin(), ext.end());
        // the pairing of the round: an attempt starts at next_start; a FIRST uniform outside (1e-7, 0.9999999) is consumed
        // alone and the next attempt starts right behind it (go_rng's `continue`)
        K4Segments segs{};
        const uint64_t end = pos0 + R;
        uint64_t cur = next_start, attempts = 0;
        bool too_many = false;
        auto add_seg = [&](uint64_t start, uint64_t n) {
            if (!n) return;
            if (segs.n >= (uint32_t)K4_MAX_SEGMENTS) { too_many = true; return; }
            segs.seg[segs.n++] = K4Segment{start - pos0, attempts, n};
            attempts += n;
        };
        for (uint64_t e : ext) {
            if (e < cur) continue;
            if (((e - cur) & 1) == 0) {                                   // a first uniform
                add_seg(cur, (e - cur) / 2);
                cur = e + 1;
            }
        }
        if (end > cur) {
            const uint64_t n = (end - cur + 1) / 2;                       // the last attempt may take its second uniform from
            add_seg(cur, n);                                              // the look-ahead
            next_start = cur + 2 * n;
        } else {
            next_start = cur;
        }
        if (too_many) return 1;
        segs.total_attempts = attempts;
        if (attempts) {
            CU(cudaMemsetAsync(d_samp, 0, sizeof(K4SampleOut), st));
            k4_sample<<<ctx->sm_count * 8, 256, 0, st>>>(d_raw, segs, d_xs, d_acc, d_ties, tie_cap, d_samp, tie_eps);
            LAUNCH_CHECK("k4_sample");
            K4SampleOut sp;
            CU(cudaMemcpyAsync(&sp, d_samp, sizeof sp, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            if (sp.n_ties > tie_cap) return 1;
            if (sp.n_ties) {                                              // the host's libm decides what is too close to call
                CU(cudaMemcpy(ties.data(), d_ties, sizeof(K4Tie) * sp.n_ties, cudaMemcpyDeviceToHost));
                for (uint32_t i = 0; i < sp.n_ties; i++) {
                    const uint8_t a = (uint8_t)hulk_b200_internal_gamma_accepts(ties[i].raw1, ties[i].raw2);
                    CU(cudaMemcpy(d_acc + ties[i].attempt, &a, 1, cudaMemcpyHostToDevice));
                }
                ctx->cws_ties += sp.n_ties;
            }
            const uint32_t nb = (uint32_t)((attempts + K4_COUNT_TPB - 1) / K4_COUNT_TPB);
            k4_count<<<nb, K4_COUNT_TPB, 0, st>>>(d_acc, attempts, d_bcount);
            LAUNCH_CHECK("k4_count");
            k4_scan_counts<<<1, 1024, 0, st>>>(d_bcount, nb, d_bprefix, d_total);
            LAUNCH_CHECK("k4_scan_counts");
            k4_scatter<<<nb, K4_COUNT_TPB, 0, st>>>(d_acc, d_xs, attempts, d_bprefix, produced, need, skip, ctx->d_r, ctx->d_c);
            LAUNCH_CHECK("k4_scatter");
            unsigned long long total = 0;
            CU(cudaMemcpyAsync(&total, d_total, sizeof total, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            produced += total;
        }
        pos0 = end;
        if (produced < need && C > 1) {
            k4_apply_poly<<<(unsigned)C, 512, 0, st>>>(d_states, d_poly, -1);
            LAUNCH_CHECK("k4_apply_poly");
        }
    }
    if (n_own) {
        k4_scale_b<<<(unsigned)((n_own + 255) / 256), 256, 0, st>>>(ctx->d_b, ctx->d_r, n_own);
        LAUNCH_CHECK("k4_scale_b");
    }
    CU(cudaStreamSynchronize(st));
    return HULK_B200_OK;
}

int hulk_b200_generate_cws_tables_device(hulk_b200_ctx *ctx) {
    if (!ctx) return HULK_B200_EARG;
    CU(cudaSetDevice(ctx->P.device));
    { const int rc = sync_all(ctx); if (rc) return rc; }
    const uint64_t n = (uint64_t)ctx->rows * (uint64_t)ctx->D;
    if (!ctx->d_r) {
        CU(dmalloc(&ctx->d_r, n));
        CU(dmalloc(&ctx->d_c, n));
        CU(dmalloc(&ctx->d_b, n));
        if (ctx->filter16) CU(dmalloc(&ctx->d_K16, (uint64_t)ctx->rows * ctx->Dp));
        else CU(dmalloc(&ctx->d_K32, (uint64_t)ctx->rows * ctx->Dp));
    }
    const int rc = draw_tables_on_device(ctx);
    if (rc == 1) return hulk_b200_generate_cws_tables(ctx);               // the rare cases the parallel draw does not cover
    if (rc) return rc;
    cudaStream_t st = ctx->stream;
    const uint64_t total = (uint64_t)ctx->rows * ctx->Dp;
    if (total) {
        if (ctx->filter16)
            k3_fold16<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ctx->d_r, ctx->d_c, ctx->d_b, ctx->rows, ctx->D,
                                                                       ctx->Dp, ctx->d_K16);
        else
            k3_fold<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ctx->d_r, ctx->d_c, ctx->d_b, ctx->rows, ctx->D,
                                                                     ctx->Dp, ctx->d_K32);
        LAUNCH_CHECK("k3_fold");
    }
    CU(cudaStreamSynchronize(st));
    ctx->tables_set = true;
    return HULK_B200_OK;
}

// parity tap: the float64 tables as they sit on the device (rows x num_bins each)
int hulk_b200_get_cws_tables(hulk_b200_ctx *ctx, double *r, double *c, double *b) {
    if (!ctx || !r || !c || !b) return HULK_B200_EARG;
    if (!ctx->tables_set) return fail(ctx, HULK_B200_ESTATE, "CWS tables not set");
    CU(cudaSetDevice(ctx->P.device));
    { const int rc = sync_all(ctx); if (rc) return rc; }
    const size_t bytes = sizeof(double) * (size_t)ctx->rows * (size_t)ctx->D;
    CU(cudaMemcpy(r, ctx->d_r, bytes, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(c, ctx->d_c, bytes, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(b, ctx->d_b, bytes, cudaMemcpyDeviceToHost));
    return HULK_B200_OK;
}

int hulk_b200_generate_cws_tables(hulk_b200_ctx *ctx) {
    if (!ctx) return HULK_B200_EARG;
    CU(cudaSetDevice(ctx->P.device));
    const uint64_t n = (uint64_t)ctx->rows * (uint64_t)ctx->D;
    std::vector<double> r, c, b;
    try {
        r.resize(n ? n : 1); c.resize(n ? n : 1); b.resize(n ? n : 1);
    } catch (const std::bad_alloc &) {
        return fail(ctx, HULK_B200_ENOMEM, "host memory for the CWS tables");
    }
    const int rc = hulk_b200_new_cws(ctx->s, ctx->D, ctx->P.slot_begin, ctx->P.slot_end, r.data(), c.data(), b.data());
    if (rc) return fail(ctx, rc);
    return upload_tables(ctx, r.data(), c.data(), b.data());
}

// ------------------------------------------------------------------------------------------
// stage 1+2
// ------------------------------------------------------------------------------------------
static void k1_geometry(const hulk_b200_ctx *ctx, uint64_t n_reads, uint64_t total_bytes, bool w9, uint32_t *tile_cap,
                        uint32_t *list_cap, size_t *smem) {
    const double avg = n_reads ? (double)total_bytes / (double)n_reads : 0.0;
    uint64_t tc = (uint64_t)(avg * K1_TPB * 1.05) + 64;
    tc = (tc + 15) & ~15ull;
    if (tc > 64 * 1024) tc = 64 * 1024;
    const double nk = std::max(1.0, avg - ctx->P.k + 1);
    uint32_t lc = (uint32_t)(0.28 * nk * (9.0 + 1.0) / (ctx->P.w + 1.0)) + 5;    // ~2/(w+1) candidates per k-mer
    // (the w = 9 kernels keep nothing but the lists in shared memory: batches of reads of several hundred bases get
    // lists of up to 192 entries -- one CTA per SM then, still an order of magnitude faster than k1_generic)
    lc = std::max(16u, std::min(lc, w9 ? ctx->list_cap_w9 : 96u));
    *tile_cap = (uint32_t)tc;
    *list_cap = lc;
    // tile (+16 bytes slack) | window buffers [w + 1][128] | lists [4 warps][lc][32] | mbarrier
    *smem = (size_t)tc + 16 + (size_t)(ctx->P.w + 1) * K1_TPB * 8 + (size_t)lc * K1_TPB * 8 + 64;
}

// enqueue the minimizer/histogram kernels over one device-resident batch
// hs: which spectrum buffer / overflow set to use; st: the stream to enqueue on
template <bool DUMP>
static int launch_k1_one(hulk_b200_ctx *ctx, int hs, cudaStream_t st, const uint8_t *d_bases, uint64_t bases_bytes,
                         const uint64_t *d_offsets, uint64_t off_base, uint32_t fixed_len, uint64_t n_reads,
                         uint64_t total_bytes, uint64_t *d_dump, uint32_t dump_cap, uint32_t *d_dump_counts,
                         uint64_t first_read);

// A batch of any size: launches of at most max_launch_reads reads each (bounds the minimizer queue and keeps
// its indices in 32 bits), back to back on the same stream.
template <bool DUMP>
static int launch_k1(hulk_b200_ctx *ctx, int hs, cudaStream_t st, const uint8_t *d_bases, uint64_t bases_bytes,
                     const uint64_t *d_offsets, uint64_t off_base, uint32_t fixed_len, uint64_t n_reads,
                     uint64_t total_bytes, uint64_t *d_dump, uint32_t dump_cap, uint32_t *d_dump_counts) {
    const uint64_t step = ctx->max_launch_reads;
    if (n_reads <= step)
        return launch_k1_one<DUMP>(ctx, hs, st, d_bases, bases_bytes, d_offsets, off_base, fixed_len, n_reads,
                                   total_bytes, d_dump, dump_cap, d_dump_counts, 0);
    const uint64_t avg = total_bytes / n_reads + 1;
    for (uint64_t r0 = 0; r0 < n_reads; r0 += step) {
        const uint64_t nr = std::min(step, n_reads - r0);
        int rc;
        if (fixed_len) {      // the sub-batch is its own byte range
            const uint64_t skip = r0 * (uint64_t)fixed_len;
            rc = launch_k1_one<DUMP>(ctx, hs, st, d_bases + skip, bases_bytes > skip ? bases_bytes - skip : 0, nullptr,
                                     off_base, fixed_len, nr, nr * (uint64_t)fixed_len,
                                     d_dump ? d_dump + r0 * dump_cap : nullptr, dump_cap,
                                     d_dump_counts ? d_dump_counts + r0 : nullptr, r0);
        } else {              // same byte range, the offsets array is entered further in
            rc = launch_k1_one<DUMP>(ctx, hs, st, d_bases, bases_bytes, d_offsets + r0, off_base, 0, nr, nr * avg,
                                     d_dump ? d_dump + r0 * dump_cap : nullptr, dump_cap,
                                     d_dump_counts ? d_dump_counts + r0 : nullptr, r0);
        }
        if (rc) return rc;
    }
    return HULK_B200_OK;
}

template <bool DUMP>
static int launch_k1_one(hulk_b200_ctx *ctx, int hs, cudaStream_t st, const uint8_t *d_bases, uint64_t bases_bytes,
                         const uint64_t *d_offsets, uint64_t off_base, uint32_t fixed_len, uint64_t n_reads,
                         uint64_t total_bytes, uint64_t *d_dump, uint32_t dump_cap, uint32_t *d_dump_counts,
                         uint64_t first_read) {
    if (n_reads == 0) return HULK_B200_OK;
    K1Params p{};
    p.bases = d_bases;
    p.bases_bytes = bases_bytes;
    p.offsets = d_offsets;
    p.fixed_len = fixed_len;
    p.n_reads = n_reads;
    p.read_base = ctx->st.n_reads + first_read;
    p.off_base = off_base;
    p.k = ctx->P.k;
    p.w = ctx->P.w;
    p.D = ctx->D;
    p.hist = ctx->d_hist[hs];
    p.n_minimizers = ctx->d_nmin;
    p.err_word = ctx->d_errword;
    p.ovf_count = ctx->d_ovf_count[hs];
    p.ovf_list = ctx->d_ovf_list[hs];
    p.ovf_cap = ctx->ovf_cap;
    p.dump = d_dump;
    p.dump_cap = dump_cap;
    p.dump_counts = d_dump_counts;
    // sequences of K1_LONG_MIN bases or more go to the sliced scan (k1_long.cuh) whenever the host cannot rule them
    // out: each takes a table of at most four entries per base behind k1_generic's allocations
    const bool long_on = ctx->long_path && (ctx->batch_max_len ? ctx->batch_max_len >= ctx->long_min
                                                               : (d_offsets != nullptr && fixed_len == 0));
    p.long_min = long_on ? ctx->long_min : ~0ull;
    p.long_seg = k1_long_seg(ctx->P.k, ctx->P.w);
    if (long_on) {
        // room for every sequence of the launch that can be that long
        const uint64_t need = std::min<uint64_t>(n_reads, total_bytes / ctx->long_min + 2);
        if (need > ctx->long_cap[hs]) {
            { const int rc = sync_all(ctx); if (rc) return rc; }
            const uint64_t cap = std::max<uint64_t>(need, 4096);
            for (int i = 0; i < ctx->nbuf; i++) {
                if (cap <= ctx->long_cap[i]) continue;
                if (ctx->d_long_tasks[i]) cudaFree(ctx->d_long_tasks[i]);
                ctx->d_long_tasks[i] = nullptr;
                ctx->long_cap[i] = 0;
                CU(dmalloc(&ctx->d_long_tasks[i], cap));
                ctx->long_cap[i] = cap;
            }
        }
    }
    p.long_cap = (uint32_t)std::min<uint64_t>(ctx->long_cap[hs], 0xffffffffull);
    p.long_tasks = ctx->d_long_tasks[hs];
    p.long_ctl = ctx->d_long_ctl[hs];
    // scratch arena of the generic path: 4 table slots per base of the batch, at least 4 Mi entries
    uint64_t want = std::max<uint64_t>(1ull << 22, 4 * total_bytes + (n_reads << 7));
    // only reserve the large arena when the generic path will take whole batches ...
    uint64_t arena_cap = 1ull << 25;
    if (!long_on && ctx->batch_max_len > (1u << 20)) {
        // ... or a very long sequence (a chromosome in --fasta mode) is in the batch and the sliced scan is off: room
        // for its open-addressing table (two slots per k-mer, a power of two), twice over
        uint64_t t = 64;
        while (t < 2 * ctx->batch_max_len) t <<= 1;
        arena_cap = std::max<uint64_t>(arena_cap, 2 * t);
    }
    if (ctx->P.w <= (uint32_t)K1_W_FAST) want = std::min<uint64_t>(want, arena_cap);
    if (long_on && ctx->batch_max_len) want += ctx->batch_long_entries + 8;
    // in front of that: one table per k1_generic thread, reused read after read -- however many reads of a few hundred
    // to a few thousand bases a batch holds, their sets never run the arena dry.  Its size follows the longest read of
    // the batch (two entries per k-mer, a power of two) up to K1_SLAB_MAX; sets that need more use the part behind
    const bool fast = ctx->P.w <= (uint32_t)K1_W_FAST;
    const unsigned generic_grid = fast ? (unsigned)(ctx->sm_count * 2)
                                       : (unsigned)std::min<uint64_t>((n_reads + 63) / 64, (uint64_t)ctx->sm_count * 8);
    p.slab_size = 2048;
    if (ctx->batch_max_len) {
        const uint64_t nk = ctx->batch_max_len >= ctx->P.k ? ctx->batch_max_len - ctx->P.k + 1 : 1;
        p.slab_size = 64;
        while (p.slab_size < K1_SLAB_MAX && p.slab_size < 2 * nk) p.slab_size <<= 1;
    }
    p.slab_entries = (uint64_t)generic_grid * 64 * p.slab_size;
    want += p.slab_entries;
    if (want > ctx->arena_entries[hs]) {
        // grow the scratch of EVERY spectrum buffer at once: the next intervals will need the same, and an
        // allocation (a device-wide synchronisation) belongs in front of the pipeline, not inside it
        // (gigabyte-sized arenas for single long sequences are grown for the buffer in use only)
        { const int rc = sync_all(ctx); if (rc) return rc; }
        for (int i = 0; i < ctx->nbuf; i++) {
            if (want > (1ull << 26) && i != hs) continue;
            if (want <= ctx->arena_entries[i]) continue;
            if (ctx->d_arena[i]) cudaFree(ctx->d_arena[i]);
            ctx->d_arena[i] = nullptr;
            ctx->arena_entries[i] = 0;
            CU(dmalloc(&ctx->d_arena[i], want));
            ctx->arena_entries[i] = want;
        }
    }
    // minimizer queue of the batch: at most list_cap keys per read (longer lists go to k1_generic)
    // (with the MinHash feed on, every kernel's minimizers go through the queue: k1_generic and k1_long_scan append
    // theirs -- at most one per k-mer -- behind the fast kernels')
    const bool feed = !DUMP && (ctx->mh_kmv || ctx->mh_khf);
    const bool use_queue = fast && !DUMP && (!ctx->fused_jump || feed);
    size_t smem = 0;
    if (fast) k1_geometry(ctx, n_reads, total_bytes, ctx->P.w == 9 && !ctx->force_tile_path, &p.tile_cap, &p.list_cap, &smem);
    p.feed_queue = feed ? 1u : 0u;
    if (use_queue || feed) {
        const uint64_t need = (use_queue ? n_reads * (uint64_t)p.list_cap : 0) + (feed ? total_bytes + n_reads : 0);
        if (need >= (1ull << 31)) return fail(ctx, HULK_B200_EARG, "batch too large for one minimizer queue");
        if (need > ctx->queue_cap[hs]) {
            { const int rc = sync_all(ctx); if (rc) return rc; }
            for (int i = 0; i < ctx->nbuf; i++) {
                if (need <= ctx->queue_cap[i]) continue;
                if (ctx->d_queue[i]) cudaFree(ctx->d_queue[i]);
                ctx->d_queue[i] = nullptr;
                ctx->queue_cap[i] = 0;
                CU(dmalloc(&ctx->d_queue[i], need));
                ctx->queue_cap[i] = need;
            }
        }
        p.queue = ctx->d_queue[hs];
        p.queue_cursor = ctx->d_queue_cursor[hs];
        p.queue_cap = ctx->queue_cap[hs];
        CU(cudaMemsetAsync(ctx->d_queue_cursor[hs], 0, 16, st));
        if (feed && ctx->mh_kmv && ctx->kmv_cand_cap[hs] < ctx->queue_cap[hs]) {
            { const int rc = sync_all(ctx); if (rc) return rc; }
            for (int i = 0; i < ctx->nbuf; i++) {
                if (ctx->kmv_cand_cap[i] >= ctx->queue_cap[i]) continue;
                if (ctx->d_kmv_cand[i]) cudaFree(ctx->d_kmv_cand[i]);
                ctx->d_kmv_cand[i] = nullptr;
                ctx->kmv_cand_cap[i] = 0;
                CU(dmalloc(&ctx->d_kmv_cand[i], ctx->queue_cap[i]));
                ctx->kmv_cand_cap[i] = ctx->queue_cap[i];
            }
        }
    }
    p.arena = ctx->d_arena[hs];
    p.arena_cursor = ctx->d_arena_cursor[hs];
    p.arena_entries = ctx->arena_entries[hs];
    CU(cudaMemsetAsync(ctx->d_ovf_count[hs], 0, 4, st));
    CU(cudaMemsetAsync(ctx->d_arena_cursor[hs], 0, 8, st));
    ProfScope prof_scope(ctx, 0, st);
    if (fast) {
        const uint64_t ntiles = (n_reads + K1_TPB - 1) / K1_TPB;
        int per_sm = (int)std::max<size_t>(1, std::min<size_t>(16, (227 * 1024) / (smem + 1024)));
        const unsigned grid = (unsigned)std::min<uint64_t>(ntiles, (uint64_t)ctx->sm_count * per_sm);
        const bool fp = k1_fp_compare_ok((int32_t)ctx->P.k, (int32_t)ctx->P.w);
        // template dispatch: <DUMP, FP-compares, QUEUE>
#define K1_DISPATCH(KERNEL, GRID, SMEM)                                                      \
        do {                                                                                  \
            if (use_queue) {                                                                  \
                if (fp) KERNEL<false, true, true><<<GRID, K1_TPB, SMEM, st>>>(p);             \
                else KERNEL<false, false, true><<<GRID, K1_TPB, SMEM, st>>>(p);               \
            } else {                                                                          \
                if (fp) KERNEL<DUMP, true, false><<<GRID, K1_TPB, SMEM, st>>>(p);             \
                else KERNEL<DUMP, false, false><<<GRID, K1_TPB, SMEM, st>>>(p);               \
            }                                                                                 \
        } while (0)
        if (ctx->P.w == 9 && !ctx->force_tile_path) {
            // register-resident window, no staged tile: shared memory is the candidate lists only
            const size_t smem9 = (size_t)p.list_cap * K1_TPB * 8;
            const uint64_t nctas = (n_reads + K1_TPB - 1) / K1_TPB;
            const int per_sm9 = use_queue ? ctx->k1_ctas_per_sm : K1_W9_CTAS_PER_SM;
            const unsigned grid9 = (unsigned)std::min<uint64_t>(nctas, (uint64_t)ctx->sm_count * per_sm9);
            // odd k >= 17: the second-generation scan, k folded in at compile time
            bool done = false;
            if (ctx->k1_v2 && (use_queue || DUMP)) {
                // one 32-read task per warp, one short-lived CTA per four tasks: the block scheduler balances them, and
                // the slots they free every few tens of microseconds go to the flush chain of the previous interval
                // (higher stream priority) instead of staying with a persistent counting CTA for the whole interval
                const unsigned gridv = ctx->k1_persistent ? grid9 : (unsigned)std::min<uint64_t>(nctas, 1u << 24);
                switch (ctx->P.k) {
#define K1_V2_CASE(KK)                                                                              \
                    case KK:                                                                        \
                        if (use_queue) k1_scan_w9_v2<false, true, KK><<<gridv, K1_TPB, smem9, st>>>(p); \
                        else k1_scan_w9_v2<true, false, KK><<<gridv, K1_TPB, smem9, st>>>(p);       \
                        done = true;                                                                \
                        break;
                    K1_V2_FOR_EACH_K(K1_V2_CASE)
#undef K1_V2_CASE
                    default: break;
                }
            }
            // the k values of the BASELINE configs get the scan with k folded in at compile time
            if (done) {
            } else if (use_queue && fp && ctx->P.k == 21)
                k1_minimizer_histogram_w9<false, true, true, 21><<<grid9, K1_TPB, smem9, st>>>(p);
            else if (use_queue && fp && ctx->P.k == 11)
                k1_minimizer_histogram_w9<false, true, true, 11><<<grid9, K1_TPB, smem9, st>>>(p);
            else if (use_queue && !fp && ctx->P.k == 31)
                k1_minimizer_histogram_w9<false, false, true, 31><<<grid9, K1_TPB, smem9, st>>>(p);
            else
                K1_DISPATCH(k1_minimizer_histogram_w9, grid9, smem9);
        } else {
            K1_DISPATCH(k1_minimizer_histogram, grid, smem);
        }
#undef K1_DISPATCH
        LAUNCH_CHECK("k1_minimizer_histogram");
        if (use_queue) {
            const unsigned gridj = (unsigned)(ctx->sm_count * ctx->jump_ctas_per_sm);
            if (ctx->jump_fx && (uint32_t)ctx->D <= JUMP_FX_MAX_BUCKETS) {          // every k^4-bin spectrum
                if (ctx->jump_smem) {
                    if (ctx->jump_batch == 3) k1_jump_queue_fx<3, true><<<gridj, K1_JUMP_TPB, 0, st>>>(p);
                    else k1_jump_queue_fx<4, true><<<gridj, K1_JUMP_TPB, 0, st>>>(p);
                } else if (ctx->jump_batch == 2) k1_jump_queue_fx<2, false><<<gridj, K1_JUMP_TPB, 0, st>>>(p);
                else if (ctx->jump_batch == 3) k1_jump_queue_fx<3, false><<<gridj, K1_JUMP_TPB, 0, st>>>(p);
                else k1_jump_queue_fx<4, false><<<gridj, K1_JUMP_TPB, 0, st>>>(p);
            } else if (ctx->jump_batch == 2) k1_jump_queue<2><<<gridj, K1_JUMP_TPB, 0, st>>>(p);
            else k1_jump_queue<4><<<gridj, K1_JUMP_TPB, 0, st>>>(p);
            LAUNCH_CHECK("k1_jump_queue");
        }
        k1_generic<DUMP><<<generic_grid, 64, 0, st>>>(p, true);
        LAUNCH_CHECK("k1_generic");
    } else {
        k1_generic<DUMP><<<generic_grid, 64, 0, st>>>(p, false);
        LAUNCH_CHECK("k1_generic");
    }
    if (long_on) {
        k1_long_plan<<<1, K1_LONG_PLAN_TPB, 0, st>>>(p, fast);
        LAUNCH_CHECK("k1_long_plan");
        k1_long_zero<<<ctx->sm_count * 4, 256, 0, st>>>(p);
        LAUNCH_CHECK("k1_long_zero");
        k1_long_scan<DUMP><<<ctx->sm_count * 8, K1_LONG_TPB, 0, st>>>(p);
        LAUNCH_CHECK("k1_long_scan");
    }
    if (feed) {
        K1MinhashParams m{};
        m.queue = p.queue;
        m.queue_cursor = p.queue_cursor;
        m.queue_cap = p.queue_cap;
        m.s = ctx->s;
        m.khf = ctx->d_khf;
        m.kmv = ctx->d_kmv_state[hs];
        m.kmv_pool = ctx->d_kmv_pool[hs];
        m.kmv_cand = ctx->d_kmv_cand[hs];
        if (ctx->mh_khf) {
            k1_khf_queue<<<ctx->sm_count * 4, K1_KHF_TPB, 0, st>>>(m);
            LAUNCH_CHECK("k1_khf_queue");
        }
        if (ctx->mh_kmv) {
            k1_kmv_filter<<<ctx->sm_count * 4, 256, 0, st>>>(m);
            LAUNCH_CHECK("k1_kmv_filter");
            k1_kmv_select<<<1, K1_KMV_SELECT_TPB, 0, st>>>(m);
            LAUNCH_CHECK("k1_kmv_select");
        }
    }
    return HULK_B200_OK;
}

static int feed_drain(hulk_b200_ctx *ctx);
// (Re)allocations free device memory, which waits for the whole device, under a lock the feeder thread's own CUDA calls
// need: every one of them is preceded by feed_drain -- no kernel is then waiting for a batch the feeder still owes.
static int ensure_stage(hulk_b200_ctx *ctx, int buf, uint64_t bytes, uint64_t n_off) {
    if (bytes + 64 <= ctx->stage_cap[buf] && n_off <= ctx->off_cap[buf]) return HULK_B200_OK;
    { const int rcd = feed_drain(ctx); if (rcd) return rcd; }
    // the next batches will be as large: grow EVERY stage buffer now -- an allocation belongs in front of the pipeline,
    // not inside its first few steps
    for (int i = 0; i < NSTAGE; i++) {
        if (bytes + 64 > ctx->stage_cap[i]) {
            CU(cudaEventSynchronize(ctx->ev_k1_prev(i)));
            if (ctx->d_stage[i]) cudaFree(ctx->d_stage[i]);
            ctx->d_stage[i] = nullptr;
            ctx->stage_cap[i] = 0;
            const uint64_t cap = ((bytes + 64 + 4095) & ~4095ull);
            CU(dmalloc(&ctx->d_stage[i], cap));
            ctx->stage_cap[i] = cap;
        }
        if (n_off > ctx->off_cap[i]) {
            CU(cudaEventSynchronize(ctx->ev_k1_prev(i)));
            if (ctx->d_off[i]) cudaFree(ctx->d_off[i]);
            ctx->d_off[i] = nullptr;
            ctx->off_cap[i] = 0;
            CU(dmalloc(&ctx->d_off[i], n_off));
            ctx->off_cap[i] = n_off;
        }
    }
    return HULK_B200_OK;
}

// The first counting kernel into spectrum buffer hs since its last flush: the buffer must be clean.  One GPU: the flush
// chain wiped it (k2_finalize) -- wait for that.  Several GPUs: the flush chain worked on the SUM; the buffer itself is
// wiped here, once every peer has gathered it.
static int before_first_k1(hulk_b200_ctx *ctx, int hs, cudaStream_t ks) {
    if (ctx->k1_pending[hs]) return HULK_B200_OK;
    if (ctx->world == 1) {
        CU(cudaStreamWaitEvent(ks, ctx->ev_hist_free[hs], 0));
        return HULK_B200_OK;
    }
    if (ctx->peer_dirty[hs]) {
        const uint32_t *flags = reinterpret_cast<const uint32_t *>(ctx->arena + ctx->off_gathered) + (size_t)hs * PEER_MAX;
        k_peer_wait<<<1, 32, 0, ks>>>(flags, ctx->world, ctx->peer_dirty[hs], ctx->d_ctl, ctx->peer_timeout_ns);
        LAUNCH_CHECK("k_peer_wait");
        CU(cudaMemsetAsync(ctx->d_hist[hs], 0, sizeof(uint32_t) * (size_t)ctx->D, ks));
        ctx->peer_dirty[hs] = 0;
    }
    return HULK_B200_OK;
}

// What the host knows about the lengths of the batch it is about to launch: the longest read, and the room the sets of
// its long sequences need (offsets == nullptr: n_reads reads of fixed_len bases).
static void note_batch_lengths(hulk_b200_ctx *ctx, const uint64_t *offsets, uint64_t n_reads, uint32_t fixed_len) {
    ctx->batch_max_len = fixed_len;
    ctx->batch_long_entries = 0;
    if (offsets) {
        for (uint64_t i = 0; i < n_reads; i++) {
            const uint64_t len = offsets[i + 1] - offsets[i];
            ctx->batch_max_len = std::max(ctx->batch_max_len, len);
            if (len >= ctx->long_min) ctx->batch_long_entries += k1_long_table_entries(len, (int32_t)ctx->P.k);
        }
    } else if (fixed_len >= ctx->long_min) {
        ctx->batch_long_entries = n_reads * k1_long_table_entries(fixed_len, (int32_t)ctx->P.k);
    }
}

static const uint64_t kMaxBatchBytes = 256ull << 20;   // staging granularity of one H2D copy + k1 launch

// ---- packed transport, device side: 2-bit codes -> the letters the scan kernels read ---------------------
// One thread per packed word = 16 bases -> one 16-byte store.  Code c becomes "ACGT"[c]; k0_patch then writes 'N'
// over the listed code-4 positions (seq_nt4_table maps every such byte to 4, minimizer.go:13-30).
__global__ void k0_unpack(const uint32_t *__restrict__ packed, uint64_t n_words, uint4 *__restrict__ ascii) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_words) return;
    const uint32_t x = packed[t];
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        uint32_t v = (x >> (8 * q)) & 0xffu;
        v = (v | (v << 4)) & 0x0F0Fu;
        v = (v | (v << 2)) & 0x3333u;                        // nibble j = code of base j
        o[q] = __byte_perm(0x54474341u /* "ACGT" */, 0u, v);
    }
    ascii[t] = make_uint4(o[0], o[1], o[2], o[3]);
}
__global__ void k0_patch(const uint32_t *__restrict__ exc, uint64_t n, uint32_t shift, uint8_t *__restrict__ ascii) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) ascii[exc[t] - shift] = (uint8_t)'N';
}

// the same two kernels for a batch the feeder thread delivers: whether it arrived packed (meta[1]) and how many
// exceptions it has (meta[2]) is only known on the device by the time they run
__global__ void k0_unpack_fed(const uint32_t *__restrict__ packed, uint64_t n_words, uint4 *__restrict__ ascii,
                              const uint32_t *__restrict__ meta) {
    // meta[3]: packed words at the head of the batch (0: it travelled as letters and is in place already); the rest of the
    // batch arrived as letters behind them
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_words || t >= (uint64_t)meta[3]) return;
    const uint32_t x = packed[t];
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        uint32_t v = (x >> (8 * q)) & 0xffu;
        v = (v | (v << 4)) & 0x0F0Fu;
        v = (v | (v << 2)) & 0x3333u;
        o[q] = __byte_perm(0x54474341u, 0u, v);
    }
    ascii[t] = make_uint4(o[0], o[1], o[2], o[3]);
}
__global__ void k0_patch_fed(const uint32_t *__restrict__ exc, const uint32_t *__restrict__ meta, uint8_t *__restrict__ ascii) {
    if (meta[1] == 0u) return;
    const uint32_t n = meta[2];
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) ascii[exc[t]] = (uint8_t)'N';
}

static uint64_t feed_exc_cap(uint64_t nb) { return std::max<uint64_t>(1024, nb / 32); }
static const bool g_feed_debug = [] { const char *e = getenv("HULK_B200_FEED_DEBUG"); return e && *e == '1'; }();
#define FEED_DBG(...) do { if (g_feed_debug) { fprintf(stderr, "[feed] " __VA_ARGS__); fputc('\n', stderr); fflush(stderr); } } while (0)

// body of the feeder thread: pack -> copy -> publish the batch's sequence number, request by request
static void feeder_main(hulk_b200_ctx *ctx) {
    cudaSetDevice(ctx->P.device);
    for (;;) {
        hulk_b200_ctx::FeedReq rq;
        auto tick = std::chrono::steady_clock::now();
        auto lap = [&](std::atomic<uint64_t> &acc) {
            const auto now = std::chrono::steady_clock::now();
            acc += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(now - tick).count();
            tick = now;
        };
        {
            std::unique_lock<std::mutex> lk(ctx->feed_mu);
            ctx->feed_cv.wait(lk, [&] { return ctx->feed_stop || !ctx->feed_q.empty(); });
            if (ctx->feed_q.empty()) return;                  // stop requested and nothing left
            rq = ctx->feed_q.front();
            ctx->feed_q.pop_front();
        }
        lap(ctx->feed_t_idle);
        int rc = HULK_B200_OK;
        std::string msg;
        auto cu = [&](cudaError_t e, const char *what) {
            if (e != cudaSuccess && rc == HULK_B200_OK) {
                rc = e == cudaErrorMemoryAllocation ? HULK_B200_ENOMEM : HULK_B200_ECUDA;
                msg = std::string(hulk_b200_strerror(rc)) + ": feeder " + what + " -> " + cudaGetErrorString(e);
            }
        };
        const int buf = rq.buf;
        const uint64_t exc_cap = feed_exc_cap(rq.nb);
        FEED_DBG("feeder: request buf %d seq %u, %llu bases: waiting for the pinned buffer", buf, rq.seq, (unsigned long long)rq.nb);
        cu(cudaEventSynchronize(ctx->ev_copy[buf]), "cudaEventSynchronize");      // the pinned buffers are free again
        lap(ctx->feed_t_evsync);
        FEED_DBG("feeder: packing");
        // (the pinned buffers were sized by the calling thread before it posted the request: nothing in this loop may
        // allocate -- an allocation can wait for the device, and the device may be waiting for this loop)
        //
        // Two resources move a batch: the host's cores (packing, bound by their memory bandwidth) and the link.  Packing
        // everything leaves the link idle three quarters of the time, so the batch is split: its TAIL travels as letters
        // -- that copy is enqueued first and runs while the HEAD is being packed -- and the head travels packed.  The
        // split point follows what is observed at the end of every pack call: if the tail had already arrived, the cores
        // are the bottleneck and the next batch packs less; if it was still on the link, more.  The step starts at 1/16
        // and halves (down to 1/64) whenever the direction turns.  (A split computed from the two measured rates,
        // f = Bp / (Bl + 0.75 Bp), was tried: both rates sag when the other side runs -- they share the host's memory
        // system -- and it settled 5-8 % slower at one rank per host.  With HULK_B200_FEED_STATS=1 the rates are still measured.)
        uint64_t n_exc = 0, moved = 0;
        if (g_host_stats && ctx->feed_last_tail >= 0 && cudaEventQuery(ctx->ev_tail[ctx->feed_last_tail]) == cudaSuccess) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, ctx->ev_tail0[ctx->feed_last_tail], ctx->ev_tail[ctx->feed_last_tail]) == cudaSuccess &&
                ms > 0.f) {
                const double bl = (double)ctx->feed_last_tail_bytes / (ms * 1e-3);
                ctx->feed_bl = ctx->feed_bl > 0.0 ? 0.75 * ctx->feed_bl + 0.25 * bl : bl;
            }
            ctx->feed_last_tail = -1;
        }
        cudaGetLastError();                                           // cudaErrorNotReady is not an error
        uint64_t n_head = (uint64_t)(ctx->feed_frac * (double)rq.nb) & ~63ull;     // bases that travel packed
        if (n_head < 4096) n_head = 0;
        if (rq.nb - n_head < 4096) n_head = rq.nb;
        // the copy stream holds nothing but this thread's copies, in request order: each waits for the kernels that last
        // read its stage buffer and for nothing else
        cu(cudaStreamWaitEvent(ctx->copy_stream, rq.stage_free, 0), "cudaStreamWaitEvent");
        const bool has_tail = n_head < rq.nb;
        if (rc == HULK_B200_OK && has_tail) {
            if (g_host_stats) cu(cudaEventRecord(ctx->ev_tail0[buf], ctx->copy_stream), "cudaEventRecord");   // link rate, for the record
            cu(cudaMemcpyAsync(ctx->d_stage[buf] + n_head, rq.src + n_head, rq.nb - n_head, cudaMemcpyHostToDevice,
                               ctx->copy_stream), "cudaMemcpyAsync");
            cu(cudaEventRecord(ctx->ev_tail[buf], ctx->copy_stream), "cudaEventRecord");
            moved += rq.nb - n_head;
            if (g_host_stats && rq.nb - n_head >= (1u << 18)) {       // long enough to time
                ctx->feed_last_tail = buf;
                ctx->feed_last_tail_bytes = rq.nb - n_head;
            }
        }
        bool packed = false;
        const uint64_t packed_head_bytes = (n_head + 3) / 4;
        if (rc == HULK_B200_OK && n_head) {
            const auto t0 = std::chrono::steady_clock::now();
            const int prc = hulk_b200_pack_bases(rq.src, n_head, ctx->h_pack[buf], ctx->h_exc[buf], exc_cap, &n_exc,
                                                 ctx->pack_threads);
            const uint64_t ns = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(
                                    std::chrono::steady_clock::now() - t0).count();
            ctx->feed_pack_ns += ns;
            if (prc) { rc = prc; msg = std::string(hulk_b200_strerror(prc)) + ": pack_bases"; }
            packed = prc == HULK_B200_OK && n_exc <= exc_cap;
            if (ns && n_head >= (1u << 18)) {
                const double bp = (double)n_head / (ns * 1e-9);
                ctx->feed_bp = ctx->feed_bp > 0.0 ? 0.75 * ctx->feed_bp + 0.25 * bp : bp;
            }
            if (ctx->feed_adapt && rc == HULK_B200_OK) {
                const bool tail_arrived = !has_tail || cudaEventQuery(ctx->ev_tail[buf]) == cudaSuccess;
                cudaGetLastError();
                const int dir = tail_arrived ? -1 : 1;
                if (dir != ctx->feed_dir) ctx->feed_step = std::max(ctx->feed_step * 0.5, 1.0 / 64.0);
                ctx->feed_dir = dir;
                ctx->feed_frac = std::min(1.0, std::max(0.0, ctx->feed_frac + dir * ctx->feed_step));
            }
        } else if (ctx->feed_adapt && n_head == 0) {
            ctx->feed_frac = 1.0 / 16.0;                              // try packing a little again
        }
        ctx->feed_frac_sum += (double)n_head / (double)std::max<uint64_t>(1, rq.nb);
        FEED_DBG("feeder: packed %llu of %llu bases (rc %d, %llu exceptions), copying", (unsigned long long)n_head,
                 (unsigned long long)rq.nb, rc, (unsigned long long)n_exc);
        tick = std::chrono::steady_clock::now();
        if (rc == HULK_B200_OK) {
            if (packed) {
                cu(cudaMemcpyAsync(ctx->d_pack[buf], ctx->h_pack[buf], packed_head_bytes, cudaMemcpyHostToDevice, ctx->copy_stream),
                   "cudaMemcpyAsync");
                if (n_exc)
                    cu(cudaMemcpyAsync(ctx->d_exc[buf], ctx->h_exc[buf], sizeof(uint32_t) * n_exc, cudaMemcpyHostToDevice,
                                       ctx->copy_stream), "cudaMemcpyAsync");
                moved += packed_head_bytes + sizeof(uint32_t) * n_exc;
                ctx->feed_batches++;
            } else if (n_head) {                              // too many foreign bytes: the letters travel as they are
                cu(cudaMemcpyAsync(ctx->d_stage[buf], rq.src, n_head, cudaMemcpyHostToDevice, ctx->copy_stream), "cudaMemcpyAsync");
                moved += n_head;
                n_exc = 0;
            }
            if (rq.offsets) {
                cu(cudaMemcpyAsync(ctx->d_off[buf], rq.offsets, sizeof(uint64_t) * rq.n_off, cudaMemcpyHostToDevice,
                                   ctx->copy_stream), "cudaMemcpyAsync");
                moved += sizeof(uint64_t) * rq.n_off;
            }
        }
        // mode and exception count, then the sequence number that releases the kernels waiting for this batch.
        // The number is published whatever happened: a device stream must never be left waiting for it.
        uint32_t *hm = ctx->h_feed + 4 * buf;
        hm[1] = (rc == HULK_B200_OK && packed) ? 1u : 0u;
        hm[2] = (uint32_t)n_exc;
        hm[3] = (rc == HULK_B200_OK && packed) ? (uint32_t)((n_head + 15) / 16) : 0u;       // packed words to unpack
        cu(cudaMemcpyAsync(ctx->d_feed + 4 * buf + 1, hm + 1, 3 * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->copy_stream),
           "cudaMemcpyAsync");
        const CUresult wr = ctx->cu_write32(reinterpret_cast<CUstream>(ctx->copy_stream),
                                            reinterpret_cast<CUdeviceptr>(ctx->d_feed + 4 * buf), rq.seq, 0);
        if (wr != CUDA_SUCCESS && rc == HULK_B200_OK) {
            rc = HULK_B200_ECUDA;
            msg = std::string(hulk_b200_strerror(rc)) + ": feeder cuStreamWriteValue32 failed";
        }
        cudaEventRecord(ctx->ev_copy[buf], ctx->copy_stream);
        lap(ctx->feed_t_enq);
        FEED_DBG("feeder: published seq %u for buf %d (rc %d)", rq.seq, buf, rc);
        ctx->feed_h2d += moved;
        {
            std::lock_guard<std::mutex> lk(ctx->feed_mu);
            ctx->feed_done++;
            if (rc != HULK_B200_OK && ctx->feed_rc == HULK_B200_OK) {
                ctx->feed_rc = rc;
                ctx->feed_err = msg;
            }
        }
        ctx->feed_cv.notify_all();
    }
}

static int ensure_pack_stage(hulk_b200_ctx *ctx, int buf, uint64_t packed_bytes, uint64_t n_exc, bool host_side) {
    if (!(packed_bytes + 16 > ctx->d_pack_cap[buf] || n_exc > ctx->d_exc_cap[buf] ||
          (host_side && (packed_bytes > ctx->h_pack_cap[buf] || n_exc > ctx->h_exc_cap[buf]))))
        return HULK_B200_OK;
    { const int rcd = feed_drain(ctx); if (rcd) return rcd; }
    for (int i = 0; i < NSTAGE; i++) {             // every stage buffer at once, like ensure_stage
        if (packed_bytes + 16 > ctx->d_pack_cap[i]) {
            CU(cudaEventSynchronize(ctx->ev_k1_prev(i)));
            if (ctx->d_pack[i]) cudaFree(ctx->d_pack[i]);
            ctx->d_pack[i] = nullptr;
            ctx->d_pack_cap[i] = 0;
            const uint64_t cap = (packed_bytes + 16 + 4095) & ~4095ull;
            CU(dmalloc(&ctx->d_pack[i], cap));
            ctx->d_pack_cap[i] = cap;
        }
        if (n_exc > ctx->d_exc_cap[i]) {
            CU(cudaEventSynchronize(ctx->ev_k1_prev(i)));
            if (ctx->d_exc[i]) cudaFree(ctx->d_exc[i]);
            ctx->d_exc[i] = nullptr;
            ctx->d_exc_cap[i] = 0;
            const uint64_t cap = std::max<uint64_t>(1024, n_exc);
            CU(dmalloc(&ctx->d_exc[i], cap));
            ctx->d_exc_cap[i] = cap;
        }
        if (!host_side) continue;
        if (packed_bytes > ctx->h_pack_cap[i]) {
            CU(cudaEventSynchronize(ctx->ev_copy[i]));
            if (ctx->h_pack[i]) cudaFreeHost(ctx->h_pack[i]);
            ctx->h_pack[i] = nullptr;
            ctx->h_pack_cap[i] = 0;
            const uint64_t cap = (packed_bytes + 4095) & ~4095ull;
            CU(cudaHostAlloc(reinterpret_cast<void **>(&ctx->h_pack[i]), cap, cudaHostAllocDefault));
            ctx->h_pack_cap[i] = cap;
        }
        if (n_exc > ctx->h_exc_cap[i]) {
            CU(cudaEventSynchronize(ctx->ev_copy[i]));
            if (ctx->h_exc[i]) cudaFreeHost(ctx->h_exc[i]);
            ctx->h_exc[i] = nullptr;
            ctx->h_exc_cap[i] = 0;
            CU(cudaHostAlloc(reinterpret_cast<void **>(&ctx->h_exc[i]), n_exc * sizeof(uint32_t), cudaHostAllocDefault));
            ctx->h_exc_cap[i] = n_exc;
        }
    }
    return HULK_B200_OK;
}

// A host batch, in one of two forms: ASCII (`bases`), or packed by the caller (`packed` + `exc`; base 0 of the stream
// is the first base of read 0).  ASCII batches are packed here when the context's input packing is on.
struct HostBatch {
    const uint8_t *bases = nullptr;
    const uint8_t *packed = nullptr;
    const uint32_t *exc = nullptr;
    uint64_t n_exc = 0;
    const uint64_t *offsets = nullptr;
    uint64_t n_reads = 0;
    uint32_t fixed_len = 0;
};

static int push_host(hulk_b200_ctx *ctx, const HostBatch &hb) {
    CU(cudaSetDevice(ctx->P.device));
    const uint64_t *offsets = hb.offsets;
    const uint64_t n_reads = hb.n_reads;
    const uint32_t fixed_len = hb.fixed_len;
    const uint64_t origin = offsets ? offsets[0] : 0;      // offsets[] value of base 0 of a caller-packed stream
    uint64_t done = 0;
    while (done < n_reads) {
        // take reads [done, upto) with at most kMaxBatchBytes bytes (at least one read)
        uint64_t upto;
        uint64_t b0, b1;
        if (offsets) {
            b0 = offsets[done];
            const uint64_t *lim = std::upper_bound(offsets + done + 1, offsets + n_reads + 1, b0 + kMaxBatchBytes);
            upto = (uint64_t)(lim - offsets) - 1;
            if (upto <= done) upto = done + 1;
            b1 = offsets[upto];
            if (b1 < b0) return fail(ctx, HULK_B200_EARG, "offsets must be non-decreasing");
        } else {
            const uint64_t per = std::max<uint64_t>(1, kMaxBatchBytes / std::max<uint32_t>(1, fixed_len));
            upto = std::min(n_reads, done + per);
            b0 = done * (uint64_t)fixed_len;
            b1 = upto * (uint64_t)fixed_len;
        }
        const uint64_t nb = b1 - b0, nr = upto - done;
        note_batch_lengths(ctx, offsets ? offsets + done : nullptr, nr, fixed_len);
        const int buf = ctx->cur_buf;
        int rc = ensure_stage(ctx, buf, nb + 16, offsets ? nr + 1 : 0);
        if (rc) return rc;

        // ---- how the bases travel: packed by the caller, packed here, or as they are
        const uint8_t *src_packed = nullptr;               // host bytes to copy into d_pack
        const uint32_t *src_exc = nullptr;
        uint64_t n_exc = 0, packed_bytes = 0, unpack_words = 0;
        uint32_t skip = 0, exc_shift = 0;
        if (hb.packed) {
            const uint64_t s0 = b0 - origin, s1 = b1 - origin, a0 = s0 & ~3ull;
            if (s1 >= (1ull << 32)) return fail(ctx, HULK_B200_EARG, "a packed batch holds fewer than 2^32 bases");
            skip = (uint32_t)(s0 - a0);
            exc_shift = (uint32_t)a0;
            src_packed = hb.packed + (a0 >> 2);
            packed_bytes = ((s1 + 3) >> 2) - (a0 >> 2);
            unpack_words = (s1 - a0 + 15) / 16;
            const uint32_t *e0 = std::lower_bound(hb.exc, hb.exc + hb.n_exc, (uint32_t)s0);
            const uint32_t *e1 = std::lower_bound(e0, hb.exc + hb.n_exc, (uint32_t)s1);
            src_exc = e0;
            n_exc = (uint64_t)(e1 - e0);
            rc = ensure_pack_stage(ctx, buf, packed_bytes, n_exc, false);
            if (rc) return rc;
        } else if (ctx->pack_threads != 0 && nb >= 4096 && nb < (1ull << 32) && (ctx->P.flags & HULK_B200_F_ASYNC_INPUT) &&
                   ctx->cu_wait32) {                   // (code-4 positions are 32-bit: a larger batch -- one giant sequence -- goes as letters)
            // ---- the feeder thread packs and copies; this thread only enqueues the consumers behind the batch's number
            if (!ctx->feeder.joinable()) ctx->feeder = std::thread(feeder_main, ctx);
            FEED_DBG("push: buf %d, %llu bases", buf, (unsigned long long)nb);
            HostLap hl;
            {
                // never more than NSTAGE - 1 requests ahead of the feeder: the wait enqueued on the copy stream below
                // (for the kernels that last used this stage buffer) must find that batch's copy already enqueued
                const auto tb = std::chrono::steady_clock::now();
                std::unique_lock<std::mutex> lk(ctx->feed_mu);
                ctx->feed_cv.wait(lk, [&] { return ctx->feed_posted - ctx->feed_done <= (uint64_t)(NSTAGE - 1); });
                ctx->feed_t_backpressure += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(
                                                std::chrono::steady_clock::now() - tb).count();
                if (ctx->feed_rc) {
                    const int frc = ctx->feed_rc;
                    ctx->err = ctx->feed_err;
                    ctx->feed_rc = 0;
                    return frc;
                }
            }
            // the feeder is done with this stage buffer's previous batch (4 requests back): size its device and pinned
            // buffers here, before anything of this batch is enqueued
            hl.lap(0);
            rc = ensure_pack_stage(ctx, buf, (nb + 3) / 4, feed_exc_cap(nb), true);
            if (rc) return rc;
            hl.lap(1);
            const uint32_t seq = ++ctx->feed_seq[buf];
            {
                std::lock_guard<std::mutex> lk(ctx->feed_mu);
                ctx->feed_q.push_back(hulk_b200_ctx::FeedReq{hb.bases + b0, nb, offsets ? offsets + done : nullptr,
                                                             offsets ? nr + 1 : 0, buf, seq, ctx->ev_k1_prev(buf)});
                ctx->feed_posted++;
            }
            ctx->feed_cv.notify_all();
            hl.lap(2);
            const int hs = ctx->cur_hist;
            cudaStream_t ks = ctx->overlap ? ctx->k1_stream[hs] : ctx->stream;
            if (ctx->cu_wait32(reinterpret_cast<CUstream>(ks), reinterpret_cast<CUdeviceptr>(ctx->d_feed + 4 * buf), seq,
                               CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
                return fail(ctx, HULK_B200_ECUDA, "cuStreamWaitValue32");
            hl.lap(3);
            { const int rcw = before_first_k1(ctx, hs, ks); if (rcw) return rcw; }
            hl.lap(4);
            const uint64_t words = (nb + 15) / 16;
            k0_unpack_fed<<<(unsigned)((words + 255) / 256), 256, 0, ks>>>(reinterpret_cast<const uint32_t *>(ctx->d_pack[buf]),
                                                                            words, reinterpret_cast<uint4 *>(ctx->d_stage[buf]),
                                                                            ctx->d_feed + 4 * buf);
            LAUNCH_CHECK("k0_unpack");
            k0_patch_fed<<<32, 256, 0, ks>>>(ctx->d_exc[buf], ctx->d_feed + 4 * buf, ctx->d_stage[buf]);
            LAUNCH_CHECK("k0_patch");
            FEED_DBG("push: wait + unpack enqueued for seq %u", seq);
            hl.lap(5);
            rc = launch_k1<false>(ctx, hs, ks, ctx->d_stage[buf], (nb + 15) & ~15ull, offsets ? ctx->d_off[buf] : nullptr,
                                  b0, fixed_len, nr, nb, nullptr, 0, nullptr);
            if (rc) return rc;
            hl.lap(6);
            CU(ctx->ev_k1_record(buf, ks));
            CU(cudaEventRecord(ctx->ev_k1_last[hs], ks));
            hl.lap(7);
            ctx->k1_pending[hs] = true;
            ctx->st.n_reads += nr;
            ctx->st.n_bases += nb;
            ctx->cur_buf = (ctx->cur_buf + 1) % NSTAGE;
            done = upto;
            continue;
        } else if (ctx->pack_threads != 0 && nb >= 4096 && nb < (1ull << 32)) {
            const uint64_t exc_cap = std::max<uint64_t>(1024, nb / 32);
            packed_bytes = (nb + 3) / 4;
            rc = ensure_pack_stage(ctx, buf, packed_bytes, exc_cap, true);
            if (rc) return rc;
            CU(cudaEventSynchronize(ctx->ev_copy[buf]));   // the copy out of this pinned buffer, NSTAGE batches ago, is done
            const auto t_pack = std::chrono::steady_clock::now();
            rc = hulk_b200_pack_bases(hb.bases + b0, nb, ctx->h_pack[buf], ctx->h_exc[buf], exc_cap, &n_exc,
                                      ctx->pack_threads);
            ctx->st.pack_ns += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(
                                   std::chrono::steady_clock::now() - t_pack).count();
            if (rc) return fail(ctx, rc, "pack_bases");
            if (n_exc <= exc_cap) {
                src_packed = ctx->h_pack[buf];
                src_exc = ctx->h_exc[buf];
                unpack_words = (nb + 15) / 16;
            } else {
                n_exc = 0;                                 // too many foreign bytes to be worth it: this batch goes as ASCII
                packed_bytes = 0;
            }
        }

        // (a batch handled on this thread while the feeder still owes copies: let it catch up first, so that the
        // wait below never lands on the copy stream in front of a copy the awaited kernels depend on)
        { const int rcd = feed_drain(ctx); if (rcd) return rcd; }
        CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_k1_prev(buf), 0));
        uint64_t moved = 0;
        if (src_packed) {
            CU(cudaMemcpyAsync(ctx->d_pack[buf], src_packed, packed_bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
            if (n_exc)
                CU(cudaMemcpyAsync(ctx->d_exc[buf], src_exc, sizeof(uint32_t) * n_exc, cudaMemcpyHostToDevice, ctx->copy_stream));
            moved = packed_bytes + sizeof(uint32_t) * n_exc;
            ctx->st.n_packed_batches++;
        } else {
            if (nb) CU(cudaMemcpyAsync(ctx->d_stage[buf], hb.bases + b0, nb, cudaMemcpyHostToDevice, ctx->copy_stream));
            moved = nb;
        }
        if (offsets)
            CU(cudaMemcpyAsync(ctx->d_off[buf], offsets + done, sizeof(uint64_t) * (nr + 1), cudaMemcpyHostToDevice,
                               ctx->copy_stream));
        CU(cudaEventRecord(ctx->ev_copy[buf], ctx->copy_stream));
        const int hs = ctx->cur_hist;
        cudaStream_t ks = ctx->overlap ? ctx->k1_stream[hs] : ctx->stream;
        CU(cudaStreamWaitEvent(ks, ctx->ev_copy[buf], 0));
        { const int rcw = before_first_k1(ctx, hs, ks); if (rcw) return rcw; }              // its last flush left it clean
        ctx->st.h2d_bytes += moved + (offsets ? sizeof(uint64_t) * (nr + 1) : 0);
        if (src_packed) {
            k0_unpack<<<(unsigned)((unpack_words + 255) / 256), 256, 0, ks>>>(reinterpret_cast<const uint32_t *>(ctx->d_pack[buf]),
                                                                                unpack_words,
                                                                                reinterpret_cast<uint4 *>(ctx->d_stage[buf]));
            LAUNCH_CHECK("k0_unpack");
            if (n_exc) {
                k0_patch<<<(unsigned)((n_exc + 255) / 256), 256, 0, ks>>>(ctx->d_exc[buf], n_exc, exc_shift, ctx->d_stage[buf]);
                LAUNCH_CHECK("k0_patch");
            }
        }
        rc = launch_k1<false>(ctx, hs, ks, ctx->d_stage[buf] + skip, (nb + 15) & ~15ull, offsets ? ctx->d_off[buf] : nullptr,
                              b0, fixed_len, nr, nb, nullptr, 0, nullptr);
        if (rc) return rc;
        CU(ctx->ev_k1_record(buf, ks));
        CU(cudaEventRecord(ctx->ev_k1_last[hs], ks));
        ctx->k1_pending[hs] = true;
        ctx->st.n_reads += nr;
        ctx->st.n_bases += nb;
        ctx->cur_buf = (ctx->cur_buf + 1) % NSTAGE;
        done = upto;
        // packed here: the caller's buffer has been read already; otherwise it is in use until the copy is done
        if (!(ctx->P.flags & HULK_B200_F_ASYNC_INPUT) && !(src_packed && !hb.packed)) CU(cudaEventSynchronize(ctx->ev_copy[buf]));
    }
    return HULK_B200_OK;
}

int hulk_b200_set_input_packing(hulk_b200_ctx *ctx, int32_t n_threads) {
    if (!ctx) return HULK_B200_EARG;
    ctx->pack_threads = n_threads;
    return HULK_B200_OK;
}

int hulk_b200_push_reads_packed(hulk_b200_ctx *ctx, const uint8_t *packed, const uint32_t *exceptions, uint64_t n_exceptions,
                                const uint64_t *offsets, uint64_t n_reads, uint32_t read_len) {
    if (!ctx) return HULK_B200_EARG;
    if (n_reads == 0) return HULK_B200_OK;
    if (!packed) return fail(ctx, HULK_B200_EARG, "packed is NULL");
    if (n_exceptions && !exceptions) return fail(ctx, HULK_B200_EARG, "exceptions is NULL");
    HostBatch hb;
    hb.packed = packed;
    hb.exc = exceptions;
    hb.n_exc = n_exceptions;
    hb.n_reads = n_reads;
    if (offsets) {
        hb.offsets = offsets;
    } else {
        if (read_len < 1) return fail(ctx, HULK_B200_EEMPTYSEQ);
        if (read_len < ctx->P.w + ctx->P.k - 1) return fail(ctx, HULK_B200_ESHORTSEQ);
        hb.fixed_len = read_len;
    }
    return push_host(ctx, hb);
}

int hulk_b200_push_reads(hulk_b200_ctx *ctx, const uint8_t *bases, const uint64_t *offsets, uint64_t n_reads) {
    if (!ctx) return HULK_B200_EARG;
    if (n_reads == 0) return HULK_B200_OK;
    if (!offsets || (!bases && offsets[n_reads] != offsets[0])) return fail(ctx, HULK_B200_EARG, "bases/offsets is NULL");
    HostBatch hb;
    hb.bases = bases;
    hb.offsets = offsets;
    hb.n_reads = n_reads;
    return push_host(ctx, hb);
}

int hulk_b200_push_reads_fixed(hulk_b200_ctx *ctx, const uint8_t *bases, uint64_t n_reads, uint32_t read_len) {
    if (!ctx) return HULK_B200_EARG;
    if (n_reads == 0) return HULK_B200_OK;
    if (read_len < 1) return fail(ctx, HULK_B200_EEMPTYSEQ);
    if (read_len < ctx->P.w + ctx->P.k - 1) return fail(ctx, HULK_B200_ESHORTSEQ);
    if (!bases) return fail(ctx, HULK_B200_EARG, "bases is NULL");
    HostBatch hb;
    hb.bases = bases;
    hb.n_reads = n_reads;
    hb.fixed_len = read_len;
    HostLap hl;
    const int rc = push_host(ctx, hb);
    hl.lap(10);
    return rc;
}

int hulk_b200_push_reads_device(hulk_b200_ctx *ctx, const uint8_t *d_bases, const uint64_t *d_offsets,
                                uint64_t n_reads, uint32_t read_len) {
    if (!ctx) return HULK_B200_EARG;
    if (n_reads == 0) return HULK_B200_OK;
    if (!d_bases) return fail(ctx, HULK_B200_EARG, "d_bases is NULL");
    CU(cudaSetDevice(ctx->P.device));
    uint64_t total_bytes, extent, off_base = 0;
    if (read_len) {
        if (read_len < ctx->P.w + ctx->P.k - 1) return fail(ctx, HULK_B200_ESHORTSEQ);
        total_bytes = extent = n_reads * (uint64_t)read_len;
        d_offsets = nullptr;
    } else {
        if (!d_offsets) return fail(ctx, HULK_B200_EARG, "d_offsets is NULL and read_len is 0");
        // the two ends of the byte range, the longest read and the room the sets of long sequences need: one round trip
        uint64_t ends[2];
        unsigned long long lens[2] = {0, 0};
        CU(cudaMemsetAsync(ctx->d_lenstat, 0, 16, ctx->stream));
        k1_length_stats<<<(unsigned)std::min<uint64_t>((n_reads + 255) / 256, (uint64_t)ctx->sm_count * 4), 256, 0, ctx->stream>>>(
            d_offsets, n_reads, ctx->long_min, (int32_t)ctx->P.k, ctx->d_lenstat);
        LAUNCH_CHECK("k1_length_stats");
        CU(cudaMemcpyAsync(&ends[0], d_offsets, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(&ends[1], d_offsets + n_reads, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(lens, ctx->d_lenstat, 16, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (ends[1] < ends[0]) return fail(ctx, HULK_B200_EARG, "offsets must be non-decreasing");
        extent = ends[1];
        total_bytes = ends[1] - ends[0];
        if (lens[0] > total_bytes) return fail(ctx, HULK_B200_EARG, "offsets must be non-decreasing");
        ctx->batch_max_len = lens[0];
        ctx->batch_long_entries = lens[1];
    }
    const int hs = ctx->cur_hist;
    if (read_len) note_batch_lengths(ctx, nullptr, n_reads, read_len);
    cudaStream_t ks = ctx->overlap ? ctx->k1_stream[hs] : ctx->stream;
    if (ctx->overlap && !(ctx->P.flags & HULK_B200_F_INPUT_READY)) {          // order behind whatever produced the input on the main stream
        CU(cudaEventRecord(ctx->ev_main, ctx->stream));
        CU(cudaStreamWaitEvent(ks, ctx->ev_main, 0));
    }
    { const int rcw = before_first_k1(ctx, hs, ks); if (rcw) return rcw; }
    // bytes that may be touched: the batch itself, rounded DOWN to 16 so wide loads never overrun it
    const int rc = launch_k1<false>(ctx, hs, ks, d_bases, extent & ~15ull, d_offsets, off_base, read_len, n_reads,
                                    total_bytes, nullptr, 0, nullptr);
    if (rc) return rc;
    CU(cudaEventRecord(ctx->ev_k1_last[hs], ks));
    ctx->k1_pending[hs] = true;
    ctx->st.n_reads += n_reads;
    ctx->st.n_bases += total_bytes;
    return HULK_B200_OK;
}

int hulk_b200_sync_inputs(hulk_b200_ctx *ctx) {
    if (!ctx) return HULK_B200_EARG;
    CU(cudaSetDevice(ctx->P.device));
    { const int rc = feed_drain(ctx); if (rc) return rc; }
    CU(cudaStreamSynchronize(ctx->copy_stream));
    return HULK_B200_OK;
}

// ------------------------------------------------------------------------------------------
// stage 3
// ------------------------------------------------------------------------------------------
static int flush_impl(hulk_b200_ctx *ctx);
int hulk_b200_flush(hulk_b200_ctx *ctx) {
    HostLap hl;
    const int rc = flush_impl(ctx);
    hl.lap(8);
    return rc;
}
static int flush_impl(hulk_b200_ctx *ctx) {
    if (!ctx) return HULK_B200_EARG;
    if (ctx->gen_pending) {
        CU(cudaSetDevice(ctx->P.device));
        { const int rcd = feed_drain(ctx); if (rcd) return rcd; }     // the upload allocates
        const int rc = finish_async_tables(ctx);
        if (rc) return rc;
    }
    if (!ctx->tables_set) return fail(ctx, HULK_B200_ESTATE, "CWS tables not set (set_cws_tables / generate_cws_tables)");
    CU(cudaSetDevice(ctx->P.device));
    cudaStream_t st = ctx->stream;                                     // CWS sweep
    cudaStream_t k2s = ctx->overlap ? ctx->k2_stream : ctx->stream;    // count-min update
    const int32_t D = ctx->D;
    const int hs = ctx->cur_hist;
    const int fi = ctx->flush_idx;
    const bool peers = ctx->world > 1;
    uint32_t *const hist = peers ? ctx->d_hist_sum : ctx->d_hist[hs];
    unsigned long long *const fbits = ctx->d_fbits[fi];
    float *const invf = ctx->d_invf[fi];
    uint32_t seq = 0;
    if (peers) {
        // tell every GPU (this one included) that this GPU's share of the interval is counted
        cudaStream_t ks = ctx->overlap ? ctx->k1_stream[hs] : ctx->stream;
        { const int rcw = before_first_k1(ctx, hs, ks); if (rcw) return rcw; }      // nothing pushed here: still a clean buffer
        seq = ++ctx->use_seq[hs];
        PeerTargets t{};
        t.n = ctx->world;
        t.seq = seq;
        for (uint32_t p = 0; p < ctx->world; p++)
            t.flag[p] = reinterpret_cast<uint32_t *>(ctx->peer_arena[p] + ctx->off_counted) + (size_t)hs * PEER_MAX + ctx->rank;
        k_peer_signal<<<1, 32, 0, ks>>>(t);
        LAUNCH_CHECK("k_peer_signal");
        CU(cudaEventRecord(ctx->ev_k1_last[hs], ks));
        ctx->k1_pending[hs] = true;
    }
    if (ctx->k1_pending[hs]) CU(cudaStreamWaitEvent(k2s, ctx->ev_k1_last[hs], 0));  // every read of the interval is counted
    if (ctx->k3_pending[fi]) CU(cudaStreamWaitEvent(k2s, ctx->ev_k3_done[fi], 0));  // the sweep two flushes back read this f
    {
    ProfScope prof_scope(ctx, 1, k2s);
    if (peers) {
        const uint32_t *flags = reinterpret_cast<const uint32_t *>(ctx->arena + ctx->off_counted) + (size_t)hs * PEER_MAX;
        k_peer_wait<<<1, 32, 0, k2s>>>(flags, ctx->world, seq, ctx->d_ctl, ctx->peer_timeout_ns);   // ... on every GPU
        LAUNCH_CHECK("k_peer_wait");
        PeerSources src{};
        PeerTargets done{};
        src.n = done.n = ctx->world;
        done.seq = seq;
        for (uint32_t p = 0; p < ctx->world; p++) {
            src.hist[p] = reinterpret_cast<const uint32_t *>(ctx->peer_arena[p] + (size_t)hs * ctx->hist_stride);
            done.flag[p] = reinterpret_cast<uint32_t *>(ctx->peer_arena[p] + ctx->off_gathered) + (size_t)hs * PEER_MAX + ctx->rank;
        }
        k2_mask_count_peers<<<ctx->nblk, K2_MASK_TPB, 0, k2s>>>(src, D, hist, ctx->d_words, ctx->d_word_prefix,
                                                         ctx->d_block_count, fbits, ctx->d_ctl, fi, done, ctx->d_ticket,
                                                         ctx->d_block_prefix);
        ctx->peer_dirty[hs] = seq;
    } else {
        k2_mask_count<<<ctx->nblk, K2_MASK_TPB, 0, k2s>>>(hist, D, ctx->d_words, ctx->d_word_prefix, ctx->d_block_count, fbits,
                                                   ctx->d_ctl, fi, ctx->d_block_prefix, ctx->d_ticket);
    }
    LAUNCH_CHECK("k2_mask_count");
    k2_cms_update<<<(CMS_CELLS * 32 + 255) / 256, 256, 0, k2s>>>(hist, ctx->d_csr_start, ctx->d_csr_bins,
                                                                 ctx->d_words, ctx->d_word_prefix, ctx->d_block_prefix,
                                                                 ctx->d_q, fbits, ctx->d_ctl, fi,
                                                                 ctx->apply_scaling ? 1 : 0, ctx->decay_weight);
    LAUNCH_CHECK("k2_cms_update");
    k2_finalize<<<(unsigned)(((uint64_t)D + 255) / 256), 256, 0, k2s>>>(hist, D, fbits, invf, ctx->d_invf16[fi],
                                                                        ctx->d_ctl, fi);
    LAUNCH_CHECK("k2_finalize");
    }
    // the buffer is wiped: the next interval but one may count into it while the CWS sweep below runs
    CU(cudaEventRecord(ctx->ev_hist_free[hs], k2s));
    ctx->k1_pending[hs] = false;
    ctx->cur_hist = (hs + 1) % ctx->nbuf;
    ctx->last_flush_idx = fi;
    ctx->flush_idx = fi ^ 1;
    if (ctx->rows) {
        if (k2s != st) {
            CU(cudaEventRecord(ctx->ev_k2_done[fi], k2s));
            CU(cudaStreamWaitEvent(st, ctx->ev_k2_done[fi], 0));
        }
        const int stages = ctx->k3_stages;
        const size_t smem = (size_t)stages * K3_SEG * 4 + 2 * stages * 8;
        const uint64_t T = (uint64_t)ctx->rows * ctx->nseg;
        const unsigned grid = (unsigned)std::min<uint64_t>(T, (uint64_t)ctx->sm_count * ctx->k3_ctas_per_sm);
        {
            ProfScope prof_scope(ctx, 2);
#define K3_FILTER_ARGS ctx->Dp, invf, ctx->d_m32, ctx->rows, ctx->nseg, ctx->d_thr32, ctx->d_cand, ctx->d_ctl, fi
#define K3_FILTER16_ARGS ctx->Dp, ctx->d_invf16[fi], ctx->d_m32, ctx->rows, ctx->nseg, ctx->d_thr32, ctx->d_cand, ctx->d_ctl, fi
            if (ctx->filter16) {
                if (stages == 4) k3_filter16<4><<<grid, K3_THREADS, smem, st>>>(ctx->d_K16, K3_FILTER16_ARGS);
                else if (stages == 6) k3_filter16<6><<<grid, K3_THREADS, smem, st>>>(ctx->d_K16, K3_FILTER16_ARGS);
                else k3_filter16<8><<<grid, K3_THREADS, smem, st>>>(ctx->d_K16, K3_FILTER16_ARGS);
            } else {
                if (stages == 4) k3_filter<4><<<grid, K3_THREADS, smem, st>>>(ctx->d_K32, K3_FILTER_ARGS);
                else if (stages == 6) k3_filter<6><<<grid, K3_THREADS, smem, st>>>(ctx->d_K32, K3_FILTER_ARGS);
                else k3_filter<8><<<grid, K3_THREADS, smem, st>>>(ctx->d_K32, K3_FILTER_ARGS);
            }
#undef K3_FILTER_ARGS
#undef K3_FILTER16_ARGS
            LAUNCH_CHECK("k3_filter");
        }
        {
            ProfScope prof_scope(ctx, 3);
            k3_resolve<<<(ctx->rows * 32 + 127) / 128, 128, 0, st>>>(ctx->d_m32, ctx->nsub_row, ctx->d_r, ctx->d_c,
                                                                     ctx->d_b, D, fbits, ctx->rows, ctx->d_sketch,
                                                                     ctx->d_weights, ctx->drift ? 1 : 0,
                                                                     ctx->decay_weight, ctx->d_cand, ctx->d_thr32,
                                                                     ctx->drift ? 1.0 / ctx->decay_weight : 1.0,
                                                                     ctx->k3_eps, ctx->d_ctl, fi);
            LAUNCH_CHECK("k3_resolve");
        }
        CU(cudaEventRecord(ctx->ev_k3_done[fi], st));
        ctx->k3_pending[fi] = true;
    } else if (k2s != st) {
        // no slots owned (a pure counting rank): keep the main stream ordered behind the flush all the same
        CU(cudaEventRecord(ctx->ev_k2_done[fi], k2s));
        CU(cudaStreamWaitEvent(st, ctx->ev_k2_done[fi], 0));
    }
    return HULK_B200_OK;
}

// read back the deferred device-side error state (requires a synchronised stream)
static int deferred_error(hulk_b200_ctx *ctx) {
    unsigned long long ew = 0;
    FlushCtl ctl;
    CU(cudaMemcpy(&ew, ctx->d_errword, 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(&ctl, ctx->d_ctl, sizeof(ctl), cudaMemcpyDeviceToHost));
    ctx->st.d2h_bytes += 8 + sizeof(ctl);
    ctx->st.n_flushes = ctl.n_flushes;
    ctx->st.n_adds = ctl.n_adds;
    ctx->st.n_rescans = ctl.n_rescans;
    if (ew != ~0ull) {
        const uint32_t code = (uint32_t)(ew & 0xff);
        const unsigned long long read = ew >> 8;
        const int rc = code == K1_ERR_EMPTY ? HULK_B200_EEMPTYSEQ : code == K1_ERR_SHORT ? HULK_B200_ESHORTSEQ
                                                                                         : HULK_B200_ENOMEM;
        return fail(ctx, rc, "read " + std::to_string(read) + (rc == HULK_B200_ENOMEM ? " (scratch arena exhausted)" : ""));
    }
    if (ctl.err) return fail(ctx, ctl.err);
    return HULK_B200_OK;
}

int hulk_b200_sync(hulk_b200_ctx *ctx) {
    if (!ctx) return HULK_B200_EARG;
    CU(cudaSetDevice(ctx->P.device));
    { const int rc = sync_all(ctx); if (rc) return rc; }
    return deferred_error(ctx);
}

int hulk_b200_finish(hulk_b200_ctx *ctx, uint64_t *mins, double *weights) {
    if (!ctx) return HULK_B200_EARG;
    const int rc = hulk_b200_sync(ctx);
    if (rc) return rc;
    if (ctx->rows) {
        if (!mins || !weights) return fail(ctx, HULK_B200_EARG, "mins/weights is NULL");
        CU(cudaMemcpy(mins, ctx->d_sketch, sizeof(uint64_t) * ctx->rows, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(weights, ctx->d_weights, sizeof(double) * ctx->rows, cudaMemcpyDeviceToHost));
        ctx->st.d2h_bytes += 16ull * ctx->rows;
    }
    return HULK_B200_OK;
}

// copies the sketch straight into page-locked host memory from a kernel: no copy-engine round trip behind
// whatever host->device batch is in flight on the same engine
__global__ void k_snapshot(const unsigned long long *__restrict__ sketch, const double *__restrict__ weights,
                           unsigned long long *__restrict__ h_mins, double *__restrict__ h_weights, uint32_t rows) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows) {
        h_mins[i] = sketch[i];
        h_weights[i] = weights[i];
    }
}
static bool device_can_write(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost && a.devicePointer == p;
}

static int snapshot_impl(hulk_b200_ctx *ctx, uint64_t *mins, double *weights);
int hulk_b200_snapshot_async(hulk_b200_ctx *ctx, uint64_t *mins, double *weights) {
    HostLap hl;
    const int rc = snapshot_impl(ctx, mins, weights);
    hl.lap(9);
    return rc;
}
static int snapshot_impl(hulk_b200_ctx *ctx, uint64_t *mins, double *weights) {
    if (!ctx) return HULK_B200_EARG;
    if (!ctx->rows) return HULK_B200_OK;
    if (!mins || !weights) return fail(ctx, HULK_B200_EARG, "mins/weights is NULL");
    CU(cudaSetDevice(ctx->P.device));
    if (device_can_write(mins) && device_can_write(weights)) {
        k_snapshot<<<(ctx->rows + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_sketch, ctx->d_weights,
                                                                     reinterpret_cast<unsigned long long *>(mins),
                                                                     weights, ctx->rows);
        LAUNCH_CHECK("k_snapshot");
        ctx->st.d2h_bytes += 16ull * ctx->rows;
        return HULK_B200_OK;
    }
    CU(cudaMemcpyAsync(mins, ctx->d_sketch, sizeof(uint64_t) * ctx->rows, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(weights, ctx->d_weights, sizeof(double) * ctx->rows, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->st.d2h_bytes += 16ull * ctx->rows;
    return HULK_B200_OK;
}

int hulk_b200_get_stats(hulk_b200_ctx *ctx, hulk_b200_stats *out) {
    if (!ctx || !out) return HULK_B200_EARG;
    CU(cudaSetDevice(ctx->P.device));
    { const int rc = sync_all(ctx); if (rc) return rc; }
    unsigned long long nm = 0;
    FlushCtl ctl;
    CU(cudaMemcpy(&nm, ctx->d_nmin, 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(&ctl, ctx->d_ctl, sizeof(ctl), cudaMemcpyDeviceToHost));
    ctx->st.n_minimizers = nm + ctx->extra_minimizers;
    ctx->st.n_flushes = ctl.n_flushes;
    ctx->st.n_adds = ctl.n_adds;
    ctx->st.n_rescans = ctl.n_rescans;
    *out = ctx->st;
    out->h2d_bytes += ctx->feed_h2d.load();
    out->pack_ns += ctx->feed_pack_ns.load();
    out->n_packed_batches += ctx->feed_batches.load();
    return HULK_B200_OK;
}

int hulk_b200_histogram_device_ptr(hulk_b200_ctx *ctx, void **d_hist_u32, int32_t *num_bins) {
    if (!ctx || !d_hist_u32) return HULK_B200_EARG;
    CU(cudaSetDevice(ctx->P.device));
    // the spectrum of the interval being counted; work enqueued on the context's stream after this
    // call (e.g. the all-reduce across GPUs) sees every read pushed so far
    const int hs = ctx->cur_hist;
    cudaStream_t k2s = ctx->overlap ? ctx->k2_stream : ctx->stream;
    if (ctx->k1_pending[hs]) CU(cudaStreamWaitEvent(k2s, ctx->ev_k1_last[hs], 0));
    *d_hist_u32 = ctx->d_hist[hs];
    if (num_bins) *num_bins = ctx->D;
    return HULK_B200_OK;
}
int hulk_b200_stream(hulk_b200_ctx *ctx, void **cuda_stream) {
    if (!ctx || !cuda_stream) return HULK_B200_EARG;
    *cuda_stream = ctx->overlap ? ctx->k2_stream : ctx->stream;    // where the flush starts: the spectrum's stream
    return HULK_B200_OK;
}
__global__ void k_merge_hist(uint32_t *__restrict__ dst, const uint32_t *__restrict__ src, int32_t D) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < D) dst[i] += src[i];
}
int hulk_b200_merge_histogram(hulk_b200_ctx *ctx, const uint32_t *hist) {
    if (!ctx || !hist) return HULK_B200_EARG;
    CU(cudaSetDevice(ctx->P.device));
    { const int rc = sync_all(ctx); if (rc) return rc; }
    uint32_t *d_tmp = nullptr;
    CU(dmalloc(&d_tmp, ctx->D));
    CU(cudaMemcpyAsync(d_tmp, hist, sizeof(uint32_t) * (size_t)ctx->D, cudaMemcpyHostToDevice, ctx->stream));
    k_merge_hist<<<(unsigned)(((uint64_t)ctx->D + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_hist[ctx->cur_hist], d_tmp,
                                                                                    ctx->D);
    LAUNCH_CHECK("k_merge_hist");
    CU(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_tmp);
    ctx->st.h2d_bytes += sizeof(uint32_t) * (size_t)ctx->D;
    return HULK_B200_OK;
}
int hulk_b200_add_minimizer_count(hulk_b200_ctx *ctx, uint64_t n) {
    if (!ctx) return HULK_B200_EARG;
    ctx->extra_minimizers += n;
    return HULK_B200_OK;
}

// ------------------------------------------------------------------------------------------
// multi-GPU: peers
// ------------------------------------------------------------------------------------------
static int peer_finish_connect(hulk_b200_ctx *ctx, uint32_t world, uint32_t rank) {
    if (!ctx->d_hist_sum) CU(dmalloc(&ctx->d_hist_sum, ctx->D));
    ctx->world = world;
    ctx->rank = rank;
    return HULK_B200_OK;
}
// debug tap: this context's two flag arrays, [NBUF][PEER_MAX] counted then [NBUF][PEER_MAX] gathered (no synchronisation)
int hulk_b200_peer_flags(hulk_b200_ctx *ctx, uint32_t *out) {
    if (!ctx || !out) return HULK_B200_EARG;
    CU(cudaSetDevice(ctx->P.device));
    cudaStream_t tmp;
    CU(cudaStreamCreateWithFlags(&tmp, cudaStreamNonBlocking));
    CU(cudaMemcpyAsync(out, ctx->arena + ctx->off_counted, sizeof(uint32_t) * 2 * NBUF * PEER_MAX, cudaMemcpyDeviceToHost, tmp));
    CU(cudaStreamSynchronize(tmp));
    cudaStreamDestroy(tmp);
    return HULK_B200_OK;
}
int hulk_b200_peer_export(hulk_b200_ctx *ctx, void *handle) {
    if (!ctx || !handle) return HULK_B200_EARG;
    static_assert(sizeof(cudaIpcMemHandle_t) <= HULK_B200_PEER_HANDLE_BYTES, "handle size");
    CU(cudaSetDevice(ctx->P.device));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, ctx->arena));
    memset(handle, 0, HULK_B200_PEER_HANDLE_BYTES);
    memcpy(handle, &h, sizeof h);
    return HULK_B200_OK;
}
int hulk_b200_peer_connect(hulk_b200_ctx *ctx, uint32_t world, uint32_t rank, const void *handles) {
    if (!ctx || !handles || world < 1 || world > PEER_MAX || rank >= world) return fail(ctx, HULK_B200_EARG, "peer world/rank");
    if (ctx->world != 1) return fail(ctx, HULK_B200_ESTATE, "peers already connected");
    CU(cudaSetDevice(ctx->P.device));
    { const int rc = sync_all(ctx); if (rc) return rc; }
    for (uint32_t p = 0; p < world; p++) {
        if (p == rank) { ctx->peer_arena[p] = ctx->arena; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const uint8_t *>(handles) + (size_t)p * HULK_B200_PEER_HANDLE_BYTES, sizeof h);
        void *ptr = nullptr;
        CU(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->peer_arena[p] = static_cast<uint8_t *>(ptr);
        ctx->peer_ipc[p] = true;
    }
    return peer_finish_connect(ctx, world, rank);
}
int hulk_b200_peer_connect_local(hulk_b200_ctx *ctx, uint32_t world, uint32_t rank, hulk_b200_ctx *const *peers) {
    if (!ctx || !peers || world < 1 || world > PEER_MAX || rank >= world || peers[rank] != ctx)
        return fail(ctx, HULK_B200_EARG, "peer world/rank");
    if (ctx->world != 1) return fail(ctx, HULK_B200_ESTATE, "peers already connected");
    CU(cudaSetDevice(ctx->P.device));
    { const int rc = sync_all(ctx); if (rc) return rc; }
    for (uint32_t p = 0; p < world; p++) {
        if (!peers[p] || peers[p]->D != ctx->D) return fail(ctx, HULK_B200_EARG, "peer context");
        if (p != rank && peers[p]->P.device != ctx->P.device) {
            int can = 0;
            CU(cudaDeviceCanAccessPeer(&can, ctx->P.device, peers[p]->P.device));
            if (!can) return fail(ctx, HULK_B200_ECUDA, "no peer access between devices " + std::to_string(ctx->P.device) +
                                                            " and " + std::to_string(peers[p]->P.device));
            const cudaError_t e = cudaDeviceEnablePeerAccess(peers[p]->P.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return fail(ctx, HULK_B200_ECUDA, std::string("cudaDeviceEnablePeerAccess -> ") + cudaGetErrorString(e));
            cudaGetLastError();
        }
        ctx->peer_arena[p] = peers[p]->arena;
    }
    return peer_finish_connect(ctx, world, rank);
}

// ------------------------------------------------------------------------------------------
// parity taps
// ------------------------------------------------------------------------------------------
int hulk_b200_get_histogram(hulk_b200_ctx *ctx, uint32_t *hist) {
    if (!ctx || !hist) return HULK_B200_EARG;
    CU(cudaSetDevice(ctx->P.device));
    { const int rc = sync_all(ctx); if (rc) return rc; }
    CU(cudaMemcpy(hist, ctx->d_hist[ctx->cur_hist], sizeof(uint32_t) * (size_t)ctx->D, cudaMemcpyDeviceToHost));
    return HULK_B200_OK;
}
int hulk_b200_get_estimates(hulk_b200_ctx *ctx, double *f) {
    if (!ctx || !f) return HULK_B200_EARG;
    CU(cudaSetDevice(ctx->P.device));
    { const int rc = sync_all(ctx); if (rc) return rc; }
    CU(cudaMemcpy(f, ctx->d_fbits[ctx->last_flush_idx], sizeof(double) * (size_t)ctx->D, cudaMemcpyDeviceToHost));
    for (int32_t i = 0; i < ctx->D; i++)
        if (std::isinf(f[i])) f[i] = NAN;
    return HULK_B200_OK;
}
int hulk_b200_get_cms(hulk_b200_ctx *ctx, double *q) {
    if (!ctx || !q) return HULK_B200_EARG;
    CU(cudaSetDevice(ctx->P.device));
    { const int rc = sync_all(ctx); if (rc) return rc; }
    CU(cudaMemcpy(q, ctx->d_q, sizeof(double) * CMS_CELLS, cudaMemcpyDeviceToHost));
    return HULK_B200_OK;
}
int hulk_b200_minimizers(hulk_b200_ctx *ctx, const uint8_t *bases, const uint64_t *offsets, uint64_t n_reads,
                         uint64_t *out, uint32_t cap, uint32_t *counts) {
    if (!ctx || !offsets || !out || !counts || cap == 0) return HULK_B200_EARG;
    if (n_reads == 0) return HULK_B200_OK;
    CU(cudaSetDevice(ctx->P.device));
    const uint64_t b0 = offsets[0], nb = offsets[n_reads] - offsets[0];
    uint8_t *d_bases = nullptr;
    uint64_t *d_offsets = nullptr, *d_dump = nullptr;
    uint32_t *d_counts = nullptr;
    const uint64_t padded = ((nb + 15) & ~15ull) + 64;
    CU(dmalloc(&d_bases, padded));
    CU(dmalloc(&d_offsets, n_reads + 1));
    CU(dmalloc(&d_dump, n_reads * cap));
    CU(dmalloc(&d_counts, n_reads));
    { const int rc0 = sync_all(ctx); if (rc0) return rc0; }
    cudaStream_t st = ctx->stream;
    if (nb) CU(cudaMemcpyAsync(d_bases, bases + b0, nb, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_offsets, offsets, sizeof(uint64_t) * (n_reads + 1), cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(d_counts, 0, sizeof(uint32_t) * n_reads, st));
    const uint64_t saved_reads = ctx->st.n_reads;
    note_batch_lengths(ctx, offsets, n_reads, 0);
    int rc = launch_k1<true>(ctx, 0, st, d_bases, (nb + 15) & ~15ull, d_offsets, b0, 0, n_reads, nb, d_dump, cap,
                             d_counts);
    ctx->st.n_reads = saved_reads;
    if (rc == HULK_B200_OK) {
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = fail(ctx, HULK_B200_ECUDA, cudaGetErrorString(e));
    }
    if (rc == HULK_B200_OK) {
        cudaMemcpy(out, d_dump, sizeof(uint64_t) * n_reads * cap, cudaMemcpyDeviceToHost);
        cudaMemcpy(counts, d_counts, sizeof(uint32_t) * n_reads, cudaMemcpyDeviceToHost);
        // the tap reports per-read errors immediately and leaves the context clean
        unsigned long long ew = 0;
        cudaMemcpy(&ew, ctx->d_errword, 8, cudaMemcpyDeviceToHost);
        if (ew != ~0ull) {
            const uint32_t code = (uint32_t)(ew & 0xff);
            cudaMemset(ctx->d_errword, 0xff, 8);
            rc = fail(ctx, code == K1_ERR_EMPTY ? HULK_B200_EEMPTYSEQ : code == K1_ERR_SHORT ? HULK_B200_ESHORTSEQ
                                                                                             : HULK_B200_ENOMEM,
                      "read " + std::to_string(ew >> 8));
        }
    }
    cudaFree(d_bases); cudaFree(d_offsets); cudaFree(d_dump); cudaFree(d_counts);
    return rc;
}
int hulk_b200_jump_hash(hulk_b200_ctx *ctx, const uint64_t *keys, uint64_t n, int32_t num_buckets, int32_t *out) {
    if (!ctx || !keys || !out) return HULK_B200_EARG;
    if (n == 0) return HULK_B200_OK;
    CU(cudaSetDevice(ctx->P.device));
    uint64_t *d_keys = nullptr;
    int32_t *d_out = nullptr;
    CU(dmalloc(&d_keys, n));
    CU(dmalloc(&d_out, n));
    CU(cudaMemcpyAsync(d_keys, keys, 8 * n, cudaMemcpyHostToDevice, ctx->stream));
    k_jump_tap<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_keys, n, num_buckets, d_out);
    LAUNCH_CHECK("k_jump_tap");
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpy(out, d_out, 4 * n, cudaMemcpyDeviceToHost));
    cudaFree(d_keys);
    cudaFree(d_out);
    return HULK_B200_OK;
}
int hulk_b200_jump_hash_fx(hulk_b200_ctx *ctx, const uint64_t *keys, uint64_t n, int32_t num_buckets, int32_t *out,
                           uint64_t *n_ambiguous) {
    if (!ctx || !keys || !out) return HULK_B200_EARG;
    if (num_buckets < 1 || (uint32_t)num_buckets > JUMP_FX_MAX_BUCKETS)
        return fail(ctx, HULK_B200_EARG, "the fixed-point step serves 1 <= num_buckets <= 2^20");
    if (n == 0) return HULK_B200_OK;
    CU(cudaSetDevice(ctx->P.device));
    uint64_t *d_keys = nullptr;
    int32_t *d_out = nullptr;
    unsigned long long *d_amb = nullptr;
    CU(dmalloc(&d_keys, n));
    CU(dmalloc(&d_out, n));
    CU(dmalloc(&d_amb, 1));
    CU(cudaMemcpyAsync(d_keys, keys, 8 * n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(d_amb, 0, 8, ctx->stream));
    k_jump_fx_tap<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_keys, n, num_buckets, d_out, d_amb);
    LAUNCH_CHECK("k_jump_fx_tap");
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpy(out, d_out, 4 * n, cudaMemcpyDeviceToHost));
    unsigned long long amb = 0;
    CU(cudaMemcpy(&amb, d_amb, 8, cudaMemcpyDeviceToHost));
    if (n_ambiguous) *n_ambiguous = amb;
    cudaFree(d_keys);
    cudaFree(d_out);
    cudaFree(d_amb);
    return HULK_B200_OK;
}
int hulk_b200_rcp_selftest(hulk_b200_ctx *ctx, uint32_t q_begin, uint32_t n, double out[2]) {
    if (!ctx || !out) return HULK_B200_EARG;
    CU(cudaSetDevice(ctx->P.device));
    unsigned long long *d_out = nullptr;
    CU(dmalloc(&d_out, 2));
    CU(cudaMemsetAsync(d_out, 0, 16, ctx->stream));
    if (n) {
        k_rcp_selftest<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(q_begin, n, d_out);
        LAUNCH_CHECK("k_rcp_selftest");
    }
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpy(out, d_out, 16, cudaMemcpyDeviceToHost));
    cudaFree(d_out);
    return HULK_B200_OK;
}
int hulk_b200_get_folded_table(hulk_b200_ctx *ctx, float *out, uint64_t *row_stride) {
    if (!ctx || !row_stride) return HULK_B200_EARG;
    *row_stride = ctx->Dp;
    if (!out) return HULK_B200_OK;
    if (!ctx->tables_set) return fail(ctx, HULK_B200_ESTATE, "CWS tables not set");
    CU(cudaSetDevice(ctx->P.device));
    CU(cudaStreamSynchronize(ctx->stream));
    const size_t ne = (size_t)ctx->rows * ctx->Dp;
    if (ctx->filter16) {                      // widen the stored bf16 values (exact)
        std::vector<uint16_t> h(ne ? ne : 1);
        CU(cudaMemcpy(h.data(), ctx->d_K16, 2 * ne, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < ne; i++) {
            const uint32_t bits = (uint32_t)h[i] << 16;
            memcpy(&out[i], &bits, 4);
        }
        return HULK_B200_OK;
    }
    CU(cudaMemcpy(out, ctx->d_K32, sizeof(float) * ne, cudaMemcpyDeviceToHost));
    return HULK_B200_OK;
}

int hulk_b200_alloc_pinned(void **ptr, uint64_t bytes) {
    if (!ptr) return HULK_B200_EARG;
    cudaError_t e = cudaHostAlloc(ptr, (size_t)(bytes ? bytes : 1), cudaHostAllocDefault);
    if (e != cudaSuccess) {
        g_create_err = std::string("cudaHostAlloc -> ") + cudaGetErrorString(e);
        return e == cudaErrorMemoryAllocation ? HULK_B200_ENOMEM : HULK_B200_ECUDA;
    }
    return HULK_B200_OK;
}
void hulk_b200_free_pinned(void *ptr) {
    if (ptr) cudaFreeHost(ptr);
}
