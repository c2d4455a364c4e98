// group.cpp -- one sketch over several GPUs of one process, behind the same C ABI.
//
// Reference semantics (paths relative to the reference checkout): the reference has ONE driver loop
// (SeqMinimizer.Run, src/pipeline/sketch.go:197-224: AddSeq every read, Flush every interval and at the end),
// ONE spectrum (src/pipeline/boss.go:54-60) and ONE HistoSketch (src/pipeline/sketch.go:277).  A group keeps
// that shape for its caller -- one handle, calls in the same order from one thread -- and spreads the work:
//   * each call's reads are split into G contiguous chunks, context g (on GPU g) counts chunk g;
//   * a flush works on the SUM of the G counting buffers, which every GPU reads from its peers over NVLink
//     (api.cu / k2_countmin.cuh: sequence flags, no collective call, no host synchronisation); integer sums, so the
//     spectrum -- and everything behind it -- is bit-identical to one GPU;
//   * the count-min update is replicated (tiny), the CWS sweep is sharded by sketch slot: context g owns slots
//     [g s / G, (g+1) s / G) and only those rows of the CWS tables;
//   * finish gathers the slots on the host.
#include <cstdlib>
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/hulk_b200.h"

struct hulk_b200_group {
    std::vector<hulk_b200_ctx *> ctx;
    std::vector<uint32_t> slot_begin;          // G + 1 entries
    hulk_b200_params P{};
    int32_t D = 0;
    std::string err;
    // newCWS (src/histosketch/histosketch.go:95-126) is ONE seeded stream consumed row by row: the tables are drawn
    // once for all s slots (on a background thread if asked) and every member gets its rows
    std::thread gen_thread;
    std::vector<double> gen_r, gen_c, gen_b;
    int gen_rc = 0;
    bool gen_pending = false;
};

static thread_local std::string g_group_create_err;

static int gfail(hulk_b200_group *g, int code, const std::string &detail = std::string()) {
    std::string msg = hulk_b200_strerror(code);
    if (!detail.empty()) msg += ": " + detail;
    if (g) g->err = msg;
    else g_group_create_err = msg;
    return code;
}
// error of member i -> error of the group
static int gfrom(hulk_b200_group *g, size_t i, int code) {
    if (code == HULK_B200_OK) return code;
    const char *t = hulk_b200_last_error(g->ctx[i]);
    g->err = (t && *t) ? std::string(t) : std::string(hulk_b200_strerror(code));
    if (g->ctx.size() > 1) g->err += " (GPU " + std::to_string(i) + " of the group)";
    return code;
}

extern "C" {

int hulk_b200_group_set_cws_tables(hulk_b200_group *g, const double *r, const double *c, const double *b);

const char *hulk_b200_group_last_error(const hulk_b200_group *g) { return g ? g->err.c_str() : g_group_create_err.c_str(); }

void hulk_b200_group_destroy(hulk_b200_group *g) {
    if (!g) return;
    if (g->gen_thread.joinable()) g->gen_thread.join();
    // every member must be idle before any arena goes away: a peer may still be reading it
    for (hulk_b200_ctx *c : g->ctx)
        if (c) hulk_b200_sync(c);
    for (hulk_b200_ctx *c : g->ctx)
        if (c) hulk_b200_destroy(c);
    delete g;
}

int hulk_b200_group_create(const hulk_b200_params *params, const int32_t *device_ids, uint32_t ngpus,
                           hulk_b200_group **out) {
    if (!params || !out) return gfail(nullptr, HULK_B200_EARG, "params/out is NULL");
    *out = nullptr;
    if (ngpus < 1 || ngpus > 16) return gfail(nullptr, HULK_B200_EARG, "a group has 1 to 16 GPUs");
    if (params->slot_begin != 0 || params->slot_end != 0)
        return gfail(nullptr, HULK_B200_EARG, "a group shards the slots itself (slot_begin = slot_end = 0)");
    hulk_b200_group *g = new (std::nothrow) hulk_b200_group();
    if (!g) return gfail(nullptr, HULK_B200_ENOMEM);
    g->P = *params;
    const uint32_t s = params->sketch_size;
    g->slot_begin.resize(ngpus + 1);
    for (uint32_t i = 0; i <= ngpus; i++) g->slot_begin[i] = (uint32_t)(((uint64_t)i * s) / ngpus);
    for (uint32_t i = 0; i < ngpus; i++) {
        hulk_b200_params p = *params;
        p.device = device_ids ? device_ids[i] : (int32_t)i;
        p.slot_begin = g->slot_begin[i];
        p.slot_end = g->slot_begin[i + 1];
        if (p.slot_begin == p.slot_end && s != 0) {       // 0,0 would mean "all slots": more GPUs than slots
            hulk_b200_group_destroy(g);
            return gfail(nullptr, HULK_B200_EARG, "more GPUs than sketch slots");
        }
        p.stream = nullptr;
        p.flags = params->flags | HULK_B200_F_ASYNC_INPUT;               // the copies of the G chunks overlap; push waits for all
        hulk_b200_ctx *c = nullptr;
        const int rc = hulk_b200_create(&p, &c);
        if (rc) {
            g_group_create_err = hulk_b200_last_error(nullptr);
            hulk_b200_group_destroy(g);
            return rc;
        }
        g->ctx.push_back(c);
    }
    if (ngpus > 1)
        for (uint32_t i = 0; i < ngpus; i++) {
            const int rc = hulk_b200_peer_connect_local(g->ctx[i], ngpus, i, g->ctx.data());
            if (rc) {
                g_group_create_err = hulk_b200_last_error(g->ctx[i]);
                hulk_b200_group_destroy(g);
                return rc;
            }
        }
    hulk_b200_stats st;
    (void)st;
    g->D = params->num_bins ? params->num_bins : (int32_t)((uint64_t)params->k * params->k * params->k * params->k);
    *out = g;
    return HULK_B200_OK;
}

uint32_t hulk_b200_group_size(const hulk_b200_group *g) { return g ? (uint32_t)g->ctx.size() : 0; }
hulk_b200_ctx *hulk_b200_group_member(hulk_b200_group *g, uint32_t i) { return (g && i < g->ctx.size()) ? g->ctx[i] : nullptr; }

int hulk_b200_group_set_cws_tables(hulk_b200_group *g, const double *r, const double *c, const double *b) {
    if (!g) return HULK_B200_EARG;
    if (g->P.sketch_size && (!r || !c || !b)) return gfail(g, HULK_B200_EARG, "table pointer is NULL");
    for (size_t i = 0; i < g->ctx.size(); i++) {
        const size_t off = (size_t)g->slot_begin[i] * (size_t)g->D;       // rows of the full s x D tables
        const int rc = hulk_b200_set_cws_tables(g->ctx[i], r + off, c + off, b + off);
        if (rc) return gfrom(g, i, rc);
    }
    return HULK_B200_OK;
}
static int group_finish_tables(hulk_b200_group *g) {               // join the draw (if any) and hand the rows out
    if (!g->gen_pending) return HULK_B200_OK;
    if (g->gen_thread.joinable()) g->gen_thread.join();
    g->gen_pending = false;
    int rc = g->gen_rc;
    if (rc) gfail(g, rc, "CWS table generation");
    else rc = hulk_b200_group_set_cws_tables(g, g->gen_r.data(), g->gen_c.data(), g->gen_b.data());
    std::vector<double>().swap(g->gen_r);
    std::vector<double>().swap(g->gen_c);
    std::vector<double>().swap(g->gen_b);
    return rc;
}
int hulk_b200_group_generate_cws_tables_device(hulk_b200_group *g) {
    if (!g) return HULK_B200_EARG;
    for (size_t i = 0; i < g->ctx.size(); i++) {                          // every GPU draws the rows of its own slots
        const int rc = hulk_b200_generate_cws_tables_device(g->ctx[i]);
        if (rc) return gfrom(g, i, rc);
    }
    return HULK_B200_OK;
}
int hulk_b200_group_generate_cws_tables(hulk_b200_group *g, int background) {
    if (!g) return HULK_B200_EARG;
    if (g->gen_pending) return gfail(g, HULK_B200_ESTATE, "table generation already running");
    {
        // HULK_B200_CWS_DEVICE=1: draw on the GPUs instead (the values can differ from the host draw in the last bit,
        // see hulk_b200_generate_cws_tables_device -- which is why it is not the default of a drop-in)
        const char *e = getenv("HULK_B200_CWS_DEVICE");
        if (e && *e == '1') return hulk_b200_group_generate_cws_tables_device(g);
    }
    const size_t n = (size_t)g->P.sketch_size * (size_t)g->D;
    try {
        g->gen_r.resize(n ? n : 1); g->gen_c.resize(n ? n : 1); g->gen_b.resize(n ? n : 1);
    } catch (const std::bad_alloc &) {
        return gfail(g, HULK_B200_ENOMEM, "host memory for the CWS tables");
    }
    g->gen_pending = true;
    g->gen_thread = std::thread([g] {
        g->gen_rc = hulk_b200_new_cws(g->P.sketch_size, g->D, 0, g->P.sketch_size, g->gen_r.data(), g->gen_c.data(),
                                      g->gen_b.data());
    });
    return background ? HULK_B200_OK : group_finish_tables(g);
}

static int push_split(hulk_b200_group *g, const uint8_t *bases, const uint64_t *offsets, uint64_t n_reads,
                      uint32_t read_len) {
    const uint64_t G = g->ctx.size();
    for (uint64_t i = 0; i < G; i++) {
        const uint64_t lo = (i * n_reads) / G, hi = ((i + 1) * n_reads) / G;
        if (hi == lo) continue;
        const int rc = offsets ? hulk_b200_push_reads(g->ctx[i], bases, offsets + lo, hi - lo)
                               : hulk_b200_push_reads_fixed(g->ctx[i], bases + lo * (uint64_t)read_len, hi - lo, read_len);
        if (rc) return gfrom(g, i, rc);
    }
    if (!(g->P.flags & HULK_B200_F_ASYNC_INPUT))                          // the caller's buffer is free on return
        for (uint64_t i = 0; i < G; i++) {
            const int rc = hulk_b200_sync_inputs(g->ctx[i]);
            if (rc) return gfrom(g, i, rc);
        }
    return HULK_B200_OK;
}
int hulk_b200_group_push_reads(hulk_b200_group *g, const uint8_t *bases, const uint64_t *offsets, uint64_t n_reads) {
    if (!g) return HULK_B200_EARG;
    if (n_reads == 0) return HULK_B200_OK;
    if (!offsets || (!bases && offsets[n_reads] != offsets[0])) return gfail(g, HULK_B200_EARG, "bases/offsets is NULL");
    return push_split(g, bases, offsets, n_reads, 0);
}
int hulk_b200_group_push_reads_fixed(hulk_b200_group *g, const uint8_t *bases, uint64_t n_reads, uint32_t read_len) {
    if (!g) return HULK_B200_EARG;
    if (n_reads == 0) return HULK_B200_OK;
    if (!bases) return gfail(g, HULK_B200_EARG, "bases is NULL");
    if (read_len < 1) return gfail(g, HULK_B200_EEMPTYSEQ);
    if (read_len < g->P.w + g->P.k - 1) return gfail(g, HULK_B200_ESHORTSEQ);
    return push_split(g, bases, nullptr, n_reads, read_len);
}
int hulk_b200_group_sync_inputs(hulk_b200_group *g) {
    if (!g) return HULK_B200_EARG;
    for (size_t i = 0; i < g->ctx.size(); i++) {
        const int rc = hulk_b200_sync_inputs(g->ctx[i]);
        if (rc) return gfrom(g, i, rc);
    }
    return HULK_B200_OK;
}

int hulk_b200_group_flush(hulk_b200_group *g) {
    if (!g) return HULK_B200_EARG;
    { const int rc = group_finish_tables(g); if (rc) return rc; }
    // nothing in a member's flush waits on the host, so enqueueing them one after the other cannot deadlock:
    // GPU i's chain waits (on the device) for the "counted" flags the later members' calls set
    for (size_t i = 0; i < g->ctx.size(); i++) {
        const int rc = hulk_b200_flush(g->ctx[i]);
        if (rc) return gfrom(g, i, rc);
    }
    return HULK_B200_OK;
}
int hulk_b200_group_sync(hulk_b200_group *g) {
    if (!g) return HULK_B200_EARG;
    int first = HULK_B200_OK;
    for (size_t i = 0; i < g->ctx.size(); i++) {
        const int rc = hulk_b200_sync(g->ctx[i]);
        if (rc && !first) first = gfrom(g, i, rc);
    }
    return first;
}
int hulk_b200_group_finish(hulk_b200_group *g, uint64_t *mins, double *weights) {
    if (!g) return HULK_B200_EARG;
    { const int rc = hulk_b200_group_sync(g); if (rc) return rc; }
    if (g->P.sketch_size && (!mins || !weights)) return gfail(g, HULK_B200_EARG, "mins/weights is NULL");
    for (size_t i = 0; i < g->ctx.size(); i++) {
        const int rc = hulk_b200_finish(g->ctx[i], mins + g->slot_begin[i], weights + g->slot_begin[i]);
        if (rc) return gfrom(g, i, rc);
    }
    return HULK_B200_OK;
}
int hulk_b200_group_reset(hulk_b200_group *g) {
    if (!g) return HULK_B200_EARG;
    { const int rc = hulk_b200_group_sync(g); if (rc && rc != HULK_B200_ESPARSE) return rc; }
    for (size_t i = 0; i < g->ctx.size(); i++) {
        const int rc = hulk_b200_reset(g->ctx[i]);
        if (rc) return gfrom(g, i, rc);
    }
    g->err.clear();
    return HULK_B200_OK;
}
int hulk_b200_group_get_stats(hulk_b200_group *g, hulk_b200_stats *out) {
    if (!g || !out) return HULK_B200_EARG;
    hulk_b200_stats sum{};
    for (size_t i = 0; i < g->ctx.size(); i++) {
        hulk_b200_stats st;
        const int rc = hulk_b200_get_stats(g->ctx[i], &st);
        if (rc) return gfrom(g, i, rc);
        sum.n_reads += st.n_reads;                       // seqCount over all chunks
        sum.n_bases += st.n_bases;
        sum.n_minimizers += st.n_minimizers;             // theBoss.GetMinimizerCount()
        sum.n_kernel_launches += st.n_kernel_launches;
        sum.n_rescans += st.n_rescans;
        sum.h2d_bytes += st.h2d_bytes;
        sum.d2h_bytes += st.d2h_bytes;
        sum.pack_ns += st.pack_ns;
        sum.n_packed_batches += st.n_packed_batches;
        if (i == 0) {                                    // every member flushes the same summed spectrum
            sum.n_flushes = st.n_flushes;
            sum.n_adds = st.n_adds;
        }
    }
    *out = sum;
    return HULK_B200_OK;
}

// ---- MinHash side sketches over the group: each member sketches its share, the parts are merged on the host
int hulk_b200_group_minhash_enable(hulk_b200_group *g, int kmv, int khf) {
    if (!g) return HULK_B200_EARG;
    for (size_t i = 0; i < g->ctx.size(); i++) {
        const int rc = hulk_b200_minhash_enable(g->ctx[i], kmv, khf);
        if (rc) return gfrom(g, i, rc);
    }
    return HULK_B200_OK;
}
int hulk_b200_group_get_khf(hulk_b200_group *g, uint64_t *mins) {
    if (!g || !mins) return HULK_B200_EARG;
    const uint32_t s = g->P.sketch_size;
    std::vector<uint64_t> part(s);
    for (size_t i = 0; i < g->ctx.size(); i++) {
        const int rc = hulk_b200_get_khf(g->ctx[i], i == 0 ? mins : part.data());
        if (rc) return gfrom(g, i, rc);
        if (i)
            for (uint32_t j = 0; j < s; j++) mins[j] = std::min(mins[j], part[j]);      // KHFsketch.Merge, khf.go:47-55
    }
    return HULK_B200_OK;
}
int hulk_b200_group_get_kmv(hulk_b200_group *g, uint64_t *mins, uint32_t *n) {
    if (!g || !mins || !n) return HULK_B200_EARG;
    const uint32_t s = g->P.sketch_size;
    std::vector<uint64_t> all, part(s);
    for (size_t i = 0; i < g->ctx.size(); i++) {
        uint32_t np = 0;
        const int rc = hulk_b200_get_kmv(g->ctx[i], part.data(), &np);
        if (rc) return gfrom(g, i, rc);
        all.insert(all.end(), part.begin(), part.begin() + np);
    }
    std::sort(all.begin(), all.end());
    if (all.size() > s) all.resize(s);
    std::copy(all.begin(), all.end(), mins);
    *n = (uint32_t)all.size();
    return HULK_B200_OK;
}

// SeqMinimizer.Run's loop over a reader (src/pipeline/sketch.go:197-224), the group's form of hulk_b200_sketch_reader
int hulk_b200_group_sketch_reader(hulk_b200_group *g, hulk_b200_reader *rd, uint64_t interval, hulk_b200_log_fn log,
                                  void *user) {
    if (!g || !rd) return HULK_B200_EARG;
    char line[128];
    uint64_t seq_count = 0, sketching_interval = 0;
    for (;;) {
        const uint8_t *bases = nullptr;
        const uint64_t *offsets = nullptr;
        uint64_t n = 0;
        const int rc = hulk_b200_reader_next(rd, &bases, &offsets, &n);
        if (rc) {
            g->err = hulk_b200_reader_error(rd);
            return rc;
        }
        if (n == 0) break;
        uint64_t done = 0;
        while (done < n) {
            uint64_t take = n - done;
            if (interval) take = std::min<uint64_t>(take, interval - (seq_count % interval));
            const int prc = push_split(g, bases, offsets + done, take, 0);               // theBoss.AddSeq  :200
            if (prc) return prc;
            const uint64_t before = seq_count;
            seq_count += take;
            done += take;
            if (log)
                for (uint64_t m = before / 100000 + 1; m * 100000 <= seq_count; m++) {   // :203-207
                    snprintf(line, sizeof line, "\tprocessed %llu sequences", (unsigned long long)(m * 100000));
                    log(user, line);
                }
            if (interval && seq_count % interval == 0) {                                 // :211-215
                sketching_interval++;
                if (log) {
                    snprintf(line, sizeof line, "\treached interval %llu -> histosketching",
                             (unsigned long long)sketching_interval);
                    log(user, line);
                }
                const int frc = hulk_b200_group_flush(g);
                if (frc) return frc;
            }
        }
        // the reader's batch is lent until the next call: every member must have copied its chunk
        const int irc = hulk_b200_group_sync_inputs(g);
        if (irc) return irc;
    }
    if (log) log(user, "generating final histosketch of k-mer spectra...");               // :220
    const int frc = hulk_b200_group_flush(g);                                            // :221
    if (frc) return frc;
    const int src = hulk_b200_group_sync(g);
    if (src) return src;
    if (seq_count == 0) return gfail(g, HULK_B200_ENOSEQ);                               // :237-239
    return HULK_B200_OK;
}

}  // extern "C"
