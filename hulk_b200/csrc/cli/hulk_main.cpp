// hulk_main.cpp -- `hulk sketch` front end over libhulk_b200.so: the reference's flags, log lines and
// JSON output, with everything numeric running on the GPU behind the C ABI (include/hulk_b200.h).
//
// The reference's host is Go (cobra); no Go toolchain exists in this image, so the tested host above
// the C ABI is this C++ program (and the Python mirror).  What it follows:
//   cmd/root.go:62-66        persistent flags  -k/--kmerSize -o/--outFile --log -p/--processors --profiling
//   cmd/sketch.go:50-59      sketch flags      -f/--fastq --fasta -w -i -s -x --stream -b --khf --kmv
//   cmd/sketch.go:64-179     runSketch: log lines, parameter checks, pipeline wiring
//   cmd/sketch.go:185-214    sketchParamCheck
//   src/pipeline/sketch.go:182-301  SeqMinimizer.Run / Sketcher.Run log lines and the output file
//   cmd/smash.go             `hulk smash`: flags, checks, the similarity matrix CSV
// Not here (out of scope): --profiling (pprof); KMV/KHF side sketches are unfed as in the reference
// unless HULK_B200_FEED_MINHASH=1
// (unwired in the reference, src/pipeline/boss.go:18-19: the flags are accepted and logged, like there).
#include <dirent.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <thread>
#include <vector>

#include "hulk_b200.h"

namespace {

FILE *g_log = stdout;

// Go's log package with LstdFlags: "2006/01/02 15:04:05 " + message + '\n'
void logf(const char *fmt, ...) {
    char msg[4096];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(msg, sizeof msg, fmt, ap);
    va_end(ap);
    struct timespec now;
    clock_gettime(CLOCK_REALTIME, &now);
    const time_t t = now.tv_sec;
    struct tm tmv;
    localtime_r(&t, &tmv);
    char ts[48];
    size_t tl = strftime(ts, sizeof ts, "%Y/%m/%d %H:%M:%S", &tmv);
    static const bool micros = getenv("HULK_LOG_MICROSECONDS") != nullptr;     // Go's log.Lmicroseconds, for timing runs
    if (micros) snprintf(ts + tl, sizeof ts - tl, ".%06ld", now.tv_nsec / 1000);
    size_t n = strlen(msg);
    while (n && msg[n - 1] == '\n') msg[--n] = 0;
    fprintf(g_log, "%s %s\n", ts, msg);
    fflush(g_log);
}
void log_line(void *, const char *line) { logf("%s", line); }

[[noreturn]] void fatal(const std::string &msg) {     // helpers.ErrorCheck -> log.Fatalf("ERROR---> %v\n")
    logf("ERROR---> %s", msg.c_str());
    exit(1);
}

struct Options {
    // root (cmd/root.go:62-66)
    unsigned long kmer_size = 21;
    std::string out_file, log_file;
    long proc = 1;
    bool profiling = false;
    // sketch (cmd/sketch.go:50-59)
    std::vector<std::string> fastq;
    bool fasta = false;
    unsigned long window_size = 9, interval = 0, sketch_size = 50;
    double decay_ratio = 1.0;
    bool streaming = false;
    std::string banner_label = "blank";
    bool add_khf = false, add_kmv = false;
    // this build only
    long device = 0;        // first CUDA device ordinal
    long gpus = 1;          // GPUs to spread the sketch over (devices device .. device + gpus - 1)
};

struct FlagDef {
    const char *name;
    char shorthand;
    enum Kind { UINT, INT, FLOAT, STRING, SLICE, BOOL } kind;
    void *dst;
    const char *usage;
};

const char *g_usage_cmd = "sketch";
void usage_sketch(const std::vector<FlagDef> &defs, FILE *out) {
    fprintf(out, "Usage:\n  hulk %s [flags]\n\nFlags:\n", g_usage_cmd);
    for (const FlagDef &d : defs) {
        if (d.shorthand) fprintf(out, "  -%c, --%-14s %s\n", d.shorthand, d.name, d.usage);
        else fprintf(out, "      --%-14s %s\n", d.name, d.usage);
    }
}

[[noreturn]] void flag_error(const std::vector<FlagDef> &defs, const std::string &msg) {
    printf("Error: %s\n", msg.c_str());
    usage_sketch(defs, stdout);
    printf("\n%s\n", msg.c_str());
    exit(1);
}

void set_flag(const std::vector<FlagDef> &defs, const FlagDef &d, const std::string &v, const std::string &shown) {
    char *end = nullptr;
    errno = 0;
    switch (d.kind) {
        case FlagDef::UINT: {
            if (v.empty() || v[0] == '-') flag_error(defs, "invalid argument \"" + v + "\" for \"" + shown + "\" flag");
            const unsigned long x = strtoul(v.c_str(), &end, 0);
            if (*end || errno) flag_error(defs, "invalid argument \"" + v + "\" for \"" + shown + "\" flag");
            *static_cast<unsigned long *>(d.dst) = x;
            break;
        }
        case FlagDef::INT: {
            const long x = strtol(v.c_str(), &end, 0);
            if (v.empty() || *end || errno) flag_error(defs, "invalid argument \"" + v + "\" for \"" + shown + "\" flag");
            *static_cast<long *>(d.dst) = x;
            break;
        }
        case FlagDef::FLOAT: {
            const double x = strtod(v.c_str(), &end);
            if (v.empty() || *end) flag_error(defs, "invalid argument \"" + v + "\" for \"" + shown + "\" flag");
            *static_cast<double *>(d.dst) = x;
            break;
        }
        case FlagDef::STRING:
            *static_cast<std::string *>(d.dst) = v;
            break;
        case FlagDef::SLICE: {                              // pflag StringSlice: comma separated, repeats append
            auto *dst = static_cast<std::vector<std::string> *>(d.dst);
            size_t at = 0;
            for (;;) {
                const size_t c = v.find(',', at);
                dst->push_back(v.substr(at, c == std::string::npos ? c : c - at));
                if (c == std::string::npos) break;
                at = c + 1;
            }
            break;
        }
        case FlagDef::BOOL: {
            bool x;
            if (v == "true" || v == "1" || v == "t" || v == "T" || v == "TRUE" || v == "True") x = true;
            else if (v == "false" || v == "0" || v == "f" || v == "F" || v == "FALSE" || v == "False") x = false;
            else flag_error(defs, "invalid argument \"" + v + "\" for \"" + shown + "\" flag");
            *static_cast<bool *>(d.dst) = x;
            break;
        }
    }
}

// pflag-style parsing: --name value, --name=value, -n value, -nvalue, -n=value, grouped bool shorthands
void parse_flags(const std::vector<FlagDef> &defs, int argc, char **argv, int first) {
    for (int i = first; i < argc; i++) {
        const std::string a = argv[i];
        if (a == "--") break;
        if (a.size() >= 3 && a[0] == '-' && a[1] == '-') {
            const size_t eq = a.find('=');
            const std::string name = a.substr(2, eq == std::string::npos ? eq : eq - 2);
            if (name == "help") { usage_sketch(defs, stdout); exit(0); }
            const FlagDef *d = nullptr;
            for (const FlagDef &x : defs)
                if (name == x.name) d = &x;
            if (!d) flag_error(defs, "unknown flag: --" + name);
            if (eq != std::string::npos) set_flag(defs, *d, a.substr(eq + 1), "--" + name);
            else if (d->kind == FlagDef::BOOL) *static_cast<bool *>(d->dst) = true;
            else if (i + 1 < argc) set_flag(defs, *d, argv[++i], "--" + name);
            else flag_error(defs, "flag needs an argument: --" + name);
        } else if (a.size() >= 2 && a[0] == '-') {
            for (size_t j = 1; j < a.size(); j++) {
                const char c = a[j];
                if (c == 'h') { usage_sketch(defs, stdout); exit(0); }
                const FlagDef *d = nullptr;
                for (const FlagDef &x : defs)
                    if (x.shorthand == c) d = &x;
                if (!d) flag_error(defs, std::string("unknown shorthand flag: '") + c + "' in " + a);
                const std::string shown = std::string("-") + c + ", --" + d->name;
                if (d->kind == FlagDef::BOOL) { *static_cast<bool *>(d->dst) = true; continue; }
                std::string rest = a.substr(j + 1);
                if (!rest.empty()) {
                    if (rest[0] == '=') rest = rest.substr(1);
                    set_flag(defs, *d, rest, shown);
                } else if (i + 1 < argc) {
                    set_flag(defs, *d, argv[++i], shown);
                } else {
                    flag_error(defs, std::string("flag needs an argument: '") + c + "' in " + a);
                }
                break;
            }
        }
        // positional arguments are ignored, as cobra does for a command without Args validation
    }
}

bool is_not_exist(const std::string &p) {
    struct stat st;
    return stat(p.c_str(), &st) != 0 && errno == ENOENT;
}
std::string dir_of(const std::string &p) {                 // filepath.Dir
    const size_t s = p.find_last_of('/');
    if (s == std::string::npos) return ".";
    std::string d = p.substr(0, s);
    while (d.size() > 1 && d.back() == '/') d.pop_back();
    if (d.empty()) return "/";
    // filepath.Clean of the common "./x" shapes
    while (d.size() > 2 && d.compare(0, 2, "./") == 0) d = d.substr(2);
    return d;
}
bool mkdir_all(const std::string &p, mode_t mode) {
    std::string cur;
    size_t at = 0;
    while (at <= p.size()) {
        const size_t s = p.find('/', at);
        cur = p.substr(0, s == std::string::npos ? p.size() : s);
        at = (s == std::string::npos) ? p.size() + 1 : s + 1;
        if (cur.empty()) continue;
        if (mkdir(cur.c_str(), mode) != 0 && errno != EEXIST) return false;
    }
    return true;
}
std::vector<std::string> split(const std::string &s, char c) {
    std::vector<std::string> out;
    size_t at = 0;
    for (;;) {
        const size_t p = s.find(c, at);
        out.push_back(s.substr(at, p == std::string::npos ? p : p - at));
        if (p == std::string::npos) break;
        at = p + 1;
    }
    return out;
}

// helpers.CheckFile / CheckExt (src/helpers/helpers.go:112-141)
void check_file(const std::string &file) {
    struct stat st;
    if (stat(file.c_str(), &st) != 0) {
        if (errno == ENOENT) fatal("file does not exist: " + file);
        fatal("can't access file (check permissions): " + file);
    }
}
void check_ext(const std::string &file) {
    const std::vector<std::string> parts = split(file, '.');
    size_t last = parts.size() - 1;
    if (parts[last] == "gz" && last > 0) last--;
    else if (parts[last] == "gz") fatal("file does not have recognised extension: " + file);
    for (const char *e : {"fastq", "fq", "fasta", "fna", "fa"})
        if (parts[last] == e) return;
    fatal("file does not have recognised extension: " + file);
}

std::string go_duration(double seconds) {                  // time.Duration.String()
    const long long ns = (long long)(seconds * 1e9);
    auto trimmed = [](long long whole, long long frac, int digits) {   // whole.frac without trailing zeros
        char buf[64];
        snprintf(buf, sizeof buf, "%lld.%0*lld", whole, digits, frac);
        std::string s = buf;
        while (s.back() == '0') s.pop_back();
        if (s.back() == '.') s.pop_back();
        return s;
    };
    if (ns < 1000) return std::to_string(ns) + "ns";
    if (ns < 1000000) return trimmed(ns / 1000, ns % 1000, 3) + "\xc2\xb5s";
    if (ns < 1000000000) return trimmed(ns / 1000000, ns % 1000000, 6) + "ms";
    const long long total_s = ns / 1000000000, h = total_s / 3600, m = (total_s / 60) % 60;
    std::string out;
    if (h) out += std::to_string(h) + "h";
    if (h || m) out += std::to_string(m) + "m";
    return out + trimmed(total_s % 60, ns % 1000000000, 9) + "s";
}

int run_sketch(int argc, char **argv) {
    Options o;
    char stamp[32];
    const time_t t0 = time(nullptr);
    struct tm tmv;
    localtime_r(&t0, &tmv);
    strftime(stamp, sizeof stamp, "%Y%m%d%H%M%S", &tmv);
    o.out_file = std::string("./hulk-") + stamp;            // cmd/root.go:35
    const std::vector<FlagDef> defs = {
        {"fastq", 'f', FlagDef::SLICE, &o.fastq, "FASTQ file(s) to sketch (can also pipe in STDIN)"},
        {"fasta", 0, FlagDef::BOOL, &o.fasta, "tells HULK that the input file is actually FASTA format (.fna/.fasta/.fa), not FASTQ (experimental feature)"},
        {"windowSize", 'w', FlagDef::UINT, &o.window_size, "minimizer window size (default 9)"},
        {"interval", 'i', FlagDef::UINT, &o.interval, "size of k-mer sampling interval (default 0 (= no interval))"},
        {"sketchSize", 's', FlagDef::UINT, &o.sketch_size, "size of sketch (default 50)"},
        {"decayRatio", 'x', FlagDef::FLOAT, &o.decay_ratio, "decay ratio used for concept drift (1.0 = concept drift disabled) (default 1)"},
        {"stream", 0, FlagDef::BOOL, &o.streaming, "prints the sketches to STDOUT after every interval is reached, whilst still writting them to disk (log file is redirected to disk))"},
        {"bannerLabel", 'b', FlagDef::STRING, &o.banner_label, "adds a label to the sketch object, for use with BANNER (default \"blank\")"},
        {"khf", 0, FlagDef::BOOL, &o.add_khf, "also generate a MinHash K-Hash Functions sketch"},
        {"kmv", 0, FlagDef::BOOL, &o.add_kmv, "also generate a MinHash K-Minimum Values (bottom-k) sketch"},
        {"kmerSize", 'k', FlagDef::UINT, &o.kmer_size, "minimizer k-mer length (default 21)"},
        {"outFile", 'o', FlagDef::STRING, &o.out_file, "directory and basename for saving the outfile(s)"},
        {"log", 0, FlagDef::STRING, &o.log_file, "filename for log file, if omitted then STDOUT used by default"},
        {"processors", 'p', FlagDef::INT, &o.proc, "number of processors to use (default 1)"},
        {"profiling", 0, FlagDef::BOOL, &o.profiling, "create the files needed to profile HULK using the go tool pprof"},
        {"device", 0, FlagDef::INT, &o.device, "CUDA device ordinal (this build; default 0)"},
        {"gpus", 0, FlagDef::INT, &o.gpus, "number of GPUs to spread the reads and the sketch slots over, starting at --device (this build; default 1)"},
    };
    parse_flags(defs, argc, argv, 2);

    // cmd/sketch.go:74-87
    if (o.streaming) {
        if (o.log_file.empty()) o.log_file = o.out_file + ".log";
        const std::string dir = dir_of(o.log_file);       // helpers.StartLogging
        if (o.log_file.find('/') != std::string::npos && is_not_exist(dir) && !mkdir_all(dir, 0700)) {
            fprintf(stderr, "can't create specified directory for log\n");
            return 1;
        }
        g_log = fopen(o.log_file.c_str(), "a");
        if (!g_log) { perror(o.log_file.c_str()); return 1; }
    }

    const auto start = std::chrono::steady_clock::now();
    logf("this is hulk (version %s)", hulk_b200_version());
    logf("please cite Rowe et al. 2019, doi: https://doi.org/10.1186/s40168-019-0653-2");
    logf("starting the sketch subcommand");
    logf("checking parameters...");

    // sketchParamCheck (cmd/sketch.go:185-214)
    const std::string out_dir = dir_of(o.out_file);
    if (out_dir != "." && is_not_exist(out_dir) && !mkdir_all(out_dir, 0700))
        fatal(std::string("can't create specified output directory: mkdir ") + out_dir + ": " + strerror(errno));
    const long ncpu = (long)std::max(1u, std::thread::hardware_concurrency());
    if (o.proc <= 0 || o.proc > ncpu) o.proc = ncpu;
    if (o.fastq.empty()) {
        struct stat st;
        if (fstat(0, &st) != 0) fatal("error with STDIN");
        if (!S_ISFIFO(st.st_mode)) fatal("no STDIN found");
        logf("\tinput file: using STDIN");
    } else {
        for (const std::string &f : o.fastq) {
            check_file(f);
            check_ext(f);
        }
    }
    logf(o.fasta ? "\tmode: FASTA" : "\tmode: FASTQ");
    logf("\tno. processors: %ld", o.proc);
    logf("\tminimizer k-mer size: %lu", o.kmer_size);
    logf("\tminimizer window size: %lu", o.window_size);
    logf("\tsketch size: %lu", o.sketch_size);
    logf(o.streaming ? "\tstreaming: enabled" : "\tstreaming: disabled");
    if (o.decay_ratio == 1.0) {
        logf("\tconcept drift: disabled");
    } else {
        logf("\tconcept drift: enabled");
        logf("\tdecay ratio: %.2f", o.decay_ratio);
    }
    // spectrumSize := int32(helpers.Pow(k, 4)) in uint (64-bit) arithmetic  cmd/sketch.go:118
    const uint64_t k64 = o.kmer_size;
    const int32_t spectrum = (int32_t)(uint32_t)(k64 * k64 * k64 * k64);
    logf("\tnumber of bins in k-mer spectrum: %d", spectrum);
    logf("\tadding KHF sketch: %s", o.add_khf ? "true" : "false");
    logf("\tadding KMV sketch: %s", o.add_kmv ? "true" : "false");

    std::string file_name = "STDIN";                        // cmd/sketch.go:147-155 (trailing comma kept)
    if (!o.fastq.empty()) {
        file_name.clear();
        for (const std::string &f : o.fastq) file_name += f + ",";
    }

    logf("initialising sketching pipeline...");
    logf("\tinitialising the processes");
    logf("\tconnecting data streams");
    logf("\tnumber of processes added to the sketching pipeline: %d", 4);
    logf("\tnumber of minions in the sketching pool: %ld", o.proc);

    // one GPU is used: hiding the others from the CUDA runtime keeps its start-up proportional to one device
    // (on an 8-GPU node the runtime otherwise initialises all of them before the first allocation returns)
    if (o.gpus < 1 || o.gpus > 16) fatal("--gpus must be between 1 and 16");
    // test switch: all members of the group on --device (the multi-GPU mechanism on a single-GPU box)
    const bool one_device = getenv("HULK_B200_GPUS_ON_ONE_DEVICE") != nullptr;
    if (one_device) {
        if (!getenv("CUDA_VISIBLE_DEVICES")) {
            setenv("CUDA_VISIBLE_DEVICES", std::to_string(o.device).c_str(), 1);
            o.device = 0;
        }
    } else if (!getenv("CUDA_VISIBLE_DEVICES")) {
        std::string vis;
        for (long g = 0; g < o.gpus; g++) vis += (g ? "," : "") + std::to_string(o.device + g);
        setenv("CUDA_VISIBLE_DEVICES", vis.c_str(), 1);
        o.device = 0;
    }
    // the reader starts first: reading/inflating overlaps context creation and the CWS table draw
    std::vector<const char *> paths;
    for (const std::string &f : o.fastq) paths.push_back(f.c_str());
    hulk_b200_reader *rd = nullptr;
    int rc = hulk_b200_reader_open(paths.data(), (uint32_t)paths.size(), o.fasta ? 1 : 0, 0, &rd);
    if (rc) fatal(hulk_b200_strerror(rc));

    logf("finding minimizers...");                          // SeqMinimizer.Run  src/pipeline/sketch.go:183
    hulk_b200_params P;
    memset(&P, 0, sizeof P);
    P.k = (uint32_t)std::min<unsigned long>(o.kmer_size, 0xffffffffu);
    P.w = (uint32_t)std::min<unsigned long>(o.window_size, 0xffffffffu);
    P.sketch_size = (uint32_t)std::min<unsigned long>(o.sketch_size, 0xffffffffu);
    P.num_bins = spectrum;
    P.decay_ratio = o.decay_ratio;
    // one handle for --gpus N (a group of one is the plain context): every interval's reads are split over the GPUs,
    // the spectrum a flush works on is summed over NVLink, the sketch slots are sharded -- same JSON as one GPU
    std::vector<int32_t> devs;
    for (long g = 0; g < o.gpus; g++) devs.push_back((int32_t)(o.device + (one_device ? 0 : g)));
    hulk_b200_group *ctx = nullptr;
    rc = hulk_b200_group_create(&P, devs.data(), (uint32_t)devs.size(), &ctx);   // findMinimizers + NewHistoSketch parameter checks
    if (rc) fatal(hulk_b200_group_last_error(nullptr));
    rc = hulk_b200_group_generate_cws_tables(ctx, 1);       // NewHistoSketch -> newCWS, drawn while the reads are counted
    if (rc) fatal(hulk_b200_group_last_error(ctx));
    // HULK_B200_FEED_MINHASH=1 (not a reference flag): --kmv / --khf get the sketches the two types were written to
    // produce -- every minimizer also goes to their AddHash -- instead of the reference's unfed ones (see below)
    const char *feed_env = getenv("HULK_B200_FEED_MINHASH");
    const bool feed_minhash = feed_env && *feed_env == '1' && (o.add_kmv || o.add_khf);
    if (feed_minhash) {
        rc = hulk_b200_group_minhash_enable(ctx, o.add_kmv ? 1 : 0, o.add_khf ? 1 : 0);
        if (rc) fatal(hulk_b200_group_last_error(ctx));
    }

    rc = hulk_b200_group_sketch_reader(ctx, rd, o.interval, log_line, nullptr);
    if (rc) {
        const char *rerr = hulk_b200_reader_error(rd);
        if (rc == HULK_B200_EFASTQ || rc == HULK_B200_ETOOLONG || rc == HULK_B200_EIO) {
            // log.Fatal(err) in the reader processes: no "ERROR--->" prefix  src/pipeline/sketch.go:53-55,75-77,150-152
            logf("%s", (rerr && *rerr) ? rerr : hulk_b200_strerror(rc));
            return 1;
        }
        if (rc == HULK_B200_ENOSEQ) fatal("no sequences received");
        fatal(hulk_b200_group_last_error(ctx));
    }
    hulk_b200_stats st;
    rc = hulk_b200_group_get_stats(ctx, &st);
    if (rc) fatal(hulk_b200_group_last_error(ctx));
    const unsigned long mean_rl = (unsigned long)((double)st.n_bases / (double)st.n_reads);
    logf("\tprocessed %llu sequences in total", (unsigned long long)st.n_reads);
    logf("\tmean sequence length: %lu", mean_rl);
    logf("\tfound %llu minimizers", (unsigned long long)st.n_minimizers);
    logf("\thistosketching across %d bins", spectrum);
    logf(o.proc > 1 ? "merging sketches and cleaning up..." : "cleaning up...");

    std::vector<uint64_t> mins(P.sketch_size);
    std::vector<double> weights(P.sketch_size);
    rc = hulk_b200_group_finish(ctx, mins.data(), weights.data());
    if (rc) fatal(hulk_b200_group_last_error(ctx));
    const std::string out_json = o.out_file + ".json";
    // --kmv / --khf as the reference behaves today: the boss builds both MinHash sketches but nothing feeds them
    // (src/pipeline/boss.go:18-19,70-71 "not used yet").  The KMV heap is therefore empty and HULKdata.Add refuses
    // it (src/sketchio/sketchio.go:59-61) -- a fatal error after the histosketch was added and before the KHF one
    // (src/pipeline/sketch.go:227-234,289-294); the KHF sketch is written with its initial MaxUint64 in every slot
    // (src/minhash/khf.go:20-32).
    if (P.sketch_size == 0) fatal("no sketch was generated by the histosketch algorithm");
    std::vector<uint64_t> khf, kmv;
    uint32_t kmv_n = 0;
    if (feed_minhash) {
        if (o.add_kmv) {
            kmv.assign(P.sketch_size, 0);
            rc = hulk_b200_group_get_kmv(ctx, kmv.data(), &kmv_n);
            if (rc) fatal(hulk_b200_group_last_error(ctx));
            if (kmv_n == 0) fatal("no sketch was generated by the kmv algorithm");
        }
        if (o.add_khf) {
            khf.assign(P.sketch_size, ~0ull);
            rc = hulk_b200_group_get_khf(ctx, khf.data());
            if (rc) fatal(hulk_b200_group_last_error(ctx));
        }
    } else {
        if (o.add_kmv) fatal("no sketch was generated by the kmv algorithm");
        if (o.add_khf) khf.assign(P.sketch_size, ~0ull);
    }
    rc = hulk_b200_write_json_minhash(out_json.c_str(), file_name.c_str(), o.banner_label.c_str(), P.k, mins.data(),
                                      weights.data(), P.sketch_size, spectrum, o.decay_ratio != 1.0,
                                      kmv_n ? kmv.data() : nullptr, kmv_n,
                                      o.add_khf ? khf.data() : nullptr, (uint32_t)khf.size());
    if (rc == HULK_B200_ENOSKETCH) fatal("no sketch was generated by the histosketch algorithm");
    if (rc == HULK_B200_EARG) fatal("json: unsupported value: a sketch weight is not finite");
    if (rc) fatal("open " + out_json + ": " + strerror(errno));
    logf("\twritten sketch to disk: %s", out_json.c_str());
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    logf("finished in %s", go_duration(secs).c_str());
    // the process ends here: unmapping gigabytes of device and pinned memory one buffer at a time
    // (hulk_b200_destroy / reader_close) would only delay the exit -- the driver reclaims it all at once
    fflush(nullptr);
    _exit(0);
}

// ---- hulk smash (cmd/smash.go) ------------------------------------------------------------------------
// encoding/csv field quoting: a field is quoted when it is empty-looking-special, or contains the delimiter,
// a quote, CR or LF, or starts with a space
std::string csv_field(const std::string &f) {
    bool q = f.empty() ? false : (f[0] == ' ' || f == "\\.");
    for (char c : f)
        if (c == ',' || c == '"' || c == '\r' || c == '\n') q = true;
    if (!q) return f;
    std::string o = "\"";
    for (char c : f) {
        if (c == '"') o += "\"\"";
        else o += c;
    }
    return o + "\"";
}
void csv_write(FILE *fh, const std::vector<std::string> &row) {
    for (size_t i = 0; i < row.size(); i++) fprintf(fh, "%s%s", i ? "," : "", csv_field(row[i]).c_str());
    fputc('\n', fh);
}
bool ends_with_json(const std::string &name) {          // filepath.Match("*.json", name)
    return name.size() >= 5 && name.compare(name.size() - 5, 5, ".json") == 0;
}
void collect_jsons(const std::string &dir, bool recursive, std::vector<std::string> &out) {   // helpers.CollectJSONs
    DIR *d = opendir(dir.c_str());
    if (!d) return;
    std::vector<std::string> names;
    while (struct dirent *e = readdir(d)) names.push_back(e->d_name);
    closedir(d);
    std::sort(names.begin(), names.end());
    for (const std::string &nm : names) {
        if (nm == "." || nm == "..") continue;
        const std::string path = dir + nm;
        struct stat st;
        if (stat(path.c_str(), &st) != 0) continue;
        if (S_ISDIR(st.st_mode)) {
            if (recursive) collect_jsons(path + "/", true, out);
        } else if (ends_with_json(nm)) {
            out.push_back(path);
        }
    }
}

int run_smash(int argc, char **argv) {
    unsigned long kmer_size = 21;
    long proc = 1, device = 0;
    bool profiling = false, recursive = false, banner_matrix = false;
    std::string out_file, log_file, sketch_dir = "./", algo = "histosketch", metric = "jaccard";
    char stamp[32];
    const time_t t0 = time(nullptr);
    struct tm tmv;
    localtime_r(&t0, &tmv);
    strftime(stamp, sizeof stamp, "%Y%m%d%H%M%S", &tmv);
    out_file = std::string("./hulk-") + stamp;
    g_usage_cmd = "smash";
    const std::vector<FlagDef> defs = {
        {"sketchDir", 'd', FlagDef::STRING, &sketch_dir, "the directory containing the sketches to smash (compare)... (default \"./\")"},
        {"recursive", 0, FlagDef::BOOL, &recursive, "recursively search the supplied sketch directory (-d)"},
        {"algorithm", 'a', FlagDef::STRING, &algo, "tells HULK which sketching algorithm to use [histosketch kmv khf] (default \"histosketch\")"},
        {"metric", 'm', FlagDef::STRING, &metric, "tells HULK which distance metric to use [jaccard weightedjaccard] (default \"jaccard\")"},
        {"bannerMatrix", 0, FlagDef::BOOL, &banner_matrix, "write a matrix file for banner"},
        {"kmerSize", 'k', FlagDef::UINT, &kmer_size, "minimizer k-mer length (default 21)"},
        {"outFile", 'o', FlagDef::STRING, &out_file, "directory and basename for saving the outfile(s)"},
        {"log", 0, FlagDef::STRING, &log_file, "filename for log file, if omitted then STDOUT used by default"},
        {"processors", 'p', FlagDef::INT, &proc, "number of processors to use (default 1)"},
        {"profiling", 0, FlagDef::BOOL, &profiling, "create the files needed to profile HULK using the go tool pprof"},
        {"device", 0, FlagDef::INT, &device, "CUDA device ordinal (this build; default 0)"},
    };
    parse_flags(defs, argc, argv, 2);
    if (!log_file.empty()) {
        g_log = fopen(log_file.c_str(), "a");
        if (!g_log) { perror(log_file.c_str()); return 1; }
    }
    logf("this is hulk (version %s)", hulk_b200_version());
    logf("starting the smash subcommand");

    // smashParamCheck (cmd/smash.go:100-180)
    if (metric != "jaccard" && metric != "weightedjaccard")
        fatal("supplied distance metric is not available: " + metric + "\nplease select one of the following: [jaccard weightedjaccard]");
    if (algo != "histosketch" && algo != "kmv" && algo != "khf")
        fatal("supplied algorithm not available: " + algo + "\nplease select one of the following: [histosketch kmv khf]");
    const std::string out_dir = dir_of(out_file);
    if (out_dir != "." && is_not_exist(out_dir) && !mkdir_all(out_dir, 0700))
        fatal(std::string("can't create specified output directory: mkdir ") + out_dir + ": " + strerror(errno));
    if (sketch_dir.empty()) fatal("no directory specified");
    {
        struct stat st;
        if (stat(sketch_dir.c_str(), &st) != 0) {
            if (errno == ENOENT) fatal("directory does not exist: " + sketch_dir);
            fatal("can't access adirectory (check permissions): " + sketch_dir);
        }
    }
    if (sketch_dir.back() != '/') sketch_dir += '/';
    std::vector<std::string> files;
    collect_jsons(sketch_dir, recursive, files);
    if (files.empty()) fatal("no JSON files found in supplied directory: " + sketch_dir + "\n");
    std::vector<hulk_b200_sketch_file *> loaded;
    char err[1024];
    for (const std::string &f : files) {
        hulk_b200_sketch_file *sf = nullptr;
        if (hulk_b200_sketch_load(f.c_str(), &sf, err, sizeof err)) fatal(err);
        loaded.push_back(sf);
    }
    if (loaded.size() < 2)
        fatal(std::to_string(loaded.size()) + " sketches found in the supplied directory, HULK needs at least 2 to smash!\n");
    logf("checking parameters and collecting sketches...");
    logf("\talgorithm: %s", algo.c_str());
    logf("\tk-mer size: %lu", kmer_size);
    logf("\tcreate matrix for banner: %s", banner_matrix ? "true" : "false");
    logf("\tnumber of sketch objects: %zu", loaded.size());
    logf("HULK SMASH!");

    // makeMatrix (cmd/smash.go:183-226); `files` is already in sort.Strings order
    std::vector<size_t> order(files.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return files[a] < files[b]; });
    const size_t n = files.size();
    uint32_t s = 0;
    std::vector<uint64_t> mins;
    std::vector<double> weights;
    for (size_t r = 0; r < n; r++) {
        const uint64_t *m = nullptr;
        const double *w = nullptr;
        uint32_t sz = 0;
        if (hulk_b200_sketch_find(loaded[order[r]], (uint32_t)kmer_size, algo.c_str(), &m, &w, &sz, err, sizeof err)) fatal(err);
        if (r == 0) s = sz;
        if (sz != s) fatal("sketch length mismatch: " + std::to_string(s) + " vs " + std::to_string(sz) + "\n");
        if (metric == "weightedjaccard" && (algo != "histosketch" || !w))
            fatal("weighted jaccard is only supported for histosketches");
        mins.insert(mins.end(), m, m + sz);
        if (w) weights.insert(weights.end(), w, w + sz);
        else weights.insert(weights.end(), sz, 0.0);
    }
    std::vector<double> sim(n * n);
    const int rc = hulk_b200_smash(mins.data(), weights.data(), (uint32_t)n, s, metric == "weightedjaccard" ? 1 : 0,
                                   (int32_t)device, sim.data());
    if (rc) fatal(std::string(hulk_b200_strerror(rc)) + (rc == HULK_B200_ECUDA ? ": no CUDA device (no CPU fallback exists)" : ""));
    const std::string matrix_path = out_file + ".hulk-matrix.csv";
    FILE *fh = fopen(matrix_path.c_str(), "wb");
    if (!fh) fatal("open " + matrix_path + ": " + strerror(errno));
    std::vector<std::string> row(n);
    for (size_t c = 0; c < n; c++) row[c] = files[order[c]];
    csv_write(fh, row);
    for (size_t r = 0; r < n; r++) {
        for (size_t c = 0; c < n; c++) {
            char buf[64];
            const double v = sim[r * n + c];
            if (v != v) snprintf(buf, sizeof buf, "NaN");                 // strconv spells the specials this way
            else if (v - v != 0) snprintf(buf, sizeof buf, v > 0 ? "+Inf" : "-Inf");
            else snprintf(buf, sizeof buf, "%.2f", v);                    // strconv.FormatFloat(v, 'f', 2, 64)
            row[c] = buf;
        }
        csv_write(fh, row);
    }
    fclose(fh);
    logf("\twritten similarity matrix to disk: %s", matrix_path.c_str());
    if (banner_matrix) {                                                  // makeBannerMatrix (cmd/smash.go:229-262)
        const std::string banner_path = out_file + ".banner-matrix.csv";  // (the reference ranges over a map: its
        fh = fopen(banner_path.c_str(), "wb");                            //  row order is random; ours is sorted)
        if (!fh) fatal("open " + banner_path + ": " + strerror(errno));
        for (size_t r = 0; r < n; r++) {
            std::vector<std::string> line;
            for (uint32_t t = 0; t < s; t++) line.push_back(std::to_string(mins[r * s + t]));
            line.push_back(hulk_b200_sketch_banner(loaded[order[r]]));
            csv_write(fh, line);
        }
        fclose(fh);
        logf("\twritten banner matrix to disk: %s", banner_path.c_str());
    }
    for (auto *sf : loaded) hulk_b200_sketch_free(sf);
    logf("finished");
    return 0;
}

void usage_root() {
    printf("\n\tHULK is a tool that creates small, fixed-size sketches from streaming microbiome sequencing data,\n"
           "\tenabling rapid metagenomic dissimilarity analysis. HULK generates an approximate k-mer spectrum from\n"
           "\ta FASTQ data stream, incrementally sketches it and makes similarity search queries against other microbiome sketches.\n\n"
           "Usage:\n  hulk [command]\n\nAvailable Commands:\n"
           "  help        Help about any command\n"
           "  sketch      Create a sketch from a set of reads\n"
           "  smash       Smash a bunch of sketches and return a distance matrix\n"
           "  version     Prints the current version and exits\n\n"
           "Use \"hulk [command] --help\" for more information about a command.\n");
}

}  // namespace

int main(int argc, char **argv) {
    if (argc < 2 || !strcmp(argv[1], "help") || !strcmp(argv[1], "-h") || !strcmp(argv[1], "--help")) {
        usage_root();
        return 0;
    }
    if (!strcmp(argv[1], "version")) {                      // cmd/version.go
        printf("%s\n", hulk_b200_version());
        return 0;
    }
    if (!strcmp(argv[1], "sketch")) return run_sketch(argc, argv);
    if (!strcmp(argv[1], "smash")) return run_smash(argc, argv);
    printf("Error: unknown command \"%s\" for \"hulk\"\nRun 'hulk --help' for usage.\n", argv[1]);
    return 1;
}
