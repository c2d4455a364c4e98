// host_fuzz.cpp -- AddressSanitizer / UBSan driver for the host-only parsers of libhulk_b200 (the JSON sketch
// reader of `hulk smash` and the FASTQ/FASTA/gzip/BGZF line reader of `hulk sketch`).  Both take files a user
// hands them, so they must reject garbage without touching memory they do not own.  Built and run by
// tools/asan/run.sh on a CPU-only box; the four device entry points the reader links against are stubbed.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/hulk_b200.h"

extern "C" {
int hulk_b200_alloc_pinned(void **p, uint64_t) { *p = nullptr; return HULK_B200_ECUDA; }   // plain memory path
void hulk_b200_free_pinned(void *) {}
int hulk_b200_push_reads(hulk_b200_ctx *, const uint8_t *, const uint64_t *, uint64_t) { return 0; }
int hulk_b200_flush(hulk_b200_ctx *) { return 0; }
int hulk_b200_sync(hulk_b200_ctx *) { return 0; }
}

static unsigned long long g_reads = 0, g_bytes = 0, g_ok = 0, g_err = 0;

static void one_json(const char *path) {
    hulk_b200_sketch_file *f = nullptr;
    char err[256] = "";
    if (hulk_b200_sketch_load(path, &f, err, sizeof err) != HULK_B200_OK) { g_err++; return; }
    g_ok++;
    for (const char *algo : {"histosketch", "kmv", "khf", "nonsense"})
        for (uint32_t k : {21u, 31u, 0u}) {
            const uint64_t *mins = nullptr;
            const double *w = nullptr;
            uint32_t s = 0;
            if (hulk_b200_sketch_find(f, k, algo, &mins, &w, &s, err, sizeof err) == HULK_B200_OK) {
                unsigned long long acc = 0;
                for (uint32_t i = 0; i < s; i++) acc += mins[i] + (w ? (w[i] != 0.0) : 0);
                g_bytes += acc & 1;
            }
        }
    (void)hulk_b200_sketch_banner(f);
    hulk_b200_sketch_free(f);
}

static void one_reads(const char *path, int fasta) {
    const unsigned long long reads_before = g_reads, err_before = g_err;
    struct Report {
        const char *path;
        unsigned long long r0, e0;
        ~Report() {
            if (getenv("HOST_FUZZ_VERBOSE")) printf("%s %llu %s\n", path, g_reads - r0, g_err != e0 ? "ERR" : "OK");
        }
    } report{path, reads_before, err_before};
    hulk_b200_reader *rd = nullptr;
    const char *paths[1] = {path};
    // HOST_FUZZ_BATCH=0: the library's default batch size, the only setting under which reader_open takes the
    // multi-threaded plain-FASTQ parser (produce_parallel); anything else keeps the small batches of the one-thread pass
    const char *hb = getenv("HOST_FUZZ_BATCH");
    const uint64_t batch = hb ? (uint64_t)atoll(hb) : (1u << 16);
    if (hulk_b200_reader_open(paths, 1, fasta, batch, &rd) != HULK_B200_OK) { g_err++; return; }
    for (;;) {
        const uint8_t *bases = nullptr;
        const uint64_t *offs = nullptr;
        uint64_t n = 0;
        const int rc = hulk_b200_reader_next(rd, &bases, &offs, &n);
        if (rc != HULK_B200_OK) { g_err++; (void)hulk_b200_reader_error(rd); break; }
        if (n == 0) { g_ok++; break; }
        unsigned long long acc = 0;
        for (uint64_t i = 0; i < n; i++)
            for (uint64_t j = offs[i]; j < offs[i + 1]; j++) acc += bases[j];     // touch every byte handed out
        g_reads += n;
        g_bytes += acc & 1;
    }
    hulk_b200_reader_close(rd);
}

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: host_fuzz json|fastq|fasta FILE...\n"); return 2; }
    const std::string mode = argv[1];
    for (int i = 2; i < argc; i++) {
        if (mode == "json") one_json(argv[i]);
        else one_reads(argv[i], mode == "fasta");
    }
    printf("%s: %d files, %llu accepted, %llu rejected, %llu reads\n", mode.c_str(), argc - 2, g_ok, g_err, g_reads);
    return 0;
}
