// ingest.cpp -- the input side of `hulk sketch`: line reader + FASTQ/FASTA framing, feeding the GPU.
//
// Reference semantics (paths relative to the reference checkout):
//   src/pipeline/sketch.go:40-79    DataStreamer.Run: every input file in order (or STDIN) through a
//                                   bufio.Scanner, gzip when the last '.'-component of the name is "gz";
//                                   `append([]byte(nil), line...)` makes an EMPTY line nil
//   src/pipeline/sketch.go:99-161   FastqHandler.Run: FASTQ = fill l1..l4 with the next non-nil lines,
//                                   FASTA = '>' starts a record, other lines are appended, an empty line
//                                   stops the input
//   src/seqio/seqio.go:37-48        NewFASTQread: only l1[0] == '@' is checked
//   src/pipeline/sketch.go:197-224  SeqMinimizer.Run: AddSeq / progress lines / Flush per interval
//
// Design: one producer thread reads (read(2) or zlib), splits lines with memchr and frames records
// straight into a ring of page-locked batches (bases + offsets), so the consumer can hand a batch to
// hulk_b200_push_reads without another copy while the producer fills the next one.  No CUDA kernels
// here; the only CUDA dependency is the pinned allocation (plain memory when no device exists, so the
// reader itself can be tested on a CPU-only box -- it computes nothing).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/hulk_b200.h"
#include "pgzip.h"

namespace {

constexpr size_t kBlock = 4u << 20;             // bytes per read()/gzread()
constexpr size_t kMaxToken = 64 * 1024;         // bufio.MaxScanTokenSize
constexpr int kRing = 4;                        // batches in flight
constexpr size_t kParChunk = 8u << 20;          // parallel FASTQ parse: bytes of file per worker task
constexpr uint64_t kParMinBytes = 256ull << 20; // ... used for plain files from this total size on

// Worker threads the reader may use next to its framing thread: all but two cores, at most `cap`;
// HULK_B200_READER_THREADS overrides (1 = everything on the framing thread's side stays single-threaded).
unsigned reader_threads(unsigned cap) {
    const unsigned hw = std::thread::hardware_concurrency();
    unsigned n = std::max(1u, std::min(cap, hw > 2 ? hw - 2 : 1u));
    if (const char *e = getenv("HULK_B200_READER_THREADS")) n = (unsigned)std::max(1, std::min(64, atoi(e)));
    return n;
}

struct HostBuf {                                // page-locked when a device exists
    void *p = nullptr;
    bool pinned = false;
    size_t bytes = 0;
    bool alloc(size_t n) {
        release();
        void *q = nullptr;
        if (hulk_b200_alloc_pinned(&q, n) == HULK_B200_OK) {
            p = q; pinned = true; bytes = n;
            return true;
        }
        if (posix_memalign(&q, 4096, n ? n : 1) != 0) return false;
        p = q; pinned = false; bytes = n;
        return true;
    }
    void release() {
        if (!p) return;
        if (pinned) hulk_b200_free_pinned(p);
        else free(p);
        p = nullptr; bytes = 0;
    }
};

struct Batch {
    HostBuf bases, offsets;
    uint64_t n_reads = 0, n_bytes = 0;          // committed
    uint64_t cap_reads() const { return offsets.bytes / 8 - 1; }
    uint8_t *b() { return static_cast<uint8_t *>(bases.p); }
    uint64_t *o() { return static_cast<uint64_t *>(offsets.p); }
};

}  // namespace

struct hulk_b200_reader {
    std::vector<std::string> paths;
    bool fasta = false;
    uint64_t batch_bytes = 0;

    Batch ring[kRing];
    std::vector<std::unique_ptr<Batch>> pring;  // parallel mode: 2 x workers batches of kParChunk bytes
    unsigned par_workers = 0;                   // > 0: parallel FASTQ parse of plain files (see produce_parallel)
    size_t par_chunk = kParChunk;               // bytes of file per worker task (HULK_B200_PARALLEL_CHUNK: tests)
    std::deque<Batch *> free_q, full_q;
    Batch *held = nullptr;                      // batch currently lent to the consumer
    std::mutex mu;
    std::condition_variable cv_free, cv_full;
    bool done = false, stop = false;
    int err = 0;                                // producer's terminal error (delivered after the full batches)
    std::string err_text, last_error;
    std::thread th;

    // producer state
    Batch *cur = nullptr;
    // FASTQ framing: which of l1..l4 are filled (the sequence itself is parked in the batch, uncommitted)
    int slot = 0;
    uint8_t l1_first = 0;
    std::string l1_text;
    uint64_t pend_bytes = 0;
    // FASTA framing
    bool have_header = false, fasta_stop = false;
    std::vector<uint8_t> fa_seq;

    bool fail(int code, const std::string &text) {
        err = code;
        err_text = text;
        return false;
    }

    Batch *take_free() {
        std::unique_lock<std::mutex> lk(mu);
        cv_free.wait(lk, [&] { return stop || !free_q.empty(); });
        if (stop) return nullptr;
        Batch *b = free_q.front();
        free_q.pop_front();
        b->n_reads = b->n_bytes = 0;
        b->o()[0] = 0;
        return b;
    }
    void publish(Batch *b) {
        std::lock_guard<std::mutex> lk(mu);
        full_q.push_back(b);
        cv_full.notify_one();
    }
    // make room for `len` more bases and one more read in the current batch; false = stopped / OOM
    bool reserve(uint64_t len) {
        if (!cur && !(cur = take_free())) return false;
        const bool full = cur->n_bytes + len > cur->bases.bytes || cur->n_reads + 1 > cur->cap_reads();
        if (full && cur->n_reads > 0) {
            publish(cur);
            if (!(cur = take_free())) return false;
        }
        if (len > cur->bases.bytes) {                       // one sequence larger than a whole batch
            if (!cur->bases.alloc((size_t)((len + 4095) & ~4095ull))) return fail(HULK_B200_ENOMEM, "batch buffer");
        }
        return true;
    }
    bool commit(uint64_t len) {
        cur->n_bytes += len;
        cur->n_reads += 1;
        cur->o()[cur->n_reads] = cur->n_bytes;
        return true;
    }

    // one bufio.Scanner line (CR already dropped); an empty line is Go's nil slice
    bool on_line(const uint8_t *p, size_t len) {
        if (fasta) {
            if (fasta_stop) return true;
            if (len == 0) { fasta_stop = true; return true; }            // sketch.go:103-105
            if (p[0] == '>') {                                           // :107
                if (have_header) {                                       // :108-119
                    if (!reserve(fa_seq.size())) return false;
                    if (!fa_seq.empty()) memcpy(cur->b() + cur->n_bytes, fa_seq.data(), fa_seq.size());
                    commit(fa_seq.size());
                }
                have_header = true;                                      // l1, l2 = line, nil   :120
                fa_seq.clear();
            } else {
                fa_seq.insert(fa_seq.end(), p, p + len);                 // l2 = append(l2, line...)  :122
            }
            return true;
        }
        if (len == 0) return true;                                       // nil line: no slot takes it  :140-147
        switch (slot) {
            case 0:
                l1_first = p[0];
                if (l1_first != '@') l1_text.assign(reinterpret_cast<const char *>(p), len);
                slot = 1;
                break;
            case 1:
                if (!reserve(len)) return false;
                memcpy(cur->b() + cur->n_bytes, p, len);                 // parked behind the committed reads
                pend_bytes = len;
                slot = 2;
                break;
            case 2:
                slot = 3;
                break;
            default:
                if (l1_first != '@')                                     // seqio.go:38-40 (checked when l4 arrives)
                    return fail(HULK_B200_EFASTQ, "read ID in fastq file does not begin with @: " + l1_text);
                commit(pend_bytes);
                slot = 0;
                break;
        }
        return true;
    }

    // split one file's byte stream into lines (bufio.ScanLines): `carry` holds an unfinished line
    bool feed(const uint8_t *data, size_t n, std::vector<uint8_t> &carry) {
        size_t pos = 0;
        while (pos < n && !fasta_stop) {
            const uint8_t *nl = static_cast<const uint8_t *>(memchr(data + pos, '\n', n - pos));
            if (!nl) {
                carry.insert(carry.end(), data + pos, data + n);
                if (carry.size() >= kMaxToken) return fail(HULK_B200_ETOOLONG, "bufio.Scanner: token too long");
                return true;
            }
            const uint8_t *lp = data + pos;
            size_t len = (size_t)(nl - lp);
            if (!carry.empty()) {
                carry.insert(carry.end(), lp, nl);
                lp = carry.data();
                len = carry.size();
            }
            if (len >= kMaxToken) return fail(HULK_B200_ETOOLONG, "bufio.Scanner: token too long");
            if (len && lp[len - 1] == '\r') len--;
            if (!on_line(lp, len)) return false;
            carry.clear();
            pos = (size_t)(nl - data) + 1;
        }
        return true;
    }
    bool end_of_file(std::vector<uint8_t> &carry) {
        if (!carry.empty() && !fasta_stop) {                 // final line without a newline is still a token
            size_t len = carry.size();
            if (carry[len - 1] == '\r') len--;
            if (!on_line(carry.data(), len)) return false;
        }
        carry.clear();
        return true;
    }

    bool read_plain(int fd, const std::string &name) {
        std::vector<uint8_t> block(kBlock), carry;
        for (;;) {
            if (stop) return false;
            const ssize_t got = ::read(fd, block.data(), block.size());
            if (got < 0) {
                if (errno == EINTR) continue;
                return fail(HULK_B200_EIO, "read " + name + ": " + strerror(errno));
            }
            if (got == 0) break;
            if (!feed(block.data(), (size_t)got, carry)) return false;
            if (fasta_stop) return true;
        }
        return end_of_file(carry);
    }
    // ---- BGZF (bgzip / htslib blocked gzip) -------------------------------------------------------------
    // A BGZF file is a series of gzip members of at most 64 KiB whose header carries the member's size
    // (extra subfield 'B','C'), so the members can be found without inflating and inflated independently.
    // Go's gzip.Reader concatenates members (multistream), so the byte stream -- and everything after it --
    // is the same as through the one-thread zlib path; only the inflating is spread over the cores.
    static bool bgzf_block(const uint8_t *d, size_t size, size_t off, size_t *bsize, size_t *hdr) {
        if (off + 18 > size) return false;
        const uint8_t *h = d + off;
        if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return false;
        const size_t xlen = h[10] | ((size_t)h[11] << 8);
        if (off + 12 + xlen > size) return false;
        for (size_t x = 12; x + 4 <= 12 + xlen;) {
            const size_t slen = h[x + 2] | ((size_t)h[x + 3] << 8);
            if (h[x] == 'B' && h[x + 1] == 'C' && slen == 2 && x + 6 <= 12 + xlen) {
                *bsize = (size_t)(h[x + 4] | ((size_t)h[x + 5] << 8)) + 1;
                *hdr = 12 + xlen;
                // no name/comment/crc fields in BGZF headers (FLG == 4)
                return h[3] == 4 && *bsize >= *hdr + 8 && off + *bsize <= size;
            }
            x += 4 + slen;
        }
        return false;
    }
    // returns 1: handled (ok), 0: handled (failed, error set), -1: not a BGZF file, 2: a plain gzip member
    // follows at *resume_off (the caller goes on with zlib from there, `carry` holds the unfinished line)
    int read_bgzf(int fd, const std::string &name, std::vector<uint8_t> &carry, size_t *resume_off) {
        struct stat st;
        if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode) || st.st_size < 28) return -1;
        const size_t size = (size_t)st.st_size;
        void *map = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (map == MAP_FAILED) return -1;
        const uint8_t *d = static_cast<const uint8_t *>(map);
        size_t bs = 0, hd = 0;
        if (!bgzf_block(d, size, 0, &bs, &hd)) { munmap(map, size); return -1; }
        madvise(map, size, MADV_SEQUENTIAL);
        // inflating is the slow stage (~0.12 GB/s of FASTQ per core against ~3.8 GB/s for the framing thread):
        // every core but the framing thread and the caller's
        const unsigned W = reader_threads(16);
        struct Blk { size_t off, bsize, hdr, out, isize; };
        // Two windows of members (~64 MiB of output each): while the framing below consumes one, the next is
        // inflated on W threads.  A window ends early at a member that is not BGZF: `tail` says why.
        enum Tail { MORE, END, RESUME, BADHDR };
        struct Win {
            std::vector<Blk> blks;
            std::vector<uint8_t> data;
            size_t total = 0;
            Tail tail = END;
            std::atomic<bool> bad{false};
            std::mutex bad_mu;                              // first bad member of the window and what is wrong with it
            size_t bad_idx = (size_t)-1;
            std::string bad_msg;
        } win[2];
        size_t off = 0, win_bytes = 64u << 20;
        if (const char *e = getenv("HULK_B200_BGZF_WINDOW")) win_bytes = std::max<size_t>(1, strtoull(e, nullptr, 10));   // tests
        auto plan = [&](Win &w) {                           // headers only: nothing is inflated here
            size_t b = 0, h = 0;
            w.blks.clear();
            w.total = 0;
            w.bad = false;
            w.bad_idx = (size_t)-1;
            w.bad_msg.clear();
            w.tail = END;
            while (off < size) {
                if (w.total >= win_bytes) { w.tail = MORE; break; }
                if (!bgzf_block(d, size, off, &b, &h)) {
                    w.tail = (off + 2 <= size && d[off] == 0x1f && d[off + 1] == 0x8b) ? RESUME : BADHDR;   // an ordinary member?
                    break;
                }
                const uint8_t *t = d + off + b - 4;
                const size_t isize = t[0] | ((size_t)t[1] << 8) | ((size_t)t[2] << 16) | ((size_t)t[3] << 24);
                if (isize > 65536) {                        // BGZF caps a member at 64 KiB of data: whatever this is, it is
                    w.tail = RESUME;                        // not a BGZF block -- let the ordinary gzip reader judge it
                    break;                                  // (and never size a buffer from an unchecked trailer)
                }
                w.blks.push_back({off, b, h, w.total, isize});
                w.total += isize;
                off += b;
            }
        };
        auto inflate_window = [&](Win &w) {
            w.data.resize(std::max<size_t>(w.total, 1));    // never a NULL next_out (a file that is only the EOF marker)
            std::atomic<size_t> next(0);
            auto work = [&] {
                // the member payloads are raw deflate streams with an empty window: pgzip.h's byte-wide decoder
                // (each into a scratch buffer first -- the decoder may write a few bytes past a match)
                pgz::Buf<uint8_t> scratch;
                std::string msg;
                for (size_t i; (i = next.fetch_add(1)) < w.blks.size();) {
                    const Blk &b = w.blks[i];
                    uint64_t got = 0;
                    size_t used = 0;
                    const pgz::Status ist = pgz::inflate_raw(d + b.off + b.hdr, b.bsize - b.hdr - 8, scratch, got, &used, msg);
                    const uint8_t *t = d + b.off + b.bsize - 8;
                    const uint32_t crc = t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
                    if (ist != pgz::ST_OK || got != b.isize || used != b.bsize - b.hdr - 8 ||
                        (uint32_t)crc32(0L, scratch.p, (uInt)got) != crc) {
                        // what compress/gzip would have said at this member: the inflater's own error first, then the
                        // trailer checks (gzip.ErrChecksum covers both the CRC and the length)
                        std::string why = ist == pgz::ST_EOF ? "unexpected EOF"
                                          : ist != pgz::ST_OK ? "flate: corrupt input (" + msg + ")"
                                                              : "gzip: invalid checksum";
                        std::lock_guard<std::mutex> lk(w.bad_mu);
                        if (i < w.bad_idx) {
                            w.bad_idx = i;
                            w.bad_msg = why;
                        }
                        w.bad = true;
                        continue;
                    }
                    memcpy(w.data.data() + b.out, scratch.p, (size_t)got);
                }
            };
            std::vector<std::thread> pool;
            for (unsigned t = 1; t < W; t++) pool.emplace_back(work);
            work();
            for (auto &x : pool) x.join();
        };
        bool ok = true, resume = false;
        int wi = 0;
        plan(win[0]);
        inflate_window(win[0]);
        for (;;) {
            Win &w = win[wi];
            std::thread ahead;
            if (w.tail == MORE) {
                plan(win[wi ^ 1]);
                ahead = std::thread([&, wi] { inflate_window(win[wi ^ 1]); });
            }
            if (stop) ok = false;
            else if (w.bad) {
                // the members in front of the first bad one are good data: they are delivered (and may raise a framing
                // error of their own) before the stream fails where a streaming reader would have failed
                const size_t good = w.blks[w.bad_idx].out;
                ok = good ? feed(w.data.data(), good, carry) : true;
                if (ok) ok = fail(HULK_B200_EIO, w.bad_msg);
            }
            else if (w.total) ok = feed(w.data.data(), w.total, carry);
            if (ahead.joinable()) ahead.join();
            if (!ok || fasta_stop) break;
            if (w.tail == BADHDR) { ok = fail(HULK_B200_EIO, "gzip: invalid header"); break; }
            if (w.tail == RESUME) { resume = true; break; }
            if (w.tail == END) break;
            wi ^= 1;
        }
        munmap(map, size);
        (void)name;
        if (!ok) return 0;
        if (resume && !fasta_stop) { *resume_off = off; return 2; }
        return (fasta_stop ? true : end_of_file(carry)) ? 1 : 0;
    }

    bool read_gz(int fd, const std::string &name) {
        std::vector<uint8_t> carry;
        size_t resume_off = 0;
        bool resumed = false;
        const char *pe = getenv("HULK_B200_PARALLEL_READER");
        const bool parallel_ok = !(pe && *pe == '0');
        if (parallel_ok) {
            const int rc = read_bgzf(fd, name, carry, &resume_off);
            if (rc == 0 || rc == 1) return rc == 1;
            if (rc == 2) {
                resumed = true;
                if (::lseek(fd, (off_t)resume_off, SEEK_SET) < 0) return fail(HULK_B200_EIO, "seek " + name);
            }
        }
        if (parallel_ok) {
            const int rc = read_gz_parallel(fd, carry, resume_off, !resumed);
            if (rc >= 0) return rc == 1;
        }
        return inflate_members(fd, name, carry);
    }

    // An ordinary (single-stream) gzip file of some size: inflated on several threads by pgzip.h -- block starts are
    // guessed per chunk, every chunk is decoded against an unknown window, and a sequential pass stitches the chunks
    // together.  Same bytes and the same errors as inflate_members.  1: ok, 0: failed, -1: not applicable.
    int read_gz_parallel(int fd, std::vector<uint8_t> &carry, size_t start, bool first) {
        struct stat st;
        if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) return -1;
        size_t min_bytes = 16u << 20;
        if (const char *e = getenv("HULK_B200_PGZ_MIN")) min_bytes = strtoull(e, nullptr, 10);
        const size_t size = (size_t)st.st_size;
        if (size < min_bytes || size <= start) return -1;
        void *map = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (map == MAP_FAILED) return -1;
        madvise(map, size, MADV_SEQUENTIAL);
        pgz::Options opt;
        opt.threads = reader_threads(16);
        if (opt.threads < 2 && !getenv("HULK_B200_PGZ_THREADS")) {           // marker decoding costs ~1.5x zlib's work:
            munmap(map, size);                                               // not worth it on one thread
            return -1;
        }
        opt.chunk_bytes = 1u << 20;
        if (const char *e = getenv("HULK_B200_PGZ_CHUNK")) opt.chunk_bytes = std::max<size_t>(64, strtoull(e, nullptr, 10));
        if (const char *e = getenv("HULK_B200_PGZ_THREADS")) opt.threads = (unsigned)std::max(1, atoi(e));
        bool stopped = false, sink_failed = false;
        const std::string text = pgz::inflate_parallel(
            static_cast<const uint8_t *>(map), size, start, first, opt, nullptr,
            [&](const uint8_t *p, size_t n) {
                if (stop) return false;
                if (!feed(p, n, carry)) { sink_failed = true; return false; }
                return !fasta_stop;
            },
            &stopped);
        munmap(map, size);
        if (sink_failed) return 0;
        if (!text.empty()) return fail(HULK_B200_EIO, text) ? 1 : 0;
        if (stopped) return fasta_stop ? 1 : 0;
        return end_of_file(carry) ? 1 : 0;
    }

    // compress/gzip.Reader in its default multistream mode: members are concatenated; after a member the next
    // header is read -- a clean EOF there ends the stream, fewer than 10 bytes is io.ErrUnexpectedEOF, anything
    // that is not a gzip header is gzip.ErrHeader (trailing garbage is an ERROR in Go; zlib's gzread would
    // silently stop), a CRC or length mismatch is gzip.ErrChecksum, a member cut short is "unexpected EOF".
    // The file position of `fd` is where the next member starts.
    bool inflate_members(int fd, const std::string &name, std::vector<uint8_t> &carry) {
        std::vector<uint8_t> in(1u << 20), block(kBlock);
        size_t in_pos = 0, in_len = 0;
        bool eof = false;
        auto refill = [&](size_t want) -> bool {            // keep unread input, top up to at least `want` bytes or EOF
            if (in_pos && in_pos < in_len) memmove(in.data(), in.data() + in_pos, in_len - in_pos);
            in_len -= in_pos;
            in_pos = 0;
            while (!eof && in_len < want) {
                const ssize_t got = ::read(fd, in.data() + in_len, in.size() - in_len);
                if (got < 0) {
                    if (errno == EINTR) continue;
                    return fail(HULK_B200_EIO, "read " + name + ": " + strerror(errno));
                }
                if (got == 0) eof = true;
                in_len += (size_t)got;
            }
            return true;
        };
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        if (inflateInit2(&zs, 15 + 16) != Z_OK) return fail(HULK_B200_ENOMEM, "inflateInit2");
        bool ok = true, first = true;
        while (ok && !fasta_stop) {
            // ---- member header (gzip.Reader.readHeader) ----
            if (!(ok = refill(10))) break;
            const size_t avail = in_len - in_pos;
            if (avail == 0 && !first) break;                                            // io.EOF between members: the end
            if (avail < 10) { ok = fail(HULK_B200_EIO, avail == 0 ? "EOF" : "unexpected EOF"); break; }
            const uint8_t *h = in.data() + in_pos;
            if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8) { ok = fail(HULK_B200_EIO, "gzip: invalid header"); break; }
            first = false;
            inflateReset(&zs);
            // ---- member body ----
            for (bool member_done = false; ok && !member_done && !fasta_stop;) {
                if (stop) { ok = false; break; }
                if (in_pos == in_len) {
                    if (!(ok = refill(1))) break;
                    if (in_len == 0) { ok = fail(HULK_B200_EIO, "unexpected EOF"); break; }
                }
                zs.next_in = in.data() + in_pos;
                zs.avail_in = (uInt)(in_len - in_pos);
                zs.next_out = block.data();
                zs.avail_out = (uInt)block.size();
                const int rc = inflate(&zs, Z_NO_FLUSH);
                in_pos = in_len - zs.avail_in;
                const size_t got = block.size() - zs.avail_out;
                if (rc == Z_STREAM_END) member_done = true;
                else if (rc == Z_DATA_ERROR || rc == Z_NEED_DICT) {
                    const std::string m = zs.msg ? zs.msg : "";
                    // data before the bad spot was handed on by Go's reader too
                    if (got && !feed(block.data(), got, carry)) { ok = false; break; }
                    ok = fail(HULK_B200_EIO, (m == "incorrect data check" || m == "incorrect length check") ? "gzip: invalid checksum"
                                             : (m == "incorrect header check" || m == "unknown compression method" ||
                                                m == "unknown header flags set" || m == "header crc mismatch") ? "gzip: invalid header"
                                             : "flate: corrupt input (" + m + ")");
                    break;
                } else if (rc != Z_OK && rc != Z_BUF_ERROR) {
                    ok = fail(rc == Z_MEM_ERROR ? HULK_B200_ENOMEM : HULK_B200_EIO, "inflate failed");
                    break;
                }
                if (got && !(ok = feed(block.data(), got, carry))) break;
            }
        }
        inflateEnd(&zs);
        if (!ok) return false;
        return fasta_stop ? true : end_of_file(carry);
    }

    // ---- parallel parse of plain FASTQ files ---------------------------------------------------------
    // FastqHandler groups the NON-EMPTY lines of the stream in fours and never resynchronises
    // (sketch.go:139-159), so the slot a line fills is (number of non-empty lines before it) mod 4.  That
    // makes the framing splittable without guessing: pass 1 counts non-empty lines per chunk in parallel, a
    // prefix sum gives every chunk its starting slot, pass 2 frames the chunks in parallel (a chunk that
    // starts inside a record looks back for the lines it needs), and the batches are delivered in file
    // order.  The result is byte-identical to the one-thread reader, errors included.
    struct ParJob {
        const uint8_t *d = nullptr;      // the mapped file
        size_t size = 0, begin = 0, end = 0;   // [begin, end): whole lines, the last may lack its newline
        uint64_t nonempty = 0;
        int phase = 0;
        Batch *out = nullptr;
        int err = 0;
        std::string err_text;
    };
    std::vector<std::string> carry_lines;       // pending lines (l1.. of the open record) at the end of the previous file

    template <class F>
    static void par_lines(const ParJob &j, F f) {      // f(ptr, len) for every line of the job, CR dropped; false stops
        size_t pos = j.begin;
        while (pos < j.end) {
            const uint8_t *nl = static_cast<const uint8_t *>(memchr(j.d + pos, '\n', j.end - pos));
            const size_t stop = nl ? (size_t)(nl - j.d) : j.end;
            size_t len = stop - pos;
            const bool too_long = len >= kMaxToken;
            if (len && j.d[pos + len - 1] == '\r') len--;
            if (!f(j.d + pos, len, too_long)) return;
            pos = stop + 1;
        }
    }
    static void par_count(ParJob &j) {
        uint64_t n = 0;
        par_lines(j, [&](const uint8_t *, size_t len, bool) { n += len != 0; return true; });
        j.nonempty = n;
    }
    void par_parse(ParJob &j) const {
        Batch &b = *j.out;
        b.n_reads = b.n_bytes = 0;
        b.o()[0] = 0;
        int ps = j.phase;
        uint8_t first_byte = '@';
        std::string id_text;
        const uint8_t *seq = nullptr;
        size_t seq_len = 0;
        std::string seq_copy;
        if (ps) {                                           // the record in progress: find its lines so far
            std::vector<std::pair<const uint8_t *, size_t>> prev;          // newest first
            size_t pos = j.begin;
            while ((int)prev.size() < ps && pos > 0) {
                const size_t line_end = pos - 1;                          // the newline that ends the previous line
                const void *q = line_end ? memrchr(j.d, '\n', line_end) : nullptr;
                const size_t start = q ? (size_t)(static_cast<const uint8_t *>(q) - j.d) + 1 : 0;
                size_t len = line_end - start;
                if (len && j.d[start + len - 1] == '\r') len--;
                if (len) prev.emplace_back(j.d + start, len);
                pos = start;
            }
            // lines still missing were the tail of the previous file
            const int missing = ps - (int)prev.size();
            auto line_at = [&](int idx, const uint8_t **p, size_t *len) {    // idx 0 = l1
                if (idx < missing) {
                    const std::string &c = carry_lines[carry_lines.size() - missing + idx];
                    *p = reinterpret_cast<const uint8_t *>(c.data());
                    *len = c.size();
                } else {
                    const auto &pr = prev[prev.size() - 1 - (idx - missing)];
                    *p = pr.first;
                    *len = pr.second;
                }
            };
            // A looked-back line of a token's length or more was fatal for the job (or file) it lies in; that job
            // reports it, and the jobs of a round run side by side, so this one must not touch the line either:
            // its batch holds par_chunk + 2 kMaxToken bytes and nothing longer than a token may be copied into it.
            for (int idx = 0; idx < ps; idx++) {
                const uint8_t *pl;
                size_t nl;
                line_at(idx, &pl, &nl);
                if (nl >= kMaxToken) {
                    j.err = HULK_B200_ETOOLONG;
                    j.err_text = "bufio.Scanner: token too long";
                    return;
                }
            }
            const uint8_t *p1;
            size_t n1;
            line_at(0, &p1, &n1);
            first_byte = p1[0];
            if (first_byte != '@') id_text.assign(reinterpret_cast<const char *>(p1), n1);
            if (ps >= 2) line_at(1, &seq, &seq_len);
        }
        par_lines(j, [&](const uint8_t *p, size_t len, bool too_long) {
            if (too_long) {
                j.err = HULK_B200_ETOOLONG;
                j.err_text = "bufio.Scanner: token too long";
                return false;
            }
            if (len == 0) return true;
            switch (ps) {
                case 0:
                    first_byte = p[0];
                    if (first_byte != '@') id_text.assign(reinterpret_cast<const char *>(p), len);
                    ps = 1;
                    break;
                case 1: seq = p; seq_len = len; ps = 2; break;
                case 2: ps = 3; break;
                default:
                    if (first_byte != '@') {
                        j.err = HULK_B200_EFASTQ;
                        j.err_text = "read ID in fastq file does not begin with @: " + id_text;
                        return false;
                    }
                    if (b.n_bytes + seq_len > b.bases.bytes) {             // cannot happen for lines below a token's length
                        j.err = HULK_B200_ETOOLONG;
                        j.err_text = "bufio.Scanner: token too long";
                        return false;
                    }
                    memcpy(b.b() + b.n_bytes, seq, seq_len);
                    b.n_bytes += seq_len;
                    b.n_reads += 1;
                    b.o()[b.n_reads] = b.n_bytes;
                    ps = 0;
                    break;
            }
            return true;
        });
    }

    bool produce_parallel() {
        uint64_t nonempty_total = 0;                        // non-empty lines of the stream so far
        const unsigned W = par_workers;
        for (size_t fi = 0; fi < paths.size(); fi++) {
            const std::string &name = paths[fi];
            const int fd = ::open(name.c_str(), O_RDONLY);
            if (fd < 0) return fail(HULK_B200_EIO, "open " + name + ": " + strerror(errno));
            struct stat st;
            if (fstat(fd, &st) != 0) { ::close(fd); return fail(HULK_B200_EIO, "stat " + name); }
            const size_t size = (size_t)st.st_size;
            if (size == 0) { ::close(fd); continue; }
            void *map = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
            ::close(fd);
            if (map == MAP_FAILED) return fail(HULK_B200_EIO, "mmap " + name + ": " + strerror(errno));
            madvise(map, size, MADV_SEQUENTIAL);
            const uint8_t *d = static_cast<const uint8_t *>(map);
            auto line_start_at_or_after = [&](size_t off) -> size_t {      // first line start >= off
                if (off == 0) return 0;
                if (off >= size) return size;
                const void *q = memchr(d + off - 1, '\n', size - (off - 1));
                return q ? (size_t)(static_cast<const uint8_t *>(q) - d) + 1 : size;
            };
            const size_t n_chunks = (size + par_chunk - 1) / par_chunk;
            bool ok = true;
            for (size_t c0 = 0; ok && c0 < n_chunks; c0 += W) {
                const unsigned nj = (unsigned)std::min<size_t>(W, n_chunks - c0);
                std::vector<ParJob> jobs(nj);
                for (unsigned t = 0; t < nj; t++) {
                    jobs[t].d = d;
                    jobs[t].size = size;
                    jobs[t].begin = line_start_at_or_after((c0 + t) * par_chunk);
                    jobs[t].end = line_start_at_or_after((c0 + t + 1) * par_chunk);
                }
                {
                    std::vector<std::thread> pool;
                    for (unsigned t = 0; t < nj; t++) pool.emplace_back([&, t] { par_count(jobs[t]); });
                    for (auto &x : pool) x.join();
                }
                for (unsigned t = 0; t < nj; t++) {
                    jobs[t].phase = (int)(nonempty_total & 3);
                    nonempty_total += jobs[t].nonempty;
                    if (!(jobs[t].out = take_free())) { munmap(map, size); return false; }     // stopped
                }
                {
                    std::vector<std::thread> pool;
                    for (unsigned t = 0; t < nj; t++) pool.emplace_back([&, t] { par_parse(jobs[t]); });
                    for (auto &x : pool) x.join();
                }
                for (unsigned t = 0; t < nj; t++) {
                    if (!ok) {                               // behind an error: hand the batch back unused
                        std::lock_guard<std::mutex> lk(mu);
                        free_q.push_back(jobs[t].out);
                        continue;
                    }
                    if (jobs[t].out->n_reads > 0) publish(jobs[t].out);
                    else { std::lock_guard<std::mutex> lk(mu); free_q.push_back(jobs[t].out); }
                    if (jobs[t].err) ok = fail(jobs[t].err, jobs[t].err_text);
                }
            }
            if (ok) {                                        // lines of a record left open at the end of this file
                const int open_lines = (int)(nonempty_total & 3);
                std::vector<std::string> tail;               // newest first
                size_t pos = size;
                if (pos > 0 && d[pos - 1] != '\n') {         // final line without a newline
                    const void *q = memrchr(d, '\n', pos);
                    const size_t start = q ? (size_t)(static_cast<const uint8_t *>(q) - d) + 1 : 0;
                    size_t len = pos - start;
                    if (len && d[start + len - 1] == '\r') len--;
                    if (len && (int)tail.size() < open_lines) tail.emplace_back(reinterpret_cast<const char *>(d + start), len);
                    pos = start;
                }
                while ((int)tail.size() < open_lines && pos > 0) {
                    const size_t line_end = pos - 1;
                    const void *q = line_end ? memrchr(d, '\n', line_end) : nullptr;
                    const size_t start = q ? (size_t)(static_cast<const uint8_t *>(q) - d) + 1 : 0;
                    size_t len = line_end - start;
                    if (len && d[start + len - 1] == '\r') len--;
                    if (len) tail.emplace_back(reinterpret_cast<const char *>(d + start), len);
                    pos = start;
                }
                std::vector<std::string> next_carry;
                const int missing = open_lines - (int)tail.size();          // still older: from the previous carry
                for (int i = 0; i < missing; i++) next_carry.push_back(carry_lines[carry_lines.size() - missing + i]);
                for (size_t i = tail.size(); i-- > 0;) next_carry.push_back(tail[i]);
                carry_lines.swap(next_carry);
            }
            munmap(map, size);
            if (!ok) return false;
        }
        return true;
    }

    void produce() {
        bool ok = true;
        if (par_workers) {
            ok = produce_parallel();
            std::lock_guard<std::mutex> lk(mu);
            done = true;
            cv_full.notify_all();
            return;
        }
        if (paths.empty()) {
            ok = read_plain(0, "STDIN");
        } else {
            for (size_t i = 0; ok && i < paths.size() && !fasta_stop; i++) {
                const std::string &name = paths[i];
                const int fd = ::open(name.c_str(), O_RDONLY);
                if (fd < 0) { ok = fail(HULK_B200_EIO, "open " + name + ": " + strerror(errno)); break; }
#ifdef POSIX_FADV_SEQUENTIAL
                posix_fadvise(fd, 0, 0, POSIX_FADV_SEQUENTIAL);
#endif
                const size_t dot = name.rfind('.');
                const bool gz = dot != std::string::npos && name.compare(dot + 1, std::string::npos, "gz") == 0;
                ok = gz ? read_gz(fd, name) : read_plain(fd, name);
                ::close(fd);
            }
        }
        if (ok && fasta) {                                               // "flush final fasta"  sketch.go:126-135
            if (!have_header) {
                ok = fail(HULK_B200_EFASTQ, "no FASTA record in the input (the reference indexes a nil header here)");
            } else if (reserve(fa_seq.size())) {
                if (!fa_seq.empty()) memcpy(cur->b() + cur->n_bytes, fa_seq.data(), fa_seq.size());
                commit(fa_seq.size());
            } else {
                ok = false;
            }
        }
        // (an unfinished FASTQ record at the end of the input is never emitted by the reference)
        if (cur && cur->n_reads > 0) { publish(cur); cur = nullptr; }
        std::lock_guard<std::mutex> lk(mu);
        done = true;
        cv_full.notify_all();
    }
};

extern "C" {

int hulk_b200_reader_open(const char *const *paths, uint32_t n_paths, int fasta, uint64_t batch_bytes,
                          hulk_b200_reader **out) {
    if (!out || (n_paths && !paths)) return HULK_B200_EARG;
    *out = nullptr;
    hulk_b200_reader *rd = new (std::nothrow) hulk_b200_reader();
    if (!rd) return HULK_B200_ENOMEM;
    for (uint32_t i = 0; i < n_paths; i++) {
        if (!paths[i]) { delete rd; return HULK_B200_EARG; }
        rd->paths.emplace_back(paths[i]);
    }
    rd->fasta = fasta != 0;
    rd->batch_bytes = batch_bytes ? batch_bytes : (32ull << 20);
    if (rd->batch_bytes < 4096) rd->batch_bytes = 4096;
    // plain FASTQ files of some size: parse them on several threads (produce_parallel); everything else
    // (gzip, STDIN, FASTA, small inputs) goes through the one-thread line reader
    if (!rd->fasta && n_paths > 0 && batch_bytes == 0) {
        uint64_t total = 0;
        bool eligible = true;
        for (const std::string &nm : rd->paths) {
            struct stat st;
            const size_t dot = nm.rfind('.');
            const bool gz = dot != std::string::npos && nm.compare(dot + 1, std::string::npos, "gz") == 0;
            if (gz || stat(nm.c_str(), &st) != 0 || !S_ISREG(st.st_mode)) { eligible = false; break; }
            total += (uint64_t)st.st_size;
        }
        const char *force = getenv("HULK_B200_PARALLEL_READER");           // "1": also for small inputs, "0": never
        const unsigned hw = std::thread::hardware_concurrency();
        if (eligible && hw >= 4 && !(force && *force == '0') && (total >= kParMinBytes || (force && *force == '1')))
            rd->par_workers = getenv("HULK_B200_READER_THREADS") ? reader_threads(8) : std::min(8u, hw / 2);
        const char *pc = getenv("HULK_B200_PARALLEL_CHUNK");
        if (pc && atoll(pc) >= 16) rd->par_chunk = (size_t)atoll(pc);
    }
    if (rd->par_workers) {
        for (unsigned i = 0; i < 2 * rd->par_workers; i++) {
            std::unique_ptr<Batch> b(new Batch());
            // a chunk's sequences plus one record carried in from the chunk before; records of >= 8 bytes
            // (a chunk ends at the first line start at or after its nominal end, so one more line may ride along)
            if (!b->bases.alloc(rd->par_chunk + 2 * kMaxToken) || !b->offsets.alloc((rd->par_chunk / 8 + 8) * 8)) {
                delete rd;
                return HULK_B200_ENOMEM;
            }
            rd->free_q.push_back(b.get());
            rd->pring.push_back(std::move(b));
        }
        rd->th = std::thread([rd] { rd->produce(); });
        *out = rd;
        return HULK_B200_OK;
    }
    // offsets: room for reads as short as 32 bases on average (shorter reads just close the batch earlier)
    const uint64_t cap_reads = std::max<uint64_t>(1024, rd->batch_bytes / 32);
    for (int i = 0; i < kRing; i++) {
        if (!rd->ring[i].bases.alloc((size_t)rd->batch_bytes) || !rd->ring[i].offsets.alloc((size_t)(cap_reads + 1) * 8)) {
            for (int j = 0; j <= i; j++) { rd->ring[j].bases.release(); rd->ring[j].offsets.release(); }
            delete rd;
            return HULK_B200_ENOMEM;
        }
        rd->free_q.push_back(&rd->ring[i]);
    }
    rd->th = std::thread([rd] { rd->produce(); });
    *out = rd;
    return HULK_B200_OK;
}

int hulk_b200_reader_next(hulk_b200_reader *rd, const uint8_t **bases, const uint64_t **offsets, uint64_t *n_reads) {
    if (!rd || !bases || !offsets || !n_reads) return HULK_B200_EARG;
    *n_reads = 0;
    *bases = nullptr;
    *offsets = nullptr;
    std::unique_lock<std::mutex> lk(rd->mu);
    if (rd->held) {
        rd->free_q.push_back(rd->held);
        rd->held = nullptr;
        rd->cv_free.notify_one();
    }
    rd->cv_full.wait(lk, [&] { return rd->done || !rd->full_q.empty(); });
    if (!rd->full_q.empty()) {
        Batch *b = rd->full_q.front();
        rd->full_q.pop_front();
        rd->held = b;
        *bases = b->b();
        *offsets = b->o();
        *n_reads = b->n_reads;
        return HULK_B200_OK;
    }
    if (rd->err) {
        rd->last_error = rd->err_text;
        return rd->err;
    }
    return HULK_B200_OK;                                                  // end of input
}

const char *hulk_b200_reader_error(const hulk_b200_reader *rd) { return rd ? rd->last_error.c_str() : ""; }

void hulk_b200_reader_close(hulk_b200_reader *rd) {
    if (!rd) return;
    {
        std::lock_guard<std::mutex> lk(rd->mu);
        rd->stop = true;
        rd->cv_free.notify_all();
    }
    if (rd->th.joinable()) rd->th.join();
    for (int i = 0; i < kRing; i++) { rd->ring[i].bases.release(); rd->ring[i].offsets.release(); }
    for (auto &b : rd->pring) { b->bases.release(); b->offsets.release(); }
    delete rd;
}

int hulk_b200_sketch_reader(hulk_b200_ctx *ctx, hulk_b200_reader *rd, uint64_t interval, hulk_b200_log_fn log,
                            void *user) {
    if (!ctx || !rd) return HULK_B200_EARG;
    char line[128];
    uint64_t seq_count = 0, sketching_interval = 0;
    for (;;) {
        const uint8_t *bases = nullptr;
        const uint64_t *offsets = nullptr;
        uint64_t n = 0;
        const int rc = hulk_b200_reader_next(rd, &bases, &offsets, &n);
        if (rc) return rc;
        if (n == 0) break;
        uint64_t done = 0;
        while (done < n) {
            uint64_t take = n - done;
            if (interval) take = std::min<uint64_t>(take, interval - (seq_count % interval));
            // the sub-range [done, done + take] of the batch's offsets addresses the same `bases`
            const int prc = hulk_b200_push_reads(ctx, bases, offsets + done, take);     // theBoss.AddSeq  :200
            if (prc) return prc;
            const uint64_t before = seq_count;
            seq_count += take;
            done += take;
            if (log)
                for (uint64_t m = before / 100000 + 1; m * 100000 <= seq_count; m++) {  // :203-207
                    snprintf(line, sizeof line, "\tprocessed %llu sequences", (unsigned long long)(m * 100000));
                    log(user, line);
                }
            if (interval && seq_count % interval == 0) {                                // :211-215
                sketching_interval++;
                if (log) {
                    snprintf(line, sizeof line, "\treached interval %llu -> histosketching",
                             (unsigned long long)sketching_interval);
                    log(user, line);
                }
                const int frc = hulk_b200_flush(ctx);
                if (frc) return frc;
            }
        }
    }
    if (log) log(user, "generating final histosketch of k-mer spectra...");              // :220
    const int frc = hulk_b200_flush(ctx);                                               // :221
    if (frc) return frc;
    const int src = hulk_b200_sync(ctx);
    if (src) return src;
    if (seq_count == 0) return HULK_B200_ENOSEQ;                                        // :237-239
    return HULK_B200_OK;
}

}  // extern "C"
