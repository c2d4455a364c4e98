// pgzip.h -- inflate ONE ordinary gzip stream on several threads (host code, no CUDA).
//
// A deflate stream is sequential for two reasons: blocks start at arbitrary BIT positions, and every block may copy
// from the 32 KiB of output in front of it.  Both are worked around the way pugz / rapidgzip do:
//   1. the compressed file is cut into chunks; for each chunk the first position that parses as the header of a
//      dynamic-Huffman block (complete code-length code, complete literal/length and distance codes, an end-of-block
//      symbol) is searched bit by bit;
//   2. every chunk is decoded from its position with an UNKNOWN window: output symbols are 16 bit, a copy that
//      reaches in front of the chunk yields a marker 0x8000 | (offset in the unknown window), markers are copied
//      around like data;
//   3. a sequential pass walks the chunks in order, checks that chunk i ended EXACTLY where chunk i+1 started (a
//      wrongly guessed start is simply never reached -- the predecessor keeps decoding through that territory and the
//      guess is dropped, so correctness never rests on the search), and hands the last 32 KiB on as the next window;
//   4. markers are replaced, symbols narrowed to bytes and CRC-32s taken in parallel; the CRCs are combined per member.
// gzip framing follows Go's compress/gzip.Reader (multistream): the messages are Go's (see ingest.cpp).
//
// Everything here is checked against zlib byte for byte (tests/test_ingest_cli.py, tools/asan).
#pragma once
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace pgz {

constexpr uint32_t kWin = 32768;
constexpr int kPB = 10;                          // primary table bits (literal/length), distance uses the same
enum Kind : uint32_t { LIT = 0, LEN = 1, EOB = 2, SUB = 3, BAD = 4, LIT2 = 5 };    // LIT2: two literals in one entry
// table entry: val << 16 | extra << 8 | kind << 4 | len
static inline uint32_t mk(uint32_t val, uint32_t extra, uint32_t kind, uint32_t len) {
    return val << 16 | extra << 8 | kind << 4 | len;
}
static inline uint32_t e_len(uint32_t e) { return e & 15; }
static inline uint32_t e_kind(uint32_t e) { return (e >> 4) & 15; }
static inline uint32_t e_extra(uint32_t e) { return (e >> 8) & 255; }
static inline uint32_t e_val(uint32_t e) { return e >> 16; }

struct Table {
    uint32_t t[(1 << kPB) + 2048];
};

// Canonical Huffman table from code lengths, with zlib's acceptance rules (inflate_table): over-subscribed sets are
// refused; an incomplete set is accepted only if it is a single code of length 1; an empty set decodes nothing.
// `sym_entry(sym, len)` gives the entry of a symbol.  Returns false for an invalid set.
template <class F>
static inline bool build_table(Table &T, const uint8_t *lens, int n, F sym_entry, bool pairs = false) {
    int count[16] = {0};
    for (int i = 0; i < n; i++) count[lens[i]]++;
    int max = 15;
    while (max > 0 && count[max] == 0) max--;
    const uint32_t bad = mk(0, 0, BAD, 1);
    for (int i = 0; i < (1 << kPB); i++) T.t[i] = bad;
    if (max == 0) return true;
    int left = 1;
    for (int len = 1; len <= 15; len++) {
        left <<= 1;
        left -= count[len];
        if (left < 0) return false;
    }
    if (left > 0 && max != 1) return false;
    uint32_t next[16];
    uint32_t code = 0;
    for (int len = 1; len <= 15; len++) {
        code = (code + (uint32_t)count[len - 1] * (len > 1)) << 1;
        next[len] = code;
    }
    // (count[0] must not enter the recurrence: handled by the (len > 1) factor for len == 1 and by construction after)
    uint8_t submax[1 << kPB];
    bool any_long = max > kPB;
    uint32_t codes[288];
    if (any_long) memset(submax, 0, sizeof submax);
    {
        uint32_t nx[16];
        memcpy(nx, next, sizeof nx);
        for (int s = 0; s < n; s++) {
            const int len = lens[s];
            if (!len) continue;
            uint32_t c = nx[len]++, r = 0;
            for (int b = 0; b < len; b++) r |= ((c >> b) & 1u) << (len - 1 - b);
            codes[s] = r;
            if (len > kPB) {
                uint8_t &m = submax[r & ((1u << kPB) - 1)];
                if (len > m) m = (uint8_t)len;
            }
        }
    }
    uint32_t free_at = 1u << kPB;
    for (int s = 0; s < n; s++) {
        const uint32_t len = lens[s];
        if (!len) continue;
        const uint32_t r = codes[s];
        if (len <= (uint32_t)kPB) {
            const uint32_t e = sym_entry(s, len);
            for (uint32_t i = r; i < (1u << kPB); i += 1u << len) T.t[i] = e;
        } else {
            const uint32_t pre = r & ((1u << kPB) - 1);
            const uint32_t sb = submax[pre] - kPB;
            if (e_kind(T.t[pre]) != SUB) {
                if (free_at + (1u << sb) > sizeof T.t / sizeof T.t[0]) return false;
                T.t[pre] = mk(free_at, sb, SUB, kPB);
                for (uint32_t i = 0; i < (1u << sb); i++) T.t[free_at + i] = bad;
                free_at += 1u << sb;
            }
            const uint32_t base = e_val(T.t[pre]);
            const uint32_t e = sym_entry(s, len - kPB);
            for (uint32_t i = r >> kPB; i < (1u << sb); i += 1u << (len - kPB)) T.t[base + i] = e;
        }
    }
    if (pairs) {
        // Literal-heavy data (FASTQ quality strings) is bound by the lookup -> shift -> lookup latency chain, one
        // literal per trip.  Where the index bits behind a literal's code hold a complete second literal code, the
        // entry yields both: val = first | second << 8, len = both lengths.  (The second lookup uses only bits the
        // index really has: an entry of length l2 is the same for every value of the bits above l2.)
        static thread_local uint32_t single[1 << kPB];
        memcpy(single, T.t, sizeof single);
        for (uint32_t i = 0; i < (1u << kPB); i++) {
            const uint32_t e1 = single[i];
            if (e_kind(e1) != LIT) continue;
            const uint32_t l1 = e_len(e1);
            const uint32_t e2 = single[i >> l1];
            if (e_kind(e2) == LIT && l1 + e_len(e2) <= (uint32_t)kPB)
                T.t[i] = mk(e_val(e1) | e_val(e2) << 8, 0, LIT2, l1 + e_len(e2));
        }
    }
    return true;
}

static const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

static inline uint32_t litlen_entry(int s, uint32_t len) {
    if (s < 256) return mk((uint32_t)s, 0, LIT, len);
    if (s == 256) return mk(0, 0, EOB, len);
    if (s <= 285) return mk(kLenBase[s - 257], kLenExtra[s - 257], LEN, len);
    return mk(0, 0, BAD, len);
}
static inline uint32_t dist_entry(int s, uint32_t len) {
    if (s < 30) return mk(kDistBase[s], kDistExtra[s], LEN, len);
    return mk(0, 0, BAD, len);
}

struct Bits {
    const uint8_t *base, *p, *end;
    uint64_t bb = 0;
    uint32_t bc = 0;
    void init(const uint8_t *b, const uint8_t *e, uint64_t bitpos) {
        base = b;
        end = e;
        p = b + (bitpos >> 3);
        bb = 0;
        bc = 0;
        const uint32_t skip = (uint32_t)(bitpos & 7);
        if (skip && p < end) {
            bb = (uint64_t)(*p++) >> skip;
            bc = 8 - skip;
        }
    }
    uint64_t pos() const { return (uint64_t)(p - base) * 8 - bc; }
    inline void refill() {
        if (end - p >= 8) {
            uint64_t w;
            memcpy(&w, p, 8);
            bb |= w << bc;
            p += (63 - bc) >> 3;
            bc |= 56;
        } else {
            while (bc <= 56 && p < end) {
                bb |= (uint64_t)(*p++) << bc;
                bc += 8;
            }
        }
    }
    inline bool need(uint32_t n) {               // make n <= 32 bits available; false at the end of the input
        if (bc < n) refill();
        return bc >= n;
    }
    inline uint32_t peek(uint32_t n) const { return (uint32_t)(bb & ((1ull << n) - 1)); }
    inline void drop(uint32_t n) {
        bb >>= n;
        bc -= n;
    }
    void align_to_byte() {                       // and give whole unread bytes back
        drop(bc & 7);
        p -= bc >> 3;
        bb = 0;
        bc = 0;
    }
};

enum Status { ST_OK = 0, ST_EOF, ST_CORRUPT, ST_HEADER, ST_TARGET, ST_END, ST_STOPPED };

// Reads a dynamic block's code description at the reader's position (just behind BFINAL/BTYPE).
static inline Status read_dynamic(Bits &in, Table &lit, Table &dist) {
    if (!in.need(14)) return ST_EOF;
    const uint32_t hlit = in.peek(5) + 257;
    in.drop(5);
    const uint32_t hdist = in.peek(5) + 1;
    in.drop(5);
    const uint32_t hclen = in.peek(4) + 4;
    in.drop(4);
    if (hlit > 286 || hdist > 30) return ST_CORRUPT;
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    uint8_t cl[19] = {0};
    for (uint32_t i = 0; i < hclen; i++) {
        if (!in.need(3)) return ST_EOF;
        cl[order[i]] = (uint8_t)in.peek(3);
        in.drop(3);
    }
    // the code-length code: 7-bit single-level table
    uint16_t clt[128];
    {
        int count[8] = {0};
        for (int i = 0; i < 19; i++) count[cl[i]]++;
        int left = 1, max = 7;
        while (max > 0 && count[max] == 0) max--;
        if (max == 0) return ST_CORRUPT;
        for (int len = 1; len <= 7; len++) {
            left <<= 1;
            left -= count[len];
            if (left < 0) return ST_CORRUPT;
        }
        if (left > 0 && max != 1) return ST_CORRUPT;        // zlib: "invalid code lengths set"
        for (int i = 0; i < 128; i++) clt[i] = 0;            // len 0 = invalid
        uint32_t next[8], code = 0;
        for (int len = 1; len <= 7; len++) {
            code = (code + (uint32_t)count[len - 1] * (len > 1)) << 1;
            next[len] = code;
        }
        for (int s = 0; s < 19; s++) {
            const int len = cl[s];
            if (!len) continue;
            uint32_t c = next[len]++, r = 0;
            for (int b = 0; b < len; b++) r |= ((c >> b) & 1u) << (len - 1 - b);
            for (uint32_t i = r; i < 128; i += 1u << len) clt[i] = (uint16_t)(s << 4 | len);
        }
    }
    uint8_t lens[286 + 30 + 140];
    uint32_t n = 0;
    const uint32_t total = hlit + hdist;
    while (n < total) {
        if (!in.need(7 + 7)) {
            if (in.bc == 0) return ST_EOF;
        }
        const uint16_t e = clt[in.peek(7)];
        const uint32_t len = e & 15, sym = e >> 4;
        if (len == 0) return ST_CORRUPT;
        if (len > in.bc) return ST_EOF;
        in.drop(len);
        if (sym < 16) {
            lens[n++] = (uint8_t)sym;
            continue;
        }
        uint32_t rep, val = 0, xb;
        if (sym == 16) {
            if (n == 0) return ST_CORRUPT;
            val = lens[n - 1];
            xb = 2;
            rep = 3;
        } else if (sym == 17) {
            xb = 3;
            rep = 3;
        } else {
            xb = 7;
            rep = 11;
        }
        if (in.bc < xb) return ST_EOF;
        rep += in.peek(xb);
        in.drop(xb);
        if (n + rep > total) return ST_CORRUPT;
        while (rep--) lens[n++] = (uint8_t)val;
    }
    if (lens[256] == 0) return ST_CORRUPT;                   // zlib: "invalid code -- missing end-of-block"
    if (!build_table(lit, lens, (int)hlit, litlen_entry, true)) return ST_CORRUPT;
    if (!build_table(dist, lens + hlit, (int)hdist, dist_entry)) return ST_CORRUPT;
    return ST_OK;
}

static inline void fixed_tables(Table &lit, Table &dist) {
    uint8_t lens[288];
    for (int i = 0; i < 144; i++) lens[i] = 8;
    for (int i = 144; i < 256; i++) lens[i] = 9;
    for (int i = 256; i < 280; i++) lens[i] = 7;
    for (int i = 280; i < 288; i++) lens[i] = 8;
    build_table(lit, lens, 288, litlen_entry, true);
    uint8_t dl[32];
    for (int i = 0; i < 32; i++) dl[i] = 5;
    build_table(dist, dl, 32, dist_entry);
}

// gzip member header at byte offset *off (RFC 1952); on success *off is the first byte of the deflate data.
// Go's readHeader: fewer than 10 bytes -> io.ErrUnexpectedEOF; wrong magic or method -> gzip.ErrHeader.
static inline Status gzip_header(const uint8_t *d, size_t size, size_t *off) {
    size_t p = *off;
    if (size - p < 10) return ST_EOF;
    if (d[p] != 0x1f || d[p + 1] != 0x8b || d[p + 2] != 8) return ST_HEADER;
    const uint8_t flg = d[p + 3];
    p += 10;
    if (flg & 4) {                                           // FEXTRA
        if (size - p < 2) return ST_EOF;
        const size_t xlen = d[p] | (size_t)d[p + 1] << 8;
        p += 2;
        if (size - p < xlen) return ST_EOF;
        p += xlen;
    }
    for (int f = 8; f <= 16; f <<= 1)                        // FNAME, FCOMMENT: zero-terminated
        if (flg & f) {
            const void *z = memchr(d + p, 0, size - p);
            if (!z) return ST_EOF;
            p = (size_t)(static_cast<const uint8_t *>(z) - d) + 1;
        }
    if (flg & 2) {                                           // FHCRC: CRC-16 of the header, checked like Go and zlib do
        if (size - p < 2) return ST_EOF;
        const uint32_t want = d[p] | (uint32_t)d[p + 1] << 8;
        if ((crc32(0L, d + *off, (uInt)(p - *off)) & 0xffffu) != want) return ST_HEADER;
        p += 2;
    }
    *off = p;
    return ST_OK;
}

// Uninitialised, reusable storage: the chunk buffers are tens of MB and live across groups, so neither zero-filling
// nor returning them to the OS between groups (page faults on the way back) is wanted.
template <class E>
struct Buf {
    E *p = nullptr;
    size_t cap = 0;
    Buf() = default;
    Buf(const Buf &) = delete;
    Buf &operator=(const Buf &) = delete;
    ~Buf() { delete[] p; }
    void reserve(size_t n, size_t keep) {        // at least n elements, the first `keep` survive
        if (n <= cap) return;
        size_t nc = cap ? cap : 1;
        while (nc < n) nc *= 2;
        E *q = new E[nc];
        if (keep) memcpy(q, p, keep * sizeof(E));
        delete[] p;
        p = q;
        cap = nc;
    }
};

struct MemberEnd {
    uint64_t out_index;                          // symbols of this chunk in front of the member's end
    uint32_t crc, isize;
};

struct Chunk {
    uint64_t start_bit = 0;                      // where decoding starts (a block header), or a member header if `at_header`
    bool at_header = false, found = false, dead = false;
    Buf<uint16_t> *out = nullptr;                // kWin marker prefix + symbols (pool slot, reused across groups)
    uint64_t n_out = 0;                          // symbols behind the prefix
    uint64_t n16 = 0;                            // the first n16 of them are 16-bit symbols in `out`, the rest bytes in `bytes`
    std::vector<MemberEnd> ends;
    Status status = ST_OK;                       // how decoding stopped
    uint64_t end_bit = 0;                        // ST_TARGET: the position reached (== start of chunk `next_live`)
    size_t next_live = 0;
    std::string corrupt_msg;
    // after stitching
    std::vector<uint8_t> window;                 // the <= 32 KiB in front of this chunk (resolved), newest last
    Buf<uint8_t> *bytes = nullptr;               // resolved output (pool slot)
    std::vector<uint32_t> piece_crc;             // crc of [prev member end, member end) pieces + the tail piece
    bool marker_error = false;
};

// The symbols of one Huffman block, up to and including its end-of-block.  The bit reader lives in locals for the
// duration (registers), `o` is the write index into out.p.  E = uint16_t while markers may be around, uint8_t after.
template <class E>
static inline Status huffman_block(Bits &in_, Buf<E> &out, uint64_t &o_, int64_t member_lo, const Table &lit,
                                   const Table &dist, std::string &msg) {
    uint64_t bb = in_.bb;
    uint32_t bc = in_.bc;
    const uint8_t *ip = in_.p;
    const uint8_t *const iend = in_.end;
    uint64_t o = o_;
    E *op = out.p;
    uint64_t cap = out.cap;
    Status st = ST_OK;
    const uint32_t PM = (1u << kPB) - 1;
#define PGZ_REFILL()                                                  \
    do {                                                              \
        if (iend - ip >= 8) {                                         \
            uint64_t w_;                                              \
            memcpy(&w_, ip, 8);                                       \
            bb |= w_ << bc;                                           \
            ip += (63 - bc) >> 3;                                     \
            bc |= 56;                                                 \
        } else {                                                      \
            while (bc <= 56 && ip < iend) {                           \
                bb |= (uint64_t)(*ip++) << bc;                        \
                bc += 8;                                              \
            }                                                         \
        }                                                             \
    } while (0)
#define PGZ_DROP(n) do { bb >>= (n); bc -= (n); } while (0)
    for (;;) {
        if (o + 320 > cap) {
            out.reserve(o + 320, o);
            op = out.p;
            cap = out.cap;
        }
        PGZ_REFILL();
        uint32_t e = lit.t[bb & PM];
        if (e_kind(e) == SUB) {
            PGZ_DROP(kPB);
            e = lit.t[e_val(e) + (uint32_t)(bb & ((1u << e_extra(e)) - 1))];
        }
        if (e_len(e) > bc) { st = ST_EOF; break; }
        PGZ_DROP(e_len(e));
        const uint32_t kind = e_kind(e);
        if (kind == LIT || kind == LIT2) {
            // up to three entries of literals per refill (>= 41 bits are left, a primary-table entry takes <= 10)
            op[o] = (E)(e_val(e) & 0xffu);
            op[o + 1] = (E)(e_val(e) >> 8);                  // (slack: overwritten when the entry held one literal)
            o += 1 + (kind == LIT2);
            uint32_t e2 = lit.t[bb & PM];
            uint32_t k2 = e_kind(e2);
            if ((k2 == LIT || k2 == LIT2) && e_len(e2) <= bc) {
                PGZ_DROP(e_len(e2));
                op[o] = (E)(e_val(e2) & 0xffu);
                op[o + 1] = (E)(e_val(e2) >> 8);
                o += 1 + (k2 == LIT2);
                e2 = lit.t[bb & PM];
                k2 = e_kind(e2);
                if ((k2 == LIT || k2 == LIT2) && e_len(e2) <= bc) {
                    PGZ_DROP(e_len(e2));
                    op[o] = (E)(e_val(e2) & 0xffu);
                    op[o + 1] = (E)(e_val(e2) >> 8);
                    o += 1 + (k2 == LIT2);
                }
            }
            continue;
        }
        if (kind == EOB) break;
        if (kind != LEN) { msg = "invalid literal/length code"; st = ST_CORRUPT; break; }
        const uint32_t lx = e_extra(e);
        if (lx > bc) { st = ST_EOF; break; }
        const uint32_t L = e_val(e) + (uint32_t)(bb & ((1u << lx) - 1));
        PGZ_DROP(lx);
        // (no refill: a match is always the first entry behind one, so >= 56 - 15 - 5 = 36 bits are left and a
        // distance takes <= 15 + 13; at the very end of the input the length checks below catch a short buffer)
        uint32_t de = dist.t[bb & PM];
        if (e_kind(de) == SUB) {
            PGZ_DROP(kPB);
            de = dist.t[e_val(de) + (uint32_t)(bb & ((1u << e_extra(de)) - 1))];
        }
        if (e_len(de) > bc) { st = ST_EOF; break; }
        if (e_kind(de) != LEN) { msg = "invalid distance code"; st = ST_CORRUPT; break; }
        PGZ_DROP(e_len(de));
        const uint32_t dx = e_extra(de);
        if (dx > bc) { st = ST_EOF; break; }
        const uint32_t D = e_val(de) + (uint32_t)(bb & ((1u << dx) - 1));
        PGZ_DROP(dx);
        if ((int64_t)o - (int64_t)D < member_lo) { msg = "invalid distance too far back"; st = ST_CORRUPT; break; }
        const E *src = op + o - D;
        E *dst = op + o;
        if (D >= 16) {                                       // 16 symbols at a time; may write up to 15 symbols past the
            memcpy(dst, src, 16 * sizeof(E));                // match, into slack that later output overwrites
            for (uint32_t i = 16; i < L; i += 16) memcpy(dst + i, src + i, 16 * sizeof(E));
        } else if (D >= 8) {
            for (uint32_t i = 0; i < L; i += 8) memcpy(dst + i, src + i, 8 * sizeof(E));
        } else {
            for (uint32_t i = 0; i < L; i++) dst[i] = src[i];
        }
        o += L;
    }
#undef PGZ_REFILL
#undef PGZ_DROP
    in_.bb = bb;
    in_.bc = bc;
    in_.p = ip;
    o_ = o;
    return st;
}

// Decodes from chunk.start_bit until a block boundary equal to one of targets[ti..] is reached (ST_TARGET), the file
// ends cleanly behind a member (ST_END), or something is wrong.  `member_lo`: index of the first symbol of the current
// member, or -kWin while the chunk runs on an unknown window.
static inline void decode_chunk(const uint8_t *d, size_t size, Chunk &c, const std::vector<Chunk> &all, size_t self,
                                uint64_t budget, const std::atomic<bool> *stop) {
    Bits in;
    Buf<uint16_t> &out = *c.out;
    out.reserve(kWin + (1u << 20), 0);
    for (uint32_t i = 0; i < kWin; i++) out.p[i] = (uint16_t)(0x8000u | i);
    uint64_t o = kWin;                           // write index into `out`
    int64_t member_lo;                           // absolute index in `out` of the oldest symbol a copy may reach
    size_t ti = self + 1;                        // next candidate target
    static thread_local Table tl_lit, tl_dist;
    Table &lit = tl_lit, &dist = tl_dist;
    if (c.at_header) {
        size_t off = (size_t)(c.start_bit >> 3);
        const Status hs = gzip_header(d, size, &off);
        if (hs != ST_OK) { c.status = hs; c.n_out = 0; return; }
        in.init(d, d + size, (uint64_t)off * 8);
        member_lo = (int64_t)o;
    } else {
        in.init(d, d + size, c.start_bit);
        member_lo = 0;
    }
    // Once no copy can reach a marker any more -- the last 32 KiB (or everything since the member's first byte) are
    // plain symbols -- the rest of the chunk is decoded straight into bytes: half the writes, nothing to resolve.
    Buf<uint8_t> &B = *c.bytes;
    bool byte_mode = false;
    uint64_t next_check = kWin;
    auto finish = [&](Status s) {
        c.status = s;
        c.n_out = o - kWin;
        if (!byte_mode) c.n16 = c.n_out;
        c.end_bit = in.pos();
    };
    for (;;) {
        // ---- block boundary ----
        const uint64_t here = in.pos();
        while (ti < all.size() && (!all[ti].found || all[ti].start_bit < here)) ti++;
        if (ti < all.size() && all[ti].start_bit == here) { c.next_live = ti; finish(ST_TARGET); return; }
        // memory bound for very compressible input: stop at this block boundary, the next group goes on from here
        // (the chunks behind this one are dropped for this group)
        if (o - kWin > budget) { c.next_live = all.size(); finish(ST_TARGET); return; }
        if (stop && stop->load(std::memory_order_relaxed)) { finish(ST_STOPPED); return; }
        if (!byte_mode && o - kWin >= next_check) {
            const uint64_t k = o - kWin;
            const int64_t lo_k = member_lo - (int64_t)kWin;              // < 0: the unknown window is still reachable
            if (lo_k > 0 || k >= kWin) {
                const uint64_t reach = lo_k > 0 ? std::min<uint64_t>(k - (uint64_t)lo_k, kWin) : kWin;
                const uint16_t *tail = out.p + kWin + k - reach;
                uint16_t any = 0;
                for (uint64_t i = 0; i < reach; i++) any |= tail[i];
                if (!(any & 0x8000u)) {
                    c.n16 = k;
                    B.reserve(k + (1u << 20), 0);
                    for (uint64_t i = 0; i < reach; i++) B.p[k - reach + i] = (uint8_t)tail[i];
                    byte_mode = true;
                }
            }
            next_check = k + kWin;
        }
        if (!in.need(3)) { finish(ST_EOF); return; }
        const uint32_t bfinal = in.peek(1), btype = (in.peek(3) >> 1);
        in.drop(3);
        if (btype == 3) { c.corrupt_msg = "invalid block type"; finish(ST_CORRUPT); return; }
        if (btype == 0) {
            in.align_to_byte();
            if (in.end - in.p < 4) { finish(ST_EOF); return; }
            const uint32_t len = in.p[0] | (uint32_t)in.p[1] << 8, nlen = in.p[2] | (uint32_t)in.p[3] << 8;
            if ((len ^ 0xffffu) != nlen) { c.corrupt_msg = "invalid stored block lengths"; finish(ST_CORRUPT); return; }
            in.p += 4;
            const bool cut = (size_t)(in.end - in.p) < len;
            const uint32_t take = cut ? (uint32_t)(in.end - in.p) : len;
            if (byte_mode) {
                B.reserve(o - kWin + take + 512, o - kWin);
                memcpy(B.p + (o - kWin), in.p, take);
            } else {
                out.reserve(o + take + 512, o);
                for (uint32_t i = 0; i < take; i++) out.p[o + i] = in.p[i];
            }
            o += take;
            in.p += take;
            if (cut) { finish(ST_EOF); return; }
        } else {
            if (btype == 1) fixed_tables(lit, dist);
            else {
                const Status s = read_dynamic(in, lit, dist);
                if (s != ST_OK) { c.corrupt_msg = "invalid code lengths"; finish(s); return; }
            }
            Status hs;
            if (byte_mode) {
                uint64_t k = o - kWin;
                hs = huffman_block<uint8_t>(in, B, k, member_lo - (int64_t)kWin, lit, dist, c.corrupt_msg);
                o = k + kWin;
            } else {
                hs = huffman_block<uint16_t>(in, out, o, member_lo, lit, dist, c.corrupt_msg);
            }
            if (hs != ST_OK) { finish(hs); return; }
        }
        if (bfinal) {
            // ---- member trailer, then the next header or the end of the file (gzip.Reader.Read, multistream) ----
            in.align_to_byte();
            if (in.end - in.p < 8) { finish(ST_EOF); return; }
            MemberEnd me;
            me.out_index = o - kWin;
            me.crc = in.p[0] | (uint32_t)in.p[1] << 8 | (uint32_t)in.p[2] << 16 | (uint32_t)in.p[3] << 24;
            me.isize = in.p[4] | (uint32_t)in.p[5] << 8 | (uint32_t)in.p[6] << 16 | (uint32_t)in.p[7] << 24;
            c.ends.push_back(me);
            in.p += 8;
            if (in.p == in.end) { finish(ST_END); return; }
            size_t off = (size_t)(in.p - d);
            const Status hs = gzip_header(d, size, &off);
            if (hs != ST_OK) { finish(hs); return; }
            in.init(d, d + size, (uint64_t)off * 8);
            member_lo = (int64_t)o;
        }
    }
}

// One complete raw deflate stream with nothing in front of it (a BGZF member's payload), straight into bytes.
// *consumed = bytes of input used, k = bytes produced.
static inline Status inflate_raw(const uint8_t *d, size_t n, Buf<uint8_t> &B, uint64_t &k, size_t *consumed,
                                 std::string &msg) {
    static thread_local Table tl_lit, tl_dist;
    Table &lit = tl_lit, &dist = tl_dist;
    Bits in;
    in.init(d, d + n, 0);
    k = 0;
    B.reserve(1u << 17, 0);
    for (;;) {
        if (!in.need(3)) return ST_EOF;
        const uint32_t bfinal = in.peek(1), btype = in.peek(3) >> 1;
        in.drop(3);
        if (btype == 3) { msg = "invalid block type"; return ST_CORRUPT; }
        if (btype == 0) {
            in.align_to_byte();
            if (in.end - in.p < 4) return ST_EOF;
            const uint32_t len = in.p[0] | (uint32_t)in.p[1] << 8, nlen = in.p[2] | (uint32_t)in.p[3] << 8;
            if ((len ^ 0xffffu) != nlen) { msg = "invalid stored block lengths"; return ST_CORRUPT; }
            in.p += 4;
            if ((size_t)(in.end - in.p) < len) return ST_EOF;
            B.reserve(k + len + 512, k);
            memcpy(B.p + k, in.p, len);
            k += len;
            in.p += len;
        } else {
            if (btype == 1) fixed_tables(lit, dist);
            else {
                const Status s = read_dynamic(in, lit, dist);
                if (s != ST_OK) { msg = "invalid code lengths"; return s; }
            }
            const Status hs = huffman_block<uint8_t>(in, B, k, 0, lit, dist, msg);
            if (hs != ST_OK) return hs;
        }
        if (bfinal) break;
    }
    in.align_to_byte();
    *consumed = (size_t)(in.p - d);
    return ST_OK;
}

// First position in [from_bit, to_bit) that parses as a non-final dynamic block header.
static inline bool find_block(const uint8_t *d, size_t size, uint64_t from_bit, uint64_t to_bit, uint64_t *found) {
    static thread_local Table tl_lit, tl_dist;
    Table &lit = tl_lit, &dist = tl_dist;
    for (uint64_t pos = from_bit; pos < to_bit; pos++) {
        // 3 + 14 header bits straight from memory
        const size_t byte = (size_t)(pos >> 3);
        if (byte + 8 > size) return false;
        uint64_t w;
        memcpy(&w, d + byte, 8);
        w >>= (pos & 7);
        if ((w & 7) != 4) continue;                          // BFINAL = 0, BTYPE = 10
        if (((w >> 3) & 31) > 29) continue;                  // HLIT <= 286
        if (((w >> 8) & 31) > 29) continue;                  // HDIST <= 30
        {
            // the code-length code must be complete (or a single 1-bit code): Kraft sum of its <= 19 three-bit
            // lengths, straight from the next 57 bits -- rejects ~99 % of what got this far without building anything
            const uint32_t hclen = 4 + (uint32_t)((w >> 13) & 15);
            const uint64_t at = pos + 17;
            if ((size_t)(at >> 3) + 8 > size) return false;
            uint64_t v;
            memcpy(&v, d + (at >> 3), 8);
            v >>= (at & 7);
            uint32_t kraft = 0, nz = 0, last = 0;
            for (uint32_t i = 0; i < hclen; i++, v >>= 3) {
                const uint32_t l = (uint32_t)(v & 7);
                if (l) { kraft += 128u >> l; nz++; last = l; }
            }
            if (!(kraft == 128 || (nz == 1 && last == 1))) continue;
        }
        Bits in;
        in.init(d, d + size, pos + 3);
        if (read_dynamic(in, lit, dist) != ST_OK) continue;
        // A complete prefix code decodes ANY bit string, so the header alone is a weak test (a few false hits per
        // MB).  Walk the block's symbols (no output) to its end-of-block and ask for a sane header behind it.
        bool good = false;
        for (uint64_t produced = 0; produced < (8u << 20);) {
            in.refill();
            uint32_t e = lit.t[in.bb & ((1u << kPB) - 1)];
            if (e_kind(e) == SUB) {
                in.drop(kPB);
                e = lit.t[e_val(e) + (uint32_t)(in.bb & ((1u << e_extra(e)) - 1))];
            }
            if (e_len(e) > in.bc) break;
            in.drop(e_len(e));
            const uint32_t kind = e_kind(e);
            if (kind == LIT || kind == LIT2) { produced += 1 + (kind == LIT2); continue; }
            if (kind == EOB) {
                if (!in.need(17)) break;
                const uint32_t h = in.peek(17), bt = (h >> 1) & 3;
                if (bt == 2) {
                    if (((h >> 3) & 31) > 29 || ((h >> 8) & 31) > 29) break;
                    Bits nx = in;
                    nx.drop(3);
                    static thread_local Table l2, d2;
                    good = read_dynamic(nx, l2, d2) == ST_OK;
                } else if (bt == 0) {
                    Bits nx = in;
                    nx.drop(3);
                    nx.align_to_byte();
                    good = nx.end - nx.p >= 4 && ((nx.p[0] | nx.p[1] << 8) ^ 0xffff) == (nx.p[2] | nx.p[3] << 8);
                } else good = false;                          // fixed blocks behind a guess: not worth trusting
                break;
            }
            if (kind != LEN) break;
            const uint32_t lx = e_extra(e);
            if (lx > in.bc) break;
            produced += e_val(e) + in.peek(lx);
            in.drop(lx);
            if (in.bc < 32) in.refill();
            uint32_t de = dist.t[in.bb & ((1u << kPB) - 1)];
            if (e_kind(de) == SUB) {
                in.drop(kPB);
                de = dist.t[e_val(de) + (uint32_t)(in.bb & ((1u << e_extra(de)) - 1))];
            }
            if (e_len(de) > in.bc || e_kind(de) != LEN) break;
            in.drop(e_len(de));
            if (e_extra(de) > in.bc) break;
            in.drop(e_extra(de));
        }
        if (!good) continue;
        *found = pos;
        return true;
    }
    return false;
}

struct Options {
    unsigned threads = 4;
    size_t chunk_bytes = 4u << 20;
};

// Inflates the gzip members of the file image [d, d + size) from byte offset `start` and hands the bytes to
// sink(ptr, len) in order; sink returns false to stop early.  Returns "" or Go's error text; *stopped is set when the
// sink asked to stop.  `first`: no member has been read from this file yet (an empty input is then io.EOF).
// One group of chunks after the parallel stages: everything the sequential hand-over needs.
struct Group {
    std::vector<Chunk> ch;
    std::vector<size_t> live;                    // the chunks that really follow each other, in order
    bool more = false;                           // the stream goes on behind this group (at next_bit)
};

template <class Sink>
static inline std::string inflate_parallel(const uint8_t *d, size_t size, size_t start, bool first, const Options &opt,
                                           const std::atomic<bool> *stop, Sink sink, bool *stopped) {
    *stopped = false;
    if (start >= size) return first ? "EOF" : "";
    const unsigned T = opt.threads ? opt.threads : 1;
    const size_t CB = opt.chunk_bytes ? opt.chunk_bytes : (4u << 20);
    // state of the stitching pass (advanced by prepare) ...
    std::vector<uint8_t> window;                 // last <= 32 KiB of the current member
    uint64_t next_bit = (uint64_t)start * 8;     // where the next chunk has to start
    bool next_is_header = true;
    unsigned barren = 0, quiet = 0;              // see prepare: groups without a usable guess / groups left without search
    // ... and of the hand-over (advanced by hand_over)
    uint32_t crc = 0;                            // of the current member so far
    uint64_t member_len = 0;
    // T - 1 workers that live as long as this call (three parallel stages per group: spawning threads for each
    // would cost a tenth of a group's time); the calling thread of run() works too
    struct Pool {
        std::mutex mu;
        std::condition_variable cv_job, cv_done;
        std::function<void(size_t)> fn;
        size_t n = 0, next = 0, active = 0;
        uint64_t epoch = 0;
        bool quit = false;
        std::vector<std::thread> th;
        void worker() {
            uint64_t seen = 0;
            std::unique_lock<std::mutex> lk(mu);
            for (;;) {
                cv_job.wait(lk, [&] { return quit || epoch != seen; });
                if (quit) return;
                seen = epoch;
                active++;
                while (next < n) {
                    const size_t i = next++;
                    lk.unlock();
                    fn(i);
                    lk.lock();
                }
                if (--active == 0) cv_done.notify_all();
            }
        }
        explicit Pool(unsigned workers) {
            for (unsigned t = 0; t < workers; t++) th.emplace_back([this] { worker(); });
        }
        ~Pool() {
            {
                std::lock_guard<std::mutex> lk(mu);
                quit = true;
            }
            cv_job.notify_all();
            for (auto &x : th) x.join();
        }
        void run(std::function<void(size_t)> f, size_t count) {
            std::unique_lock<std::mutex> lk(mu);
            fn = std::move(f);
            n = count;
            next = 0;
            epoch++;
            active++;                                        // the caller
            cv_job.notify_all();
            while (next < n) {
                const size_t i = next++;
                lk.unlock();
                fn(i);
                lk.lock();
            }
            if (--active == 0) cv_done.notify_all();
            cv_done.wait(lk, [&] { return active == 0 && next >= n; });
            n = 0;
        }
    } pool(T > 1 ? T - 1 : 0);
    auto run = [&](std::function<void(size_t)> fn, size_t n) { pool.run(std::move(fn), n); };
    const bool timing = getenv("HULK_B200_PGZ_TIMING") != nullptr;
    double t_phase[7] = {0, 0, 0, 0, 0, 0, 0};          // [5] chunks decoded, [6] chunks used
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    struct Report {
        bool on;
        double *t;
        ~Report() {
            if (on) fprintf(stderr, "pgzip phases: search %.3f decode %.3f windows %.3f resolve %.3f sink %.3f s; chunks decoded %.0f used %.0f\n", t[0], t[1], t[2], t[3], t[4], t[5], t[6]);
        }
    } report{timing, t_phase};
    // two sets of chunk buffers: one group is handed over while the next one is decoded
    std::vector<Buf<uint16_t>> pool16[2] = {std::vector<Buf<uint16_t>>(T), std::vector<Buf<uint16_t>>(T)};
    std::vector<Buf<uint8_t>> pool8[2] = {std::vector<Buf<uint8_t>>(T), std::vector<Buf<uint8_t>>(T)};

    // ---- stages 1-4 for the group that starts at next_bit ----
    auto prepare = [&](Group &g, int set) {
        double t0 = now();
        // chunk 0 continues exactly at next_bit, chunks 1..T start at searched block headers
        std::vector<Chunk> &ch = g.ch;
        ch.clear();
        ch.resize(T + 1);
        g.live.clear();
        for (unsigned i = 0; i < T; i++) { ch[i].out = &pool16[set][i]; ch[i].bytes = &pool8[set][i]; }
        ch[0].start_bit = next_bit;
        ch[0].at_header = next_is_header;
        ch[0].found = true;
        const size_t base = (size_t)(next_bit >> 3);
        // input without dynamic blocks (stored or fixed-Huffman data) offers no starts: after two groups in which
        // only chunk 0 was usable the search is skipped for a while instead of scanning every chunk in vain
        const bool search = quiet == 0;
        if (quiet) quiet--;
        run([&](size_t i) {
            if (i == 0 || !search) return;
            const size_t from = base + i * CB;
            if (from >= size) return;
            const size_t to = std::min(size, from + std::min<size_t>(CB, 512u << 10));   // blocks are tens of KB: no start in 512 KB = give up
            uint64_t pos;
            if (find_block(d, size, (uint64_t)from * 8, (uint64_t)to * 8, &pos)) {
                ch[i].start_bit = pos;
                ch[i].found = true;
            }
        }, T + 1);
        t_phase[0] += now() - t0;
        t0 = now();
        // the last chunk of a group is only a stop mark for its predecessor
        run([&](size_t i) { if (ch[i].found) decode_chunk(d, size, ch[i], ch, i, std::max<uint64_t>(32 * (uint64_t)CB, 1u << 20), stop); }, T);
        t_phase[1] += now() - t0;
        t0 = now();
        // sequential pass: which chunks are real, windows, member ends
        std::vector<size_t> &live = g.live;
        size_t i = 0;
        for (;;) {
            live.push_back(i);
            if (ch[i].status != ST_TARGET) break;
            if (ch[i].next_live >= T) break;                 // reached the stop mark: the next group starts there
            i = ch[i].next_live;
        }
        if (search && T > 1) {
            barren = live.size() == 1 ? barren + 1 : 0;
            if (barren >= 2) { quiet = 16; barren = 0; }
        }
        for (size_t q = 0; q < T; q++) t_phase[5] += ch[q].found;
        t_phase[6] += (double)live.size();
        // windows: chunk `live[j]` needs the window in front of it
        for (size_t j = 0; j < live.size(); j++) {
            Chunk &c = ch[live[j]];
            c.window = window;
            // resolve this chunk's tail to get the next window (only the part that reaches the end matters)
            const uint64_t last_end = c.ends.empty() ? 0 : c.ends.back().out_index;
            const uint64_t tail = c.n_out - last_end;        // symbols of the member still open at the chunk's end
            const uint64_t take = std::min<uint64_t>(tail, kWin);
            std::vector<uint8_t> nw;
            auto sym = [&](uint64_t k) -> uint16_t { return k < c.n16 ? c.out->p[kWin + k] : (uint16_t)c.bytes->p[k]; };
            if (!c.ends.empty()) {
                // a member started inside this chunk: its window is only what this chunk wrote behind that start
                nw.resize(take);
                for (uint64_t k = 0; k < take; k++) {
                    const uint16_t sy = sym(c.n_out - take + k);
                    if (sy & 0x8000u) { c.marker_error = true; nw[k] = 0; }
                    else nw[k] = (uint8_t)sy;
                }
            } else {
                const uint64_t keep = std::min<uint64_t>(window.size(), kWin - take);
                nw.assign(window.end() - (ptrdiff_t)keep, window.end());
                nw.resize(keep + take);
                for (uint64_t k = 0; k < take; k++) {
                    const uint16_t sy = sym(c.n_out - take + k);
                    if (sy & 0x8000u) {
                        const uint32_t back = kWin - (sy & 0x7fffu);         // 1 = the byte right in front of the chunk
                        if (back > window.size()) { c.marker_error = true; nw[keep + k] = 0; }
                        else nw[keep + k] = window[window.size() - back];
                    } else nw[keep + k] = (uint8_t)sy;
                }
            }
            window.swap(nw);
        }
        t_phase[2] += now() - t0;
        t0 = now();
        // parallel: resolve markers, narrow, crc per piece
        run([&](size_t j) {
            Chunk &c = ch[live[j]];
            c.bytes->reserve(c.n_out + 64, c.n16 < c.n_out ? c.n_out : 0);     // (bytes behind n16 are already there)
            const uint16_t *src = c.out->p + kWin;
            const uint8_t *w = c.window.data();
            const size_t wn = c.window.size();
            uint8_t *dst = c.bytes->p;
            const uint64_t first_end = c.ends.empty() ? c.n_out : c.ends[0].out_index;
            uint64_t k = 0;
            // With a full window in front of the chunk every marker is valid, and a 64 K-entry table turns symbols
            // into bytes without a branch (FASTQ keeps markers alive in every record: the constant parts are always
            // copied from the record before).  Blocks of 64 symbols without any marker are just narrowed.
            static thread_local std::vector<uint8_t> lut;
            const bool use_lut = wn == kWin && c.ends.empty();
            if (use_lut) {
                lut.resize(65536);
                for (int q = 0; q < 256; q++) lut[q] = (uint8_t)q;
                memcpy(lut.data() + 0x8000, w, kWin);        // marker 0x8000 | i  ->  window[i]
            }
            const uint8_t *lt = lut.data();
            for (; k + 64 <= c.n16; k += 64) {
                uint16_t any = 0;
                for (int q = 0; q < 64; q++) any |= src[k + q];
                if (!(any & 0x8000u)) {
                    for (int q = 0; q < 64; q++) dst[k + q] = (uint8_t)src[k + q];
                    continue;
                }
                if (use_lut) {
                    for (int q = 0; q < 64; q++) dst[k + q] = lt[src[k + q]];
                    continue;
                }
                for (int q = 0; q < 64; q++) {
                    const uint16_t sy = src[k + q];
                    if (sy & 0x8000u) {
                        const uint32_t back = kWin - (sy & 0x7fffu);
                        if (k + q >= first_end || back > wn) { c.marker_error = true; dst[k + q] = 0; }
                        else dst[k + q] = w[wn - back];
                    } else dst[k + q] = (uint8_t)sy;
                }
            }
            for (; k < c.n16; k++) {
                const uint16_t sy = src[k];
                if (sy & 0x8000u) {
                    const uint32_t back = kWin - (sy & 0x7fffu);
                    if (k >= first_end || back > wn) { c.marker_error = true; dst[k] = 0; }
                    else dst[k] = w[wn - back];
                } else dst[k] = (uint8_t)sy;
            }
            uint64_t from = 0;
            for (size_t m = 0; m <= c.ends.size(); m++) {
                const uint64_t to = m < c.ends.size() ? c.ends[m].out_index : c.n_out;
                uint32_t pc = 0;
                for (uint64_t a = from; a < to; a += 1u << 30)
                    pc = (uint32_t)crc32(pc, dst + a, (uInt)std::min<uint64_t>(1u << 30, to - a));
                c.piece_crc.push_back(pc);
                from = to;
            }
        }, live.size());
        t_phase[3] += now() - t0;
        const Chunk &lastc = ch[live.back()];
        g.more = lastc.status == ST_TARGET;
        if (g.more) {
            next_bit = lastc.end_bit;
            next_is_header = false;
        }
    };

    // ---- stage 5: hand the bytes on, close members, report the first problem in stream order ----
    auto hand_over = [&](Group &g, bool *done) -> std::string {
        *done = true;
        for (size_t j = 0; j < g.live.size(); j++) {
            Chunk &c = g.ch[g.live[j]];
            uint64_t from = 0;
            // a copy reached in front of the member's first byte
            if (c.marker_error) return std::string("flate: corrupt input (invalid distance too far back)");
            for (size_t m = 0; m <= c.ends.size(); m++) {
                const uint64_t to = m < c.ends.size() ? c.ends[m].out_index : c.n_out;
                if (to > from && !sink(c.bytes->p + from, (size_t)(to - from))) { *stopped = true; return ""; }
                crc = (uint32_t)crc32_combine(crc, c.piece_crc[m], (z_off_t)(to - from));
                member_len += to - from;
                if (m < c.ends.size()) {
                    if (crc != c.ends[m].crc || (uint32_t)member_len != c.ends[m].isize) return "gzip: invalid checksum";
                    crc = 0;
                    member_len = 0;
                }
                from = to;
            }
        }
        const Chunk &lastc = g.ch[g.live.back()];
        switch (lastc.status) {
            case ST_TARGET: *done = false; return "";
            case ST_END: return "";
            case ST_EOF: return "unexpected EOF";
            case ST_HEADER: return "gzip: invalid header";
            case ST_STOPPED: *stopped = true; return "";
            default: return "flate: corrupt input (" + lastc.corrupt_msg + ")";
        }
    };

    Group grp[2];
    int cur = 0;
    prepare(grp[0], 0);
    for (;;) {
        if (stop && stop->load()) { *stopped = true; return ""; }
        std::thread ahead;
        if (grp[cur].more) ahead = std::thread([&, cur] { prepare(grp[cur ^ 1], cur ^ 1); });
        bool done = true;
        const double t0 = now();
        const std::string text = hand_over(grp[cur], &done);
        t_phase[4] += now() - t0;
        if (ahead.joinable()) ahead.join();
        if (done || !text.empty() || *stopped) return text;
        cur ^= 1;
    }
}

}  // namespace pgz
