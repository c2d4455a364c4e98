// sketch_json.cpp -- the read side of the JSON sketch format: sketchio.LoadHULKdata and HULKdata.FindSketch
// (reference src/sketchio/sketchio.go:98-261), used by `hulk smash`.  Host code only.
//
// Checks kept from the reference, with its messages: class must be "hulk_sketch", version must equal 1.0.0,
// at least one signature, known algorithm, a stored md5sum that matches helpers.MD5sum over the mins.
#include <sys/stat.h>

#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/hulk_b200.h"

namespace {

// ---- a small JSON reader: just enough for documents written by encoding/json ----
struct JValue {
    enum Kind { NUL, BOOL, NUM, STR, ARR, OBJ } kind = NUL;
    bool b = false;
    std::string text;                                   // STR: decoded; NUM: the literal
    std::vector<JValue> arr;
    std::vector<std::pair<std::string, JValue>> obj;
    const JValue *get(const char *key) const {
        for (const auto &kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
};

struct JParser {
    const char *p, *end;
    bool ok = true;
    int depth = 0;                                      // a sketch document nests 5 deep; encoding/json itself stops at 10000
    void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; }
    bool lit(const char *s) {
        const size_t n = strlen(s);
        if ((size_t)(end - p) >= n && !memcmp(p, s, n)) { p += n; return true; }
        return false;
    }
    static void utf8(std::string &o, unsigned cp) {
        if (cp < 0x80) o += (char)cp;
        else if (cp < 0x800) { o += (char)(0xC0 | (cp >> 6)); o += (char)(0x80 | (cp & 0x3F)); }
        else if (cp < 0x10000) { o += (char)(0xE0 | (cp >> 12)); o += (char)(0x80 | ((cp >> 6) & 0x3F)); o += (char)(0x80 | (cp & 0x3F)); }
        else { o += (char)(0xF0 | (cp >> 18)); o += (char)(0x80 | ((cp >> 12) & 0x3F)); o += (char)(0x80 | ((cp >> 6) & 0x3F)); o += (char)(0x80 | (cp & 0x3F)); }
    }
    std::string str() {
        std::string o;
        p++;                                             // opening quote
        while (p < end && *p != '"') {
            if (*p == '\\' && p + 1 < end) {
                p++;
                switch (*p) {
                    case 'n': o += '\n'; break;
                    case 't': o += '\t'; break;
                    case 'r': o += '\r'; break;
                    case 'b': o += '\b'; break;
                    case 'f': o += '\f'; break;
                    case 'u': {
                        if (end - p < 5) { ok = false; return o; }
                        unsigned cp = (unsigned)strtoul(std::string(p + 1, p + 5).c_str(), nullptr, 16);
                        p += 4;
                        if (cp >= 0xD800 && cp < 0xDC00 && end - p >= 7 && p[1] == '\\' && p[2] == 'u') {
                            const unsigned lo = (unsigned)strtoul(std::string(p + 3, p + 7).c_str(), nullptr, 16);
                            cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                            p += 6;
                        }
                        utf8(o, cp);
                        break;
                    }
                    default: o += *p;
                }
                p++;
            } else {
                o += *p++;
            }
        }
        if (p >= end) ok = false;
        else p++;
        return o;
    }
    JValue value() {
        JValue v;
        ws();
        if (p >= end) { ok = false; return v; }
        struct Depth {
            int &d;
            explicit Depth(int &x) : d(x) { d++; }
            ~Depth() { d--; }
        } guard(depth);
        if (depth > 256) { ok = false; return v; }
        if (*p == '{') {
            v.kind = JValue::OBJ;
            p++;
            ws();
            if (p < end && *p == '}') { p++; return v; }
            while (ok) {
                ws();
                if (p >= end || *p != '"') { ok = false; break; }
                std::string k = str();
                ws();
                if (p >= end || *p != ':') { ok = false; break; }
                p++;
                v.obj.emplace_back(std::move(k), value());
                ws();
                if (p < end && *p == ',') { p++; continue; }
                if (p < end && *p == '}') { p++; break; }
                ok = false;
            }
        } else if (*p == '[') {
            v.kind = JValue::ARR;
            p++;
            ws();
            if (p < end && *p == ']') { p++; return v; }
            while (ok) {
                v.arr.push_back(value());
                ws();
                if (p < end && *p == ',') { p++; continue; }
                if (p < end && *p == ']') { p++; break; }
                ok = false;
            }
        } else if (*p == '"') {
            v.kind = JValue::STR;
            v.text = str();
        } else if (lit("true")) { v.kind = JValue::BOOL; v.b = true; }
        else if (lit("false")) { v.kind = JValue::BOOL; }
        else if (lit("null")) { v.kind = JValue::NUL; }
        else {
            const char *q = p;
            while (p < end && (strchr("+-.eE", *p) || (*p >= '0' && *p <= '9'))) p++;
            if (p == q) { ok = false; return v; }
            v.kind = JValue::NUM;
            v.text.assign(q, p);
        }
        return v;
    }
};

struct Sig {
    std::string algorithm, md5;
    uint32_t ksize = 0;
    std::vector<uint64_t> mins;
    std::vector<double> weights;
};

std::string str_of(const JValue *v) { return (v && v->kind == JValue::STR) ? v->text : std::string(); }

}  // namespace

struct hulk_b200_sketch_file {
    std::string klass, filename, version, banner, path, err;
    std::vector<Sig> sigs;
};

extern "C" {

// sketchio.LoadHULKdata (src/sketchio/sketchio.go:98-195)
int hulk_b200_sketch_load(const char *path, hulk_b200_sketch_file **out, char *err, uint64_t errcap) {
    auto fail = [&](int code, const std::string &m) {
        if (err && errcap) snprintf(err, (size_t)errcap, "%s", m.c_str());
        return code;
    };
    if (!path || !out) return fail(HULK_B200_EARG, "path/out is NULL");
    *out = nullptr;
    struct stat st;
    if (stat(path, &st) != 0)                                                    // helpers.CheckFile
        return fail(HULK_B200_EIO, (errno == ENOENT ? std::string("file does not exist: ")
                                                    : std::string("can't access file (check permissions): ")) + path);
    FILE *fh = fopen(path, "rb");
    if (!fh) return fail(HULK_B200_EIO, std::string("open ") + path + ": " + strerror(errno));
    std::string data;
    char buf[1 << 16];
    size_t got;
    while ((got = fread(buf, 1, sizeof buf, fh)) > 0) data.append(buf, got);
    fclose(fh);
    JParser jp{data.data(), data.data() + data.size()};
    const JValue root = jp.value();
    if (!jp.ok || root.kind != JValue::OBJ) return fail(HULK_B200_EIO, std::string("not a JSON sketch: ") + path);
    std::unique_ptr<hulk_b200_sketch_file> f(new hulk_b200_sketch_file());
    f->path = path;
    f->klass = str_of(root.get("class"));
    f->filename = str_of(root.get("filename"));
    f->version = str_of(root.get("version"));
    f->banner = str_of(root.get("banner_label"));
    const JValue *sigs = root.get("signatures");
    if (sigs && sigs->kind == JValue::ARR) {
        for (const JValue &sv : sigs->arr) {
            Sig s;
            s.algorithm = str_of(sv.get("Algorithm"));
            if (s.algorithm != "histosketch" && s.algorithm != "kmv" && s.algorithm != "khf")
                return fail(HULK_B200_EIO, "unknown sketching algorithm: " + s.algorithm);
            const JValue *sk = sv.get("Sketch");
            if (sk && sk->kind == JValue::OBJ) {
                if (const JValue *v = sk->get("ksize")) s.ksize = (uint32_t)strtoul(v->text.c_str(), nullptr, 10);
                s.md5 = str_of(sk->get("md5sum"));
                if (const JValue *v = sk->get("mins"))
                    for (const JValue &e : v->arr) s.mins.push_back(strtoull(e.text.c_str(), nullptr, 10));
                if (const JValue *v = sk->get("weights"))
                    for (const JValue &e : v->arr) s.weights.push_back(strtod(e.text.c_str(), nullptr));
            }
            f->sigs.push_back(std::move(s));
        }
    }
    if (f->sigs.empty()) return fail(HULK_B200_EIO, std::string("no signatures found in supplied file: ") + path);
    if (f->klass != "hulk_sketch") return fail(HULK_B200_EIO, std::string("JSON not created by HULK: ") + path);
    if (f->version != HULK_B200_VERSION)
        return fail(HULK_B200_EIO, "the loaded sketch was created with a different version of HULK: " + f->version);
    for (const Sig &s : f->sigs) {
        if (s.md5.empty()) return fail(HULK_B200_EIO, "no MD5 was stored for a sketch: " + f->filename);
        char md5[33];
        hulk_b200_md5_mins(s.mins.data(), (uint32_t)s.mins.size(), md5);
        if (s.md5 != md5) return fail(HULK_B200_EIO, "md5sum mismatch: " + s.md5 + " vs. " + md5);
    }
    *out = f.release();
    return HULK_B200_OK;
}

// HULKdata.FindSketch (src/sketchio/sketchio.go:197-260)
int hulk_b200_sketch_find(const hulk_b200_sketch_file *f, uint32_t k, const char *algo, const uint64_t **mins,
                          const double **weights, uint32_t *s, char *err, uint64_t errcap) {
    auto fail = [&](const std::string &m) {
        if (err && errcap) snprintf(err, (size_t)errcap, "%s", m.c_str());
        return HULK_B200_EARG;
    };
    if (!f || !algo || !mins || !s) return fail("argument is NULL");
    const std::string a = algo;
    if (a != "histosketch" && a != "kmv" && a != "khf")
        return fail("specified algorithm (" + a + ") not found in the supplied sketch: " + f->filename);
    const Sig *hit = nullptr;
    unsigned n_algo = 0, n_hit = 0;
    for (const Sig &sg : f->sigs) {
        if (sg.algorithm != a) continue;
        n_algo++;
        if (sg.ksize == k) { hit = &sg; n_hit++; }
    }
    if (n_algo == 0) return fail("no sketches were produced using the " + a + " algorithm in file: " + f->filename);
    if (n_hit > 1) return fail("found " + std::to_string(n_hit) + " possible duplicate sketches in the supplied sketch file: " + f->filename);
    if (n_hit == 0) return fail("specified k-mer size (" + std::to_string(k) + ") not found in the supplied sketch file: " + f->filename);
    *mins = hit->mins.data();
    if (weights) *weights = hit->weights.size() == hit->mins.size() ? hit->weights.data() : nullptr;
    *s = (uint32_t)hit->mins.size();
    return HULK_B200_OK;
}

const char *hulk_b200_sketch_banner(const hulk_b200_sketch_file *f) { return f ? f->banner.c_str() : ""; }
void hulk_b200_sketch_free(hulk_b200_sketch_file *f) { delete f; }

}  // extern "C"
