// jf_format.cc — reader / writer of jellyfish "binary/sorted" hash files (see jf_format.hpp).
#include "jf_format.hpp"

#include <unistd.h>

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <istream>
#include <map>
#include <numeric>
#include <ostream>
#include <sstream>
#include <atomic>
#include <cstdlib>
#include <thread>
#include <utility>

namespace kat_host {
namespace {

// ---- a JSON value, just enough for the header (jsoncpp's terse writer output and anything a
// hand-edited header may contain: nested objects/arrays, strings with escapes, numbers) ----
struct JVal {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false;
    double d = 0;
    uint64_t u = 0; // exact value of a non-negative integer literal
    bool is_uint = false;
    std::string s;
    std::vector<JVal> a;
    std::map<std::string, JVal> o;
    const JVal *get(const std::string &k) const {
        auto it = o.find(k);
        return it == o.end() ? nullptr : &it->second;
    }
};
struct JParser {
    const char *p, *end;
    std::string err;
    void ws() {
        while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) p++;
    }
    bool fail(const char *m) {
        if (err.empty()) err = m;
        return false;
    }
    bool str(std::string &out) {
        if (p >= end || *p != '"') return fail("expected string");
        p++;
        while (p < end && *p != '"') {
            if (*p == '\\') {
                if (++p >= end) return fail("bad escape");
                switch (*p) {
                case 'n': out += '\n'; break;
                case 't': out += '\t'; break;
                case 'r': out += '\r'; break;
                case 'b': out += '\b'; break;
                case 'f': out += '\f'; break;
                case 'u': { // keep the low byte; headers are ASCII
                    if (end - p < 5) return fail("bad \\u escape");
                    out += (char)strtol(std::string(p + 1, p + 5).c_str(), nullptr, 16);
                    p += 4;
                    break;
                }
                default: out += *p;
                }
                p++;
            } else
                out += *p++;
        }
        if (p >= end) return fail("unterminated string");
        p++;
        return true;
    }
    bool value(JVal &v, int depth = 0) {
        if (depth > 32) return fail("nesting too deep");
        ws();
        if (p >= end) return fail("unexpected end");
        if (*p == '{') {
            v.kind = JVal::Obj;
            p++;
            ws();
            if (p < end && *p == '}') return p++, true;
            for (;;) {
                ws();
                std::string k;
                if (!str(k)) return false;
                ws();
                if (p >= end || *p != ':') return fail("expected ':'");
                p++;
                if (!value(v.o[k], depth + 1)) return false;
                ws();
                if (p < end && *p == ',') { p++; continue; }
                if (p < end && *p == '}') return p++, true;
                return fail("expected ',' or '}'");
            }
        }
        if (*p == '[') {
            v.kind = JVal::Arr;
            p++;
            ws();
            if (p < end && *p == ']') return p++, true;
            for (;;) {
                v.a.emplace_back();
                if (!value(v.a.back(), depth + 1)) return false;
                ws();
                if (p < end && *p == ',') { p++; continue; }
                if (p < end && *p == ']') return p++, true;
                return fail("expected ',' or ']'");
            }
        }
        if (*p == '"') {
            v.kind = JVal::Str;
            return str(v.s);
        }
        if (end - p >= 4 && !strncmp(p, "true", 4)) return v.kind = JVal::Bool, v.b = true, p += 4, true;
        if (end - p >= 5 && !strncmp(p, "false", 5)) return v.kind = JVal::Bool, v.b = false, p += 5, true;
        if (end - p >= 4 && !strncmp(p, "null", 4)) return v.kind = JVal::Null, p += 4, true;
        const char *q = p;
        while (q < end && (isdigit((unsigned char)*q) || *q == '-' || *q == '+' || *q == '.' || *q == 'e' || *q == 'E')) q++;
        if (q == p) return fail("unexpected character");
        const std::string num(p, q);
        v.kind = JVal::Num;
        v.d = strtod(num.c_str(), nullptr);
        v.is_uint = num.find_first_not_of("0123456789") == std::string::npos;
        if (v.is_uint) v.u = strtoull(num.c_str(), nullptr, 10);
        p = q;
        return true;
    }
};
uint64_t as_u64(const JVal *v, uint64_t dflt) {
    if (!v) return dflt;
    if (v->kind == JVal::Num) return v->is_uint ? v->u : (uint64_t)v->d;
    if (v->kind == JVal::Bool) return v->b;
    return dflt;
}
std::string as_str(const JVal *v) { return v && v->kind == JVal::Str ? v->s : std::string(); }

void json_string(std::string &out, const std::string &s) { // jsoncpp valueToQuotedString
    out += '"';
    for (unsigned char c : s) {
        switch (c) {
        case '"': out += "\\\""; break;
        case '\\': out += "\\\\"; break;
        case '\n': out += "\\n"; break;
        case '\t': out += "\\t"; break;
        case '\r': out += "\\r"; break;
        case '\b': out += "\\b"; break;
        case '\f': out += "\\f"; break;
        default:
            if (c < 0x20) {
                char b[8];
                snprintf(b, sizeof b, "\\u%04X", c);
                out += b;
            } else
                out += (char)c;
        }
    }
    out += '"';
}

} // namespace

bool jf_read_header(std::istream &is, JfHeader &h, std::string &err) {
    // generic_file_header::read: up to 9 decimal digits of length, then '{'
    std::string len;
    for (int i = 0; i < 9 && isdigit(is.peek()); i++) len += (char)is.get();
    if (is.peek() != '{') return err = "not a jellyfish hash: no JSON header", false;
    const unsigned long hlen = strtoul(len.c_str(), nullptr, 10);
    if (hlen < 2 || hlen > (64u << 20)) return err = "bad header length", false;
    std::string buf(hlen, '\0');
    is.read(&buf[0], (std::streamsize)hlen);
    if (!is.good()) return err = "truncated header", false;
    size_t e = buf.size();
    while (e > 0 && buf[e - 1] == '\0') e--;
    JParser jp{buf.data(), buf.data() + e, {}};
    JVal root;
    if (!jp.value(root) || root.kind != JVal::Obj) return err = "header JSON: " + (jp.err.empty() ? std::string("not an object") : jp.err), false;
    h = JfHeader();
    h.offset = 9 + hlen;
    h.alignment = (unsigned)as_u64(root.get("alignment"), 0);
    h.canonical = as_u64(root.get("canonical"), 0) != 0;
    h.counter_len = (unsigned)as_u64(root.get("counter_len"), 0);
    h.format = as_str(root.get("format"));
    h.key_len = (unsigned)as_u64(root.get("key_len"), 0);
    h.val_len = (unsigned)as_u64(root.get("val_len"), 0);
    h.max_reprobe = (unsigned)as_u64(root.get("max_reprobe"), 0);
    h.size = as_u64(root.get("size"), 0);
    h.exe_path = as_str(root.get("exe_path"));
    h.hostname = as_str(root.get("hostname"));
    h.pwd = as_str(root.get("pwd"));
    h.time = as_str(root.get("time"));
    if (const JVal *c = root.get("cmdline"))
        if (c->kind == JVal::Arr)
            for (auto &x : c->a) h.cmdline += (h.cmdline.empty() ? "" : " ") + as_str(&x);
    if (const JVal *m = root.get("matrix1")) {
        h.matrix_r = (unsigned)as_u64(m->get("r"), 0);
        h.matrix_c = (unsigned)as_u64(m->get("c"), 0);
        if (const JVal *cols = m->get("columns"))
            for (auto &x : cols->a) h.matrix_columns.push_back(as_u64(&x, 0));
        h.matrix_columns.resize(h.matrix_c, 0);
    }
    if (const JVal *r = root.get("reprobes"))
        for (auto &x : r->a) h.reprobes.push_back(as_u64(&x, 0));
    return true;
}

void jf_write_header(std::ostream &os, JfHeader &h) {
    if (h.hostname.empty()) {
        char b[256] = "";
        if (gethostname(b, sizeof b - 1) == 0) h.hostname = b;
    }
    if (h.pwd.empty()) {
        char b[4096];
        if (getcwd(b, sizeof b)) h.pwd = b;
    }
    if (h.time.empty()) {
        time_t t = time(nullptr);
        char b[64];
        struct tm tmv;
        localtime_r(&t, &tmv);
        strftime(b, sizeof b, "%a %b %e %H:%M:%S %Y", &tmv); // ctime() layout, as get_localtime()
        h.time = b;
    }
    if (h.exe_path.empty()) {
        char b[4096];
        const ssize_t l = readlink("/proc/self/exe", b, sizeof b - 1);
        if (l > 0) h.exe_path.assign(b, (size_t)l);
    }
    // Json::FastWriter: one line, members in key order
    std::string j = "{\"alignment\":" + std::to_string(h.alignment);
    j += std::string(",\"canonical\":") + (h.canonical ? "true" : "false");
    j += ",\"counter_len\":" + std::to_string(h.counter_len);
    j += ",\"exe_path\":";
    json_string(j, h.exe_path);
    j += ",\"format\":";
    json_string(j, h.format);
    j += ",\"hostname\":";
    json_string(j, h.hostname);
    j += ",\"key_len\":" + std::to_string(h.key_len);
    j += ",\"matrix1\":{\"c\":" + std::to_string(h.matrix_c) + ",\"columns\":[";
    for (size_t i = 0; i < h.matrix_columns.size(); i++) j += (i ? "," : "") + std::to_string(h.matrix_columns[i]);
    j += "],\"r\":" + std::to_string(h.matrix_r) + "}";
    j += ",\"max_reprobe\":" + std::to_string(h.max_reprobe);
    j += ",\"pwd\":";
    json_string(j, h.pwd);
    j += ",\"reprobes\":[";
    for (size_t i = 0; i < h.reprobes.size(); i++) j += (i ? "," : "") + std::to_string(h.reprobes[i]);
    j += "],\"size\":" + std::to_string(h.size);
    j += ",\"time\":";
    json_string(j, h.time);
    j += ",\"val_len\":" + std::to_string(h.val_len) + "}";
    // generic_file_header::write: 9-digit length (JSON + padding), JSON, zero padding to `alignment`
    size_t hlen = j.size(), pad = 0;
    if (h.alignment > 0 && (9 + j.size()) % h.alignment) pad = h.alignment - (9 + j.size()) % h.alignment;
    hlen += pad;
    char lenbuf[16];
    snprintf(lenbuf, sizeof lenbuf, "%09zu", hlen);
    os.write(lenbuf, 9);
    os.write(j.data(), (std::streamsize)j.size());
    for (size_t i = 0; i < pad; i++) os.put('\0');
    h.offset = 9 + hlen;
}

uint64_t jf_position(const JfHeader &h, uint64_t key) {
    uint64_t res = 0;
    const unsigned c = h.matrix_c;
    for (unsigned j = 0; j < c && j < 64; j++)
        if ((key >> j) & 1) res ^= h.matrix_columns[c - 1 - j];
    return h.size ? res & (h.size - 1) : res;
}

uint64_t jf_position(const JfHeader &h, uint64_t key_hi, uint64_t key_lo) {
    uint64_t res = 0;
    const unsigned c = h.matrix_c;
    for (unsigned j = 0; j < c && j < 128; j++)
        if (((j < 64 ? key_lo >> j : key_hi >> (j - 64)) & 1) != 0) res ^= h.matrix_columns[c - 1 - j];
    return h.size ? res & (h.size - 1) : res;
}

bool jf_read_records_wide(std::istream &is, const JfHeader &h, size_t batch,
                          const std::function<void(const uint64_t *, const uint64_t *, const uint64_t *, size_t)> &sink, std::string &err) {
    const unsigned kb = h.key_bytes(), cb = h.counter_len, rb = kb + cb;
    if (kb == 0 || kb > 16 || cb == 0 || cb > 8) return err = "unsupported record layout (key_len " + std::to_string(h.key_len) + " bits, counter_len " + std::to_string(cb) + ")", false;
    if (batch == 0) batch = 1;
    std::vector<unsigned char> raw(batch * rb);
    std::vector<uint64_t> hi(batch), lo(batch), counts(batch);
    for (;;) {
        is.read(reinterpret_cast<char *>(raw.data()), (std::streamsize)raw.size());
        const size_t got = (size_t)is.gcount();
        if (got % rb) return err = "Size of database must be a multiple of the length of a record (" + std::to_string(rb) + ")", false;
        const size_t n = got / rb;
        for (size_t i = 0; i < n; i++) {
            uint64_t w[2] = {0, 0}, c = 0;
            memcpy(w, &raw[i * rb], kb); // little-endian words of mer_dna::data(): the low word first
            memcpy(&c, &raw[i * rb + kb], cb);
            lo[i] = w[0], hi[i] = w[1], counts[i] = c;
        }
        if (n) sink(hi.data(), lo.data(), counts.data(), n);
        if (got < raw.size()) break;
    }
    return true;
}

bool jf_read_records(std::istream &is, const JfHeader &h, size_t batch,
                     const std::function<void(const uint64_t *, const uint64_t *, size_t)> &sink, std::string &err) {
    const unsigned kb = h.key_bytes(), cb = h.counter_len, rb = kb + cb;
    if (kb == 0 || kb > 8 || cb == 0 || cb > 8) return err = "unsupported record layout (key_len " + std::to_string(h.key_len) + " bits, counter_len " + std::to_string(cb) + ")", false;
    if (batch == 0) batch = 1;
    std::vector<unsigned char> raw(batch * rb);
    std::vector<uint64_t> keys(batch), counts(batch);
    for (;;) {
        is.read(reinterpret_cast<char *>(raw.data()), (std::streamsize)raw.size());
        const size_t got = (size_t)is.gcount();
        if (got % rb) // HashLoader::loadHash refuses such a file (lib/src/jellyfish_helper.cc:162-167)
            return err = "Size of database must be a multiple of the length of a record (" + std::to_string(rb) + ")", false;
        const size_t n = got / rb;
        for (size_t i = 0; i < n; i++) {
            uint64_t k = 0, c = 0;
            memcpy(&k, &raw[i * rb], kb);      // little-endian words of mer_dna::data()
            memcpy(&c, &raw[i * rb + kb], cb);
            keys[i] = k;
            counts[i] = c;
        }
        if (n) sink(keys.data(), counts.data(), n);
        if (got < raw.size()) break;
    }
    return true;
}

// Sort (key, count) records by key.  Large tables (a 10 Gbp data set dumps 1.2 G records) are
// split by the top key bits into 4096 buckets in one counting pass, and the buckets are sorted
// by all host threads; small ones go through one std::sort.
static void sort_by_key(std::vector<uint64_t> &keys, std::vector<uint64_t> &counts, unsigned k) {
    const size_t n = keys.size();
    size_t parallel_from = (size_t)1 << 20;
    if (const char *e = getenv("KAT_SORT_PARALLEL_MIN")) parallel_from = (size_t)strtoull(e, nullptr, 10);
    std::vector<std::pair<uint64_t, uint64_t>> rec(n);
    if (n < parallel_from) {
        for (size_t i = 0; i < n; i++) rec[i] = {keys[i], counts[i]};
        std::sort(rec.begin(), rec.end());
    } else {
        const unsigned key_bits = 2 * k, bb = std::min(12u, key_bits), shift = key_bits - bb;
        const size_t nb = (size_t)1 << bb;
        std::vector<size_t> start(nb + 1, 0);
        for (size_t i = 0; i < n; i++) start[(keys[i] >> shift) + 1]++;
        for (size_t b = 0; b < nb; b++) start[b + 1] += start[b];
        std::vector<size_t> at(start.begin(), start.end() - 1);
        for (size_t i = 0; i < n; i++) rec[at[keys[i] >> shift]++] = {keys[i], counts[i]};
        unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), 64u));
        std::vector<std::thread> th;
        std::atomic<size_t> next{0};
        auto work = [&]() {
            for (size_t b; (b = next.fetch_add(1)) < nb;) std::sort(rec.begin() + (std::ptrdiff_t)start[b], rec.begin() + (std::ptrdiff_t)start[b + 1]);
        };
        for (unsigned t = 1; t < nt; t++) th.emplace_back(work);
        work();
        for (auto &x : th) x.join();
    }
    for (size_t i = 0; i < n; i++) keys[i] = rec[i].first, counts[i] = rec[i].second;
}

static void fill_dump_header(JfHeader &h, unsigned k, bool canonical, uint64_t n);
void jf_prepare_dump(JfHeader &h, unsigned k, bool canonical, uint64_t n, std::vector<uint64_t> &keys, std::vector<uint64_t> &counts) {
    fill_dump_header(h, k, canonical, n);
    // ascending (position, key) == ascending key under this matrix
    sort_by_key(keys, counts, k);
}
void jf_prepare_dump_wide(JfHeader &h, unsigned k, bool canonical, uint64_t n, std::vector<uint64_t> &keys_hi, std::vector<uint64_t> &keys_lo,
                          std::vector<uint64_t> &counts) {
    fill_dump_header(h, k, canonical, n);
    struct Rec {
        uint64_t hi, lo, count;
    };
    std::vector<Rec> rec(keys_lo.size());
    for (size_t i = 0; i < rec.size(); i++) rec[i] = {keys_hi[i], keys_lo[i], counts[i]};
    std::sort(rec.begin(), rec.end(), [](const Rec &a, const Rec &b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); });
    for (size_t i = 0; i < rec.size(); i++) keys_hi[i] = rec[i].hi, keys_lo[i] = rec[i].lo, counts[i] = rec[i].count;
}
static void fill_dump_header(JfHeader &h, unsigned k, bool canonical, uint64_t n) {
    h.format = "binary/sorted";
    h.key_len = 2 * k;
    h.canonical = canonical;
    h.counter_len = 4; // InputHandler::count: header->counter_len(4) (lib/src/input_handler.cc:198)
    h.val_len = 7;
    h.max_reprobe = 126;
    h.alignment = 8;
    unsigned lsize = 1;
    while (lsize < 63 && (1ULL << lsize) < 2 * n) lsize++;
    h.size = 1ULL << lsize;
    h.matrix_r = lsize;
    h.matrix_c = h.key_len;
    h.matrix_columns.assign(h.matrix_c, 0);
    const unsigned shift = h.key_len > lsize ? h.key_len - lsize : 0;
    for (unsigned j = shift; j < h.matrix_c; j++) h.matrix_columns[h.matrix_c - 1 - j] = 1ULL << (j - shift);
    h.reprobes.resize(h.max_reprobe + 1);
    for (unsigned i = 0; i <= h.max_reprobe; i++) h.reprobes[i] = i == 0 ? 1 : (uint64_t)i * (i + 1) / 2; // JF/lib/storage.cc quadratic_reprobes
}

void jf_write_file_wide(std::ostream &os, JfHeader &h, const std::vector<uint64_t> &keys_hi, const std::vector<uint64_t> &keys_lo,
                        const std::vector<uint64_t> &counts) {
    jf_write_header(os, h);
    const unsigned kb = h.key_bytes(), cb = h.counter_len, rb = kb + cb;
    const uint64_t vmax = cb >= 8 ? ~0ULL : ((1ULL << (8 * cb)) - 1);
    std::vector<unsigned char> buf;
    buf.reserve((size_t)rb << 16);
    for (size_t i = 0; i < keys_lo.size(); i++) {
        const uint64_t w[2] = {keys_lo[i], keys_hi[i]}, c = std::min(counts[i], vmax);
        const size_t at = buf.size();
        buf.resize(at + rb);
        memcpy(&buf[at], w, kb);
        memcpy(&buf[at + kb], &c, cb);
        if (buf.size() >= ((size_t)rb << 16)) {
            os.write(reinterpret_cast<const char *>(buf.data()), (std::streamsize)buf.size());
            buf.clear();
        }
    }
    if (!buf.empty()) os.write(reinterpret_cast<const char *>(buf.data()), (std::streamsize)buf.size());
}

void jf_write_file(std::ostream &os, JfHeader &h, const std::vector<uint64_t> &keys, const std::vector<uint64_t> &counts) {
    jf_write_header(os, h);
    const unsigned kb = h.key_bytes(), cb = h.counter_len, rb = kb + cb;
    const uint64_t vmax = cb >= 8 ? ~0ULL : ((1ULL << (8 * cb)) - 1);
    std::vector<unsigned char> buf;
    buf.reserve((size_t)rb << 16);
    for (size_t i = 0; i < keys.size(); i++) {
        const uint64_t k = keys[i], c = std::min(counts[i], vmax);
        const size_t at = buf.size();
        buf.resize(at + rb);
        memcpy(&buf[at], &k, kb);
        memcpy(&buf[at + kb], &c, cb);
        if (buf.size() >= ((size_t)rb << 16)) {
            os.write(reinterpret_cast<const char *>(buf.data()), (std::streamsize)buf.size());
            buf.clear();
        }
    }
    if (!buf.empty()) os.write(reinterpret_cast<const char *>(buf.data()), (std::streamsize)buf.size());
}

} // namespace kat_host
