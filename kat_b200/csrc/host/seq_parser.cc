// seq_parser.cc — host-side FASTA/FASTQ reader of the `kat` front-end.
//
// Produces, per input file, exactly the base stream the reference hands to its
// k-mer iterator: jellyfish's mer_overlap_sequence_parser
// (JFI/mer_overlap_sequence_parser.hpp:132-289) as KAT configures it
// (4096-byte buffers, lib/src/jellyfish_helper.cc:233).  Record rules:
//   * first byte '>' = FASTA, '@' = FASTQ, anything else "Unsupported format" (:171-185);
//   * header lines skipped, '\n' stripped (multi-line records fine), one 'N'
//     between consecutive records (:202,234);
//   * FASTQ qualities are skipped by LENGTH, so '@' inside them is harmless (:274-289);
//   * `--5ptrim X` drops the first X characters of every record (:248-253) — and,
//     a reference quirk kept for bit-exactness, of every 4 KiB buffer of a FASTA
//     record (`newread` is re-armed per buffer, :197);
//   * a malformed FASTQ record (e.g. no newline after the last quality line)
//     raises "Invalid fastq sequence" (:288), which the reference swallows
//     (JFI/cooperative_pool2.hpp:249-260): the buffer being filled is dropped and
//     the stream ends.  We reproduce that and report it.
// The (k-1)-base seam the reference carries between buffers only exists so
// that no k-mer is lost at a buffer edge; the stream returned here is simply
// contiguous.
#include "seq_parser.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>

namespace kat_host {

bool read_file(const std::string &path, std::vector<unsigned char> &out, std::string &err) {
    // Plain regular files are read directly (the bytes zlib would pass through untouched, at
    // page-cache speed); everything else goes through zlib as the reference's stream_manager
    // does (JFI/stream_manager.hpp:133-145).
    {
        const int fd = open(path.c_str(), O_RDONLY);
        struct stat st;
        unsigned char magic[2] = {0, 0};
        if (fd >= 0 && fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && pread(fd, magic, 2, 0) >= 0 &&
            !(magic[0] == 0x1f && magic[1] == 0x8b)) {
            out.resize((size_t)st.st_size);
            size_t got = 0;
            while (got < out.size()) {
                const ssize_t r = pread(fd, out.data() + got, std::min<size_t>(out.size() - got, (size_t)1 << 30), (off_t)got);
                if (r < 0) {
                    err = "Read error on input file: " + path;
                    close(fd);
                    return false;
                }
                if (r == 0) break;
                got += (size_t)r;
            }
            out.resize(got);
            close(fd);
            return true;
        }
        if (fd >= 0) close(fd);
    }
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) {
        err = "Could not open input file: " + path;
        return false;
    }
    gzbuffer(f, 1 << 20);
    out.clear();
    std::vector<unsigned char> buf(8 << 20);
    for (;;) {
        const int n = gzread(f, buf.data(), (unsigned)buf.size());
        if (n < 0) {
            err = "Read error on input file: " + path;
            gzclose(f);
            return false;
        }
        if (n == 0) break;
        out.insert(out.end(), buf.begin(), buf.begin() + n);
    }
    gzclose(f);
    return true;
}

namespace {

// The slice of std::istream behaviour the reference parser relies on.
class CharStream {
public:
    CharStream(const unsigned char *p, size_t n) : p_(p), n_(n) {}
    bool good() const { return !eof_ && !fail_; }
    bool ok() const { return !fail_; } // operator bool
    int peek() {                       // sentry: peeking a stream that is not good() fails it
        if (!good()) {
            fail_ = true;
            return -1;
        }
        if (pos_ >= n_) {
            eof_ = true;
            return -1;
        }
        return p_[pos_];
    }
    size_t ignore(size_t nmax, int delim) {
        if (!good()) {
            fail_ = true;
            return 0;
        }
        size_t cnt = 0;
        while (cnt < nmax) {
            if (pos_ >= n_) {
                eof_ = true;
                break;
            }
            const int c = p_[pos_++];
            ++cnt;
            if (c == delim) break;
        }
        return cnt;
    }
    void ignore_line() { ignore(static_cast<size_t>(-1), '\n'); }
    void skip_newlines() {
        while (peek() == '\n') ++pos_;
    }
    // istream::get(char*, n): at most n-1 characters, stops before '\n'; nothing read = failbit
    size_t get_line_part(unsigned char *dst, size_t n) {
        if (!good()) {
            fail_ = true;
            return 0;
        }
        size_t cnt = 0;
        while (cnt + 1 < n) {
            if (pos_ >= n_) {
                eof_ = true;
                break;
            }
            const unsigned char c = p_[pos_];
            if (c == '\n') break;
            dst[cnt++] = c;
            ++pos_;
        }
        if (cnt == 0) fail_ = true;
        return cnt;
    }

private:
    const unsigned char *p_;
    size_t n_, pos_ = 0;
    bool eof_ = false, fail_ = false;
};

constexpr size_t kBuf = 4096; // lib/src/jellyfish_helper.cc:233

size_t read_sequence(CharStream &is, size_t read, unsigned char *buf, int stop, unsigned trim5p) {
    size_t nread = read;
    if (trim5p > 0) {
        is.skip_newlines();
        is.ignore(trim5p, -2);
    }
    while (is.ok() && nread < kBuf - 1 && is.peek() != stop) {
        is.skip_newlines();
        nread += is.get_line_part(buf + nread, kBuf - nread);
        is.skip_newlines();
    }
    return nread - read;
}

bool skip_quals(CharStream &is, size_t read_len) {
    is.ignore_line();
    size_t quals = 0;
    while (is.good() && quals < read_len) {
        is.skip_newlines();
        quals += is.ignore(read_len - quals + 1, '\n');
        if (is.ok()) ++read_len;
    }
    is.skip_newlines();
    if (quals != read_len) return false;
    const int c = is.peek();
    return c == '@' || c == -1;
}

// The reference's parser, character by character (see the header of this file).
ParseStatus exact_impl(const unsigned char *data, size_t n, int k, unsigned trim5p, std::string &out, size_t *fastq_carry, LastBuffer *last) {
    CharStream is(data, n);
    const size_t out_base = out.size();
    if (last) *last = LastBuffer{};
    const int first = is.peek();
    if (first == -1) return ParseStatus::Ok; // empty file: nothing to count
    if (first != '>' && first != '@') return ParseStatus::UnsupportedFormat;
    const bool fastq = first == '@';
    is.ignore_line();
    unsigned char buf[kBuf + 8];
    const size_t seam = static_cast<size_t>(k - 1);
    bool have_seam = false;
    size_t seq_len = fastq_carry ? *fastq_carry : 0; // not reset between the files of a group (see seq_parser.hpp)
    out.reserve(out.size() + n / 2 + 16);
    while (is.good()) { // one pass = one 4 KiB buffer of the reference
        size_t read = have_seam ? seam : 0;
        const size_t start = read;
        if (!fastq) {
            bool newread = true;
            while (is.good() && read < kBuf - k - 1) {
                read += read_sequence(is, read, buf, '>', newread ? trim5p : 0);
                if (is.peek() == '>') {
                    buf[read++] = 'N';
                    is.ignore_line();
                    newread = true;
                } else
                    newread = false;
            }
        } else {
            while (is.good() && read < kBuf - k - 1) {
                const size_t nread = read_sequence(is, read, buf, '+', seq_len == 0 ? trim5p : 0);
                read += nread;
                seq_len += nread;
                if (is.peek() == '+') {
                    if (!skip_quals(is, seq_len + trim5p)) return ParseStatus::InvalidFastq;
                    if (is.good()) {
                        buf[read++] = 'N';
                        is.ignore_line();
                    }
                    seq_len = 0;
                }
            }
        }
        if (last) last->at = out.size() - out_base, last->seam = start; // the pass that runs into the end of the file is the last one
        out.append(reinterpret_cast<const char *>(buf + start), read - start);
        have_seam = read >= seam;
    }
    if (fastq_carry) *fastq_carry = seq_len;
    return ParseStatus::Ok;
}


// ---- fast paths -------------------------------------------------------------------------------
// The character-by-character emulation above defines the result; it runs at ~0.25 GB/s.  For
// the two regular layouts below the stream it produces can be written down directly, record by
// record, with memchr — and in parallel, because record boundaries can be found from any
// offset.  Every precondition is CHECKED while scanning; the first violation anywhere makes the
// caller fall back to the exact emulation for the whole file.
//
//   FASTQ, 4 lines per record, every line '\n'-terminated (also the last), no empty line:
//     '@' header / sequence (length > trim5p, first byte not '@' or '+', no '+' inside) /
//     '+' line / quality of the same length.
//     Stream = seq_1[trim5p:] 'N' seq_2[trim5p:] 'N' ... seq_last[trim5p:]
//     (read_sequence + skip_quals, JFI/mer_overlap_sequence_parser.hpp:218-289: the 'N' is only
//     appended while the stream is still good, i.e. not after the last record).
//   FASTA, trim5p == 0, '>' only as the first byte of header lines, no empty line:
//     Stream = the sequence lines of each record joined, records separated by 'N' (:189-216).
struct Piece {
    size_t begin = 0, end = 0;   // byte range of whole records
    size_t out_bytes = 0;        // stream bytes incl. one 'N' after every record
    bool ok = true;
};

inline const unsigned char *find_nl(const unsigned char *p, const unsigned char *e) {
    return p < e ? static_cast<const unsigned char *>(memchr(p, '\n', (size_t)(e - p))) : nullptr;
}

// Validate the FASTQ records in [p.begin, p.end) and, when dst != nullptr, emit them.
void fastq_piece(const unsigned char *data, Piece &p, unsigned trim, char *dst) {
    const unsigned char *q = data + p.begin, *e = data + p.end;
    size_t outb = 0;
    while (q < e) {
        const unsigned char *h_nl = (*q == '@') ? find_nl(q, e) : nullptr;
        const unsigned char *s = h_nl ? h_nl + 1 : nullptr;
        const unsigned char *s_nl = s ? find_nl(s, e) : nullptr;
        const unsigned char *pl = s_nl ? s_nl + 1 : nullptr;
        const unsigned char *p_nl = (pl && pl < e && *pl == '+') ? find_nl(pl, e) : nullptr;
        const unsigned char *ql = p_nl ? p_nl + 1 : nullptr;
        const unsigned char *q_nl = ql ? find_nl(ql, e) : nullptr;
        if (!q_nl) { p.ok = false; return; }
        const size_t slen = (size_t)(s_nl - s);
        if (slen <= trim || slen != (size_t)(q_nl - ql) || *s == '@' || *s == '+' || memchr(s, '+', slen) != nullptr) { p.ok = false; return; }
        if (dst) {
            memcpy(dst + outb, s + trim, slen - trim);
            dst[outb + slen - trim] = 'N';
        }
        outb += slen - trim + 1;
        q = q_nl + 1;
    }
    p.out_bytes = outb;
}

// Same for FASTA records (each piece starts at a '>' line).
void fasta_piece(const unsigned char *data, size_t n_total, Piece &p, char *dst) {
    const unsigned char *q = data + p.begin, *e = data + p.end;
    size_t outb = 0;
    while (q < e) {
        if (*q != '>') { p.ok = false; return; }
        const unsigned char *nl = find_nl(q, e);
        if (!nl) { p.ok = false; return; } // header without newline: leave it to the emulation
        q = nl + 1;
        while (q < e && *q != '>') { // sequence lines
            const unsigned char *l_nl = find_nl(q, e);
            const unsigned char *l_end = l_nl ? l_nl : e;
            const size_t len = (size_t)(l_end - q);
            if (len == 0 || memchr(q, '>', len) != nullptr) { p.ok = false; return; }
            if (!l_nl && (size_t)(e - data) != n_total) { p.ok = false; return; } // only the file's last line may lack '\n'
            if (dst) memcpy(dst + outb, q, len);
            outb += len;
            q = l_nl ? l_nl + 1 : e;
        }
        if (dst) dst[outb] = 'N';
        outb += 1;
    }
    p.out_bytes = outb;
}

// Cut [0, n) into pieces of whole records, starting near multiples of `target` bytes.
bool split_pieces(const unsigned char *data, size_t n, bool fastq, size_t target, std::vector<Piece> &pieces) {
    std::vector<size_t> starts{0};
    for (size_t want = target; want < n; want += target) {
        const unsigned char *e = data + n;
        const unsigned char *nl = find_nl(data + want, e);
        size_t found = 0;
        for (int tries = 0; nl && tries < 64; tries++) { // candidate record starts: the lines after `want`
            const unsigned char *c = nl + 1;
            if (c >= e) break;
            if (!fastq) {
                if (*c == '>') { found = (size_t)(c - data); break; }
            } else if (*c == '@') { // a record start has a '+' line two lines down and a sequence line that does not start with '@'
                const unsigned char *l1 = find_nl(c, e);
                const unsigned char *l2 = l1 ? find_nl(l1 + 1, e) : nullptr;
                if (l1 && l2 && l2 + 1 < e && l2[1] == '+' && l1[1] != '@' && l1[1] != '+') { found = (size_t)(c - data); break; }
            }
            nl = find_nl(c, e);
        }
        if (!fastq && !found) { // long FASTA records: look further for the next header
            const unsigned char *c = data + want;
            while ((nl = find_nl(c, e)) != nullptr) {
                if (nl + 1 < e && nl[1] == '>') { found = (size_t)(nl + 1 - data); break; }
                c = nl + 1;
            }
        }
        if (found > starts.back()) {
            starts.push_back(found);
            want = std::max(want, found);
        } else if (fastq && !found)
            return false; // could not find a record start: not the regular layout
    }
    pieces.clear();
    for (size_t i = 0; i < starts.size(); i++) {
        Piece p;
        p.begin = starts[i];
        p.end = i + 1 < starts.size() ? starts[i + 1] : n;
        pieces.push_back(p);
    }
    return true;
}

bool parse_regular_layout(const unsigned char *data, size_t n, bool fastq, unsigned trim, std::string &out) {
    if (n == 0) return false;
    if (fastq && data[n - 1] != '\n') return false;
    if (!fastq && trim != 0) return false; // per-buffer trim quirk of FASTA records: emulation only
    size_t target = (size_t)16 << 20;
    if (const char *e = getenv("KAT_PARSE_PIECE_BYTES")) target = std::max<size_t>(64, strtoull(e, nullptr, 10));
    unsigned n_threads = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    if (const char *e = getenv("KAT_PARSE_THREADS")) n_threads = std::max(1, atoi(e));
    std::vector<Piece> pieces;
    if (!split_pieces(data, n, fastq, target, pieces)) return false;
    auto run = [&](bool emit, char *base, const std::vector<size_t> *offs) {
        auto work = [&](unsigned t) {
            for (size_t i = t; i < pieces.size(); i += n_threads) {
                char *dst = emit ? base + (*offs)[i] : nullptr;
                if (fastq) fastq_piece(data, pieces[i], trim, dst);
                else fasta_piece(data, n, pieces[i], dst);
            }
        };
        const unsigned nt = (unsigned)std::min<size_t>(n_threads, pieces.size());
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nt; t++) th.emplace_back(work, t);
        work(0);
        for (auto &x : th) x.join();
        if (nt < n_threads) // (pieces.size() < n_threads: the strided loop above already covered everything)
            return;
    };
    run(false, nullptr, nullptr); // pass 1: validate, size
    std::vector<size_t> offs(pieces.size());
    size_t total = 0;
    for (size_t i = 0; i < pieces.size(); i++) {
        if (!pieces[i].ok) return false;
        offs[i] = total;
        total += pieces[i].out_bytes;
    }
    if (total == 0) return false;
    const size_t at = out.size();
    out.resize(at + total);
    run(true, &out[at], &offs); // pass 2: emit
    out.resize(at + total - 1); // no 'N' after the last record
    return true;
}

} // namespace

static bool g_last_fast = false;
bool last_parse_was_fast() { return g_last_fast; }

ParseStatus parse_sequence_file(const unsigned char *data, size_t n, int k, unsigned trim5p, std::string &out, size_t *fastq_carry) {
    // a FASTQ file that inherits a non-zero seq_len is not what the regular layout describes: emulation only
    // (a regular file leaves seq_len at 0, so the carry is untouched by the fast path)
    const bool carried = fastq_carry && *fastq_carry != 0 && n > 0 && data[0] == '@';
    g_last_fast = !carried && n > 0 && (data[0] == '>' || data[0] == '@') && !getenv("KAT_PARSE_EXACT_ONLY") &&
                  parse_regular_layout(data, n, data[0] == '@', trim5p, out);
    if (g_last_fast) return ParseStatus::Ok;
    return exact_impl(data, n, k, trim5p, out, fastq_carry, nullptr);
}

ParseStatus parse_sequence_file_exact(const unsigned char *data, size_t n, int k, unsigned trim5p, std::string &out, size_t *fastq_carry, LastBuffer *last) {
    g_last_fast = false;
    return exact_impl(data, n, k, trim5p, out, fastq_carry, last);
}

int first_byte_of_file(const std::string &path) {
    struct stat st;
    if (stat(path.c_str(), &st) != 0 || !S_ISREG(st.st_mode)) return -2; // a pipe must not be read twice
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) return -1;
    unsigned char c = 0;
    const int n = gzread(f, &c, 1);
    gzclose(f);
    return n == 1 ? (int)c : -1;
}

int stream_regular_file(const std::string &path, unsigned trim5p, size_t block_bytes,
                        const std::function<bool(const char *, size_t)> &sink, std::string &err) {
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) return 0;
    struct stat st;
    unsigned char magic[2] = {0, 0};
    if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode) || pread(fd, magic, 2, 0) < 1 || (magic[0] == 0x1f && magic[1] == 0x8b) ||
        (magic[0] != '@' && magic[0] != '>')) {
        close(fd);
        return 0;
    }
    const bool fastq = magic[0] == '@';
    if (!fastq && trim5p != 0) { // per-buffer trim quirk: emulation only
        close(fd);
        return 0;
    }
    block_bytes = std::max<size_t>(block_bytes, 256);
    std::vector<unsigned char> buf;
    std::string bases;
    size_t file_off = 0, emitted = 0;
    const size_t file_size = (size_t)st.st_size;
    int rc = 1;
    while (file_off < file_size || !buf.empty()) {
        // top the buffer up by one block
        const size_t have = buf.size(), want = std::min(block_bytes, file_size - file_off);
        buf.resize(have + want);
        size_t got = 0;
        while (got < want) {
            const ssize_t r = pread(fd, buf.data() + have + got, want - got, (off_t)(file_off + got));
            if (r <= 0) break;
            got += (size_t)r;
        }
        if (got != want) {
            err = "Read error on input file: " + path;
            rc = emitted ? -1 : 0;
            break;
        }
        file_off += want;
        const bool last = file_off >= file_size;
        // the records that are complete in the buffer end where the last record START in it begins
        size_t boundary = buf.size();
        if (!last) {
            boundary = 0;
            const unsigned char *b = buf.data(), *e = b + buf.size();
            size_t p = buf.size();
            if (!fastq) { // the last header line: '>' is only allowed as the first byte of a line
                for (const void *q = memrchr(b, '>', buf.size()); q != nullptr;) {
                    const size_t at = (size_t)(static_cast<const unsigned char *>(q) - b);
                    if (at > 0 && b[at - 1] == '\n') {
                        boundary = at;
                        break;
                    }
                    q = at > 0 ? memrchr(b, '>', at) : nullptr;
                }
            } else {
                for (int lines = 0; p > 0 && lines < 64; lines++) { // walk line starts backwards: a record start is at most 4 lines away
                    const void *nl = p >= 2 ? memrchr(b, '\n', p - 1) : nullptr; // newline before the line that ends at p
                    const size_t ls = nl ? (size_t)(static_cast<const unsigned char *>(nl) - b) + 1 : 0;
                    const unsigned char *c = b + ls;
                    if (ls > 0 && *c == '@') {
                        const unsigned char *l1 = find_nl(c, e);
                        const unsigned char *l2 = l1 ? find_nl(l1 + 1, e) : nullptr;
                        if (l1 && l2 && l2 + 1 < e && l2[1] == '+' && l1[1] != '@' && l1[1] != '+') { boundary = ls; break; }
                    }
                    if (ls == 0) break;
                    p = ls;
                }
            }
            if (boundary == 0) { // no record start in a whole block (a giant record, or not the regular layout)
                if (buf.size() > 8 * block_bytes + (64u << 20)) {
                    rc = emitted ? -1 : 0;
                    break;
                }
                continue; // read more
            }
        }
        bases.clear();
        if (boundary > 0 && !parse_regular_layout(buf.data(), boundary, fastq, trim5p, bases)) {
            rc = emitted ? -1 : 0;
            break;
        }
        if (!bases.empty()) {
            if (!sink(bases.data(), bases.size())) {
                rc = -1;
                break;
            }
            emitted++;
        }
        buf.erase(buf.begin(), buf.begin() + (std::ptrdiff_t)boundary);
        if (last) {
            if (!buf.empty()) rc = emitted ? -1 : 0; // (cannot happen: the last block is parsed to its end)
            break;
        }
    }
    close(fd);
    return rc;
}

// First FASTQ record start at or after `from` (the line after a '\n' that starts with '@', whose
// next line starts with neither '@' nor '+' and whose line after that starts with '+'); `n` if none.
static size_t next_fastq_record(const unsigned char *data, size_t n, size_t from) {
    if (from == 0) return 0;
    const unsigned char *e = data + n;
    const unsigned char *nl = find_nl(data + from - 1, e);
    for (int tries = 0; nl && tries < 64; tries++) {
        const unsigned char *c = nl + 1;
        if (c >= e) break;
        if (*c == '@') {
            const unsigned char *l1 = find_nl(c, e);
            const unsigned char *l2 = l1 ? find_nl(l1 + 1, e) : nullptr;
            if (l1 && l2 && l2 + 1 < e && l2[1] == '+' && l1[1] != '@' && l1[1] != '+') return (size_t)(c - data);
        }
        nl = find_nl(c, e);
    }
    return n;
}

int stream_fastq_raw(const std::string &path, size_t block_bytes, int n_parts,
                     const std::function<int(const char *const *parts, const size_t *lens)> &sink, std::string &err) {
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) return 0;
    struct stat st;
    unsigned char magic[2] = {0, 0};
    if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode) || st.st_size < 8 || pread(fd, magic, 2, 0) < 1 || magic[0] != '@') {
        close(fd);
        return 0;
    }
    const size_t n = (size_t)st.st_size;
    void *map = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (map == MAP_FAILED) return 0;
    const unsigned char *data = static_cast<const unsigned char *>(map);
    int rc = 1;
    if (data[n - 1] != '\n') rc = 0; // the regular layout ends every line with '\n'
    block_bytes = std::max<size_t>(std::min<size_t>(block_bytes, (size_t)768 << 20), 4096);
    n_parts = std::max(1, std::min(n_parts, 8));
    size_t at = 0, taken = 0;
    while (rc == 1 && at < n) {
        const size_t want = at + block_bytes;
        const size_t end = want >= n ? n : next_fastq_record(data, n, want);
        if (end - at > ((size_t)1 << 30) * (size_t)n_parts) { // no record start for a long stretch: not this layout
            rc = taken ? -1 : 0;
            break;
        }
        madvise(const_cast<unsigned char *>(data) + (end & ~(size_t)4095), std::min(block_bytes, n - (end & ~(size_t)4095)), MADV_WILLNEED);
        const char *parts[8];
        size_t lens[8], lo = at;
        for (int p = 0; p < n_parts; p++) {
            size_t hi = p + 1 == n_parts ? end : next_fastq_record(data, end, at + (end - at) * (size_t)(p + 1) / (size_t)n_parts);
            hi = std::max(std::min(hi, end), lo);
            parts[p] = reinterpret_cast<const char *>(data + lo);
            lens[p] = hi - lo;
            lo = hi;
        }
        const int r = sink(parts, lens);
        if (r < 0) {
            err = "raw FASTQ feed failed for " + path;
            rc = -1;
        } else if (r == 0)
            rc = taken ? -1 : 0;
        else
            taken++;
        at = end;
    }
    munmap(map, n);
    return rc;
}

} // namespace kat_host
