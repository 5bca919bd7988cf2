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

#include <zlib.h>

#include <cstring>

namespace kat_host {

bool read_file(const std::string &path, std::vector<unsigned char> &out, std::string &err) {
    // every input goes through zlib, as the reference's stream_manager does
    // (JFI/stream_manager.hpp:133-145): gzread passes plain files through untouched.
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

} // namespace

ParseStatus parse_sequence_file(const unsigned char *data, size_t n, int k, unsigned trim5p, std::string &out) {
    CharStream is(data, n);
    const int first = is.peek();
    if (first == -1) return ParseStatus::Ok; // empty file: nothing to count
    if (first != '>' && first != '@') return ParseStatus::UnsupportedFormat;
    const bool fastq = first == '@';
    is.ignore_line();
    unsigned char buf[kBuf + 8];
    const size_t seam = static_cast<size_t>(k - 1);
    bool have_seam = false;
    size_t seq_len = 0;
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
        out.append(reinterpret_cast<const char *>(buf + start), read - start);
        have_seam = read >= seam;
    }
    return ParseStatus::Ok;
}

} // namespace kat_host
