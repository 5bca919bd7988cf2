// kat_front.hpp — what the tools of the `kat` front-end share: option parsing, input groups
// (InputHandler: lib/src/input_handler.cc), the GPUs of a run (`--gpus N`), counting / loading /
// dumping an input through libkatgpu, matrix and distance printing.  Header-only (inline), one
// translation unit per tool: kat_hist.cc, kat_gcp.cc, kat_comp.cc, kat_sect.cc, kat_main.cc.
#pragma once
#include <glob.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "katgpu.h"
#include "jf_format.hpp"
#include "seq_parser.hpp"

using std::cerr;
using std::cout;
using std::endl;
using std::string;
using std::vector;

namespace kat_front {


inline const char *const kVersion = "2.4.2-katgpu0.2";

struct KatError : std::runtime_error {
    int code;
    KatError(int c, const string &m) : std::runtime_error(m), code(c) {}
};
inline void gpu_check(katgpu_status st) {
    if (st == KATGPU_ERR_HASH_FULL) throw KatError(5, "Hash full"); // std::runtime_error("Hash full") -> exit 5 (src/kat.cc:292)
    if (st != KATGPU_OK) throw KatError(4, string("katgpu: ") + katgpu_last_error());
}

// ---- tiny option parser (boost::program_options stand-in) ----------------------
struct OptSpec {
    char shortname;     // 0 = none
    const char *longname;
    bool takes_value;
};
struct Parsed {
    std::map<string, string> val;
    vector<string> positional;
    bool has(const string &k) const { return val.count(k) != 0; }
    string get(const string &k, const string &dflt) const {
        auto it = val.find(k);
        return it == val.end() ? dflt : it->second;
    }
};
inline void check_u16_options(const struct Parsed &p);
inline Parsed parse_args(int argc, char **argv, const vector<OptSpec> &specs) {
    Parsed p;
    for (int i = 1; i < argc; i++) {
        const string a = argv[i];
        const OptSpec *sp = nullptr;
        string value;
        bool have_value = false;
        if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
            string name = a.substr(2);
            const size_t eq = name.find('=');
            if (eq != string::npos) {
                value = name.substr(eq + 1);
                name = name.substr(0, eq);
                have_value = true;
            }
            for (auto &s : specs)
                if (name == s.longname) sp = &s;
            if (!sp) throw KatError(1, "Error: Parsing Command Line: unrecognised option '--" + name + "'");
        } else if (a.size() >= 2 && a[0] == '-' && a != "-") {
            for (auto &s : specs)
                if (s.shortname && a[1] == s.shortname) sp = &s;
            if (!sp) throw KatError(1, "Error: Parsing Command Line: unrecognised option '" + a + "'");
            if (a.size() > 2) {
                if (sp->takes_value) {
                    value = a.substr(2);
                    have_value = true;
                } else { // bundled switches: -nv
                    for (size_t j = 1; j < a.size(); j++) {
                        const OptSpec *q = nullptr;
                        for (auto &s : specs)
                            if (s.shortname && a[j] == s.shortname) q = &s;
                        if (!q || q->takes_value) throw KatError(1, "Error: Parsing Command Line: bad option bundle '" + a + "'");
                        p.val[q->longname] = "1";
                    }
                    continue;
                }
            }
        } else {
            p.positional.push_back(a);
            continue;
        }
        if (sp->takes_value) {
            if (!have_value) {
                if (i + 1 >= argc) throw KatError(1, string("Error: Parsing Command Line: the required argument for option '--") + sp->longname + "' is missing");
                value = argv[++i];
            }
            p.val[sp->longname] = value;
        } else
            p.val[sp->longname] = "1";
    }
    check_u16_options(p);
    return p;
}
inline uint64_t to_u64(const string &s, const char *what);
// Options the reference declares as uint16_t (po::value<uint16_t>: threads, mer_len, the bin counts, filter kmer's GC bounds):
// a value beyond 65535 is a command-line error there, exit code 1 (src/kat.cc:286-288).
inline void check_u16_options(const Parsed &p) {
    for (const char *name : {"threads", "mer_len", "cvg_bins", "gc_bins", "d1_bins", "d2_bins", "low_gc", "high_gc"}) {
        const auto it = p.val.find(name);
        if (it != p.val.end() && to_u64(it->second, name) > 65535)
            throw KatError(1, string("Error: Parsing Command Line: the argument ('") + it->second + "') for option '--" + name + "' is invalid");
    }
    for (const char *name : {"min_repeat", "max_repeat"}) { // uint32_t in kat sect
        const auto it = p.val.find(name);
        if (it != p.val.end() && to_u64(it->second, name) > 0xFFFFFFFFull)
            throw KatError(1, string("Error: Parsing Command Line: the argument ('") + it->second + "') for option '--" + name + "' is invalid");
    }
}
// Path-typed options (po::value<path> / vector<path>): boost reads a path up to the first white space, so a file name with a
// blank in it is "invalid" for the reference, a command-line error with exit code 1.  positional[0] is the tool's
// `first` hidden option, the others its `rest`.
inline void check_path_words(const Parsed &p, const char *first, const char *rest, std::initializer_list<const char *> options = {"output_prefix"},
                             bool positionals = true) {
    auto bad = [](const string &v) { return v.find_first_of(" \t\n\r\f\v") != string::npos; };
    auto fail = [](const string &v, const char *name) {
        throw KatError(1, string("Error: Parsing Command Line: the argument ('") + v + "') for option '--" + name + "' is invalid");
    };
    for (const char *name : options) {
        const auto it = p.val.find(name);
        if (it != p.val.end() && bad(it->second)) fail(it->second, name);
    }
    for (size_t i = 0; positionals && i < p.positional.size(); i++)
        if (bad(p.positional[i])) fail(p.positional[i], i == 0 ? first : rest);
}
inline uint64_t to_u64(const string &s, const char *what) {
    try {
        size_t pos = 0;
        if (s.empty() || s[0] == '-') throw std::invalid_argument(s);
        // boost::lexical_cast-like: accept plain integers and 1e8-style values
        const double d = std::stod(s, &pos);
        if (pos != s.size() || d < 0) throw std::invalid_argument(s);
        if (s.find_first_of(".eE") == string::npos) return std::stoull(s);
        return (uint64_t)d;
    } catch (...) {
        throw KatError(1, string("Error: Parsing Command Line: the argument ('") + s + "') for option '--" + what + "' is invalid");
    }
}
inline double to_double(const string &s, const char *what) {
    try {
        size_t pos = 0;
        const double d = std::stod(s, &pos);
        if (pos != s.size()) throw std::invalid_argument(s);
        return d;
    } catch (...) {
        throw KatError(1, string("Error: Parsing Command Line: the argument ('") + s + "') for option '--" + what + "' is invalid");
    }
}
inline vector<unsigned> parse_trim(const string &s) {
    vector<unsigned> v;
    std::stringstream ss(s);
    string item;
    while (std::getline(ss, item, ',')) {
        try {
            v.push_back((unsigned)to_u64(item, "5ptrim"));
        } catch (const KatError &) { // boost::lexical_cast<uint16_t> of the list items (src/histogram.cc:345-350): a boost::exception, exit code 4
            throw KatError(4, "bad lexical cast: source type value could not be interpreted as target");
        }
    }
    if (v.empty()) v.push_back(0);
    return v;
}

// mer lengths: one 64-bit word per k-mer up to 32, two words for 33..64 (jellyfish::mer_dna has no limit; larger k
// is refused).  Two-word k-mers are counted on one GPU.
constexpr int kMaxMerLen = 64;
inline void check_mer_len(int k, uint64_t hash_size = 100000000ULL) {
    if (k == 0) { // HashCounter(hashSize, 0, ...): jellyfish refuses a hash matrix without columns, std::runtime_error -> exit code 5
        int rows = 0;
        while (rows < 64 && (1ULL << rows) < hash_size) rows++;
        throw KatError(5, "Invalid matrix size " + std::to_string(hash_size < 2 ? 0 : rows) + "x0");
    }
    if (k < 1 || k > kMaxMerLen) throw KatError(4, "This GPU build supports mer lengths 1..64 (one or two 64-bit words per k-mer); got " + std::to_string(k));
}
inline void check_wide_limits(int k, int gpus) {
    if (k > 32 && gpus > 1) throw KatError(4, "mer lengths above 32 are counted on one GPU: run without --gpus");
}

// ---- InputHandler (lib/src/input_handler.cc) ------------------------------------
inline string leaf(const string &p) {
    const size_t s = p.find_last_of('/');
    return s == string::npos ? p : p.substr(s + 1);
}
// JellyfishHelper::isPipe (lib/src/jellyfish_helper.cc:258-260): anything under /proc or /dev —
// including regular files in /dev/shm — is "a pipe": printed as <pipe>, assumed to be a sequence
// file, exempt from the existence check.
inline bool is_pipe(const string &p) { return p.rfind("/proc", 0) == 0 || p.rfind("/dev", 0) == 0; }
inline string join(const vector<string> &v, bool leaves) { // leaves: InputHandler::fileName, else InputHandler::pathString
    string s;
    for (auto &p : v) s += (leaves ? leaf(p) : is_pipe(p) ? string("<pipe>") : p) + " ";
    while (!s.empty() && s.back() == ' ') s.pop_back(); // boost::trim_right_copy
    return s;
}
inline bool exists(const string &p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0;
}
inline string lower(string s) {
    for (auto &c : s) c = (char)tolower((unsigned char)c);
    return s;
}
inline string ext_of(const string &p) {
    const string l = leaf(p);
    const size_t d = l.find_last_of('.');
    return d == string::npos || d == 0 ? "" : l.substr(d);
}
// JellyfishHelper::isSequenceFile (lib/src/jellyfish_helper.cc:270-307): by extension, else by first byte
inline bool is_sequence_file(const string &path) {
    if (is_pipe(path)) return true; // lib/src/jellyfish_helper.cc:273
    string e = lower(ext_of(path));
    string stem = path;
    if (e == ".gz") {
        stem = path.substr(0, path.size() - 3);
        e = lower(ext_of(stem));
    }
    static const char *kExt[] = {".fastq", ".fq", ".fasta", ".fa", ".fna", ".fas", ".scafseq"};
    for (auto x : kExt)
        if (e == x) return true;
    std::ifstream f(path);
    char ch = 0;
    f >> ch;
    return ch == '>' || ch == '@';
}
// InputHandler::globFiles (lib/src/input_handler.cc:245-316): space-separated patterns, brace + tilde expansion
inline vector<string> glob_group(const string &group) {
    vector<string> pats;
    std::stringstream ss(group);
    string item;
    while (std::getline(ss, item, ' '))
        pats.push_back(item);
    if (pats.empty()) throw KatError(4, "No input provided for this input group");
    glob_t g;
    memset(&g, 0, sizeof g);
    int i = 0;
    for (auto &p : pats) {
        int flags = GLOB_TILDE | GLOB_NOCHECK | GLOB_BRACE;
        if (i++ > 0) flags |= GLOB_APPEND;
        if (glob(p.c_str(), flags, nullptr, &g) != 0) throw KatError(4, "Problem globbing input pattern: " + p);
    }
    vector<string> out;
    for (size_t j = 0; j < g.gl_pathc; j++) out.push_back(g.gl_pathv[j]);
    if (g.gl_pathc) globfree(&g);
    if (out.empty()) out.push_back(pats[0]);
    return out;
}

struct Input {
    vector<string> files;
    vector<unsigned> trim5p;
    bool canonical = true;
    uint64_t hash_size = 100000000ULL; // DEFAULT_HASH_SIZE, lib/include/kat/jellyfish_helper.hpp:76
    int index = 1;
    unsigned threads = 1;   // -t: decides how many of the group's files the reference reads side by side (count_input)
    bool load_mode = false; // InputMode::LOAD: the (first) input is a jellyfish hash
    katgpu_table *table = nullptr;        // rank 0's (the only one on a single GPU)
    vector<katgpu_table *> tables;        // one per GPU of the job
    uint64_t n_distinct = 0, n_total = 0;

    void set_trim(const vector<unsigned> &t) { // InputHandler::set5pTrim
        if (t.empty()) trim5p.assign(files.size(), 0);
        else if (t.size() == 1) trim5p.assign(files.size(), t[0]);
        else if (t.size() == files.size()) trim5p = t;
        else throw KatError(4, "Inconsistent number of inputs and trimming settings.  Please establish your inputs before trying to set trimming vector.  Also ensure you have the same number of input files to trimming settings.");
    }
    void validate() { // InputHandler::validateInput (lib/src/input_handler.cc:99-136)
        for (auto &p : files) {
            if (!is_pipe(p) && !exists(p)) throw KatError(4, "Could not find input file at: " + p + "; please check the path and try again.");
            // `start` is never cleared in the reference, so the mode is simply that of the last
            // file and the "Cannot mix sequence files and jellyfish hashes" check cannot fire
            load_mode = !is_sequence_file(p);
        }
    }
};

struct Timer {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    double secs() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};
inline void phase_done(const Timer &t) {
    char b[64];
    snprintf(b, sizeof b, "  Time taken: %.1fs\n\n", t.secs()); // auto_cpu_timer(1, "  Time taken: %ws\n\n")
    cout << " done." << b;
    cout.flush();
}

// Files at least this large (KAT_STREAM_BYTES) are read, parsed and counted block by block
// (KAT_STREAM_BLOCK bytes at a time) when they are in a regular layout: host memory stays at a
// few blocks instead of file + base stream.
inline size_t env_size(const char *name, size_t dflt) {
    const char *e = getenv(name);
    return e ? (size_t)strtoull(e, nullptr, 10) : dflt;
}
inline size_t file_size_or_zero(const string &p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode) ? (size_t)st.st_size : 0;
}

// ---- GPUs -------------------------------------------------------------------------
// `--gpus N` (or KATGPU_GPUS): one libkatgpu context and one host thread per GPU in this process.
// The contexts form a job (katgpu_dist_init: NCCL communicator + peer-mapped receive areas); every
// libkatgpu call on a job is collective, so run_on_all() makes the same call from N threads.
// Reads are sharded by position in the parsed base stream; k-mers travel to their owner GPU
// inside libkatgpu; every GPU receives the all-reduced result and rank 0's copy is written out.
[[noreturn]] inline void die(const KatError &e) {
    cout.flush();
    cerr << (e.code == 5 ? "Error: " : "") << e.what() << endl;
    _exit(e.code); // other GPU threads may be parked in a collective: no orderly unwinding
}
template <class F>
inline void run_on_all(int n, F f) { // f(rank)
    if (n == 1) {
        f(0);
        return;
    }
    vector<std::thread> th;
    for (int r = 0; r < n; r++)
        th.emplace_back([&f, r] {
            try {
                f(r);
            } catch (const KatError &e) {
                static std::mutex m;
                std::lock_guard<std::mutex> g(m);
                die(e);
            }
        });
    for (auto &t : th) t.join();
}
struct Ctx {
    vector<katgpu_ctx *> all;
    katgpu_ctx *c = nullptr; // rank 0
    explicit Ctx(int gpus) {
        int dev0 = 0;
        if (const char *e = getenv("KATGPU_DEVICE")) dev0 = atoi(e);
        if (gpus < 1 || gpus > 8) throw KatError(4, "--gpus must be 1..8");
        if (gpus > 1 && katgpu_device_count() < dev0 + gpus)
            throw KatError(4, "--gpus " + std::to_string(gpus) + ": only " + std::to_string(katgpu_device_count()) + " CUDA devices visible");
        all.assign(gpus, nullptr);
        for (int r = 0; r < gpus; r++) gpu_check(katgpu_init(dev0 + r, &all[r]));
        c = all[0];
        if (gpus > 1) {
            unsigned char id[128];
            gpu_check(katgpu_dist_unique_id(id));
            run_on_all(gpus, [&](int r) { gpu_check(katgpu_dist_init(all[r], r, gpus, id)); });
        }
    }
    int n() const { return (int)all.size(); }
    ~Ctx() {
        run_on_all(n(), [&](int r) { katgpu_shutdown(all[r]); });
    }
};
inline int gpus_option(const Parsed &a) {
    const char *e = getenv("KATGPU_GPUS");
    return (int)to_u64(a.get("gpus", e ? e : "1"), "gpus");
}
inline void free_tables(Ctx &ctx, Input &in) {
    run_on_all(ctx.n(), [&](int r) {
        if (r < (int)in.tables.size()) katgpu_table_free(in.tables[r]);
    });
    in.tables.clear();
    in.table = nullptr;
}
// Feed one parsed base stream: on a job, GPU r takes the r-th share of the k-mer START positions;
// its bytes run k-1 past the share's end so that every k-mer is complete on exactly one GPU (a
// start inside that tail has fewer than k bases left and yields nothing).
inline void feed_all(Ctx &ctx, Input &in, int k, const char *b, size_t n) {
    if (ctx.n() == 1) {
        gpu_check(katgpu_count_feed(in.table, b, n));
        return;
    }
    run_on_all(ctx.n(), [&](int r) {
        const size_t lo = n * (size_t)r / ctx.n(), hi = n * (size_t)(r + 1) / ctx.n();
        const size_t end = std::min(n, hi + (size_t)k - 1);
        gpu_check(katgpu_count_feed(in.tables[r], b + lo, hi > lo ? end - lo : 0));
    });
}

// InputHandler::count (lib/src/input_handler.cc:180-202) on the GPU
inline void count_input(Ctx &ctx, Input &in, int k, bool allow_grow, bool allow_streaming = true) {
    Timer t;
    if (allow_streaming) {
        cout << "Input " << in.index << " is a sequence file.  Counting kmers for input " << in.index << " (" << join(in.files, false) << ") ...";
        cout.flush();
    }
    // HashCounter(hashSize, ...) (lib/src/input_handler.cc:184): a table of 2^ceilLog2(size) slots hashes with a matrix of
    // ceilLog2(size) rows, and jellyfish refuses one without rows (deps/jellyfish-2.2.0/lib/rectangular_binary_matrix.cc:24-28):
    // std::runtime_error, exit code 5
    if (in.hash_size < 2) throw KatError(5, "Invalid matrix size 0x" + std::to_string(2 * k));
    const size_t n_streams = std::max<size_t>(1, std::min<size_t>(in.files.size(), std::max(1u, in.threads))); // parser streams, see below
    // A file in no supported format anywhere in the group: fatal among the files the constructor opens (checked before
    // anything is counted); otherwise the group is read in "careful" mode — every file through the emulation, which knows
    // the reference's buffers, and the last buffer of each file held back until its stream has opened its next file.
    bool careful = false;
    {
        size_t opened = 0; // streams that hold their first file
        for (size_t i = 0; i < in.files.size(); i++) {
            const int c = kat_host::first_byte_of_file(in.files[i]);
            if (c == -1) continue;
            const bool supported = c == '>' || c == '@' || c == -2; // (a pipe is taken for a sequence stream; the parser has the last word)
            if (opened < n_streams) {
                if (!supported) throw KatError(5, "Unsupported format");
                opened++;
            } else if (!supported)
                careful = true;
        }
    }
    in.tables.assign(ctx.n(), nullptr);
    // a job shares the k-mers out: every GPU's table needs room for its 1/N of the distinct ones
    const uint64_t per_gpu_slots = ctx.n() > 1 ? std::max<uint64_t>(in.hash_size / ctx.n(), 1024) : in.hash_size;
    run_on_all(ctx.n(), [&](int r) { gpu_check(katgpu_count_begin(ctx.all[r], k, in.canonical ? 1 : 0, per_gpu_slots, allow_grow ? 1 : 0, &in.tables[r])); });
    in.table = in.tables[0];
    { // size the library's pending-key buffers once, for the whole input (about one k-mer per two file bytes of FASTQ)
        uint64_t bytes = 0;
        for (auto &f : in.files) bytes += file_size_or_zero(f);
        if (bytes >= ((uint64_t)64 << 20)) run_on_all(ctx.n(), [&](int r) { gpu_check(katgpu_count_reserve(in.tables[r], bytes / 2 / (uint64_t)ctx.n())); });
    }
    const size_t stream_from = env_size("KAT_STREAM_BYTES", (size_t)2 << 30), block = env_size("KAT_STREAM_BLOCK", (size_t)512 << 20);
    std::vector<unsigned char> raw;
    std::string bases;
    // KAT_GPU_PARSE=0 keeps FASTQ parsing on the host
    const bool gpu_parse = allow_streaming && env_size("KAT_GPU_PARSE", 1) != 0;
    // The reference reads the group through S = min(files, threads) parser streams (lib/src/jellyfish_helper.cc:231-233).  Files
    // are handed out in order: the first S while the parser is built, each later one to the stream that reaches the end of its
    // file first (an empty file is skipped by the stream that opened it, which takes the next one as well).  What a stream
    // keeps from one file to its next is `seq_len` (seq_parser.hpp: non-zero only after a file that ended inside a FASTQ
    // record; the regular-layout routes below then do not apply) — and whether it is still alive: an exception inside
    // produce() ends THAT stream only (JFI/cooperative_pool2.hpp:249-260), the others go on taking files; with one stream the
    // rest of the group is lost.  Two exceptions exist:
    //   * "Invalid fastq sequence": the buffer being filled is dropped with it (the emulation in seq_parser.cc does that);
    //   * "Unsupported format" when the stream opens its NEXT file: produce() is still holding the last 4 KiB buffer of the
    //     file it just finished, which is lost with it (JFI/mer_overlap_sequence_parser.hpp:132-152).  From the files opened
    //     while the parser is built the same exception is not caught at all: the tool fails with exit code 5
    //     (:119-122,158-187; src/kat.cc:293-295).
    // Which live stream takes a later file is a race in the reference; it shows in the result only when a stream carries
    // seq_len or runs into an unsupported file, and is modelled by bytes read so far.
    struct RefStream {
        bool alive = true, has_file = false;
        size_t carry = 0;
        uint64_t clock = 0;
        std::string held; // careful mode: the last buffer of the stream's previous file (with its seam), not yet counted
    };
    vector<RefStream> streams(n_streams);
    const bool fast_routes = allow_streaming && !careful;
    size_t built = 0; // streams that hold their first file: the parser's constructor is done when built == streams.size()
    for (size_t i = 0; i < in.files.size(); i++) {
        string err;
        const bool ctor = built < streams.size();
        size_t s = built;
        if (!ctor) {
            s = streams.size();
            for (size_t j = 0; j < streams.size(); j++)
                if (streams[j].alive && (s == streams.size() || streams[j].clock < streams[s].clock)) s = j;
            if (s == streams.size()) break; // every stream has ended: the remaining files are never opened
        }
        RefStream &st = streams[s];
        st.clock += file_size_or_zero(in.files[i]);
        auto took_file = [&]() {
            if (ctor && !st.has_file) st.has_file = true, built++;
        };
        if (fast_routes && gpu_parse && st.carry == 0) { // plain 4-line FASTQ: the raw bytes go to the GPU(s) and are parsed there
            const int r = kat_host::stream_fastq_raw(in.files[i], env_size("KAT_RAW_BLOCK", (size_t)256 << 20) * (size_t)ctx.n(), ctx.n(),
                                                     [&](const char *const *parts, const size_t *lens) {
                                                         vector<int> regular(ctx.n(), 0);
                                                         run_on_all(ctx.n(), [&](int rk) {
                                                             gpu_check(katgpu_count_feed_fastq(in.tables[rk], parts[rk], lens[rk], in.trim5p[i], &regular[rk]));
                                                         });
                                                         int n_reg = 0;
                                                         for (int rk = 0; rk < ctx.n(); rk++) n_reg += regular[rk] != 0;
                                                         // some GPUs counted their share, others refused theirs: start over
                                                         return n_reg == ctx.n() ? 1 : n_reg == 0 ? 0 : -1;
                                                     }, err);
            if (r == 1) {
                took_file();
                continue;
            }
            if (r < 0) { // irregular after some blocks were counted: start this input over with host parsing
                free_tables(ctx, in);
                count_input(ctx, in, k, allow_grow, false);
                return;
            }
        }
        if (fast_routes && st.carry == 0 && file_size_or_zero(in.files[i]) >= stream_from) {
            const int r = kat_host::stream_regular_file(in.files[i], in.trim5p[i], block, [&](const char *b, size_t n) {
                feed_all(ctx, in, k, b, n);
                return true;
            }, err);
            if (r == 1) {
                took_file();
                continue;
            }
            if (r < 0) { // irregular after some blocks were counted: start this input over, whole files at a time
                free_tables(ctx, in);
                count_input(ctx, in, k, allow_grow, false);
                return;
            }
        }
        if (!kat_host::read_file(in.files[i], raw, err)) throw KatError(4, err);
        if (raw.empty()) continue; // open_next_file, `case EOF`: the same stream opens the next file
        took_file();
        bases.clear();
        kat_host::LastBuffer last;
        const auto status = careful ? kat_host::parse_sequence_file_exact(raw.data(), raw.size(), k, in.trim5p[i], bases, &st.carry, &last)
                                    : kat_host::parse_sequence_file(raw.data(), raw.size(), k, in.trim5p[i], bases, &st.carry);
        if (status == kat_host::ParseStatus::UnsupportedFormat) { // (never while the parser is built: checked above)
            cerr << "warning: " << in.files[i] << ": Unsupported format; like the reference, the parser stream that opened it loses the last buffer of its previous file and reads no further files" << endl;
            st.alive = false;
            st.held.clear();
            continue;
        }
        if (!st.held.empty()) { // the stream has opened its next file: the buffer it was holding is queued
            feed_all(ctx, in, k, st.held.data(), st.held.size());
            st.held.clear();
        }
        if (status == kat_host::ParseStatus::InvalidFastq) {
            cerr << "warning: " << in.files[i] << ": Invalid fastq sequence; like the reference, the parser stream that read it reads no further files" << endl;
            st.alive = false;
            if (!bases.empty()) feed_all(ctx, in, k, bases.data(), bases.size());
            continue;
        }
        if (careful) { // count up to the last buffer; hold that one (and the seam its k-mers reach back over) back
            if (last.at > 0) feed_all(ctx, in, k, bases.data(), last.at);
            st.held.assign(bases, last.at - last.seam, std::string::npos);
        } else if (!bases.empty())
            feed_all(ctx, in, k, bases.data(), bases.size());
    }
    for (auto &st : streams) // no file left to open: every stream queues what it holds and is done
        if (!st.held.empty()) feed_all(ctx, in, k, st.held.data(), st.held.size());
    run_on_all(ctx.n(), [&](int r) {
        uint64_t nd = 0, nt = 0;
        gpu_check(katgpu_count_end(in.tables[r], &nd, &nt)); // on a job: the job's totals, on every rank
        if (r == 0) in.n_distinct = nd, in.n_total = nt;
    });
    phase_done(t);
}

// InputHandler::loadHeader + loadHash (lib/src/input_handler.cc:138-142,204-219) via
// HashLoader::loadHash (lib/src/jellyfish_helper.cc:97-187): only the FIRST file of the group is
// read; k and the canonical flag come from its header.  Returns the file's mer length.
inline kat_host::JfHeader read_jf_header(const string &path, std::ifstream &f) {
    f.open(path.c_str(), std::ios::in | std::ios::binary);
    kat_host::JfHeader h;
    string err;
    if (!f.good() || !kat_host::jf_read_header(f, h, err)) throw KatError(4, "Failed to parse header of file: " + path);
    if (h.format == "bloomcounter")
        throw KatError(4, "KAT does not currently support bloom counted kmer hashes.  Please create a binary hash with jellyfish or KAT and use that instead.");
    if (h.format == "text/sorted")
        throw KatError(4, "Processing a text format hash will be painfully slow, so we don't support it.  Please create a binary hash with jellyfish or KAT and use that instead.");
    if (h.format != "binary/sorted") throw KatError(4, "Unknown format '" + h.format + "'");
    return h;
}
inline int load_input(katgpu_ctx *ctx, Input &in) {
    std::ifstream f;
    const kat_host::JfHeader h = read_jf_header(in.files[0], f);
    const int k = (int)(h.key_len / 2);
    if (k < 1 || k > kMaxMerLen) throw KatError(4, "This GPU build supports mer lengths 1..64 (one or two 64-bit words per k-mer); the hash holds " + std::to_string(k) + "-mers");
    struct stat st;
    uint64_t n_records = 0;
    if (stat(in.files[0].c_str(), &st) == 0 && (uint64_t)st.st_size >= h.offset) {
        const uint64_t data = (uint64_t)st.st_size - h.offset;
        if (data % h.record_bytes())
            throw KatError(4, "Size of database (" + std::to_string(data) + ") must be a multiple of the length of a record (" + std::to_string(h.record_bytes()) + ")");
        n_records = data / h.record_bytes();
    }
    // a hash without records: the reference asks for an array of 2^ceilLog2(0) slots, whose hash matrix has no rows —
    // std::runtime_error("Invalid matrix size 0x<key bits>"), exit code 5 (lib/src/jellyfish_helper.cc:149-165,
    // deps/jellyfish-2.2.0/lib/rectangular_binary_matrix.cc:24-28)
    if (n_records == 0) throw KatError(5, "Invalid matrix size 0x" + std::to_string(h.key_len));
    in.canonical = h.canonical;
    // the reference sizes the new array at 2 * nbRecords rounded up to a power of two (:149-150)
    gpu_check(katgpu_count_begin(ctx, k, in.canonical ? 1 : 0, std::max<uint64_t>(2 * n_records, 1024), 1, &in.table));
    string err;
    katgpu_status first_bad = KATGPU_OK;
    const bool ok = k > 32 ? kat_host::jf_read_records_wide(f, h, (size_t)1 << 22, [&](const uint64_t *keys_hi, const uint64_t *keys_lo, const uint64_t *counts, size_t n) {
        if (first_bad == KATGPU_OK) first_bad = katgpu_count_add_pairs_wide(in.table, keys_hi, keys_lo, counts, n);
    }, err) : kat_host::jf_read_records(f, h, (size_t)1 << 22, [&](const uint64_t *keys, const uint64_t *counts, size_t n) {
        if (first_bad == KATGPU_OK) first_bad = katgpu_count_add_pairs(in.table, keys, counts, n);
    }, err);
    gpu_check(first_bad);
    if (!ok) throw KatError(4, err);
    gpu_check(katgpu_count_end(in.table, &in.n_distinct, &in.n_total));
    return k;
}
inline void load_inputs_phase(Ctx &gpus, Input *ins, int n_in, int &k) { // the "Loading hashes into memory..." phase
    if (gpus.n() > 1) throw KatError(4, "jellyfish hash inputs are loaded on one GPU: run without --gpus");
    katgpu_ctx *ctx = gpus.c;
    Timer t;
    cout << "Loading hashes into memory...";
    cout.flush();
    for (int i = 0; i < n_in; i++)
        if (ins[i].load_mode) {
            const int kk = load_input(ctx, ins[i]);
            ins[i].tables.assign(1, ins[i].table);
            if (n_in == 1) k = kk; // hist / gcp take merLen from the loaded hash (InputHandler::loadHash)
        }
    phase_done(t);
}

// InputHandler::dump (lib/src/input_handler.cc:221-243): a counted input is written as a
// jellyfish binary/sorted hash (JellyfishHelper::dumpHash: 4-byte counters, one file), a
// loaded one is symlinked.
inline void dump_input(Input &in, int k, const string &out_path) {
    struct stat st;
    if (lstat(out_path.c_str(), &st) == 0) unlink(out_path.c_str());
    if (in.load_mode) {
        if (symlink(in.files[0].c_str(), out_path.c_str()) != 0) throw KatError(4, "Could not create symlink " + out_path);
        return;
    }
    Timer t;
    cout << "Dumping hash to " << out_path << " ...";
    cout.flush();
    vector<uint64_t> keys(std::max<uint64_t>(in.n_distinct, 1)), counts(keys.size()), keys_hi(k > 32 ? keys.size() : 0);
    uint64_t n = 0;
    if (k > 32) gpu_check(katgpu_table_export_wide(in.table, keys_hi.data(), keys.data(), counts.data(), keys.size(), &n));
    else gpu_check(katgpu_table_export(in.table, keys.data(), counts.data(), keys.size(), &n));
    keys.resize((size_t)std::min<uint64_t>(n, keys.size()));
    counts.resize(keys.size());
    kat_host::JfHeader h;
    std::ofstream out(out_path.c_str(), std::ios::out | std::ios::binary);
    if (k > 32) { // two words per key: 2k bits = ceil(2k / 8) bytes per record key, low word first
        keys_hi.resize(keys.size());
        kat_host::jf_prepare_dump_wide(h, (unsigned)k, in.canonical, keys.size(), keys_hi, keys, counts);
        kat_host::jf_write_file_wide(out, h, keys_hi, keys, counts);
    } else {
        kat_host::jf_prepare_dump(h, (unsigned)k, in.canonical, keys.size(), keys, counts);
        kat_host::jf_write_file(out, h, keys, counts);
    }
    out.close();
    if (!out.good()) throw KatError(4, "Could not write " + out_path);
    phase_done(t);
}

inline void ensure_parent_dir(const string &prefix) { // KatFS::ensureDirectoryExists(absolute(prefix).parent_path())
    string p = prefix;
    if (p.empty() || p[0] != '/') {
        char cwd[4096];
        if (!getcwd(cwd, sizeof cwd)) return;
        p = string(cwd) + "/" + p;
    }
    const size_t s = p.find_last_of('/');
    if (s == string::npos || s == 0) return;
    const string dir = p.substr(0, s);
    string cur;
    std::stringstream ss(dir);
    string part;
    while (std::getline(ss, part, '/')) {
        if (part.empty()) continue;
        cur += "/" + part;
        if (!exists(cur) && mkdir(cur.c_str(), 0777) != 0 && !exists(cur))
            throw KatError(4, "Could not create output directory: " + dir);
    }
}

// SparseMatrix::printMatrix / getMaxVal (lib/include/kat/sparse_matrix.hpp:162-173,269-277)
inline void print_matrix(std::ostream &out, const vector<uint64_t> &mx, uint32_t rows, uint32_t cols) {
    for (uint32_t i = 0; i < rows; i++) {
        out << mx[(size_t)i * cols];
        for (uint32_t j = 1; j < cols; j++) out << " " << mx[(size_t)i * cols + j];
        out << endl;
    }
}
inline uint64_t max_val(const vector<uint64_t> &mx) {
    uint64_t m = 0;
    for (auto v : mx) m = std::max(m, v);
    return m;
}
// boost::filesystem::path::operator<< : quoted, '&' escapes '"' and '&'
inline string quoted_path(const string &p) {
    string s = "\"";
    for (char c : p) {
        if (c == '"' || c == '&') s += '&';
        s += c;
    }
    return s + "\"";
}

// Distance metrics with the reference's arithmetic (lib/include/kat/distance_metrics.hpp:50-126)
inline double minkowski(const vector<uint64_t> &a, const vector<uint64_t> &b, int p) {
    uint64_t sum = 0;
    for (size_t i = 0; i < a.size(); i++) {
        const uint64_t diff = a[i] < b[i] ? b[i] - a[i] : a[i] - b[i];
        sum += std::pow(diff, p); // double -> uint64 at every step, as the reference
    }
    return p == 1 ? (double)sum : std::pow((double)sum, 1.0 / (double)p);
}
inline double cosine(const vector<uint64_t> &a, const vector<uint64_t> &b) {
    double dot = 0.0, da = 0.0, db = 0.0;
    for (size_t i = 0; i < a.size(); i++) {
        dot += a[i] * b[i]; // uint64 product (wraps) before the double add, as the reference
        da += std::pow(a[i], 2);
        db += std::pow(b[i], 2);
    }
    return 1.0 - (dot / (std::sqrt(da) * std::sqrt(db)));
}
inline double canberra(const vector<uint64_t> &a, const vector<uint64_t> &b) {
    double sum = 0.0;
    for (size_t i = 0; i < a.size(); i++) {
        const double diff = (double)a[i] - (double)b[i];
        const double s = a[i] + b[i];
        if (s > 0) sum += std::abs(diff) / s;
    }
    return sum;
}
inline double jaccard(const vector<uint64_t> &a, const vector<uint64_t> &b) {
    double mn = 0.0, mx = 0.0;
    for (size_t i = 0; i < a.size(); i++) mn += std::min(a[i], b[i]);
    for (size_t i = 0; i < a.size(); i++) mx += std::max(a[i], b[i]);
    return 1.0 - (mn / mx);
}
inline void print_distances(std::ostream &out, const vector<uint64_t> &a, const vector<uint64_t> &b) {
    out << " - Manhattan distance: " << minkowski(a, b, 1) << endl;
    out << " - Euclidean distance: " << minkowski(a, b, 2) << endl;
    out << " - Cosine distance: " << cosine(a, b) << endl;
    out << " - Canberra distance: " << canberra(a, b) << endl;
    out << " - Jaccard distance: " << jaccard(a, b) << endl;
}



} // namespace kat_front
