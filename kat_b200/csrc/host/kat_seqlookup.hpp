// kat_seqlookup.hpp — what `kat sect` and `kat cold` share: reading the records of a sequence file the way SeqAn's
// readRecords does, and looking up every k-mer position of every record on the GPU(s) in batches
// (JellyfishHelper::getCount per position in the reference: src/sect.cc:524-541, src/cold.cc:338-356).
#pragma once
#include "kat_front.hpp"

namespace kat_front {

struct SeqRecord {
    string name, seq, qual; // qual: FASTQ only (kat filter seq writes the records back out)
};

inline bool is_ws(unsigned char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }

// SeqAn 2.0 readRecord for FASTA / FASTQ into CharStrings (seqan/seq_io/fasta_fastq.h): the id is the
// header line without its marker and line end; the sequence is every non-whitespace character up to
// the next '>' (FASTA) or up to the '+' that opens the quality block (FASTQ); qualities are skipped
// by count, ignoring whitespace.  *is_fastq (optional) reports the detected format.
inline vector<SeqRecord> read_records(const string &path, bool *is_fastq = nullptr) {
    std::vector<unsigned char> raw;
    string err;
    if (!kat_host::read_file(path, raw, err)) throw KatError(4, err);
    vector<SeqRecord> out;
    size_t i = 0;
    const size_t n = raw.size();
    while (i < n && is_ws(raw[i])) i++;
    if (i >= n) return out;
    const bool fastq = raw[i] == '@';
    if (is_fastq) *is_fastq = fastq;
    if (!fastq && raw[i] != '>') throw KatError(4, "Unknown sequence file format: " + path);
    while (i < n) {
        if (raw[i] != (fastq ? '@' : '>')) throw KatError(4, "Malformed record in " + path);
        i++;
        SeqRecord r;
        while (i < n && raw[i] != '\n' && raw[i] != '\r') r.name += (char)raw[i++];
        while (i < n && (raw[i] == '\n' || raw[i] == '\r')) i++;
        const unsigned char stop = fastq ? '+' : '>';
        while (i < n && raw[i] != stop) {
            if (!is_ws(raw[i])) r.seq += (char)raw[i];
            i++;
        }
        if (fastq) {
            while (i < n && raw[i] != '\n') i++; // the '+' line
            while (i < n && r.qual.size() < r.seq.size()) {
                if (!is_ws(raw[i])) r.qual += (char)raw[i];
                i++;
            }
            while (i < n && is_ws(raw[i])) i++;
        }
        out.push_back(std::move(r));
    }
    return out;
}

inline bool valid_base(char c) { // validKmer, lib/include/kat/str_utils.hpp:183-201
    switch (c) {
    case 'A': case 'a': case 'C': case 'c': case 'G': case 'g': case 'T': case 't': return true;
    default: return false;
    }
}
inline uint64_t base_code(char c) { // mer_dna::code, JFI/mer_dna.hpp:46-63
    switch (c) {
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 0;
    }
}


// Every k-mer position of every record: packed key (0 where the window holds a non-ACGT character) and validity;
// first[i] = index of record i's first position.  keys = the last 32 bases of the k-mer, keys_hi = the first K-32
// (all zero for K <= 32) — mer_dna's packing, first base most significant.
struct SeqPositions {
    vector<size_t> first; // relative to the first record of the range
    vector<uint64_t> keys, keys_hi;
    vector<uint8_t> valid;
};
// positions of records [begin, end)
inline SeqPositions kmer_positions(const vector<SeqRecord> &recs, size_t begin, size_t end, uint32_t K) {
    SeqPositions P;
    const size_t nr = end - begin;
    P.first.assign(nr + 1, 0);
    for (size_t i = 0; i < nr; i++) P.first[i + 1] = P.first[i] + (recs[begin + i].seq.size() >= K ? recs[begin + i].seq.size() - K + 1 : 0);
    P.keys.assign(P.first.back(), 0);
    if (K > 32) P.keys_hi.assign(P.first.back(), 0);
    P.valid.assign(P.first.back(), 0);
    const uint64_t lo_mask = K >= 32 ? ~0ULL : ((1ULL << (2 * K)) - 1);
    const uint64_t hi_mask = K <= 32 ? 0ULL : K == 64 ? ~0ULL : ((1ULL << (2 * K - 64)) - 1);
    for (size_t i = 0; i < nr; i++) {
        const string &s = recs[begin + i].seq;
        uint64_t lo = 0, hi = 0;
        uint32_t run = 0;
        for (size_t j = 0; j < s.size(); j++) {
            if (valid_base(s[j])) {
                hi = ((hi << 2) | (lo >> 62)) & hi_mask;
                lo = ((lo << 2) | base_code(s[j])) & lo_mask;
                run++;
            } else
                run = 0;
            if (j + 1 >= K) {
                const size_t p = P.first[i] + (j + 1 - K);
                P.keys[p] = run >= K ? lo : 0;
                if (K > 32) P.keys_hi[p] = run >= K ? hi : 0;
                P.valid[p] = run >= K;
            }
        }
    }
    return P;
}
// JellyfishHelper::getCount(hash, mer, canonical) for every position, batched on the GPU(s) (collective on a job)
inline vector<uint64_t> lookup_positions(Ctx &ctx, Input &in, const SeqPositions &P) {
    const vector<uint64_t> &keys = P.keys;
    vector<uint64_t> found(keys.size(), 0);
    const size_t batch = (size_t)1 << 26;
    for (size_t off = 0; off < keys.size(); off += batch) {
        const size_t m = std::min(batch, keys.size() - off);
        vector<vector<uint64_t>> other(ctx.n());
        run_on_all(ctx.n(), [&](int r) {
            if (r) other[r].resize(m);
            gpu_check(katgpu_lookup_wide(in.tables[r], P.keys_hi.empty() ? nullptr : P.keys_hi.data() + off, keys.data() + off, m, in.canonical ? 1 : 0,
                                         r ? other[r].data() : found.data() + off));
        });
    }
    return found;
}

// The tools walk the records of a sequence file in order; the k-mer positions of a WINDOW of records (at most
// KAT_LOOKUP_WINDOW positions, 2^26 by default: ~1.7 GB of keys, flags and counts on the host instead of 25 bytes per
// base of the whole file) are resident at a time, looked up in every hash of the tool in one batch each.
struct LookupWindow {
    Ctx &ctx;
    vector<Input *> inputs;
    const vector<SeqRecord> &recs;
    uint32_t K;
    size_t max_pos;
    size_t lo = 0, hi = 0; // records [lo, hi) are resident
    SeqPositions P;
    vector<vector<uint64_t>> found; // [input][position]
    LookupWindow(Ctx &c, vector<Input *> ins, const vector<SeqRecord> &r, uint32_t k)
        : ctx(c), inputs(std::move(ins)), recs(r), K(k), max_pos(env_size("KAT_LOOKUP_WINDOW", (size_t)1 << 26)) {}
    void ensure(size_t i) { // make record i resident (records are visited in increasing order)
        if (i >= lo && i < hi) return;
        lo = i;
        size_t n_pos = 0;
        for (hi = lo; hi < recs.size(); hi++) {
            const size_t np = recs[hi].seq.size() >= K ? recs[hi].seq.size() - K + 1 : 0;
            if (hi > lo && n_pos + np > max_pos) break;
            n_pos += np;
        }
        P = kmer_positions(recs, lo, hi, K);
        found.assign(inputs.size(), {});
        for (size_t t = 0; t < inputs.size(); t++) found[t] = lookup_positions(ctx, *inputs[t], P);
    }
    size_t at(size_t i) const { return P.first[i - lo]; }        // index of record i's first position in the window
    size_t n_positions(size_t i) const { return P.first[i - lo + 1] - P.first[i - lo]; }
};

} // namespace kat_front
