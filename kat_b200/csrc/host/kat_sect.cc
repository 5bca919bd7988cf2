// kat_sect.cc — `kat sect` of the GPU front-end: k-mer coverage along every sequence of a file
// (src/sect.cc:86-605).  The hash is counted (or loaded) exactly as for hist/gcp/comp; the per-position
// lookups — the reference's hot loop, one JellyfishHelper::getCount per k-mer (src/sect.cc:524-541,
// lib/src/jellyfish_helper.cc:189-194) — go to the GPU in batches through katgpu_lookup.  Everything
// else (record reading, the statistics, the text files) stays on the host and reproduces the
// reference's files byte for byte:
//   <prefix>-counts.cvg   FASTA-like, one count per k-mer position           (printCounts, :342-359)
//   <prefix>-stats.tsv    per-sequence median / mean / gc% / ... table       (printStatTable, :423-440)
//   <prefix>-counts.gc    with -g: GC% of every k-mer                        (printGCCounts, :365-383)
//   <prefix>-non_repetitive.fa / -repetitive.fa  with -E / -F                (printRegions, :385-420)
// The reference builds a contamination matrix as well but never writes it (Sect::save is not called
// from Sect::main, :733-735), so neither do we.
#include "kat_seqlookup.hpp"

#include <iomanip>

namespace kat_front {

namespace {

void print_regions(std::ostream &out, const vector<SeqRecord> &recs, const vector<vector<uint64_t>> &counts, uint32_t k, uint32_t min_count, uint32_t max_count) {
    for (size_t i = 0; i < recs.size(); i++) {
        uint32_t index = 1, start = 0;
        const vector<uint64_t> &sc = counts[i];
        const string &seq = recs[i].seq;
        const string maxcnt = max_count > 0 ? string("-") + std::to_string(max_count) : "+";
        if (sc.empty()) continue;
        bool in_region = false;
        string ss;
        auto header = [&](uint32_t end) {
            out << ">" << recs[i].name << "___region:" << index++ << "_length:" << end - start - 1 << "_pos:" << start + 1 << ":" << end << "_cov:" << min_count << maxcnt << endl;
        };
        for (size_t j = 0; j < sc.size(); j++) {
            const uint64_t c = sc[j];
            if (c >= min_count && (c <= max_count || max_count == 0)) {
                if (!in_region) {
                    start = (uint32_t)j;
                    in_region = true;
                }
                ss += seq[j];
            } else if (in_region) {
                const uint32_t end = (uint32_t)j + k - 1;
                header(end);
                out << ss;
                for (size_t q = j + 1; q < end; q++) out << seq[q];
                out << endl;
                in_region = false;
                ss.clear();
            }
        }
        if (in_region) {
            const uint32_t end = (uint32_t)sc.size() + k - 1;
            header(end);
            out << ss;
            for (size_t q = sc.size(); q < end; q++) out << seq[q];
            out << endl;
        }
    }
}

} // namespace

int sect_main(int argc, char **argv) {
    const vector<OptSpec> specs = {{'o', "output_prefix", true}, {'x', "gc_bins", true}, {'y', "cvg_bins", true}, {'l', "cvg_logscale", false},
                                   {'t', "threads", true}, {0, "5ptrim", true}, {'N', "non_canonical", false}, {'m', "mer_len", true},
                                   {'H', "hash_size", true}, {'n', "no_count_stats", false}, {'g', "output_gc_stats", false},
                                   {'E', "extract_nr", false}, {'F', "extract_r", false}, {'M', "min_repeat", true}, {'G', "max_repeat", true},
                                   {'d', "dump_hash", false}, {'v', "verbose", false}, {0, "gpus", true}, {0, "help", false}};
    const Parsed a = parse_args(argc, argv, specs);
    check_path_words(a, "seq_file", "counts_files");
    if (a.has("help") || argc <= 1) {
        cout << "Usage: kat sect [options] <sequence_file> (<input>)+\n\nEstimates coverage levels across sequences in the provided input sequence file.\n\n"
                "Options: -o prefix (kat-sect) -t threads --5ptrim list -N -m mer_len (27) -H hash_size (1e8) -n -g -E -F -M min_repeat (2) -G max_repeat (0) -d -v --help\n" << endl;
        return 1;
    }
    Timer total;
    cout << "Running KAT in SECT mode" << endl << "------------------------" << endl << endl;
    if (a.positional.size() < 2) throw KatError(4, "kat sect needs a sequence file and at least one input (sequence files to count, or a jellyfish hash)");
    const string seq_file = a.positional[0];
    int k = (int)to_u64(a.get("mer_len", "27"), "mer_len");
    check_mer_len(k, to_u64(a.get("hash_size", "100000000"), "hash_size"));
    Input in;
    in.files.assign(a.positional.begin() + 1, a.positional.end());
    in.set_trim(parse_trim(a.get("5ptrim", "0")));
    in.canonical = !a.has("non_canonical");
    in.threads = (unsigned)to_u64(a.get("threads", "1"), "threads");
    in.hash_size = to_u64(a.get("hash_size", "100000000"), "hash_size");
    const string prefix = a.get("output_prefix", "kat-sect");
    const uint32_t min_repeat = (uint32_t)to_u64(a.get("min_repeat", "2"), "min_repeat"), max_repeat = (uint32_t)to_u64(a.get("max_repeat", "0"), "max_repeat");
    if (!exists(seq_file)) throw KatError(4, "Could not find sequence file at: " + seq_file + "; please check the path and try again.");
    in.validate();
    ensure_parent_dir(prefix);
    Ctx ctx(gpus_option(a));
    check_wide_limits(k, ctx.n());
    if (a.has("dump_hash") && ctx.n() > 1) throw KatError(4, "--dump_hash writes one GPU's table: run without --gpus");
    if (in.load_mode) load_inputs_phase(ctx, &in, 1, k);
    else count_input(ctx, in, k, true);

    Timer t;
    cout << "Calculating kmer coverage across sequences ...";
    cout.flush();
    const vector<SeqRecord> recs = read_records(seq_file);
    const uint32_t K = (uint32_t)k;
    LookupWindow W(ctx, {&in}, recs, K);
    std::ofstream stats((prefix + "-stats.tsv").c_str());
    stats << "seq_name\tmedian\tmean\tgc%\tseq_length\tkmers_in_seq\tinvalid_kmers\t%_invalid\tnon_zero_kmers\t%_non_zero\t%_non_zero_corrected" << endl;
    stats << std::fixed << std::setprecision(5);
    std::ofstream cvg, gcs, nr, rp;
    if (!a.has("no_count_stats")) cvg.open((prefix + "-counts.cvg").c_str());
    if (a.has("output_gc_stats")) gcs.open((prefix + "-counts.gc").c_str());
    if (a.has("extract_nr")) nr.open((prefix + "-non_repetitive.fa").c_str());
    if (a.has("extract_r")) rp.open((prefix + "-repetitive.fa").c_str());
    // the reference works through the file in batches of 1024 records and writes each batch's outputs in turn
    const size_t kBatch = 1024;
    for (size_t b0 = 0; b0 < recs.size(); b0 += kBatch) {
        const size_t b1 = std::min(recs.size(), b0 + kBatch);
        vector<SeqRecord> brecs(recs.begin() + b0, recs.begin() + b1);
        vector<vector<uint64_t>> counts(b1 - b0);
        vector<vector<int>> gc_counts(b1 - b0);
        struct Row {
            uint32_t median, length, invalid, non_zero;
            double mean, gc, p_invalid, p_non_zero, p_non_zero_corr;
        };
        vector<Row> rows(b1 - b0);
        for (size_t i = b0; i < b1; i++) { // Sect::processSeq, src/sect.cc:490-605
            W.ensure(i);
            const size_t first_i = W.at(i);
            const vector<uint8_t> &valid = W.P.valid;
            const vector<uint64_t> &found = W.found[0];
            const string &s = recs[i].seq;
            const uint64_t len = s.size();
            const int64_t nb = (int64_t)len - k + 1;
            Row &r = rows[i - b0];
            uint64_t non_zero = 0, invalid = 0, sum = 0;
            r.median = 0;
            r.mean = 0.0;
            if (nb > 0) {
                vector<uint64_t> &sc = counts[i - b0];
                vector<int> &gc = gc_counts[i - b0];
                sc.assign((size_t)nb, 0);
                gc.assign((size_t)nb, 0);
                int g = 0; // G/C characters in the current window
                for (int64_t j = 0; j < (int64_t)len; j++) {
                    const char c = s[(size_t)j];
                    g += (c == 'G' || c == 'g' || c == 'C' || c == 'c');
                    if (j >= k) {
                        const char o = s[(size_t)(j - k)];
                        g -= (o == 'G' || o == 'g' || o == 'C' || o == 'c');
                    }
                    if (j + 1 >= k) {
                        const size_t p = (size_t)(j + 1 - k);
                        if (!valid[first_i + p]) {
                            sc[p] = 0;
                            gc[p] = -1;
                            invalid++;
                        } else {
                            const uint64_t cnt = found[first_i + p];
                            sum += cnt;
                            sc[p] = cnt;
                            gc[p] = g; // gcCount(merstr), lib/include/kat/str_utils.hpp:151-161
                            if (cnt != 0) non_zero++;
                        }
                    }
                }
                vector<uint64_t> sorted = sc;
                std::sort(sorted.begin(), sorted.end());
                r.median = (uint32_t)(double)sorted[sorted.size() / 2];
                r.mean = (double)sum / (double)nb;
            }
            r.length = (uint32_t)len;
            r.non_zero = (uint32_t)non_zero;
            r.p_non_zero = non_zero == 0 || nb <= 0 ? 0.0 : ((double)non_zero / (double)nb) * 100.0;
            r.invalid = (uint32_t)invalid;
            r.p_invalid = invalid == 0 || nb <= 0 ? 0.0 : ((double)invalid / (double)nb) * 100.0;
            const uint64_t not_invalid = (uint64_t)nb - invalid; // (unsigned, like the reference: never <= 0 unless equal)
            r.p_non_zero_corr = non_zero == 0 || not_invalid <= 0 ? 0.0 : ((double)non_zero / (double)not_invalid) * 100.0;
            uint64_t gs = 0, cs = 0, ns = 0;
            for (char c : s) {
                if (c == 'G' || c == 'g') gs++;
                else if (c == 'C' || c == 'c') cs++;
                else if (c == 'N' || c == 'n') ns++;
            }
            r.gc = ((double)(gs + cs)) / ((double)(len - ns));
        }
        if (cvg.is_open())
            for (size_t i = 0; i < counts.size(); i++) {
                cvg << ">" << brecs[i].name << endl;
                if (!counts[i].empty()) {
                    cvg << counts[i][0];
                    for (size_t j = 1; j < counts[i].size(); j++) cvg << " " << counts[i][j];
                    cvg << endl;
                } else
                    cvg << "0" << endl;
            }
        if (gcs.is_open())
            for (size_t i = 0; i < gc_counts.size(); i++) {
                gcs << ">" << brecs[i].name << std::fixed << std::setprecision(1) << endl;
                auto pct = [&](int c) { return c == -1 ? -0.1 : (((double)c / (double)k) * 100.0); };
                if (!gc_counts[i].empty()) {
                    gcs << pct(gc_counts[i][0]);
                    for (size_t j = 1; j < gc_counts[i].size(); j++) gcs << " " << pct(gc_counts[i][j]);
                    gcs << endl;
                } else
                    gcs << "0.0" << endl;
            }
        if (nr.is_open()) print_regions(nr, brecs, counts, K, 1, min_repeat);
        if (rp.is_open()) print_regions(rp, brecs, counts, K, min_repeat, max_repeat);
        for (size_t i = 0; i < rows.size(); i++) {
            const Row &r = rows[i];
            stats << brecs[i].name << "\t" << r.median << "\t" << r.mean << "\t" << r.gc << "\t" << r.length << "\t" << (uint32_t)(r.length - (uint32_t)k + 1) << "\t"
                  << r.invalid << "\t" << r.p_invalid << "\t" << r.non_zero << "\t" << r.p_non_zero << "\t" << r.p_non_zero_corr << endl;
        }
    }
    phase_done(t);
    if (a.has("dump_hash")) dump_input(in, k, prefix + "-hash.jf" + std::to_string(k));
    {
        Timer tm;
        cout << "Merging matrices ...";
        phase_done(tm);
    }
    free_tables(ctx, in);
    char b[96];
    snprintf(b, sizeof b, "KAT SECT completed.\nTotal runtime: %.1fs\n\n", total.secs());
    cout << b;
    return 0;
}

} // namespace kat_front
