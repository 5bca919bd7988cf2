// kat_cold.cc — `kat cold` of the GPU front-end (src/cold.cc:68-425): for every contig of an assembly the median and
// mean read-k-mer coverage, the assembly copy number (median count of its k-mers in the assembly's own hash), GC% and the
// validity statistics -> <prefix>-stats.tsv.  Both hashes are counted (or loaded) like any other input — NON-canonically:
// the reference never sets InputHandler::canonical for cold (lib/include/kat/input_handler.hpp:48 defaults it to false) —
// and the two JellyfishHelper::getCount calls per k-mer position (src/cold.cc:348-352) become two batched katgpu_lookup
// passes.  Plotting is outside this build.
#include "kat_seqlookup.hpp"

#include <iomanip>

namespace kat_front {

int cold_main(int argc, char **argv) {
    const vector<OptSpec> specs = {{'o', "output_prefix", true}, {'x', "gc_bins", true}, {'y', "cvg_bins", true}, {'t', "threads", true},
                                   {0, "5ptrim", true}, {'m', "mer_len", true}, {'H', "hash_size", true}, {'d', "dump_hashes", false},
                                   {'g', "disable_hash_grow", false}, {'p', "output_type", true}, {'v', "verbose", false}, {0, "gpus", true},
                                   {0, "help", false}};
    const Parsed a = parse_args(argc, argv, specs);
    check_path_words(a, "asm_file", "reads_files");
    if (a.has("help") || argc <= 1) {
        cout << "Usage: kat cold [options] <assembly> (<reads>)+\n\nCalculates contig level coverage and GC% from reads and an assembly.\n\n"
                "Options: -o prefix (kat-cold) -t threads --5ptrim list -m mer_len (27) -H hash_size (1e8) -d -v --help\n" << endl;
        return 1;
    }
    Timer total;
    cout << "Running KAT in Cold mode" << endl << "------------------------" << endl << endl;
    if (a.positional.size() < 2) throw KatError(4, "kat cold needs an assembly and at least one reads input");
    int k = (int)to_u64(a.get("mer_len", "27"), "mer_len");
    check_mer_len(k, to_u64(a.get("hash_size", "100000000"), "hash_size"));
    const uint64_t hash_size = to_u64(a.get("hash_size", "100000000"), "hash_size");
    Input reads, assembly;
    reads.files.assign(a.positional.begin() + 1, a.positional.end());
    reads.set_trim(parse_trim(a.get("5ptrim", "0")));
    reads.canonical = false;
    reads.threads = (unsigned)to_u64(a.get("threads", "1"), "threads");
    reads.hash_size = hash_size; // Cold::setHashSize, src/cold.hpp:151-154
    reads.index = 1;
    assembly.files.assign(1, a.positional[0]);
    assembly.set_trim({});
    assembly.canonical = false;
    assembly.threads = (unsigned)to_u64(a.get("threads", "1"), "threads");
    assembly.hash_size = hash_size / 2;
    assembly.index = 1;
    const string prefix = a.get("output_prefix", "kat-cold");
    reads.validate();
    assembly.validate();
    ensure_parent_dir(prefix);
    Ctx ctx(gpus_option(a));
    check_wide_limits(k, ctx.n());
    if (a.has("dump_hashes") && ctx.n() > 1) throw KatError(4, "--dump_hashes writes one GPU's table: run without --gpus");
    int k_reads = k, k_asm = k;
    if (reads.load_mode) load_inputs_phase(ctx, &reads, 1, k_reads);
    else count_input(ctx, reads, k, true);
    if (assembly.load_mode) load_inputs_phase(ctx, &assembly, 1, k_asm);
    else count_input(ctx, assembly, k, true);
    if (k_reads != k_asm) throw KatError(4, "Cannot process hashes that were created with different K-mer lengths.");
    k = k_reads;

    Timer t;
    cout << "Calculating kmer coverage across sequences ...";
    cout.flush();
    const vector<SeqRecord> recs = read_records(assembly.files[0]);
    LookupWindow W(ctx, {&reads, &assembly}, recs, (uint32_t)k);
    std::ofstream stats((prefix + "-stats.tsv").c_str());
    stats << "seq_name\tread_median_cvg\tread_mean_cvg\tasm_cn\tgc%\tseq_length\tkmers_in_seq\tinvalid_kmers\t%_invalid\tnon_zero_kmers\t%_non_zero\t%_non_zero_corrected" << endl;
    stats << std::fixed << std::setprecision(5);
    for (size_t i = 0; i < recs.size(); i++) { // Cold::processSeq, src/cold.cc:303-410; printStatTable, :253-271
        W.ensure(i);
        const vector<uint64_t> &in_reads = W.found[0], &in_asm = W.found[1];
        const string &s = recs[i].seq;
        const uint64_t len = s.size();
        const int64_t nb = (int64_t)len - k + 1;
        uint64_t non_zero = 0, invalid = 0, sum = 0;
        uint32_t median = 0, asm_cn = 0;
        double mean = 0.0;
        if (nb > 0) {
            vector<uint64_t> rc((size_t)nb, 0), ac((size_t)nb, 0);
            for (size_t p = 0; p < (size_t)nb; p++) {
                const size_t q = W.at(i) + p;
                if (!W.P.valid[q]) {
                    invalid++;
                    continue;
                }
                rc[p] = in_reads[q];
                ac[p] = in_asm[q];
                sum += rc[p];
                if (rc[p] != 0) non_zero++;
            }
            std::sort(rc.begin(), rc.end());
            median = (uint32_t)(double)rc[rc.size() / 2];
            mean = (double)sum / (double)nb;
            std::sort(ac.begin(), ac.end());
            asm_cn = (uint32_t)(double)ac[ac.size() / 2];
        }
        const double p_non_zero = non_zero == 0 || nb <= 0 ? 0.0 : ((double)non_zero / (double)nb) * 100.0;
        const double p_invalid = invalid == 0 || nb <= 0 ? 0.0 : ((double)invalid / (double)nb) * 100.0;
        const uint64_t not_invalid = (uint64_t)nb - invalid;
        const double p_corr = non_zero == 0 || not_invalid <= 0 ? 0.0 : ((double)non_zero / (double)not_invalid) * 100.0;
        uint64_t gs = 0, cs = 0, ns = 0;
        for (char c : s) {
            if (c == 'G' || c == 'g') gs++;
            else if (c == 'C' || c == 'c') cs++;
            else if (c == 'N' || c == 'n') ns++;
        }
        const double gc = ((double)(gs + cs)) / ((double)(len - ns));
        stats << recs[i].name << "\t" << median << "\t" << mean << "\t" << asm_cn << "\t" << gc << "\t" << (uint32_t)len << "\t"
              << (uint32_t)((uint32_t)len - (uint32_t)k + 1) << "\t" << (uint32_t)invalid << "\t" << p_invalid << "\t" << (uint32_t)non_zero << "\t" << p_non_zero
              << "\t" << p_corr << endl;
    }
    phase_done(t);
    if (a.has("dump_hashes")) {
        dump_input(reads, k, prefix + "-reads_hash.jf" + std::to_string(k));
        dump_input(assembly, k, prefix + "-asm_hash.jf" + std::to_string(k));
    }
    free_tables(ctx, reads);
    free_tables(ctx, assembly);
    char b[96];
    snprintf(b, sizeof b, "KAT CoLD completed.\nTotal runtime: %.1fs\n\n", total.secs());
    cout << b;
    return 0;
}

} // namespace kat_front
