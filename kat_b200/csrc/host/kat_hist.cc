// kat_hist.cc — `kat hist` of the GPU front-end (src/histogram.cc:257-369: options, Histogram ctor, output file).
#include "kat_front.hpp"

namespace kat_front {

// =============================== kat hist ========================================
int hist_main(int argc, char **argv) {
    const vector<OptSpec> specs = {{'o', "output_prefix", true}, {'t', "threads", true}, {'l', "low", true}, {'h', "high", true},
                                   {'i', "inc", true}, {0, "5ptrim", true}, {'N', "non_canonical", false}, {'m', "mer_len", true},
                                   {'H', "hash_size", true}, {'d', "dump_hash", false}, {'p', "output_type", true},
                                   {'v', "verbose", false}, {0, "gpus", true}, {0, "help", false}};
    const Parsed a = parse_args(argc, argv, specs);
    check_path_words(a, "inputs", "inputs");
    if (a.has("help") || argc <= 1) {
        cout << "Usage: kat hist [options] (<input>)+\n\nCreate an histogram of k-mer occurrences from the input.\n\n"
                "Options: -o prefix (kat.hist) -t threads -l low (1) -h high (10000) -i inc (1) --5ptrim list -N -m mer_len (27) -H hash_size (1e8) -v --help\n" << endl;
        return 1;
    }
    const uint64_t low = to_u64(a.get("low", "1"), "low"), high = to_u64(a.get("high", "10000"), "high"), inc = to_u64(a.get("inc", "1"), "inc");
    int k = (int)to_u64(a.get("mer_len", "27"), "mer_len");
    Timer total;
    cout << "Running KAT in HIST mode" << endl << "------------------------" << endl << endl;
    Input in;
    in.files = a.positional;
    in.set_trim(parse_trim(a.get("5ptrim", "0")));
    in.canonical = !a.has("non_canonical");
    in.threads = (unsigned)to_u64(a.get("threads", "1"), "threads");
    in.hash_size = to_u64(a.get("hash_size", "100000000"), "hash_size");
    const string prefix = a.get("output_prefix", "kat.hist");
    // Histogram ctor, src/histogram.cc:57-71, src/histogram.hpp:172-178
    const uint64_t base = low > 1 ? (1 >= low ? 1 : low - 1) : 1, ceil = high + 1, nb = ceil + 1 - base;
    if (high < low) throw KatError(4, "High count value must be >= to low count value.  High: " + std::to_string(high) + "; Low: " + std::to_string(low));
    if (inc == 0) throw KatError(4, "Increment must be >= 1");
    check_mer_len(k, to_u64(a.get("hash_size", "100000000"), "hash_size"));
    in.validate();
    ensure_parent_dir(prefix);
    Ctx ctx(gpus_option(a));
    check_wide_limits(k, ctx.n());
    if (a.has("dump_hash") && ctx.n() > 1) throw KatError(4, "--dump_hash writes one GPU's table: run without --gpus");
    if (in.load_mode) load_inputs_phase(ctx, &in, 1, k);
    else count_input(ctx, in, k, true);
    vector<uint64_t> bins(nb);
    {
        Timer t;
        cout << "Bining kmers ...";
        cout.flush();
        vector<vector<uint64_t>> other(ctx.n()); // every GPU of a job receives the all-reduced bins
        run_on_all(ctx.n(), [&](int r) {
            if (r) other[r].resize(nb);
            gpu_check(katgpu_hist(in.tables[r], base, ceil, inc, nb, r ? other[r].data() : bins.data()));
        });
        phase_done(t);
    }
    if (a.has("dump_hash")) dump_input(in, k, prefix + "-hash.jf" + std::to_string(k));
    {
        Timer t;
        cout << "Saving results to disk ...";
        cout.flush();
        std::ofstream out(prefix.c_str());
        out << "# Title:" << k << "-mer spectra for: " << join(in.files, true) << endl;
        out << "# XLabel:" << k << "-mer frequency" << endl;
        out << "# YLabel:# distinct " << k << "-mers" << endl;
        out << "# Kmer value:" << k << endl;
        out << "# Input 1:" << join(in.files, false) << endl;
        out << "###" << endl;
        uint64_t col = base;
        for (uint64_t i = 0; i < nb; i++, col += inc) out << col << " " << bins[i] << "\n";
        out.close();
        phase_done(t);
    }
    free_tables(ctx, in);
    char b[96];
    snprintf(b, sizeof b, "KAT HIST completed.\nTotal runtime: %.1fs\n\n", total.secs());
    cout << b;
    return 0;
}

} // namespace kat_front
