// kat_gcp.cc — `kat gcp` of the GPU front-end (src/gcp.cc:255-370).
#include "kat_front.hpp"

namespace kat_front {

// =============================== kat gcp =========================================
int gcp_main(int argc, char **argv) {
    const vector<OptSpec> specs = {{'o', "output_prefix", true}, {'t', "threads", true}, {'x', "cvg_scale", true}, {'y', "cvg_bins", true},
                                   {0, "5ptrim", true}, {'N', "non_canonical", false}, {'m', "mer_len", true}, {'H', "hash_size", true},
                                   {'d', "dump_hash", false}, {'p', "output_type", true}, {'v', "verbose", false}, {0, "gpus", true}, {0, "help", false}};
    const Parsed a = parse_args(argc, argv, specs);
    check_path_words(a, "inputs", "inputs");
    if (a.has("help") || argc <= 1) {
        cout << "Usage: kat gcp [options] (<input>)+\n\nCompares GC content and K-mer coverage from the input.\n\n"
                "Options: -o prefix (kat-gcp) -t threads -x cvg_scale (1.0) -y cvg_bins (1000) --5ptrim list -N -m mer_len (27) -H hash_size (1e8) -v --help\n" << endl;
        return 1;
    }
    const double scale = to_double(a.get("cvg_scale", "1.0"), "cvg_scale");
    const uint32_t bins = (uint32_t)to_u64(a.get("cvg_bins", "1000"), "cvg_bins");
    int k = (int)to_u64(a.get("mer_len", "27"), "mer_len");
    Timer total;
    cout << "Running KAT in GCP mode" << endl << "------------------------" << endl << endl;
    Input in;
    in.files = a.positional;
    in.set_trim(parse_trim(a.get("5ptrim", "0")));
    in.canonical = !a.has("non_canonical");
    in.threads = (unsigned)to_u64(a.get("threads", "1"), "threads");
    in.hash_size = to_u64(a.get("hash_size", "100000000"), "hash_size");
    const string prefix = a.get("output_prefix", "kat-gcp");
    check_mer_len(k, to_u64(a.get("hash_size", "100000000"), "hash_size"));
    in.validate();
    ensure_parent_dir(prefix);
    Ctx ctx(gpus_option(a));
    check_wide_limits(k, ctx.n());
    if (a.has("dump_hash") && ctx.n() > 1) throw KatError(4, "--dump_hash writes one GPU's table: run without --gpus");
    if (in.load_mode) load_inputs_phase(ctx, &in, 1, k);
    else count_input(ctx, in, k, true);
    vector<uint64_t> mx((size_t)k * (bins + 1));
    {
        Timer t;
        cout << "Analysing kmers in hash ...";
        cout.flush();
        vector<vector<uint64_t>> other(ctx.n());
        run_on_all(ctx.n(), [&](int r) {
            if (r) other[r].resize(mx.size());
            gpu_check(katgpu_gcp(in.tables[r], scale, bins, r ? other[r].data() : mx.data()));
        });
        phase_done(t);
        if (a.has("dump_hash")) dump_input(in, k, prefix + "-hash.jf" + std::to_string(k));
    }
    {
        Timer t;
        cout << "Saving results to disk ...";
        cout.flush();
        std::ofstream out((prefix + ".mx").c_str());
        out << "# Title:K-mer coverage vs GC count plot for: " << join(in.files, true) << endl;
        out << "# XLabel:" << k << "-mer frequency" << endl;
        out << "# YLabel:GC count" << endl;
        out << "# ZLabel:# distinct " << k << "-mers" << endl;
        // the matrix has `cvgBins + 1` columns in 16-bit arithmetic (src/gcp.cc:93) — none for 65535: nothing is ever found in it
        // (inc() has no bounds check), the header says so, and printMatrix's first get() throws SparseMatrixException
        // (boost::exception, exit code 4) with the header already in the file
        const bool no_columns = bins == 65535;
        out << "# Columns:" << (no_columns ? 0 : bins + 1) << endl;
        out << "# Rows:" << k << endl;
        out << "# MaxVal:" << (no_columns ? 0 : max_val(mx)) << endl;
        out << "# Transpose:0" << endl;
        out << "# Kmer value:" << k << endl;
        out << "# Input 1:" << join(in.files, false) << endl;
        out << "###" << endl;
        if (no_columns) {
            out.close();
            throw KatError(4, "Requested coords exceed limits of matrix.  Coords: 0,0.  Limits: " + std::to_string(k) + ",0");
        }
        print_matrix(out, mx, (uint32_t)k, bins + 1);
        out.close();
        phase_done(t);
    }
    free_tables(ctx, in);
    char b[96];
    snprintf(b, sizeof b, "KAT GCP completed.\nTotal runtime: %.1fs\n\n", total.secs());
    cout << b;
    return 0;
}

} // namespace kat_front
