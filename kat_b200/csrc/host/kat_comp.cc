// kat_comp.cc — `kat comp` of the GPU front-end (src/comp.cc:633-843, lib/src/comp_counters.cc:144-206).
#include "kat_front.hpp"

namespace kat_front {

// =============================== kat comp ========================================
void print_comp_counts(std::ostream &out, const katgpu_comp_counters &c, const string &p1, const string &p2, const string &p3,
                       const vector<uint64_t> &sp1, const vector<uint64_t> &sp2, const vector<uint64_t> &ssp1,
                       const vector<uint64_t> &ssp2) { // CompCounters::printCounts, lib/src/comp_counters.cc:144-206
    out << "K-mer statistics for: " << endl;
    out << " - Hash 1: " << quoted_path(p1) << endl;
    out << " - Hash 2: " << quoted_path(p2) << endl;
    if (c.hash3_total > 0) out << " - Hash 3: " << quoted_path(p3) << endl;
    out << endl;
    out << "Total K-mers in: " << endl;
    out << " - Hash 1: " << c.hash1_total << endl;
    out << " - Hash 2: " << c.hash2_total << endl;
    if (c.hash3_total > 0) out << " - Hash 3: " << c.hash3_total << endl;
    out << endl;
    out << "Distinct K-mers in:" << endl;
    out << " - Hash 1: " << c.hash1_distinct << endl;
    out << " - Hash 2: " << c.hash2_distinct << endl;
    if (c.hash3_total > 0) out << " - Hash 3: " << c.hash3_distinct << endl;
    out << endl;
    out << "Total K-mers only found in:" << endl;
    out << " - Hash 1: " << c.hash1_only_total << endl;
    out << " - Hash 2: " << c.hash2_only_total << endl;
    out << endl;
    out << "Distinct K-mers only found in:" << endl;
    out << " - Hash 1: " << c.hash1_only_distinct << endl;
    out << " - Hash 2: " << c.hash2_only_distinct << endl << endl;
    out << "Shared K-mers:" << endl;
    out << " - Total shared found in hash 1: " << c.shared_hash1_total << endl;
    out << " - Total shared found in hash 2: " << c.shared_hash2_total << endl;
    out << " - Distinct shared K-mers: " << c.shared_distinct << endl << endl;
    out << "Distance between spectra 1 and 2 (all k-mers):" << endl;
    print_distances(out, sp1, sp2);
    out << endl;
    out << "Distance between spectra 1 and 2 (shared k-mers):" << endl;
    print_distances(out, ssp1, ssp2);
    out << endl;
}

int comp_main(int argc, char **argv) {
    const vector<OptSpec> specs = {{'o', "output_prefix", true}, {'t', "threads", true}, {'x', "d1_scale", true}, {'y', "d2_scale", true},
                                   {'i', "d1_bins", true}, {'j', "d2_bins", true}, {0, "d1_5ptrim", true}, {0, "d2_5ptrim", true},
                                   {'N', "non_canonical_1", false}, {'O', "non_canonical_2", false}, {'P', "non_canonical_3", false},
                                   {'m', "mer_len", true}, {'H', "hash_size_1", true}, {'I', "hash_size_2", true}, {'J', "hash_size_3", true},
                                   {'d', "dump_hashes", false}, {'g', "disable_hash_grow", false}, {'n', "density_plot", false},
                                   {'p', "output_type", true}, {'h', "output_hists", false}, {'v', "verbose", false}, {0, "gpus", true}, {0, "help", false}};
    const Parsed a = parse_args(argc, argv, specs);
    if (a.has("help") || argc <= 1) {
        cout << "Usage: kat comp [options] <input_1> <input_2> [<input_3>]\n\nCompares jellyfish K-mer count hashes.\n\n"
                "Options: -o prefix (kat-comp) -t threads -x d1_scale -y d2_scale -i d1_bins (1001) -j d2_bins (1001) --d1_5ptrim --d2_5ptrim "
                "-N -O -P -m mer_len (27) -H -I -J hash sizes (1e8) -g -n -h -v --help\n" << endl;
        return 1;
    }
    Timer total;
    cout << "Running KAT in COMP mode" << endl << "------------------------" << endl << endl;
    if (a.positional.size() > 3) throw KatError(1, "Error: Parsing Command Line: too many positional options have been specified on the command line");
    if (a.positional.size() < 1) throw KatError(4, "Nothing specified for input group 1");
    if (a.positional.size() < 2) throw KatError(4, "Nothing specified for input group 2");
    const bool three = a.positional.size() >= 3;
    int k = (int)to_u64(a.get("mer_len", "27"), "mer_len");
    check_mer_len(k, to_u64(a.get("hash_size_1", "100000000"), "hash_size_1"));
    const double d1s = to_double(a.get("d1_scale", "1.0"), "d1_scale"), d2s = to_double(a.get("d2_scale", "1.0"), "d2_scale");
    const uint32_t d1b = (uint32_t)to_u64(a.get("d1_bins", "1001"), "d1_bins"), d2b = (uint32_t)to_u64(a.get("d2_bins", "1001"), "d2_bins");
    if (d1b == 0 || d2b == 0) throw KatError(4, "d1_bins and d2_bins must be >= 1");
    Input in[3];
    const char *nc[3] = {"non_canonical_1", "non_canonical_2", "non_canonical_3"};
    const char *hs[3] = {"hash_size_1", "hash_size_2", "hash_size_3"};
    for (int i = 0; i < (three ? 3 : 2); i++) {
        in[i].files = glob_group(a.positional[i]);
        in[i].index = i + 1;
        in[i].canonical = !a.has(nc[i]);
        in[i].threads = (unsigned)to_u64(a.get("threads", "1"), "threads");
        in[i].hash_size = to_u64(a.get(hs[i], "100000000"), hs[i]);
    }
    in[0].set_trim(parse_trim(a.get("d1_5ptrim", "0")));
    in[1].set_trim(parse_trim(a.get("d2_5ptrim", "0")));
    if (three) in[2].set_trim({});
    const string prefix = a.get("output_prefix", "kat-comp");
    for (int i = 0; i < (three ? 3 : 2); i++) in[i].validate();
    ensure_parent_dir(prefix);
    Ctx ctx(gpus_option(a));
    check_wide_limits(k, ctx.n());
    if (a.has("dump_hashes") && ctx.n() > 1) throw KatError(4, "--dump_hashes writes one GPU's table: run without --gpus");
    if (ctx.n() > 1)
        for (int i = 0; i < (three ? 3 : 2); i++)
            if (!in[i].canonical) throw KatError(4, "with --gpus every input must be counted canonically (no -N/-O/-P): the shared-k-mer walk is local to the GPU that owns a k-mer");
    const int n_in = three ? 3 : 2;
    // Comp::execute (src/comp.cc:140-167): count the sequence inputs, then read the headers of
    // the hash inputs; when every input is a hash the mer length comes from the first header;
    // every hash must hold k-mers of the mer length in use (InputHandler::validateMerLen).
    for (int i = 0; i < n_in; i++)
        if (!in[i].load_mode) count_input(ctx, in[i], k, !a.has("disable_hash_grow"));
    bool any_load = false, all_load = true;
    vector<unsigned> hash_k(n_in, 0);
    for (int i = 0; i < n_in; i++) {
        if (in[i].load_mode) {
            std::ifstream f;
            hash_k[i] = read_jf_header(in[i].files[0], f).key_len / 2;
            any_load = true;
        } else
            all_load = false;
    }
    if (all_load) k = (int)hash_k[0];
    for (int i = 0; i < n_in; i++)
        if (in[i].load_mode && (int)hash_k[i] != k)
            throw KatError(4, "Cannot process hashes that were created with different K-mer lengths.  Expected: " + std::to_string(k) +
                                  ".  Key length was " + std::to_string(hash_k[i]) + " for : " + in[i].files[0]);
    if (any_load) {
        int k_unused = k;
        load_inputs_phase(ctx, in, n_in, k_unused);
    }
    const size_t cells = (size_t)d1b * d2b, len = std::min(d1b, d2b);
    vector<uint64_t> main_mx(cells), ends, middle, mixed, spectra(4 * len);
    if (three) {
        ends.resize(cells);
        middle.resize(cells);
        mixed.resize(cells);
    }
    katgpu_comp_counters cc;
    {
        Timer t;
        cout << "Comparing hashes ...";
        cout.flush();
        run_on_all(ctx.n(), [&](int r) {
            if (r == 0) {
                gpu_check(katgpu_comp(in[0].tables[0], in[1].tables[0], three ? in[2].tables[0] : nullptr, d1s, d2s, d1b, d2b, main_mx.data(),
                                      three ? ends.data() : nullptr, three ? middle.data() : nullptr, three ? mixed.data() : nullptr, &cc,
                                      spectra.data()));
                return;
            }
            vector<uint64_t> m(cells), e, mi, mx2, sp(4 * len); // the other GPUs' copies of the all-reduced result
            if (three) e.resize(cells), mi.resize(cells), mx2.resize(cells);
            katgpu_comp_counters c2;
            gpu_check(katgpu_comp(in[0].tables[r], in[1].tables[r], three ? in[2].tables[r] : nullptr, d1s, d2s, d1b, d2b, m.data(),
                                  three ? e.data() : nullptr, three ? mi.data() : nullptr, three ? mx2.data() : nullptr, &c2, sp.data()));
        });
        phase_done(t);
    }
    if (a.has("dump_hashes")) // src/comp.cc:174-179
        for (int i = 0; i < n_in; i++) dump_input(in[i], k, prefix + "-hash" + std::to_string(in[i].index) + ".jf" + std::to_string(k));
    const vector<uint64_t> sp1(spectra.begin(), spectra.begin() + len), sp2(spectra.begin() + len, spectra.begin() + 2 * len),
        ssp1(spectra.begin() + 2 * len, spectra.begin() + 3 * len), ssp2(spectra.begin() + 3 * len, spectra.end());
    const string p1 = in[0].files[0], p2 = in[1].files[0], p3 = three ? in[2].files[0] : "";
    {
        Timer t;
        cout << "Saving results to disk ...";
        cout.flush();
        {
            std::ofstream out((prefix + "-main.mx").c_str()); // Comp::printMainMatrix, src/comp.cc:308-326
            out << "# Title:K-mer comparison plot" << endl
                << "# XLabel:" << k << "-mer frequency for: " << join(in[0].files, true) << endl
                << "# YLabel:" << k << "-mer frequency for: " << join(in[1].files, true) << endl
                << "# ZLabel:# distinct " << k << "-mers" << endl
                << "# Columns:" << d2b << endl
                << "# Rows:" << d1b << endl
                << "# MaxVal:" << max_val(main_mx) << endl
                << "# Transpose:1" << endl
                << "# Kmer value:" << k << endl
                << "# Input 1:" << join(in[0].files, false) << endl
                << "# Input 2:" << join(in[1].files, false) << endl
                << "###" << endl;
            print_matrix(out, main_mx, d1b, d2b);
        }
        if (three) { // src/comp.cc:330-356
            std::ofstream e((prefix + "-ends.mx").c_str());
            e << "# Each row represents K-mer frequency for: " << p1 << endl;
            e << "# Each column represents K-mer frequency for sequence ends: " << p3 << endl;
            print_matrix(e, ends, d1b, d2b);
            std::ofstream m((prefix + "-middle.mx").c_str());
            m << "# Each row represents K-mer frequency for: " << p1 << endl;
            m << "# Each column represents K-mer frequency for sequence middles: " << p2 << endl;
            print_matrix(m, middle, d1b, d2b);
            std::ofstream x((prefix + "-mixed.mx").c_str());
            x << "# Each row represents K-mer frequency for hash file 1: " << p1 << endl;
            x << "# Each column represents K-mer frequency for mixed: " << p2 << " and " << p3 << endl;
            print_matrix(x, mixed, d1b, d2b);
        }
        {
            std::ofstream out((prefix + ".stats").c_str());
            print_comp_counts(out, cc, p1, p2, p3, sp1, sp2, ssp1, ssp2);
        }
        if (a.has("output_hists")) { // Comp::printHist, src/comp.cc:235-246
            for (int i = 0; i < 2; i++) {
                std::ofstream out((prefix + "." + std::to_string(i + 1) + ".hist").c_str());
                out << "# Title:" << k << "-mer spectra for: " << join(in[i].files, false) << endl;
                out << "# XLabel:" << k << "-mer frequency" << endl;
                out << "# YLabel:# distinct " << k << "-mers" << endl;
                out << "###" << endl;
                const vector<uint64_t> &h = i == 0 ? sp1 : sp2;
                for (uint64_t j = 0; j < h.size(); j++) out << j << " " << h[j] << "\n";
            }
        }
        phase_done(t);
    }
    cout << endl << "Summary statistics" << endl << "------------------" << endl << endl;
    print_comp_counts(cout, cc, p1, p2, p3, sp1, sp2, ssp1, ssp2);
    for (int i = 0; i < (three ? 3 : 2); i++) free_tables(ctx, in[i]);
    char b[96];
    snprintf(b, sizeof b, "KAT COMP completed.\nTotal runtime: %.1fs\n\n", total.secs());
    cout << b;
    return 0;
}

} // namespace kat_front
