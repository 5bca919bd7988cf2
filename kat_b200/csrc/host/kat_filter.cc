// kat_filter.cc — `kat filter kmer` and `kat filter seq` of the GPU front-end (src/filter.cc:66-131 dispatch).
//
//   filter kmer  (src/filter_kmer.cc:119-309): count or load a hash, keep the k-mers whose count and GC count lie
//                inside (or, with -i, outside) the bounds, write them as a jellyfish hash <prefix>-in.jf<k>
//                (and the rest as <prefix>-out.jf<k> with -s).  The walk over the table with the bounds test
//                (FilterKmer::filterSlice / inBounds, :257-309) runs on the GPU: katgpu_table_export_filtered(_wide).
//   filter seq   (src/filter_sequence.cc:106-368): count or load a hash, look up every k-mer position of every
//                record of one (or two, paired) sequence file(s) — FilterSeq::getProfile, one
//                JellyfishHelper::getCount per position (:353-366) — as batched katgpu_lookup calls, and write
//                the records whose fraction of found k-mers reaches the threshold (or not, -i) to
//                <prefix>.in[.R1|.R2]<ext>, the others to <prefix>.out… with -s, per-record statistics to
//                <prefix>.stats with --stats.  Records are written the way SeqAn 2.0's writeRecord does
//                (FASTA wrapped at 70 columns, FASTQ unwrapped with a bare '+' line: seq_io/fasta_fastq.h:448-505).
#include "kat_seqlookup.hpp"

#include <random>

namespace kat_front {

namespace {

// ---------------------------------------------------------------- filter kmer --
int filter_kmer_main(int argc, char **argv) {
    const vector<OptSpec> specs = {{'o', "output_prefix", true}, {'t', "threads", true}, {'c', "low_count", true}, {'d', "high_count", true},
                                   {'g', "low_gc", true}, {'h', "high_gc", true}, {'i', "invert", false}, {'s', "separate", false},
                                   {'N', "non_canonical", false}, {'m', "mer_len", true}, {'H', "hash_size", true}, {'v', "verbose", false},
                                   {0, "gpus", true}, {0, "help", false}};
    const Parsed a = parse_args(argc, argv, specs);
    check_path_words(a, "inputs", "inputs");
    if (a.has("help") || argc <= 1) {
        cout << "Usage: kat filter kmer [options] (<input>)+\n\nFilter k-mers to those within defined count and GC bounds.\n\n"
                "Options: -o prefix (kat.filter.kmer) -t threads -c low_count (1) -d high_count (10000) -g low_gc (1) -h high_gc (100) -i -s -N "
                "-m mer_len (27) -H hash_size (1e8) -v --help\n" << endl;
        return 1;
    }
    Timer total;
    cout << "Running KAT in filter kmer mode" << endl << "-------------------------------" << endl << endl;
    int k = (int)to_u64(a.get("mer_len", "27"), "mer_len");
    check_mer_len(k, to_u64(a.get("hash_size", "100000000"), "hash_size"));
    const uint64_t low_count = to_u64(a.get("low_count", "1"), "low_count"), high_count = to_u64(a.get("high_count", "10000"), "high_count");
    const uint32_t low_gc = (uint32_t)to_u64(a.get("low_gc", "1"), "low_gc"), high_gc = (uint32_t)to_u64(a.get("high_gc", "100"), "high_gc");
    const bool invert = a.has("invert"), separate = a.has("separate");
    if (high_count < low_count) throw KatError(4, "High kmer count value must be >= to low kmer count value");
    if (high_gc < low_gc) throw KatError(4, "High GC count value must be >= to low GC count value");
    if (a.positional.empty()) throw KatError(4, "No input provided for this input group");
    Input in;
    in.files = a.positional; // setMultipleInputs: every positional argument is a file of the one input group
    in.set_trim({});
    in.canonical = !a.has("non_canonical");
    in.threads = (unsigned)to_u64(a.get("threads", "1"), "threads");
    in.hash_size = to_u64(a.get("hash_size", "100000000"), "hash_size");
    const string prefix = a.get("output_prefix", "kat.filter.kmer");
    in.validate();
    ensure_parent_dir(prefix);
    Ctx ctx(gpus_option(a));
    if (in.load_mode) load_inputs_phase(ctx, &in, 1, k);
    else count_input(ctx, in, k, true);

    // FilterKmer::filter: one walk per output set; on a job every GPU hands back its own part of the k-mers
    Timer t;
    cout << "Filtering kmers ...";
    cout.flush();
    struct Set {
        vector<uint64_t> keys_hi, keys, counts; // keys_hi: the first k-32 bases, all zero for k <= 32
        uint64_t total = 0;
    };
    auto select = [&](bool inside) {
        vector<Set> part(ctx.n());
        run_on_all(ctx.n(), [&](int r) {
            uint64_t n = 0;
            gpu_check(katgpu_table_export_filtered_wide(in.tables[r], low_count, high_count, low_gc, high_gc, inside ? 1 : 0, nullptr, nullptr, nullptr, 0, &n));
            part[r].keys.resize((size_t)std::max<uint64_t>(n, 1));
            part[r].keys_hi.resize(part[r].keys.size());
            part[r].counts.resize(part[r].keys.size());
            gpu_check(katgpu_table_export_filtered_wide(in.tables[r], low_count, high_count, low_gc, high_gc, inside ? 1 : 0, part[r].keys_hi.data(),
                                                        part[r].keys.data(), part[r].counts.data(), (size_t)n, &n));
            part[r].keys.resize((size_t)n);
            part[r].keys_hi.resize((size_t)n);
            part[r].counts.resize((size_t)n);
        });
        Set s;
        for (auto &p : part) {
            s.keys_hi.insert(s.keys_hi.end(), p.keys_hi.begin(), p.keys_hi.end());
            s.keys.insert(s.keys.end(), p.keys.begin(), p.keys.end());
            s.counts.insert(s.counts.end(), p.counts.begin(), p.counts.end());
        }
        for (uint64_t c : s.counts) s.total += c;
        return s;
    };
    // not separating: the kept set is the in-bounds one unless inverted; separating: in-bounds -> in, the rest -> out,
    // and --invert has no effect (FilterKmer::filterSlice, src/filter_kmer.cc:266-286)
    Set keep = select(separate ? true : !invert), discard;
    if (separate) discard = select(false);
    phase_done(t);
    cout << "K-mers in input   : " << in.n_distinct << " distinct; " << in.n_total << " total." << endl
         << "K-mers to keep    : " << keep.keys.size() << " distinct; " << keep.total << " total." << endl;
    if (separate) cout << "K-mers to discard : " << discard.keys.size() << " distinct; " << discard.total << " total." << endl;
    cout << endl;

    auto dump = [&](Set &s, const string &path) { // FilterKmer::dump, :203-219
        struct stat st;
        if (lstat(path.c_str(), &st) == 0) unlink(path.c_str());
        Timer td;
        cout << "Dumping hash to " << path << " ...";
        cout.flush();
        kat_host::JfHeader h;
        std::ofstream out(path.c_str(), std::ios::out | std::ios::binary);
        if (k > 32) {
            kat_host::jf_prepare_dump_wide(h, (unsigned)k, in.canonical, s.keys.size(), s.keys_hi, s.keys, s.counts);
            kat_host::jf_write_file_wide(out, h, s.keys_hi, s.keys, s.counts);
        } else {
            kat_host::jf_prepare_dump(h, (unsigned)k, in.canonical, s.keys.size(), s.keys, s.counts);
            kat_host::jf_write_file(out, h, s.keys, s.counts);
        }
        out.close();
        if (!out.good()) throw KatError(4, "Could not write " + path);
        phase_done(td);
    };
    dump(keep, prefix + "-in.jf" + std::to_string(k));
    if (separate) dump(discard, prefix + "-out.jf" + std::to_string(k));
    free_tables(ctx, in);
    char b[96];
    snprintf(b, sizeof b, "KAT filter kmer completed.\nTotal runtime: %.1fs\n\n", total.secs());
    cout << b;
    return 0;
}

// ----------------------------------------------------------------- filter seq --
// seqan::SeqFileOut picks the format from the file extension (seq_io/sequence_file.h FileExtensions)
enum class SeqFormat { Fasta, Fastq };
SeqFormat format_of_extension(const string &ext) {
    const string e = lower(ext);
    for (const char *x : {".fa", ".fasta", ".faa", ".ffn", ".fna", ".frn"})
        if (e == x) return SeqFormat::Fasta;
    for (const char *x : {".fq", ".fastq"})
        if (e == x) return SeqFormat::Fastq;
    throw KatError(4, "Cannot write filtered sequences: unsupported file extension '" + ext + "' (fa, fasta, fna, ffn, faa, frn, fq, fastq)");
}
void write_wrapped(std::ostream &out, const string &s, size_t line) { // writeWrappedString, stream/tokenization.h:464-482
    size_t left = s.size(), at = 0;
    do {
        const size_t n = line == 0 ? left : std::min(left, line);
        out.write(s.data() + at, (std::streamsize)n);
        out.put('\n');
        at += n, left -= n;
    } while (left != 0);
}
void write_record(std::ostream &out, SeqFormat f, const SeqRecord &r) {
    if (f == SeqFormat::Fasta) {
        out << '>' << r.name << '\n';
        write_wrapped(out, r.seq, 70);
    } else {
        out << '@' << r.name << '\n';
        write_wrapped(out, r.seq, 0);
        out << "+\n";
        write_wrapped(out, r.qual, 0);
    }
}

int filter_seq_main(int argc, char **argv) {
    const vector<OptSpec> specs = {{'o', "output_prefix", true}, {'t', "threads", true}, {'T', "threshold", true}, {'i', "invert", false},
                                   {'s', "separate", false}, {0, "seq", true}, {0, "seq2", true}, {'f', "frequency", true}, {0, "stats", false},
                                   {'N', "non_canonical", false}, {'m', "mer_len", true}, {'H', "hash_size", true}, {'v', "verbose", false},
                                   {0, "gpus", true}, {0, "help", false}};
    const Parsed a = parse_args(argc, argv, specs);
    check_path_words(a, "inputs", "inputs", {"output_prefix", "seq", "seq2"});
    if (a.has("help") || argc <= 1) {
        cout << "Usage: kat filter seq [options] --seq <seq_file_to_filter> [--seq2 <seq_file_2>] (<input>)+\n\n"
                "Filter sequences based on whether those sequences contain specific k-mers.\n\n"
                "Options: -o prefix (kat.filter.kmer) -t threads -T threshold (0.1) -i -s --seq file --seq2 file -f frequency (0) --stats -N "
                "-m mer_len (27) -H hash_size (1e8) -v --help\n" << endl;
        return 1;
    }
    const string seq1 = a.get("seq", ""), seq2 = a.get("seq2", "");
    if (seq1.empty()) throw KatError(4, "You must specify at least one sequence file to filter");
    Timer total;
    cout << "Running KAT in filter sequence mode" << endl << "-----------------------------------" << endl << endl;
    const bool paired = !seq2.empty();
    if (!exists(seq1)) throw KatError(4, "Could not find input file at: " + seq1 + "; please check the path and try again.");
    if (paired && !exists(seq2)) throw KatError(4, "Could not find sequence file at: " + seq2 + "; please check the path and try again.");
    int k = (int)to_u64(a.get("mer_len", "27"), "mer_len");
    check_mer_len(k, to_u64(a.get("hash_size", "100000000"), "hash_size"));
    const double threshold = to_double(a.get("threshold", "0.1"), "threshold"), frequency = to_double(a.get("frequency", "0"), "frequency");
    const bool invert = a.has("invert"), separate = a.has("separate"), do_stats = a.has("stats");
    if (a.positional.empty()) throw KatError(4, "No input provided for this input group");
    Input in;
    in.files = a.positional;
    in.set_trim({});
    in.canonical = !a.has("non_canonical");
    in.threads = (unsigned)to_u64(a.get("threads", "1"), "threads");
    in.hash_size = to_u64(a.get("hash_size", "100000000"), "hash_size");
    const string prefix = a.get("output_prefix", "kat.filter.kmer");
    in.validate();
    ensure_parent_dir(prefix);
    const string ext = ext_of(seq1); // both mates are written with the first file's extension (:189-205)
    const SeqFormat fmt = format_of_extension(ext);
    Ctx ctx(gpus_option(a));
    check_wide_limits(k, ctx.n());
    if (in.load_mode) load_inputs_phase(ctx, &in, 1, k);
    else count_input(ctx, in, k, true);

    Timer t;
    cout << "Filtering sequences ..." << endl;
    const vector<SeqRecord> r1 = read_records(seq1);
    vector<SeqRecord> r2;
    if (paired) r2 = read_records(seq2);
    LookupWindow w1(ctx, {&in}, r1, (uint32_t)k), w2(ctx, {&in}, r2, (uint32_t)k);
    std::ofstream stats, in1, out1, in2, out2;
    if (do_stats) {
        stats.open((prefix + ".stats").c_str());
        stats << "index\tnb_bases\tnb_kmers\tnb_hits\tratio" << endl;
    }
    in1.open((prefix + ".in" + (paired ? ".R1" : "") + ext).c_str());
    if (separate) out1.open((prefix + ".out" + (paired ? ".R1" : "") + ext).c_str());
    if (paired) {
        in2.open((prefix + ".in.R2" + ext).c_str());
        if (separate) out2.open((prefix + ".out.R2" + ext).c_str());
    }
    std::random_device rd;
    std::mt19937 gen(rd());
    std::uniform_real_distribution<> urd;
    auto hits = [](LookupWindow &w, size_t i, uint64_t &n_pos) { // FilterSeq::getProfile, :330-368
        w.ensure(i);
        uint64_t h = 0;
        const size_t q0 = w.at(i), n = w.n_positions(i);
        for (size_t q = q0; q < q0 + n; q++) h += w.P.valid[q] && w.found[0][q] > 0;
        n_pos += n;
        return h;
    };
    uint64_t keepers = 0, n_total = 0;
    for (size_t i = 0; i < r1.size(); i++) { // FilterSeq::processSeq, :267-328
        // a second file that runs out first: SeqAn's readRecord throws UnexpectedEnd, a std::exception -> "Error: ...", exit code 5
        // (src/filter_sequence.cc:217-219, src/kat.cc:293-295); the records written so far stay in the output files
        if (paired && i >= r2.size()) throw KatError(5, "Unexpected end of input.");
        const double random_val = urd(gen);
        uint64_t nb_kmers = 0;
        uint64_t matches = hits(w1, i, nb_kmers);
        if (paired) matches += hits(w2, i, nb_kmers);
        const double ratio = (double)matches / (double)nb_kmers; // 0 / 0 = NaN: such a record is never kept
        bool keep = true;
        if ((ratio >= threshold && !invert) || (invert && ratio < threshold)) {
            if (frequency > 0.0 && frequency < random_val) keep = false;
            else {
                keepers++;
                write_record(in1, fmt, r1[i]);
                if (paired) write_record(in2, fmt, r2[i]);
            }
        } else
            keep = false;
        if (separate && !keep) {
            write_record(out1, fmt, r1[i]);
            if (paired) write_record(out2, fmt, r2[i]);
        }
        if (do_stats) {
            const size_t len = r1[i].seq.size() + (paired ? r2[i].seq.size() : 0);
            stats << i << "\t" << len << "\t" << nb_kmers << "\t" << matches << "\t" << ratio << endl;
        }
        n_total++;
        if (n_total % 100000 == 0) cout << "Processed " << n_total << (paired ? " pairs" : " entries") << endl;
    }
    if (paired && r2.size() > r1.size()) throw KatError(4, "Second sequence file appears to be longer than the first.");
    char b[96];
    snprintf(b, sizeof b, "Finished filtering.  Time taken: %.1fs\n\n", t.secs());
    cout << b;
    cout << "Found " << keepers << " / " << n_total << " to keep" << endl << endl;
    free_tables(ctx, in);
    snprintf(b, sizeof b, "KAT filter seq completed.\nTotal runtime: %.1fs\n\n", total.secs());
    cout << b;
    return 0;
}

} // namespace

int filter_main(int argc, char **argv) { // Filter::main, src/filter.cc:66-131: argv[0] = "filter", argv[1] = mode
    if (argc == 1 || (argc == 2 && (!strcmp(argv[1], "--help") || !strcmp(argv[1], "-v") || !strcmp(argv[1], "--verbose")))) {
        cout << "Usage: kat filter <mode>\n\nFiltering tools.  Modes:\n  kmer  Filter k-mers to those within defined count and GC bounds.\n"
                "  seq   Filter sequences based on whether they contain k-mers of a hash.\n" << endl;
        return 1;
    }
    string mode = argv[1];
    for (auto &c : mode) c = (char)toupper((unsigned char)c);
    if (mode == "KMER") return filter_kmer_main(argc - 1, argv + 1);
    if (mode == "SEQ") return filter_seq_main(argc - 1, argv + 1);
    throw KatError(4, string("Could not recognise mode string: ") + argv[1]);
}

} // namespace kat_front
