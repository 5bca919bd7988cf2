// kat_main.cc — entry point of the `kat hist | gcp | comp | sect | cold | filter` front-end over libkatgpu (tools: kat_hist.cc,
// kat_gcp.cc, kat_comp.cc, kat_sect.cc, kat_cold.cc, kat_filter.cc; shared helpers: kat_front.hpp, kat_seqlookup.hpp).
//
// Mirrors the reference's CLI surface and output files for the three spectra
// tools (src/kat.cc:178-305, src/histogram.cc:257-369, src/gcp.cc:255-370,
// src/comp.cc:633-843) and its libkat helpers (InputHandler:
// lib/src/input_handler.cc; SparseMatrix printing: lib/include/kat/sparse_matrix.hpp;
// CompCounters::printCounts: lib/src/comp_counters.cc:144-206; distance metrics:
// lib/include/kat/distance_metrics.hpp).  Counting and the table walks run on
// the GPU through the C-ABI of include/katgpu.h; this file only parses
// arguments, reads sequence files, and writes the text outputs byte for byte
// as the reference does.  Jellyfish-hash inputs (LOAD mode, InputHandler::loadHash) and
// -d/--dump_hash (InputHandler::dump) read and write the reference's "binary/sorted" files
// (jf_format.cc).  Not supported (outside the hot path): plotting.
#include "kat_front.hpp"

namespace kat_front {
int hist_main(int argc, char **argv);
int gcp_main(int argc, char **argv);
int comp_main(int argc, char **argv);
int sect_main(int argc, char **argv);
int cold_main(int argc, char **argv);
int filter_main(int argc, char **argv);
} // namespace kat_front
using namespace kat_front;

int main(int argc, char **argv) {
    try {
        if (argc < 2 || !strcmp(argv[1], "--help") || !strcmp(argv[1], "-h")) {
            cout << "The K-mer Analysis Toolkit (KAT), GPU spectra engine.\n\nUsage: kat <mode> [options]\n\n"
                    "Modes: hist, gcp, comp, sect, cold, filter (counting, spectra, coverage lookups and filtering on the GPU through libkatgpu).\n"
                    "plot is not part of this build.\n" << endl;
            return 1;
        }
        if (!strcmp(argv[1], "--version") || !strcmp(argv[1], "-V")) {
            cout << "kat " << kVersion << endl;
            return 0;
        }
        const string mode = argv[1];
        if (mode == "_parse") { // test hook, no GPU involved: dump the base stream of one file to stdout
            const Parsed a = parse_args(argc - 1, argv + 1, {{'m', "mer_len", true}, {0, "5ptrim", true}, {0, "carry", true}});
            if (a.positional.size() != 1) throw KatError(1, "usage: kat _parse -m K [--5ptrim X] [--carry SEQ_LEN] <file>");
            size_t carry = (size_t)to_u64(a.get("carry", "0"), "carry"); // seq_len state inherited from the previous file of a group
            std::vector<unsigned char> raw;
            string err, bases;
            Timer t0;
            if (file_size_or_zero(a.positional[0]) >= env_size("KAT_STREAM_BYTES", (size_t)2 << 30)) { // the streaming route, blocks joined by 'N'
                size_t n_blocks = 0;
                const int r = kat_host::stream_regular_file(a.positional[0], (unsigned)to_u64(a.get("5ptrim", "0"), "5ptrim"),
                                                            env_size("KAT_STREAM_BLOCK", (size_t)512 << 20), [&](const char *b, size_t n) {
                                                                if (n_blocks++) bases += 'N';
                                                                bases.append(b, n);
                                                                return true;
                                                            }, err);
                if (getenv("KAT_TRACE")) fprintf(stderr, "[kat] stream route: rc %d, %zu blocks, %.3f s\n", r, n_blocks, t0.secs());
                if (r == 1) {
                    fwrite(bases.data(), 1, bases.size(), stdout);
                    return 0;
                }
                bases.clear();
            }
            if (!kat_host::read_file(a.positional[0], raw, err)) throw KatError(4, err);
            const double t_read = t0.secs();
            const auto st = kat_host::parse_sequence_file(raw.data(), raw.size(), (int)to_u64(a.get("mer_len", "27"), "mer_len"),
                                                          (unsigned)to_u64(a.get("5ptrim", "0"), "5ptrim"), bases, &carry);
            fprintf(stderr, "carry_out=%zu\n", carry);
            if (getenv("KAT_TRACE"))
                fprintf(stderr, "[kat] read %.3f s, parse %.3f s (%zu -> %zu bytes), parser: %s\n", t_read, t0.secs() - t_read, raw.size(), bases.size(),
                        kat_host::last_parse_was_fast() ? "fast" : "exact");
            fwrite(bases.data(), 1, bases.size(), stdout);
            return st == kat_host::ParseStatus::Ok ? 0 : st == kat_host::ParseStatus::UnsupportedFormat ? 2 : 3;
        }
        if (mode == "_rawsplit") { // test hook, no GPU: how stream_fastq_raw cuts a file into blocks and per-GPU shares of whole records
            if (argc != 5) throw KatError(1, "usage: kat _rawsplit <file.fastq> <block_bytes> <n_parts>");
            string err;
            size_t total = 0, blocks = 0;
            const int n_parts = atoi(argv[4]);
            const int r = kat_host::stream_fastq_raw(argv[2], (size_t)strtoull(argv[3], nullptr, 10), n_parts,
                                                     [&](const char *const *parts, const size_t *lens) {
                                                         printf("block %zu:", blocks++);
                                                         for (int p = 0; p < n_parts; p++) {
                                                             const bool whole = lens[p] == 0 || (parts[p][0] == '@' && parts[p][lens[p] - 1] == '\n');
                                                             printf(" %zu%s", lens[p], whole ? "" : "!");
                                                             total += lens[p];
                                                             if (p + 1 < n_parts && parts[p] + lens[p] != parts[p + 1]) printf("?");
                                                         }
                                                         printf("\n");
                                                         return 1;
                                                     }, err);
            printf("rc %d total %zu\n", r, total);
            return 0;
        }
        if (mode == "_jfdump") { // test hook, no GPU: header summary + records of a jellyfish hash, in file order
            if (argc != 3) throw KatError(1, "usage: kat _jfdump <hash.jf>");
            std::ifstream f;
            const kat_host::JfHeader h = read_jf_header(argv[2], f);
            const unsigned k = h.key_len / 2;
            printf("# k=%u canonical=%d counter_len=%u size=%llu format=%s offset=%zu\n", k, h.canonical ? 1 : 0, h.counter_len,
                   (unsigned long long)h.size, h.format.c_str(), h.offset);
            string err;
            uint64_t prev_pos = 0, prev_key = 0, prev_hi = 0, n_seen = 0, unsorted = 0;
            const bool ok = kat_host::jf_read_records_wide(f, h, 1 << 16, [&](const uint64_t *keys_hi, const uint64_t *keys, const uint64_t *counts, size_t n) {
                string line;
                for (size_t i = 0; i < n; i++) {
                    line.clear();
                    for (unsigned j = 0; j < k; j++) {
                        const unsigned bit = 2 * (k - 1 - j);
                        line += "ACGT"[((bit < 64 ? keys[i] >> bit : keys_hi[i] >> (bit - 64))) & 3];
                    }
                    printf("%s %llu\n", line.c_str(), (unsigned long long)counts[i]);
                    const uint64_t pos = kat_host::jf_position(h, keys_hi[i], keys[i]);
                    if (n_seen && (pos < prev_pos || (pos == prev_pos && (keys_hi[i] < prev_hi || (keys_hi[i] == prev_hi && keys[i] <= prev_key))))) unsorted++;
                    prev_pos = pos, prev_key = keys[i], prev_hi = keys_hi[i], n_seen++;
                }
            }, err);
            if (!ok) throw KatError(4, err);
            printf("# records=%llu out_of_order=%llu\n", (unsigned long long)n_seen, (unsigned long long)unsorted);
            return 0;
        }
        if (mode == "_jfwrite") { // test hook, no GPU: "KMER count" lines -> jellyfish binary/sorted hash
            const Parsed a = parse_args(argc - 1, argv + 1, {{'m', "mer_len", true}, {'N', "non_canonical", false}});
            if (a.positional.size() != 2) throw KatError(1, "usage: kat _jfwrite -m K [-N] <kmers.txt> <out.jf>");
            const unsigned k = (unsigned)to_u64(a.get("mer_len", "27"), "mer_len");
            std::ifstream in(a.positional[0].c_str());
            vector<uint64_t> keys_hi, keys, counts;
            string km;
            uint64_t c;
            while (in >> km >> c) {
                if (km.size() != k) throw KatError(4, "k-mer of the wrong length: " + km);
                uint64_t lo = 0, hi = 0;
                for (char ch : km) {
                    hi = (hi << 2) | (lo >> 62);
                    lo = (lo << 2) | (uint64_t)(ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : 3);
                }
                keys_hi.push_back(hi);
                keys.push_back(lo);
                counts.push_back(c);
            }
            kat_host::JfHeader h;
            std::ofstream out(a.positional[1].c_str(), std::ios::out | std::ios::binary);
            if (k > 32) {
                kat_host::jf_prepare_dump_wide(h, k, !a.has("non_canonical"), keys.size(), keys_hi, keys, counts);
                kat_host::jf_write_file_wide(out, h, keys_hi, keys, counts);
            } else {
                kat_host::jf_prepare_dump(h, k, !a.has("non_canonical"), keys.size(), keys, counts);
                kat_host::jf_write_file(out, h, keys, counts);
            }
            return out.good() ? 0 : 4;
        }
        cout << "Kmer Analysis Toolkit (KAT) V" << kVersion << endl << endl;
        // the reference's main drops what a tool's main returns (src/kat.cc:259-284): a tool that only printed its help
        // ("return 1" in the tool) still ends the process with 0; failures travel as exceptions
        int (*const tool)(int, char **) = mode == "hist" ? hist_main : mode == "gcp" ? gcp_main : mode == "comp" ? comp_main : mode == "sect" ? sect_main
                                          : mode == "cold" ? cold_main : mode == "filter" ? filter_main : nullptr;
        if (tool) {
            tool(argc - 1, argv + 1);
            return 0;
        }
        if (mode == "plot") throw KatError(4, "KAT mode 'plot' is outside this GPU build's scope (hist, gcp, comp, sect, cold, filter)");
        throw KatError(4, "Unrecognised KAT mode: " + mode);
    } catch (const KatError &e) {
        cerr << (e.code == 5 ? "Error: " : "") << e.what() << endl;
        return e.code;
    } catch (const std::exception &e) {
        cerr << "Error: " << e.what() << endl;
        return 5;
    } catch (...) {
        cerr << "Error: Exception of unknown type!" << endl;
        return 7;
    }
}
