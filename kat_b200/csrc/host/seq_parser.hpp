// seq_parser.hpp — FASTA/FASTQ -> base stream, host side of the `kat` front-end.
#pragma once
#include <functional>
#include <string>
#include <vector>

namespace kat_host {

enum class ParseStatus { Ok, UnsupportedFormat, InvalidFastq };

// Whole file into memory, through zlib (plain and .gz alike).
bool read_file(const std::string &path, std::vector<unsigned char> &out, std::string &err);

// Append the base stream of one file image to `out` (records separated by 'N').
ParseStatus parse_sequence_file(const unsigned char *data, size_t n, int k, unsigned trim5p, std::string &out);
// Large plain files in one of the regular layouts (see seq_parser.cc) block by block, without
// holding the file or its whole base stream in memory: `sink` gets the base stream of each block
// of whole records (no k-mer spans two records, so blocks can be counted independently).
//   returns  1  the whole file went through `sink`
//            0  not applicable (gzip, not a regular file, irregular at the very start): nothing was
//               emitted, use read_file + parse_sequence_file
//           -1  an irregularity (or a sink failure) AFTER blocks were emitted: the caller must start over
//               with the whole-file route
int stream_regular_file(const std::string &path, unsigned trim5p, size_t block_bytes,
                        const std::function<bool(const char *, size_t)> &sink, std::string &err);

// Which route the last parse_sequence_file call took: true = the validated regular-layout fast path.
bool last_parse_was_fast();

} // namespace kat_host
