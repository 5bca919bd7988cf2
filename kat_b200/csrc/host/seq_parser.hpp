// seq_parser.hpp — FASTA/FASTQ -> base stream, host side of the `kat` front-end.
#pragma once
#include <string>
#include <vector>

namespace kat_host {

enum class ParseStatus { Ok, UnsupportedFormat, InvalidFastq };

// Whole file into memory, through zlib (plain and .gz alike).
bool read_file(const std::string &path, std::vector<unsigned char> &out, std::string &err);

// Append the base stream of one file image to `out` (records separated by 'N').
ParseStatus parse_sequence_file(const unsigned char *data, size_t n, int k, unsigned trim5p, std::string &out);
// Which route the last parse_sequence_file call took: true = the validated regular-layout fast path.
bool last_parse_was_fast();

} // namespace kat_host
