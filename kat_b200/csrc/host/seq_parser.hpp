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
// `fastq_carry` (optional, in/out): the reference parser's per-stream `seq_len` — the bases of the FASTQ record being
// read — survives the move to the next file of an input group (JFI/mer_overlap_sequence_parser.hpp:158-187 resets type
// and trim only).  A file that ends inside a record (truncated, or an empty last record) leaves it non-zero; the next
// FASTQ file's first record is then measured too long, "Invalid fastq sequence" is thrown and swallowed, and the rest
// of the group is dropped.  A caller that parses a group passes the same variable for every file, in order.
ParseStatus parse_sequence_file(const unsigned char *data, size_t n, int k, unsigned trim5p, std::string &out, size_t *fastq_carry = nullptr);
// The character-level emulation alone (no fast path), which also knows the reference's 4 KiB buffers.  `last`: where, in
// the bytes this call appended, the LAST buffer's own bytes start, and how many bytes before them (the k-1 base seam) its
// k-mers may reach back over.  The reference loses exactly that buffer when the stream that read the file then opens a
// file in no supported format: produce() throws out of open_next_file before the buffer is queued
// (JFI/mer_overlap_sequence_parser.hpp:132-152, JFI/cooperative_pool2.hpp:249-260) — see kat_front.hpp::count_input.
struct LastBuffer {
    size_t at = 0, seam = 0;
};
ParseStatus parse_sequence_file_exact(const unsigned char *data, size_t n, int k, unsigned trim5p, std::string &out, size_t *fastq_carry = nullptr,
                                      LastBuffer *last = nullptr);
// The first byte of a regular file as the reference's streams see it (through zlib): -1 = empty or unreadable,
// -2 = not a regular file (a pipe can be read only once: not looked at).
int first_byte_of_file(const std::string &path);
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

// Plain 4-line FASTQ files WITHOUT host parsing: the file is memory-mapped and handed out in blocks
// of whole records, each block cut into `n_parts` shares of whole records (one per GPU); `sink`
// gets the raw bytes (libkatgpu parses them on the device, katgpu_count_feed_fastq) and returns
// 1 = taken, 0 = not in the regular layout (nothing counted), < 0 = error.  Record starts are found
// with the same test as the host fast path; a wrong guess shows up as an irregular block.
//   returns 1 whole file taken, 0 not applicable / first block irregular (nothing was counted),
//          -1 irregular (or error) after blocks were counted: the caller must start over on the host route
int stream_fastq_raw(const std::string &path, size_t block_bytes, int n_parts,
                     const std::function<int(const char *const *parts, const size_t *lens)> &sink, std::string &err);

// Which route the last parse_sequence_file call took: true = the validated regular-layout fast path.
bool last_parse_was_fast();

} // namespace kat_host
