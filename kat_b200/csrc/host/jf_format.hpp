// jf_format.hpp — jellyfish "binary/sorted" hash files (the format of `kat ... -d` dumps and
// of jellyfish-hash inputs, LOAD mode): JSON file header + fixed-size (key, count) records.
// Follows JFI/generic_file_header.hpp:96-151 (framing), JFI/file_header.hpp:25-108 (fields),
// JFI/binary_dumper.hpp:28-121 (records), JFI/sorted_dumper.hpp:52-103 + JFI/mer_heap.hpp:25-40
// (record order), JFI/rectangular_binary_matrix.hpp:205-243 (position = M * key).
#pragma once
#include <cstdint>
#include <functional>
#include <iosfwd>
#include <string>
#include <vector>

namespace kat_host {

struct JfHeader {
    // fields of jellyfish::file_header
    uint64_t size = 0;         // slots of the hash the file was dumped from (a power of two)
    unsigned key_len = 0;      // bits (2k)
    unsigned val_len = 7;      // bits of the in-memory count field (informational)
    unsigned max_reprobe = 126;
    unsigned counter_len = 4;  // bytes of the count in every record
    unsigned alignment = 8;
    bool canonical = false;
    std::string format;        // "binary/sorted"
    unsigned matrix_r = 0, matrix_c = 0;
    std::vector<uint64_t> matrix_columns; // c words; M * v = XOR of columns[c-1-j] over the set bits j of v
    std::vector<uint64_t> reprobes;
    std::string exe_path, hostname, pwd, time, cmdline;
    size_t offset = 0;         // bytes before the first record
    unsigned key_bytes() const { return key_len / 8 + (key_len % 8 != 0); }
    unsigned record_bytes() const { return key_bytes() + counter_len; }
};

// Parse the header at the start of `is`; on success the stream is positioned on the first record.
bool jf_read_header(std::istream &is, JfHeader &h, std::string &err);
// Write the header (sets h.offset).  fill_standard() analogue: hostname/pwd/time/exe_path are
// filled in when empty.
void jf_write_header(std::ostream &os, JfHeader &h);
// Position of a key in the dumped hash: (M * key) & (size - 1).  Keys longer than 32 bases carry their upper
// 2k - 64 bits in key_hi (mer_dna keeps ceil(2k/64) words, JFI/mer_dna.hpp:513-531).
uint64_t jf_position(const JfHeader &h, uint64_t key);
uint64_t jf_position(const JfHeader &h, uint64_t key_hi, uint64_t key_lo);

// Read every record (key_len <= 64), handing them out in batches of at most `batch` records.
// Returns false (err set) on a truncated file.
bool jf_read_records(std::istream &is, const JfHeader &h, size_t batch,
                     const std::function<void(const uint64_t *keys, const uint64_t *counts, size_t n)> &sink, std::string &err);

// The same for any key_len <= 128 (k <= 64): two words per key.
bool jf_read_records_wide(std::istream &is, const JfHeader &h, size_t batch,
                          const std::function<void(const uint64_t *keys_hi, const uint64_t *keys_lo, const uint64_t *counts, size_t n)> &sink,
                          std::string &err);
void jf_prepare_dump_wide(JfHeader &h, unsigned k, bool canonical, uint64_t n, std::vector<uint64_t> &keys_hi, std::vector<uint64_t> &keys_lo,
                          std::vector<uint64_t> &counts);
void jf_write_file_wide(std::ostream &os, JfHeader &h, const std::vector<uint64_t> &keys_hi, const std::vector<uint64_t> &keys_lo,
                        const std::vector<uint64_t> &counts);

// Prepare `h` for a dump of `n` k-mers of `k` bases and sort the records the way sorted_dumper
// emits them: ascending (position, key).  The matrix is the top-bits selector
// (position = key >> (2k - lsize)), under which that order is plain ascending key order —
// readers (binary_reader, binary_query, KAT's HashLoader) take the matrix from the header.
void jf_prepare_dump(JfHeader &h, unsigned k, bool canonical, uint64_t n, std::vector<uint64_t> &keys, std::vector<uint64_t> &counts);
// Header + records.  Counts saturate at 2^(8 * counter_len) - 1 (binary_writer::write).
void jf_write_file(std::ostream &os, JfHeader &h, const std::vector<uint64_t> &keys, const std::vector<uint64_t> &counts);

} // namespace kat_host
