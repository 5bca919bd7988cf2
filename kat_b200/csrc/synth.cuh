// synth.cuh — deterministic synthetic read generator (bench/test tooling, not
// part of the reference's path).  SURVEY.md §8(d): genome of i.i.d. bases,
// reads at uniform positions and strands, substitution and N noise, every draw
// from a counter-based hash so that any shard of reads is reproducible alone.
// The same inline function runs on the GPU (bench inputs) and on the CPU
// (inputs for the oracle), so both sides see identical bytes.
#pragma once
#include <cstdint>

namespace katgpu {

struct SynthParams {
    uint64_t genome_len, seed_g, seed_r, first_read, n_reads;
    uint32_t read_len;
    uint32_t thr_sub; // substitution probability * 2^24
    uint32_t thr_n;   // N probability * 2^24
};

__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
__host__ __device__ inline uint32_t genome_base(uint64_t seed_g, uint64_t pos) {
    const uint64_t w = splitmix64(seed_g * 0xD1342543DE82EF95ULL + (pos >> 5));
    return (uint32_t)(w >> (2 * (pos & 31))) & 3u;
}
__host__ __device__ inline uint64_t mulhi64(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}
// Byte j (0..read_len) of read `idx`; j == read_len is the 'N' separator.
__host__ __device__ inline char synth_byte(const SynthParams &p, uint64_t idx, uint32_t j) {
    if (j >= p.read_len) return 'N';
    const uint64_t r0 = splitmix64(p.seed_r ^ splitmix64(idx));
    const uint64_t pos = mulhi64(r0, p.genome_len - p.read_len + 1);
    const bool rev = splitmix64(r0 ^ 0x5851F42D4C957F2DULL) & 1;
    uint32_t b = rev ? 3u - genome_base(p.seed_g, pos + (p.read_len - 1 - j)) : genome_base(p.seed_g, pos + j);
    const uint64_t e = splitmix64(r0 + 0x632BE59BD9B4E019ULL * (uint64_t)(j + 1));
    if ((uint32_t)(e >> 40) < p.thr_n) return 'N';
    if ((uint32_t)((e >> 16) & 0xFFFFFFu) < p.thr_sub) b = (b + 1u + (uint32_t)(e & 0xFFFFu) % 3u) & 3u;
    return (char)(0x54474341u >> (8 * b)); // "ACGT"[b]
}

} // namespace katgpu
