// extract_emul.cc — CPU emulation of the one-word (k <= 32) extraction arithmetic of kat_b200/csrc/extract.cuh and the hash
// bijection of table.cuh.  TEST INFRASTRUCTURE ONLY (tests/test_extract_host_emulation.py).
//
// The device headers are compiled here AS THEY ARE, by g++: the CUDA qualifiers are empty outside nvcc, the handful of
// integer intrinsics they use get host stand-ins below, and the functions that hold PTX (mbarrier / TMA wrappers, bucket
// loads, the RED fast path) are force-inlined and never called, so they are never emitted.  What runs is what
// bin_bases_kernel / route_bases_kernel do per thread (tile_kernels.cu:279-318): pack 16 ASCII bytes to 2-bit codes and
// invalid bits (pack4), take the 48-base window of a 16-base group, cut the 16 k-mers and their reverse complements out
// of it with constant funnel shifts (Cut16), canonicalise, hash (mix_words) — in the kernel's tiling (8 KiB tiles + 32
// bytes of halo, bytes past the end read as 0) — and, beside it, the rolling form the other kernels use (Roll16) and
// cut_kmer.  Every hashed key is taken back through unmix_key and printed; the test compares the multiset with the oracle's.
//
//   extract_emul <k> <canonical> <stream file> <n_starts or -1>   ->  "key" per k-mer instance (decimal), then "#checks ..."
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>

#include <cuda_runtime.h> // vector types; __host__ / __device__ are empty for a host compiler
#define __CUDACC__ 1      // open the device half of the headers for this host build
#define __noinline__ __attribute__((noinline))

// ---- host stand-ins for the device intrinsics the headers use ----
static inline uint32_t __vcmpeq4(uint32_t a, uint32_t b) {
    uint32_t r = 0;
    for (int i = 0; i < 4; i++)
        if (((a >> (8 * i)) & 0xFF) == ((b >> (8 * i)) & 0xFF)) r |= 0xFFu << (8 * i);
    return r;
}
static inline uint32_t __brev(uint32_t x) {
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
    return r;
}
static inline uint64_t __brevll(uint64_t x) {
    uint64_t r = 0;
    for (int i = 0; i < 64; i++) r |= ((x >> i) & 1ull) << (63 - i);
    return r;
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) { // shift wraps at 32
    return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> (s & 31));
}
static inline uint32_t __funnelshift_rc(uint32_t lo, uint32_t hi, uint32_t s) { // shift clamps at 32
    return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> (s > 32 ? 32 : s));
}
static inline size_t __cvta_generic_to_shared(const void *p) { return (size_t)p; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
static inline unsigned long long atomicOr(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; *p = o | v; return o; }
static inline unsigned long long atomicCAS(unsigned long long *p, unsigned long long c, unsigned long long v) { const unsigned long long o = *p; if (o == c) *p = v; return o; }
static inline unsigned int atomicExch(unsigned int *p, unsigned int v) { const unsigned int o = *p; *p = v; return o; }
static inline ulonglong2 __ldg(const ulonglong2 *p) { return *p; }

#include "extract.cuh"
using namespace katgpu;

static const int kTile = 8192, kHalo = 32; // tile_kernels.cuh: kBin1Tile, kHalo

static uint64_t inv_mod_2_64(uint64_t a) {
    uint64_t x = a;
    for (int i = 0; i < 6; i++) x *= 2 - a * x;
    return x;
}

int main(int argc, char **argv) {
    if (argc < 5) return 2;
    const uint32_t k = (uint32_t)atoi(argv[1]);
    const bool canonical = atoi(argv[2]) != 0;
    std::vector<uint8_t> bases;
    {
        FILE *f = fopen(argv[3], "rb");
        if (!f) return 2;
        uint8_t buf[65536];
        size_t n;
        while ((n = fread(buf, 1, sizeof buf, f)) > 0) bases.insert(bases.end(), buf, buf + n);
        fclose(f);
    }
    const uint64_t n_valid = bases.size();
    const long long ns = atoll(argv[4]);
    const uint64_t n_starts = ns < 0 ? n_valid : (uint64_t)ns;
    TableGeom g;
    memset(&g, 0, sizeof g);
    g.k = k, g.b = 2 * k, g.sh = (g.b + 1) / 2, g.canonical = canonical;
    g.kmask = g.b == 64 ? ~0ULL : ((1ULL << g.b) - 1);
    g.m1inv = inv_mod_2_64(kMixM1);
    const uint32_t kmask_lo = (uint32_t)g.kmask, kmask_hi = (uint32_t)(g.kmask >> 32);
    const bool wide = g.b >= 32; // launch_bin_bases: the low word needs no mask
    unsigned long long n_keys = 0, roll_mismatch = 0, cut_mismatch = 0, mix_mismatch = 0;

    const uint64_t n_tiles = (n_starts + kTile - 1) / kTile;
    std::vector<uint8_t> ascii(kTile + kHalo);
    std::vector<uint32_t> W(kTile / 16 + 2), B(kTile / 16 + 2);
    for (uint64_t tile = 0; tile < n_tiles; tile++) {
        const uint64_t tile_base = tile * kTile;
        for (int i = 0; i < kTile + kHalo; i++) ascii[i] = tile_base + i < n_valid ? bases[tile_base + i] : (uint8_t)0;
        for (int gi = 0; gi < kTile / 16 + 2; gi++) { // the pack phase
            uint32_t v[4], c[4], b[4];
            memcpy(v, ascii.data() + 16 * gi, 16);
            for (int j = 0; j < 4; j++) pack4(v[j], c[j], b[j]);
            W[gi] = (c[0] << 24) | (c[1] << 16) | (c[2] << 8) | c[3];
            B[gi] = (b[0] << 12) | (b[1] << 8) | (b[2] << 4) | b[3];
        }
        for (uint32_t grp = 0; grp < (uint32_t)kTile / 16; grp++) { // one thread per 16-base group
            const uint64_t p0 = tile_base + (uint64_t)grp * 16;
            const uint64_t bad48 = ((uint64_t)B[grp] << 32) | ((uint64_t)B[grp + 1] << 16) | B[grp + 2];
            const uint32_t nst = p0 >= n_starts ? 0u : (uint32_t)std::min<uint64_t>(16, n_starts - p0);
            Cut16 cut;
            cut.init(W[grp], W[grp + 1], W[grp + 2], bad48, k, nst);
            // the rolling form over the same window (count_kernels.cu / partition_kernels.cu)
            const uint64_t hi = ((uint64_t)W[grp] << 32) | W[grp + 1], lo = (uint64_t)W[grp + 2] << 32;
            Roll16 roll;
            roll.init(hi, lo, bad48, k, nst);
            for (uint32_t r = 0; r < 16; r++) {
                if (r) roll.step(r, g.kmask);
                uint32_t klo, khi;
                if (canonical) {
                    if (wide) cut.key<true, true>(r, kmask_lo, kmask_hi, klo, khi);
                    else cut.key<true, false>(r, kmask_lo, kmask_hi, klo, khi);
                } else {
                    if (wide) cut.key<false, true>(r, kmask_lo, kmask_hi, klo, khi);
                    else cut.key<false, false>(r, kmask_lo, kmask_hi, klo, khi);
                }
                const uint64_t h = wide ? mix_words<true>(klo, khi, g.sh, kmask_lo, kmask_hi) : mix_words<false>(klo, khi, g.sh, kmask_lo, kmask_hi);
                if (roll.ok(r) != cut.ok(r)) roll_mismatch++;
                if (!cut.ok(r)) continue;
                const uint64_t key = ((uint64_t)khi << 32) | klo;
                if (roll.key(canonical) != key) roll_mismatch++;
                uint64_t fwd;
                if (!cut_kmer(hi, lo, bad48, r, k, fwd)) cut_mismatch++;
                else {
                    const uint64_t rc = revcomp_kmer(fwd, k);
                    if (((canonical && rc < fwd) ? rc : fwd) != key) cut_mismatch++;
                }
                if (mix_key(g, key) != h || unmix_key(g, h) != key || h > g.kmask) mix_mismatch++;
                printf("%llu\n", (unsigned long long)unmix_key(g, h));
                n_keys++;
            }
        }
    }
    printf("#checks keys=%llu roll_mismatch=%llu cut_mismatch=%llu mix_mismatch=%llu\n", n_keys, roll_mismatch, cut_mismatch, mix_mismatch);
    return 0;
}
