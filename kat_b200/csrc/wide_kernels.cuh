// wide_kernels.cuh — declarations of the kernels for two-word keys (k = 33..64); see wide.cuh for the table.
#pragma once
#include "reduce_kernels.cuh"
#include "wide.cuh"

namespace katgpu {

constexpr int kWideThreads = 256;
constexpr int kWidePer = 36;                       // k-mer start positions per thread and tile (9 words: conflict-free stride)
constexpr int kWideTile = kWideThreads * kWidePer; // 9216 start positions per tile
constexpr int kWideHalo = 64;                      // >= k-1 bases of look-ahead; keeps TMA sizes 16-byte multiples

struct WideExtractParams {
    WideGeom g;
    const uint8_t *bases; // base stream in HBM
    uint64_t n_valid;     // readable bytes
    uint64_t n_starts;    // k-mers starting at p < n_starts belong to this launch
    unsigned int *tile_counter;
    int n_tiles;
    int tma_ok; // bases 16-byte aligned
};

struct WideCompParams { // CompParams (reduce_kernels.cuh) over wide tables
    WideGeom ga, gb, gc;
    int pass, have_c, canon_b, canon_c, spec_smem;
    double d1_scale, d2_scale;
    uint32_t d1_bins, d2_bins, len;
    unsigned long long *main_mx, *ends_mx, *middle_mx, *mixed_mx;
    unsigned long long *counters;
    unsigned long long *spec_walk, *spec_shared1, *spec_shared2;
};

void launch_wide_extract(const WideExtractParams &P, int grid, cudaStream_t st);
void launch_wide_add_pairs(const WideGeom &g, const uint64_t *keys_hi, const uint64_t *keys_lo, const uint64_t *counts, uint64_t n, int grid,
                           cudaStream_t st);
void launch_wide_rehash(const WideGeom &og, const WideGeom &ng, int grid, cudaStream_t st);
void launch_wide_stats(const WideGeom &g, unsigned long long *out, int grid, cudaStream_t st);
void launch_wide_hist(const WideGeom &g, uint64_t base, uint64_t ceil_, uint64_t inc, uint32_t nb, unsigned long long *out, int grid,
                      cudaStream_t st);
void launch_wide_gcp(const WideGeom &g, double scale, uint32_t cvg_bins, unsigned long long *out, int grid, cudaStream_t st);
void launch_wide_comp(WideCompParams P, int grid, cudaStream_t st);
void launch_wide_lookup(const WideGeom &g, const uint64_t *keys_hi, const uint64_t *keys_lo, uint64_t n, int canonicalise, uint64_t *out,
                        int grid, cudaStream_t st);
void launch_wide_export(const WideGeom &g, uint64_t *keys_hi, uint64_t *keys_lo, uint64_t *counts, uint64_t cap, unsigned long long *cursor,
                        int grid, cudaStream_t st, const ExportSel &sel = ExportSel());

} // namespace katgpu
