// reduce_kernels.cuh — declarations of the table-walk kernels.
#pragma once
#include "table.cuh"

namespace katgpu {

constexpr uint32_t kCompSub = 128;                 // shared-memory corner of the comp matrix
constexpr size_t kMaxReduceSmem = 200 * 1024;      // dynamic shared memory budget of a reduce CTA

struct CompParams {
    TableGeom ga, gb, gc;   // walked table, probed table, optional third table
    uint64_t n_slots;       // slots of ga
    int pass;               // 1: hash1 vs hash2(,3); 2: hash2 vs hash1; 3: hash3 totals
    int have_c;
    int canon_b, canon_c;   // canonicalise the probe key (JellyfishHelper::getCount 3rd arg)
    int spec_smem;
    double d1_scale, d2_scale;
    uint32_t d1_bins, d2_bins, len;
    unsigned long long *main_mx, *ends_mx, *middle_mx, *mixed_mx;
    unsigned long long *counters;      // 13 x u64, katgpu_comp_counters order
    unsigned long long *spec_walk;     // spectrum of the walked table (spectrum1 or spectrum2)
    unsigned long long *spec_shared1, *spec_shared2;
};

void launch_hist(const TableGeom &g, uint64_t n_slots, uint64_t base, uint64_t ceil_, uint64_t inc, uint32_t nb,
                 unsigned long long *out, int grid, cudaStream_t st);
void launch_gcp(const TableGeom &g, uint64_t n_slots, double scale, uint32_t cvg_bins, unsigned long long *out,
                int grid, cudaStream_t st);
void launch_comp(CompParams P, int grid, cudaStream_t st);
void launch_lookup(const TableGeom &g, const uint64_t *keys, uint64_t n, int canonicalise, uint64_t *out, int grid,
                   cudaStream_t st);
struct ExportSel { // which k-mers an export keeps (kat filter kmer, src/filter_kmer.cc:296-309)
    int mode = 0;       // 0 all, +1 inside the bounds, -1 outside
    uint64_t low_count = 0, high_count = ~0ULL;
    uint32_t low_gc = 0, high_gc = ~0u;
};
void launch_export(const TableGeom &g, uint64_t n_slots, uint64_t *keys, uint64_t *counts, uint64_t cap,
                   unsigned long long *cursor, int grid, cudaStream_t st, const ExportSel &sel = ExportSel());

} // namespace katgpu
