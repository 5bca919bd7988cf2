// count_kernels.cuh — shared declarations of the counting kernels.
#pragma once
#include "table.cuh"

namespace katgpu {

constexpr int kTile = 8192;          // k-mer start positions per tile
constexpr int kHalo = 32;            // >= k-1 bases of look-ahead, keeps TMA sizes 16-byte multiples
constexpr int kGroups = kTile / 16;  // 16-base groups per tile
constexpr int kThreads = 256;
#ifndef KATGPU_UNROLL
#define KATGPU_UNROLL 4
#endif
#ifndef KATGPU_MIN_CTAS
#define KATGPU_MIN_CTAS 3
#endif
constexpr int kUnroll = KATGPU_UNROLL; // independent bucket loads in flight per thread
constexpr int kMaxRanks = 8;         // one NVSwitch box
constexpr int kModeInsert = 0;
constexpr int kModeRoute = 1;

struct ExtractParams {
    TableGeom g;
    const uint8_t *bases;     // base stream in HBM
    uint64_t n_valid;         // readable bytes
    uint64_t n_starts;        // k-mers starting at p < n_starts belong to this launch
    unsigned int *tile_counter;
    int n_tiles;
    int tma_ok;               // bases 16-byte aligned
    // route mode
    uint32_t n_ranks;
    uint64_t *out[kMaxRanks]; // per-owner buckets (local or peer-mapped)
    unsigned long long *cursors;
    uint64_t out_cap;
    unsigned int *route_overflow;
};

size_t extract_smem_bytes(int mode);
cudaError_t extract_configure(int mode, int *blocks_per_sm);
void launch_extract(int mode, const ExtractParams &P, int grid, cudaStream_t st);
void launch_insert_hashed(const TableGeom &g, const uint64_t *hs, uint64_t n, int grid, cudaStream_t st);
void launch_insert_pairs(const TableGeom &g, const uint64_t *pairs, uint64_t n, int grid, cudaStream_t st);
void launch_add_pairs(const TableGeom &g, const uint64_t *keys, const uint64_t *counts, uint64_t n, int grid, cudaStream_t st);
void launch_rehash(const TableGeom &old_g, const TableGeom &new_g, uint64_t n_slots, int grid, cudaStream_t st);

} // namespace katgpu
