// dist_kernels.cuh — the multi-GPU exchange fused into the extraction kernel (SURVEY.md §8(e)
// steps 2-3): every rank cuts, canonicalises and hashes the k-mers of ITS reads, sorts them by
// owner rank inside the CTA and stores each owner's run straight into that owner's HBM through
// peer-mapped memory (NVLink 5 / NVSwitch) — no staging copy, no NCCL all-to-all, no host in the
// loop.  The owner then bins the received keys into its level-1 segments (bin_keys_kernel, segment
// source) and the tiled path takes over unchanged.  See dist_kernels.cu and DESIGN.md §5.
#pragma once
#include "tile_kernels.cuh"

namespace katgpu {

// 512 threads x 8 k-mer starts = one kBinTile of the base stream per iteration; two CTAs per SM = 32 warps
// (the extraction loop is latency bound: the same change took bin_bases_kernel from 57.3 to 52.1 ms)
#ifndef KATGPU_ROUTE_PER_THREAD
#define KATGPU_ROUTE_PER_THREAD 8
#endif
constexpr int kRoutePerThread = KATGPU_ROUTE_PER_THREAD; // 8 or 16
constexpr int kRouteThreads = kBinTile / kRoutePerThread;

// Owner rank of a hashed key (any rank count 1..kMaxRanks; independent of the bucket bits).
// Same function as table.cuh owner_of; restated in kat_b200/routing.py for the CPU tests.

struct RouteParams {
    TableGeom g;                // k, canonical flag, hash constants (the table pointers are not used)
    const uint8_t *bases;
    uint64_t n_valid, n_starts;
    int n_tiles, tma_ok;
    uint32_t n_ranks;
    // This sender's receive area on every owner (peer-mapped; [o == own rank] is local memory):
    // seg[o] = [gridDim.x][cap] hashed keys, len[o] = [gridDim.x] keys CTA j wrote in this launch.
    uint64_t *seg[kMaxRanks];
    uint32_t *len[kMaxRanks];
    uint32_t cap;               // >= k-mer starts one CTA processes in this launch: a segment cannot overflow
};

size_t route_smem_bytes();
cudaError_t route_configure(int *blocks_per_sm);
void launch_route_bases(const RouteParams &P, int grid, cudaStream_t st);

} // namespace katgpu
