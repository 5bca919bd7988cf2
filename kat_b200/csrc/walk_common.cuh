// walk_common.cuh — device helpers shared by the table-walk kernels of both table layouts
// (reduce_kernels.cu: one-word keys, wide_kernels.cu: two-word keys): warp-aggregated increments,
// warp sums, the count scaling of gcp / comp and the spectrum binning of comp.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace katgpu {

// Add 1 to cells[cell] for every active lane, one atomic per distinct cell of the warp.
__device__ __forceinline__ void warp_agg_inc_u32(uint32_t *cells, uint32_t cell, bool active) {
    const uint32_t key = active ? cell : 0xFFFFFFFFu;
    const uint32_t peers = __match_any_sync(0xffffffffu, key);
    if (active && (__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(cells + cell, (uint32_t)__popc(peers));
}
__device__ __forceinline__ void warp_agg_inc_u64(unsigned long long *cells, uint64_t cell, bool active) {
    const uint64_t key = active ? cell : ~0ULL;
    const uint32_t peers = __match_any_sync(0xffffffffu, key);
    if (active && (__ffs(peers) - 1) == (int)(threadIdx.x & 31))
        atomicAdd(cells + cell, (unsigned long long)__popc(peers));
}
__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    return v;
}

constexpr int kWalkUnroll = 4; // 16-byte loads a walking thread keeps in flight

// scaleCounter (src/comp.hpp:303-306) / the cvg_pos expression of src/gcp.cc:188
__device__ __forceinline__ uint64_t scale_count(uint64_t c, double scale, bool unit) {
    if (unit || c == 0) return c;
    return (uint64_t)ceil((double)c * scale);
}

__device__ __forceinline__ uint32_t spectrum_bin(uint64_t c, uint32_t len) {
    return c >= len ? len - 1 : (uint32_t)c; // CompCounters::updateSpectrum, comp_counters.cc:130-140
}


} // namespace katgpu
