// parse_kernels.cuh — device-side FASTQ record parsing (parse_kernels.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace katgpu {

uint32_t nl_tiles(uint64_t n); // 4 KiB tiles covering n bytes
void launch_nl_count(const uint8_t *data, uint64_t n, uint32_t *tile_counts, uint32_t n_tiles, cudaStream_t st);
void launch_nl_scan(uint32_t *tile_counts, uint32_t n_tiles, uint32_t *total, cudaStream_t st);
void launch_nl_scatter(const uint8_t *data, uint64_t n, const uint32_t *tile_prefix, uint32_t n_tiles, uint32_t *nl, uint32_t nl_cap, cudaStream_t st);
void launch_fastq_mask(uint8_t *data, const uint32_t *nl, uint32_t n_records, uint32_t trim, uint32_t *flags, int grid, cudaStream_t st);

} // namespace katgpu
