// partition_kernels.cuh — declarations of the L2-resident counting path.
#pragma once
#include "count_kernels.cuh"

namespace katgpu {

constexpr int kPartThreads = 512;  // == kGroups: one 16-base group per thread per tile
constexpr int kMaxParts = 512;     // table slices (Pb <= 9)
constexpr int kHllBits = 10;       // HyperLogLog registers for the distinct-key estimate (3.3 % std error)
constexpr int kHllM = 1 << kHllBits;
constexpr int kHllSampleShift = 10; // 1/1024 of the key space is sampled
#ifndef KATGPU_SWEEP_THREADS
#define KATGPU_SWEEP_THREADS 256
#endif
#ifndef KATGPU_SWEEP_MIN_CTAS
#define KATGPU_SWEEP_MIN_CTAS 4
#endif
constexpr int kSweepThreads = KATGPU_SWEEP_THREADS;
constexpr int kSweepMinCtas = KATGPU_SWEEP_MIN_CTAS;
#ifndef KATGPU_SWEEP_UNROLL
#define KATGPU_SWEEP_UNROLL 2
#endif
constexpr int kSweepUnroll = KATGPU_SWEEP_UNROLL;
#ifndef KATGPU_SWEEP_TWOSTAGE
#define KATGPU_SWEEP_TWOSTAGE 0
#endif

struct PartitionParams {
    TableGeom g;
    const uint8_t *bases;
    uint64_t n_valid, n_starts;
    int n_tiles, tma_ok;
    uint32_t cta_rot;   // CTA that takes tile 0 (rotates between launches so segments fill evenly)
    uint32_t Pb;        // log2(partitions)
    uint32_t seg_cap;   // keys per (partition, CTA) segment, multiple of 16
    uint64_t *keys;     // [partition][cta][seg_cap]
    uint32_t *seglen;   // [partition][cta] keys appended so far (multiples of 4)
    uint32_t *hll;      // [kHllM]
    const uint64_t *in_keys; // non-null: partition these hashed keys instead of the base stream
    uint64_t n_keys;
};

struct SweepParams {
    TableGeom g;
    const uint64_t *keys;
    const uint32_t *seglen;
    uint32_t seg_cap, n_segs, part_lo, part_hi;
    uint32_t split; // CTAs per segment
    uint32_t window;      // slices the CTAs may be apart (0 = no barrier)
    uint32_t *slice_done; // [partitions] CTAs that finished each slice
    uint64_t slice_slots; // table slots per slice
    uint32_t prefetch;    // bulk-prefetch the next slice into L2
    uint32_t experiment;  // 0 = normal; 1/2 = measurement aids (results are WRONG), see sweep_kernel
};

size_t partition_smem_bytes();
cudaError_t partition_configure(int *blocks_per_sm);
void launch_partition(const PartitionParams &P, int grid, cudaStream_t st);
cudaError_t sweep_occupancy(int *blocks_per_sm);
void launch_sweep(const SweepParams &P, cudaStream_t st);

} // namespace katgpu
