// tile_kernels.cuh — the shared-memory counting path ("tiled"): two binning passes bring every
// hashed key next to the 64 KB piece of the table (a "tile") it belongs to, then each tile is
// updated in shared memory and written back once.  See tile_kernels.cu and DESIGN.md §3.
#pragma once
#include "count_kernels.cuh"

namespace katgpu {

constexpr int kBinTile = 4096;        // keys (k-mer starts) per tile of the route / hashed-key binning kernels
constexpr int kBinGroups = kBinTile / 16;
// bin_bases_kernel comes in two shapes (KATGPU_BIN1_WIDE):
//   0  256 threads x 16 k-mer starts, 4 KiB tiles, two CTAs per SM (16 warps), 32-byte granules;
//   1  1024 threads x 8 starts, 8 KiB tiles, one CTA per SM (32 warps, <= 64 registers): twice the warps for the
//      latency-bound extraction loop, and a ring deep enough (16 keys per bin and tile) for 64-byte granules,
//      i.e. half the scattered-write requests.
#ifndef KATGPU_BIN1_WIDE
#define KATGPU_BIN1_WIDE 1
#endif
#if KATGPU_BIN1_WIDE
constexpr int kBin1Tile = 8192, kBin1Threads = 1024, kBin1PerThread = 8;
#define KATGPU_BIN1_MIN_CTAS 1
#ifndef KATGPU_BIN1_GRANULE
#define KATGPU_BIN1_GRANULE 8
#endif
#else
constexpr int kBin1Tile = 4096, kBin1Threads = 256, kBin1PerThread = 16;
#endif
constexpr int kBin1Groups = kBin1Tile / 16;
#ifndef KATGPU_BIN1_MIN_CTAS
#define KATGPU_BIN1_MIN_CTAS 2
#endif
constexpr int kBin1MinCtas = KATGPU_BIN1_MIN_CTAS;
constexpr int kBin2Threads = 1024;    // one CTA per SM: 32 warps hide the key-load latency, one bin per thread in the flush
#ifndef KATGPU_BIN2_UNROLL
#define KATGPU_BIN2_UNROLL 4
#endif
#ifndef KATGPU_BIN2_PREFETCH
#define KATGPU_BIN2_PREFETCH 1
#endif
#ifndef KATGPU_BIN2_CPF
#define KATGPU_BIN2_CPF 3
#endif
constexpr int kBin2Unroll = KATGPU_BIN2_UNROLL; // keys per thread and chunk of bin_keys_kernel (level 2)
// Chunks bin_keys_kernel bins between two flushes of its staging ring, at most: a flush is two CTA-wide
// barriers and a pass over all bins, so fewer, fuller flushes are cheaper (measured 70.9 / 59.1 / 57.6 ms
// for 1 / 2 / 3 at 512 bins); the host picks the largest count whose ring still fits shared memory.
constexpr int kBin2MaxCpf = KATGPU_BIN2_CPF;
// bin_keys_kernel reads its source segments through a ring of TMA bulk copies (ncu of the register-prefetch form:
// long-scoreboard stalls 9.9 per issue, issue active 29 %): kBin2Stages buffers of kBin2Piece keys, full/empty mbarriers.
constexpr int kBin2Piece = 2048;
constexpr int kBin2Stages = 3;
constexpr int kBin2StageBytes = kBin2Stages * kBin2Piece * 8 + 64; // key buffers + 2 x kBin2Stages mbarriers
#ifndef KATGPU_TILE_THREADS
#define KATGPU_TILE_THREADS 384
#endif
#ifndef KATGPU_TILE_MIN_CTAS
#define KATGPU_TILE_MIN_CTAS 3 // 3 x 64 KB tiles per SM: one CTA's tile load / write-back hides behind the others' key loops (measured 6 % over 512 x 2)
#endif
#ifndef KATGPU_TILE_UNROLL
#define KATGPU_TILE_UNROLL 4
#endif
constexpr int kTileThreads = KATGPU_TILE_THREADS;
// Keys leave the staging rings in granules of this many (x 8 bytes, aligned in the segment):
// 4 = one 32-byte sector per request; 8 = lane pairs write 64 bytes per request (BinCore::flush).
#ifndef KATGPU_BIN1_GRANULE
#define KATGPU_BIN1_GRANULE 4
#endif
#ifndef KATGPU_BIN2_GRANULE
#define KATGPU_BIN2_GRANULE 4
#endif
#ifndef KATGPU_OLD_G8
#define KATGPU_OLD_G8 0 // 1: 64-byte granules also when a CTA has fewer than two threads per bin (round 1's shuffle-based flush; slower)
#endif
#ifndef KATGPU_BINH_GRANULE
#define KATGPU_BINH_GRANULE 4
#endif
constexpr uint32_t kBin1Granule = KATGPU_BIN1_GRANULE; // level-1 segments written by bin_bases_kernel
constexpr uint32_t kBinHGranule = KATGPU_BINH_GRANULE; // level-1 segments written by bin_hashed_kernel (keys received from other ranks)
constexpr uint32_t kBin2Granule = KATGPU_BIN2_GRANULE; // level-2 (per-tile) segments
constexpr int kTileBucketsLog2 = 11;  // a table tile = 2048 buckets = 8192 slots = 64 KB
constexpr int kTileBytes = (1 << kTileBucketsLog2) * 32;
constexpr int kTileProbes = 4;        // probes (home, +1, +3, +6 buckets) the tile kernel follows before it parks a key
constexpr int kMaxBins = 1024;        // fan-out of one binning pass
constexpr int kTileHllBits = 10;      // HyperLogLog registers of the level-1 kernel (as partition_kernel)
constexpr int kTileHllM = 1 << kTileHllBits;
constexpr int kTileHllSampleShift = 10;

// Keys that cannot take the fast way (a full staging ring or segment; in the tile kernel a full
// home bucket or a count field about to carry) are parked here and inserted with the general
// random-access kernel afterwards.  `reserved` lets the tile kernel refuse a tile up front when
// the list might not hold all of its keys (the tile is then deferred, never dropped).
struct SideList {
    uint64_t *keys;
    unsigned long long *n;        // keys pushed
    unsigned long long *reserved; // tile kernel: keys that may still be pushed by tiles in flight + n
    uint64_t cap;
    uint32_t *deferred;           // tile kernel: ids of tiles left for the general kernel
    uint32_t *n_deferred;
    uint32_t deferred_cap;
};

// Destination of one binning pass: bin d of this CTA goes to segment (d * stride + offset).
struct BinDst {
    uint64_t *keys;   // [segment][cap]
    uint32_t *len;    // [segment] keys present
    uint32_t cap;     // keys per segment, multiple of 16
    uint32_t stride;  // see above; offset depends on the kernel (CTA id, or (slice, part))
    uint32_t shift;   // bin = (h >> shift) & mask
    uint32_t mask;
    uint32_t F, R;    // bins, staging-ring slots per bin
};

struct BinBasesParams {
    TableGeom g;
    const uint8_t *bases;
    uint64_t n_valid, n_starts;
    int n_tiles, tma_ok;
    uint32_t cta_rot;
    BinDst dst;
    SideList side;
    uint32_t *hll;
};

// Level 2 (and the owner side of the multi-GPU exchange): keys in, keys out.
struct BinKeysParams {
    TableGeom g;
    // source A: a flat array of hashed keys, tiled over the grid (level 1 of received keys)
    const uint64_t *in_keys;
    uint64_t n_keys;
    uint32_t cta_rot;
    // source C (multi-GPU): dense segments [n_src][seg_cap] with seg_len[n_src] keys each, written by the
    // route_bases_kernel of every rank (dist_kernels.cu); CTA j bins segments j, j + grid, ...
    const uint64_t *seg_keys;
    const uint32_t *seg_len;
    uint32_t n_src, seg_cap;
    // source B: the level-1 segments of slice (first_slice + blockIdx.x / parts), share blockIdx.x % parts
    const uint64_t *src_keys; // [slice][src_segs][src_cap]
    const uint32_t *src_len;
    uint32_t src_segs, src_cap, first_slice, parts;
    uint32_t tile_keys;       // keys binned between two flushes
    uint32_t cpf;             // pieces (source B: of kBin2Piece keys; sources A, C: of kBinTile keys) between two flushes
    BinDst dst;
    SideList side;
    uint32_t *hll;            // source A only
};

struct TileParams {
    TableGeom g;
    const uint64_t *keys;  // [tile][parts][cap]
    const uint32_t *len;
    uint32_t cap, parts;
    uint32_t first_tile;   // table tile of local tile 0
    uint32_t n_tiles;
    SideList side;
};

size_t bin_bases_smem(uint32_t F, uint32_t R);
size_t bin_keys_smem(uint32_t F, uint32_t R);   // bin_hashed_kernel: HyperLogLog registers + ring
size_t bin_keys_b_smem(uint32_t F, uint32_t R); // bin_keys_kernel: TMA stages + ring
cudaError_t tile_configure(int *bps_bases, int *bps_tile);
void launch_bin_bases(const BinBasesParams &P, int grid, cudaStream_t st);
void launch_bin_keys(const BinKeysParams &P, int grid, cudaStream_t st);
void launch_tile_insert(const TileParams &P, int grid, cudaStream_t st);
// general insertion of every key of the listed (deferred) tiles' segments
void launch_insert_deferred(const TileParams &P, const uint32_t *tiles, uint32_t n, cudaStream_t st);

} // namespace katgpu
