// dist_kernels.cu — route_bases_kernel: extraction + the multi-GPU exchange in one kernel.
//
// Why k-mers must move at all: a count histogram is not additive over read shards, so every
// distinct k-mer has to be counted on exactly one rank.  Reads are sharded by
// record; this kernel does for rank q what bin_bases_kernel does on one GPU up to the hash —
// TMA bulk copies of 4 KiB base tiles, SIMD-in-register 2-bit packing, 16 k-mers per thread cut
// with constant funnel shifts (extract.cuh: Cut16), canonical strand, mix_words — and then, instead
// of binning by table slice, sorts the tile's keys BY OWNER RANK in shared memory (one warp-
// aggregated shared atomic per owner per warp instruction hands out the ranks) and copies each
// owner's run with coalesced stores into this CTA's private segment of the receive area in the
// OWNER's HBM.  The segment pointers are peer-mapped device memory: the stores of the (N-1)/N
// remote keys travel over NVLink / NVSwitch as whole 128-byte lines while the next tile is being
// hashed.  A segment holds every k-mer start its CTA processes in one launch (host: cap), so there
// is no overflow path.  What follows on the owner is bin_keys_kernel (segment source) -> level-2
// binning -> tile_insert_kernel: the single-GPU pipeline, unchanged.
//
// Replaces, for N > 1, the same reference loop as bin_bases_kernel (lib/src/jellyfish_helper.cc:
// 202-211, JFI/mer_iterator.hpp:59-89) plus the exchange the reference never needs (one host).
#include "dist_kernels.cuh"
#include "extract.cuh"

namespace katgpu {

// Keys leave the SM in whole, aligned 128-byte lines (16 keys): NVLink takes ~6.5 G write requests
// per second whatever their size (tools/p2p_bench.cu: 32 B pieces 215 GB/s, 128 B pieces 710 GB/s),
// so each owner's run is stored from its line-aligned position and the < 16 keys that do not fill
// a line wait in shared memory for the next tile.
constexpr uint32_t kRouteLine = 16;
// KATGPU_ROUTE_TMA_STORE=1: the runs leave the SM as TMA bulk stores (cp.async.bulk.global.shared::cta) from a
// double-buffered pool — asynchronous, so the wire works while the next tile is hashed — instead of per-thread
// 32-byte stores in a burst after every tile.
#ifndef KATGPU_ROUTE_TMA_STORE
#define KATGPU_ROUTE_TMA_STORE 0
#endif
constexpr uint32_t kRoutePools = KATGPU_ROUTE_TMA_STORE ? 2 : 1;
struct __align__(128) RouteSmem {
    uint8_t ascii[kBinTile + kHalo];
    uint32_t W[2][kBinGroups + 2];
    uint32_t B[2][kBinGroups + 2];
    uint64_t mbar;
    uint32_t cnt[kMaxRanks];     // keys of the current tile per owner
    uint32_t nc[kMaxRanks];      // keys carried over from the previous tile per owner (< kRouteLine)
    uint32_t off[kMaxRanks + 1]; // start of each owner's run in `pool` (carried keys first, then this tile's)
    uint32_t goff[kMaxRanks + 1];// exclusive prefix of the 4-key groups each owner's run stores in this tile
    uint32_t cur[kMaxRanks];     // keys this CTA has already written to each owner's segment (multiple of kRouteLine)
    uint64_t carry[kMaxRanks][kRouteLine];
    alignas(128) uint64_t pool[kRoutePools][kBinTile + kMaxRanks * (kRouteLine + 2)]; // the tile's keys, sorted by owner
};

template <bool CANON, bool WIDE>
__global__ void __launch_bounds__(kRouteThreads, 2) route_bases_kernel(const __grid_constant__ RouteParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    RouteSmem &S = *reinterpret_cast<RouteSmem *>(smem_raw);
    const TableGeom &g = P.g;
    const int tid = threadIdx.x;
    const uint32_t lane = tid & 31u, lt_mask = (1u << lane) - 1u;
    const uint32_t k = g.k, sh = g.sh, n_ranks = P.n_ranks;
    const uint32_t kmask_lo = (uint32_t)g.kmask, kmask_hi = (uint32_t)(g.kmask >> 32);

    auto tile_tma = [&](int t) { return P.tma_ok && ((uint64_t)(t + 1) * kBinTile + kHalo <= P.n_valid); };
    int tile = (int)blockIdx.x; // static striding: CTA j owns tiles j, j + grid, ... (the host sizes `cap` from that)
    if (tid < kMaxRanks) {
        S.cnt[tid] = 0;
        S.cur[tid] = 0;
        S.nc[tid] = 0;
    }
    if (tid == 0) {
        mbar_init(&S.mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (tile < P.n_tiles && tile_tma(tile)) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&S.mbar, kBinTile + kHalo);
            tma_load_1d(S.ascii, P.bases + (uint64_t)tile * kBinTile, kBinTile + kHalo, &S.mbar);
        }
    }
    __syncthreads();
    uint32_t parity = 0, buf = 0;

    for (; tile < P.n_tiles; tile += gridDim.x, buf ^= 1) {
        uint64_t *pool = S.pool[KATGPU_ROUTE_TMA_STORE ? buf : 0];
        const uint64_t tile_base = (uint64_t)tile * kBinTile;
        if (tile_tma(tile)) {
            while (!mbar_try_wait(&S.mbar, parity)) {
            }
            parity ^= 1;
        } else { // ragged tail / unaligned input: guarded loads, bytes past n_valid read as 0 (= not a base)
            for (int i = tid; i < (kBinTile + kHalo); i += kRouteThreads) {
                const uint64_t p = tile_base + i;
                S.ascii[i] = p < P.n_valid ? P.bases[p] : (uint8_t)0;
            }
            __syncthreads();
        }
        uint32_t *Wb = S.W[buf], *Bb = S.B[buf];
        for (int gi = tid; gi < kBinGroups + 2; gi += kRouteThreads) {
            const uint4 v = reinterpret_cast<const uint4 *>(S.ascii)[gi];
            uint32_t c0, c1, c2, c3, b0, b1, b2, b3;
            pack4(v.x, c0, b0);
            pack4(v.y, c1, b1);
            pack4(v.z, c2, b2);
            pack4(v.w, c3, b3);
            Wb[gi] = (c0 << 24) | (c1 << 16) | (c2 << 8) | c3;
            Bb[gi] = (b0 << 12) | (b1 << 8) | (b2 << 4) | b3;
        }
        __syncthreads();
        if (tid == 0) { // ascii is free again: fetch the next tile while this one is hashed
            const int nt = tile + gridDim.x;
            if (nt < P.n_tiles && tile_tma(nt)) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&S.mbar, kBinTile + kHalo);
                tma_load_1d(S.ascii, P.bases + (uint64_t)nt * kBinTile, kBinTile + kHalo, &S.mbar);
            }
        }
        // kRoutePerThread k-mers per thread, kept in registers until every key of the tile has its place; a thread
        // owns (half of) one 16-base group and sees its 48-base window
        const uint32_t grp = kRoutePerThread == 16 ? (uint32_t)tid : ((uint32_t)tid & (uint32_t)(kBinGroups - 1));
        const bool second_half = kRoutePerThread != 16 && (uint32_t)tid >= (uint32_t)kBinGroups; // warp-uniform
        const uint64_t p0 = tile_base + (uint64_t)grp * 16;
        const uint64_t bad48 = ((uint64_t)Bb[grp] << 32) | ((uint64_t)Bb[grp + 1] << 16) | Bb[grp + 2];
        const uint32_t nst = p0 >= P.n_starts ? 0u : (uint32_t)min((uint64_t)16, P.n_starts - p0);
        Cut16 cut;
        cut.init(Wb[grp], Wb[grp + 1], Wb[grp + 2], bad48, k, nst);
        uint64_t hk[kRoutePerThread];
        uint32_t meta[kRoutePerThread / 2]; // two keys per word: bit 15 valid, bits 12-14 owner, bits 0-11 rank inside the owner's run
#pragma unroll
        for (uint32_t i = 0; i < (uint32_t)kRoutePerThread / 2; i++) meta[i] = 0;
        auto one = [&](uint32_t r, uint32_t i) { // r = start inside the group, i = slot in this thread's registers: both constants after unrolling
            uint32_t lo, hi;
            cut.key<CANON, WIDE>(r, kmask_lo, kmask_hi, lo, hi);
            const uint64_t h = mix_words<WIDE>(lo, hi, sh, kmask_lo, kmask_hi);
            const bool ok = cut.ok(r);
            const uint32_t o = ok ? owner_of(h, n_ranks) : 0xFFu;
            const uint32_t peers = __match_any_sync(0xffffffffu, o); // lanes whose key goes to the same owner
            const int leader = __ffs(peers) - 1;
            uint32_t base = 0;
            if (ok && (int)lane == leader) base = atomicAdd(&S.cnt[o], (uint32_t)__popc(peers));
            base = __shfl_sync(0xffffffffu, base, leader);
            hk[i] = h;
            if (ok) meta[i >> 1] |= (0x8000u | (o << 12) | (base + (uint32_t)__popc(peers & lt_mask))) << (16 * (i & 1));
        };
        if (!second_half) {
#pragma unroll
            for (uint32_t i = 0; i < (uint32_t)kRoutePerThread; i++) one(i, i);
        } else {
#pragma unroll
            for (uint32_t i = 0; i < (uint32_t)kRoutePerThread; i++) one(16 - kRoutePerThread + i, i);
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t a = 0, ga = 0;
#if KATGPU_ROUTE_TMA_STORE
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); // the stores out of THIS pool (two tiles ago) have read it
#endif
            for (uint32_t d = 0; d < n_ranks; d++) {
                a = (a + 1u) & ~1u; // runs start 16-byte aligned in the pool (bulk-copy source)
                S.off[d] = a;
                S.goff[d] = ga;
                const uint32_t m = S.nc[d] + S.cnt[d];
                a += m;
                ga += m / kRouteLine * (kRouteLine / 4);
            }
            S.off[n_ranks] = a;
            S.goff[n_ranks] = ga;
        }
        __syncthreads();
        if (tid < (int)(n_ranks * kRouteLine)) { // carried keys open each run
            const uint32_t d = tid / kRouteLine, j = tid % kRouteLine;
            if (j < S.nc[d]) pool[S.off[d] + j] = S.carry[d][j];
        }
#pragma unroll
        for (uint32_t r = 0; r < (uint32_t)kRoutePerThread; r++) {
            const uint32_t m = (meta[r >> 1] >> (16 * (r & 1))) & 0xFFFFu;
            const uint32_t d = (m >> 12) & 7u;
            if (m & 0x8000u) pool[S.off[d] + S.nc[d] + (m & 0xFFFu)] = hk[r];
        }
        __syncthreads();
#if KATGPU_ROUTE_TMA_STORE
        if (tid == 0) { // one bulk store per owner: whole lines, local HBM for the own rank, NVLink for the others
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            for (uint32_t d = 0; d < n_ranks; d++) {
                const uint32_t w = (S.nc[d] + S.cnt[d]) / kRouteLine * kRouteLine;
                if (w)
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(P.seg[d] + (size_t)blockIdx.x * P.cap + S.cur[d]),
                                 "r"(smem_u32(pool + S.off[d])), "r"(w * 8u)
                                 : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
#else
        // copy-out: every thread stores 32 bytes, a warp 1 KiB of consecutive lines — local HBM for
        // the own rank, NVLink for the others
        const uint32_t groups = S.goff[n_ranks];
        for (uint32_t gi = tid; gi < groups; gi += kRouteThreads) {
            uint32_t d = 0;
#pragma unroll
            for (uint32_t j = 1; j < (uint32_t)kMaxRanks; j++) d += (j < n_ranks && gi >= S.goff[j]) ? 1u : 0u;
            const uint32_t q = (gi - S.goff[d]) * 4;
            const uint64_t *src = pool + S.off[d] + q;
            uint64_t *dst = P.seg[d] + (size_t)blockIdx.x * P.cap + S.cur[d] + q;
            asm volatile("st.global.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "l"(src[0]), "l"(src[1]), "l"(src[2]), "l"(src[3]) : "memory");
        }
#endif
        __syncthreads();
        if (tid < (int)(n_ranks * kRouteLine)) { // what did not fill a line waits for the next tile
            const uint32_t d = tid / kRouteLine, j = tid % kRouteLine;
            const uint32_t m = S.nc[d] + S.cnt[d], w = m / kRouteLine * kRouteLine;
            if (j < m - w) S.carry[d][j] = pool[S.off[d] + w + j];
        }
        __syncthreads();
        if (tid < (int)n_ranks) {
            const uint32_t m = S.nc[tid] + S.cnt[tid], w = m / kRouteLine * kRouteLine;
            S.cur[tid] += w;
            S.nc[tid] = m - w;
            S.cnt[tid] = 0;
        }
        // (the pack phase of the next iteration ends with a barrier before cnt / nc are touched again)
    }
    __syncthreads();
#if KATGPU_ROUTE_TMA_STORE
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); // every run has landed before the lengths are published
#endif
    if (tid < (int)(n_ranks * kRouteLine)) { // the last, partial line of every run
        const uint32_t d = tid / kRouteLine, j = tid % kRouteLine;
        if (j < S.nc[d]) P.seg[d][(size_t)blockIdx.x * P.cap + S.cur[d] + j] = S.carry[d][j];
    }
    if (tid < (int)n_ranks) P.len[tid][blockIdx.x] = S.cur[tid] + S.nc[tid];
}

size_t route_smem_bytes() { return sizeof(RouteSmem); }

cudaError_t route_configure(int *blocks_per_sm) {
    cudaError_t e;
    const int smem = (int)sizeof(RouteSmem);
    if ((e = cudaFuncSetAttribute(route_bases_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(route_bases_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(route_bases_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(route_bases_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
    int a = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, route_bases_kernel<true, true>, kRouteThreads, sizeof(RouteSmem));
    *blocks_per_sm = a < 2 ? a : 2;
    return e;
}

void launch_route_bases(const RouteParams &P, int grid, cudaStream_t st) {
    const bool canon = P.g.canonical != 0, wide = P.g.b >= 32;
    const size_t smem = sizeof(RouteSmem);
    if (canon && wide) route_bases_kernel<true, true><<<grid, kRouteThreads, smem, st>>>(P);
    else if (canon) route_bases_kernel<true, false><<<grid, kRouteThreads, smem, st>>>(P);
    else if (wide) route_bases_kernel<false, true><<<grid, kRouteThreads, smem, st>>>(P);
    else route_bases_kernel<false, false><<<grid, kRouteThreads, smem, st>>>(P);
}

} // namespace katgpu
