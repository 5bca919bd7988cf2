// partition_kernels.cu — the L2-resident counting path (selectable with KATGPU_COUNT_PATH=partition).
//
// Measured on B200: inserting k-mers in read order makes every probe a random
// HBM line fetch and tops out at ~15.6 G inserts/s whatever the occupancy
// (profiles/r1_v0_count_kernel_ncu.txt).  So counting is split in two
// streaming phases, the GPU analogue of the reference's slice-wise table
// ownership (JFI/large_hash_array.hpp:487-503) applied to *insertion*:
//
//   P  partition_kernel  extract + canonicalise + hash (as extract_kernel) and
//      bucket the 64-bit hashed keys by the table slice they will land in
//      (top Pb bits of the hash).  Per tile: warp-level match.any ranking ->
//      per-warp counters -> one block scan -> counting sort in shared memory
//      -> coalesced 32-byte-aligned appends to this CTA's private segment of
//      every slice (no global atomics).
//   S  sweep_kernel  walks the slices in order; all CTAs work on the same
//      slice at (nearly) the same time, so the 32-64 MB of table it covers
//      stays in the 126 MB L2: bucket loads and atomics hit L2, HBM only sees
//      the key stream (8 B/k-mer, sequential) and one read + write-back of the
//      table.
#include "partition_kernels.cuh"
#include "extract.cuh"

namespace katgpu {

constexpr uint64_t kSentinel = ~0ULL; // padding key (segments are appended in multiples of 4 keys)

constexpr int kStageCap = kTile + 3 * kMaxParts; // every slice run is padded to a multiple of 4 keys

struct __align__(128) PartSmem {
    uint8_t ascii[kTile + kHalo];
    uint32_t W[kGroups + 2];
    uint32_t B[kGroups + 2];
    uint64_t mbar;
    uint32_t off[kMaxParts + 2];                     // exclusive prefix of the PADDED per-slice counts
    uint32_t gcur[kMaxParts];                        // this CTA's append cursor per slice (keys)
    int64_t sbase[kMaxParts];                        // this tile: key index in P.keys of stage[0]-relative slot 0 of the slice
    uint32_t cnt[kMaxParts];                         // keys of this tile per slice (shared atomics hand out ranks)
    uint32_t wscan[kPartThreads / 32];
    uint8_t ovf[kMaxParts];                          // this tile's run does not fit the slice segment
    uint32_t any_ovf;
    uint32_t hll[kHllM];
    uint64_t stage[kStageCap];
};

__device__ __forceinline__ uint64_t sel16(const uint64_t (&a)[16], int i) {
    uint64_t v = a[0];
#pragma unroll
    for (int j = 1; j < 16; j++) v = (i == j) ? a[j] : v;
    return v;
}
__device__ __forceinline__ uint64_t hll_hash(uint64_t h) {
    uint64_t x = h * 0xD6E8FEB86659FD93ULL;
    x ^= x >> 32;
    x *= 0xD6E8FEB86659FD93ULL;
    x ^= x >> 32;
    return x;
}

// KEYS = false: tiles of the base stream (extract + hash);  KEYS = true: tiles of already
// hashed keys (the owner side of the multi-GPU exchange).
template <bool KEYS>
__global__ void __launch_bounds__(kPartThreads, 2) partition_kernel(const __grid_constant__ PartitionParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    PartSmem &S = *reinterpret_cast<PartSmem *>(smem_raw);
    const TableGeom &g = P.g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t k = g.k;
    const uint32_t nparts = 1u << P.Pb;
    const uint32_t pshift = g.b - P.Pb;
    const bool sentinel_possible = g.b == 64; // only then can a real hash equal the padding value

    for (uint32_t i = tid; i < nparts; i += kPartThreads) S.gcur[i] = P.seglen[(size_t)i * gridDim.x + blockIdx.x];
    for (uint32_t i = tid; i < (uint32_t)kHllM; i += kPartThreads) S.hll[i] = 0;
    for (uint32_t i = tid; i < (uint32_t)kMaxParts; i += kPartThreads) S.cnt[i] = 0;
    auto tile_tma = [&](int t) { return P.tma_ok && ((uint64_t)(t + 1) * kTile + kHalo <= P.n_valid); };
    // static striding (per-CTA segment fill is predictable), rotated so that successive launches
    // hand their first tiles to different CTAs
    int tile = (int)((blockIdx.x + gridDim.x - P.cta_rot % gridDim.x) % gridDim.x);
    if (!KEYS && tid == 0) {
        mbar_init(&S.mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (tile < P.n_tiles && tile_tma(tile)) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&S.mbar, kTile + kHalo);
            tma_load_1d(S.ascii, P.bases + (uint64_t)tile * kTile, kTile + kHalo, &S.mbar);
        }
    }
    __syncthreads();
    uint32_t parity = 0;
    unsigned long long direct_new = 0, direct_total = 0;

    for (; tile < P.n_tiles; tile += gridDim.x) {
        const uint64_t tile_base = (uint64_t)tile * kTile;
        uint64_t hh[16];
        uint32_t rk[8]; // 16-bit rank inside (warp, slice), two per register
        uint32_t hll_mask = 0, sent_mask = 0;
        if (!KEYS) {
            if (tile_tma(tile)) {
                while (!mbar_try_wait(&S.mbar, parity)) {
                }
                parity ^= 1;
            } else {
                for (int i = tid; i < (kTile + kHalo); i += kPartThreads) {
                    const uint64_t p = tile_base + i;
                    S.ascii[i] = p < P.n_valid ? P.bases[p] : (uint8_t)0;
                }
                __syncthreads();
            }
            for (int gi = tid; gi < kGroups + 2; gi += kPartThreads) {
                const uint4 v = reinterpret_cast<const uint4 *>(S.ascii)[gi];
                uint32_t c0, c1, c2, c3, b0, b1, b2, b3;
                pack4(v.x, c0, b0);
                pack4(v.y, c1, b1);
                pack4(v.z, c2, b2);
                pack4(v.w, c3, b3);
                S.W[gi] = (c0 << 24) | (c1 << 16) | (c2 << 8) | c3;
                S.B[gi] = (b0 << 12) | (b1 << 8) | (b2 << 4) | b3;
            }
            __syncthreads();
            if (tid == 0) {
                const int nt = tile + gridDim.x;
                if (nt < P.n_tiles && tile_tma(nt)) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_expect_tx(&S.mbar, kTile + kHalo);
                    tma_load_1d(S.ascii, P.bases + (uint64_t)nt * kTile, kTile + kHalo, &S.mbar);
                }
            }
            // ---- A1: 16 rolled canonical k-mers per thread -> hashed keys ----
            const int gi = tid; // kGroups == kPartThreads
            const uint64_t p0 = tile_base + (uint64_t)gi * 16;
            const uint64_t hi = ((uint64_t)S.W[gi] << 32) | S.W[gi + 1];
            const uint64_t lo = (uint64_t)S.W[gi + 2] << 32;
            const uint64_t bad48 = ((uint64_t)S.B[gi] << 32) | ((uint64_t)S.B[gi + 1] << 16) | S.B[gi + 2];
            const uint32_t nst = p0 >= P.n_starts ? 0u : (uint32_t)min((uint64_t)16, P.n_starts - p0);
            Roll16 roll;
            roll.init(hi, lo, bad48, k, nst);
#pragma unroll
            for (uint32_t r = 0; r < 16; r++) {
                if (r > 0) roll.step(r, g.kmask);
                const uint64_t h = mix_key(g, roll.key(g.canonical));
                const bool ok = roll.ok(r);
                if (sentinel_possible) sent_mask |= (uint32_t)(ok && h == kSentinel) << r;
                hh[r] = ok ? h : kSentinel;
            }
        } else {
#pragma unroll
            for (uint32_t r = 0; r < 16; r++) { // coalesced: consecutive threads read consecutive keys
                const uint64_t i = tile_base + (uint64_t)r * kPartThreads + tid;
                const bool ok = i < P.n_keys;
                const uint64_t h = ok ? __ldg(P.in_keys + i) : kSentinel;
                if (sentinel_possible) sent_mask |= (uint32_t)(ok && h == kSentinel) << r;
                hh[r] = h;
            }
        }
        // ---- A2: rank every key inside its slice: one shared-memory atomic per key ----
#pragma unroll
        for (uint32_t r = 0; r < 16; r++) {
            const uint64_t h = hh[r];
            const bool real = h != kSentinel;
            hll_mask |= (uint32_t)(real && ((uint32_t)h & ((1u << kHllSampleShift) - 1)) == 0) << r;
            uint32_t rank = 0;
            if (real) rank = atomicAdd(&S.cnt[(uint32_t)(h >> pshift)], 1u);
            if (r & 1) rk[r >> 1] |= rank << 16;
            else rk[r >> 1] = rank;
        }
        // rare side work kept out of the unrolled loops
        if (hll_mask) { // distinct-key estimate: HyperLogLog over a 1/2^kHllSampleShift sample of the key space
#pragma unroll 1
            for (uint32_t m = hll_mask; m; m &= m - 1) {
                const uint64_t x = hll_hash(sel16(hh, __ffs(m) - 1));
                atomicMax(&S.hll[x >> (64 - kHllBits)], (uint32_t)__clzll((x << kHllBits) | (1ULL << (kHllBits - 1))) + 1);
            }
        }
        if (sent_mask) { // a real key that equals the padding value (k = 32 only): insert it directly
            for (uint32_t m = sent_mask; m; m &= m - 1) {
                Bucket bk = {};
                direct_new += insert_hashed(g, kSentinel, 1, bk, false) == 1;
                direct_total++;
            }
        }
        if (tid == 0) S.any_ovf = 0;
        __syncthreads();
        // ---- B: per-slice totals over warps, padded to 4 keys, then block scan over slices ----
        uint32_t mine = 0;
        if ((uint32_t)tid < nparts) { // one slice column per thread (nparts <= kPartThreads)
            const uint32_t run = S.cnt[tid];
            mine = (run + 3) & ~3u;
            const bool o = S.gcur[tid] + mine > P.seg_cap;
            S.ovf[tid] = o;
            if (o) S.any_ovf = 1;
            mine |= run << 16; // low 16 bits: padded count, high 16: real count
        }
        uint32_t inc = mine & 0xFFFFu;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) S.wscan[warp] = inc;
        __syncthreads();
        uint32_t wbase = 0;
#pragma unroll
        for (int w = 0; w < kPartThreads / 32; w++)
            if (w < warp) wbase += S.wscan[w];
        if ((uint32_t)tid < nparts) {
            const uint32_t padded = mine & 0xFFFFu, real_n = mine >> 16;
            const uint32_t o = wbase + inc - padded;
            S.off[tid] = o;
            S.sbase[tid] = (int64_t)(((size_t)tid * gridDim.x + blockIdx.x) * P.seg_cap + S.gcur[tid]) - (int64_t)o;
            for (uint32_t j = real_n; j < padded; j++) S.stage[o + j] = kSentinel; // <= 3 pad keys
            if ((uint32_t)tid == nparts - 1) S.off[nparts] = wbase + inc; // total (padded)
        }
        __syncthreads();
        // ---- C: counting-sort scatter into the stage ----
#pragma unroll
        for (uint32_t r = 0; r < 16; r++) {
            const uint64_t h = hh[r];
            if (h != kSentinel) {
                const uint32_t part = (uint32_t)(h >> pshift);
                const uint32_t rank = (rk[r >> 1] >> ((r & 1) * 16)) & 0xFFFFu;
                S.stage[S.off[part] + rank] = h;
            }
        }
        __syncthreads();
        // ---- D: flat, coalesced append: stage[i] goes to its slice's segment of this CTA ----
        {
            const uint32_t total = S.off[nparts];
            for (uint32_t i = tid; i < total; i += kPartThreads) {
                const uint64_t h = S.stage[i];
                if (h == kSentinel) continue; // padding slot: its slice's column thread writes it below
                const uint32_t part = (uint32_t)(h >> pshift);
                if (!S.ovf[part]) P.keys[S.sbase[part] + (int64_t)i] = h;
            }
            if ((uint32_t)tid < nparts && !S.ovf[tid]) { // <= 3 sentinel keys keep every append a multiple of 32 bytes
                const uint32_t padded = mine & 0xFFFFu, real_n = mine >> 16;
                uint64_t *dst = P.keys + ((size_t)tid * gridDim.x + blockIdx.x) * P.seg_cap + S.gcur[tid];
                for (uint32_t j = real_n; j < padded; j++) dst[j] = kSentinel;
            }
        }
        if (S.any_ovf) { // some segment is full: insert that run the slow (random-access) way; never lose a k-mer
            for (uint32_t part = warp; part < nparts; part += kPartThreads / 32) {
                if (!S.ovf[part]) continue;
                const uint32_t o = S.off[part], n = S.off[part + 1] - o;
                for (uint32_t j = lane; j < n; j += 32) {
                    const uint64_t h = S.stage[o + j];
                    if (h == kSentinel) continue;
                    Bucket bk = {};
                    direct_new += insert_hashed(g, h, 1, bk, false) == 1;
                    direct_total++;
                }
            }
        }
        __syncthreads();
        if ((uint32_t)tid < nparts) {
            if (!S.ovf[tid]) S.gcur[tid] += mine & 0xFFFFu;
            S.cnt[tid] = 0; // for the next tile
        }
        __syncthreads();
    }
    for (uint32_t i = tid; i < nparts; i += kPartThreads) P.seglen[(size_t)i * gridDim.x + blockIdx.x] = S.gcur[i];
    for (uint32_t i = tid; i < (uint32_t)kHllM; i += kPartThreads)
        if (S.hll[i]) atomicMax(P.hll + i, S.hll[i]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        direct_total += __shfl_down_sync(0xffffffffu, direct_total, d);
        direct_new += __shfl_down_sync(0xffffffffu, direct_new, d);
    }
    if (lane == 0) {
        if (direct_total) atomicAdd(&g.ctr->n_total, direct_total);
        if (direct_new) atomicAdd(&g.ctr->n_distinct, direct_new);
    }
}

// Streaming load of pending keys: do not let them displace the table slice in L2.
__device__ __forceinline__ uint64_t ld_key_stream(const uint64_t *p, uint64_t policy) {
    uint64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(policy));
    return v;
}

__global__ void __launch_bounds__(kSweepThreads, kSweepMinCtas) sweep_kernel(const __grid_constant__ SweepParams P) {
    const TableGeom &g = P.g;
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    unsigned long long my_new = 0, my_total = 0;
    // CTA -> (segment, piece): the same for every partition, so all CTAs advance
    // through the slices in step and the slice being filled stays in L2.
    const uint32_t seg = blockIdx.x / P.split, piece = blockIdx.x % P.split;
    for (uint32_t part = P.part_lo; part < P.part_hi; part++) {
        // Slice barrier: nobody starts slice `part` before every CTA has finished slice
        // part - window.  Without it per-SM speed differences let the CTAs drift tens of
        // slices apart and the table stops being L2-resident (measured: 77 % of bucket
        // loads missed L2).  All CTAs are co-resident (host checks), so spinning is safe.
        if (P.window && part >= P.part_lo + P.window) {
            if (threadIdx.x == 0) {
                const volatile uint32_t *c = P.slice_done + (part - P.window);
                while (*c < gridDim.x) __nanosleep(200);
            }
            __syncthreads();
        }
        // Stream the NEXT slice of the table into L2 with bulk prefetches (one share per CTA):
        // its first-touch misses would otherwise be random 32-byte HBM reads.
        if (threadIdx.x == 0 && P.prefetch) {
            const uint32_t nxt = part == P.part_lo ? part : part + 1; // first iteration: this slice, then one ahead
            for (uint32_t q = nxt; q <= part + 1 && q < P.part_hi; q++) {
                const uint64_t slice_bytes = P.slice_slots * 8;
                const uint64_t share = (slice_bytes / gridDim.x + 15) & ~15ULL;
                const uint64_t lo_b = (uint64_t)blockIdx.x * share;
                if (lo_b < slice_bytes) {
                    const char *pb = reinterpret_cast<const char *>(g.slots + (uint64_t)q * P.slice_slots) + lo_b;
                    uint64_t left = min(share, slice_bytes - lo_b);
                    while (left) {
                        const uint32_t n = left < 32768 ? (uint32_t)left : 32768u;
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pb), "r"(n) : "memory");
                        pb += n;
                        left -= n;
                    }
                }
            }
        }
        const size_t sidx = (size_t)part * P.n_segs + seg;
        const uint32_t len = P.seglen[sidx];
        const uint32_t per = ((len + P.split - 1) / P.split + 3) & ~3u;
        const uint32_t lo = min(len, piece * per), hi = min(len, lo + per);
        const uint64_t *keys = P.keys + sidx * P.seg_cap;
        for (uint32_t base = lo; base < hi; base += kSweepThreads * kSweepUnroll) {
            uint64_t h[kSweepUnroll];
            bool ok[kSweepUnroll];
            Bucket bk[kSweepUnroll];
#pragma unroll
            for (int u = 0; u < kSweepUnroll; u++) {
                const uint32_t i = base + u * kSweepThreads + threadIdx.x;
                h[u] = i < hi ? ld_key_stream(keys + i, pol) : kSentinel;
                ok[u] = h[u] != kSentinel;
                if (ok[u]) bk[u] = load_bucket(g.slots + (((h[u] >> g.rbits) & g.bmask) << 2));
            }
#if !KATGPU_SWEEP_TWOSTAGE
            if (P.experiment) { // measurement aid: 1 = key + bucket loads only, 2 = plus a blind RED on slot 0
#pragma unroll
                for (int u = 0; u < kSweepUnroll; u++) {
                    if (ok[u]) {
                        my_total += (bk[u].s[0] ^ bk[u].s[1] ^ bk[u].s[2] ^ bk[u].s[3]) == 0x123456789ULL;
                        if (P.experiment == 2)
                            asm volatile("red.global.add.u64 [%0], %1;" ::"l"(g.slots + (((h[u] >> g.rbits) & g.bmask) << 2)), "l"(1ULL << g.cshift) : "memory");
                    }
                }
                continue;
            }
#pragma unroll
            for (int u = 0; u < kSweepUnroll; u++) {
                if (ok[u]) {
                    my_total++;
                    my_new += insert_home(g, h[u], bk[u]) == 1;
                }
            }
#else
            InsertTicket tk[kSweepUnroll];
#pragma unroll
            for (int u = 0; u < kSweepUnroll; u++) { // issue every key's RED / CAS ...
                tk[u].state = 0;
                if (ok[u]) {
                    my_total++;
                    tk[u] = insert_issue(g, h[u], bk[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < kSweepUnroll; u++) // ... then look at the CAS results
                if (ok[u] && tk[u].state) my_new += insert_finish(g, h[u], tk[u]) == 1;
#endif
        }
        if (P.window) {
            __syncthreads();
            if (threadIdx.x == 0) atomicAdd(P.slice_done + part, 1u);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        my_total += __shfl_down_sync(0xffffffffu, my_total, d);
        my_new += __shfl_down_sync(0xffffffffu, my_new, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (my_total) atomicAdd(&g.ctr->n_total, my_total);
        if (my_new) atomicAdd(&g.ctr->n_distinct, my_new);
    }
}

size_t partition_smem_bytes() { return sizeof(PartSmem); }

cudaError_t partition_configure(int *blocks_per_sm) {
    cudaError_t e = cudaFuncSetAttribute(partition_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PartSmem));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(partition_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PartSmem));
    if (e != cudaSuccess) return e;
    int a = 0, b = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, partition_kernel<false>, kPartThreads, sizeof(PartSmem));
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, partition_kernel<true>, kPartThreads, sizeof(PartSmem));
    *blocks_per_sm = a < b ? a : b;
    return e;
}
void launch_partition(const PartitionParams &P, int grid, cudaStream_t st) {
    if (P.in_keys) partition_kernel<true><<<grid, kPartThreads, sizeof(PartSmem), st>>>(P);
    else partition_kernel<false><<<grid, kPartThreads, sizeof(PartSmem), st>>>(P);
}
cudaError_t sweep_occupancy(int *blocks_per_sm) {
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, sweep_kernel, kSweepThreads, 0);
}
void launch_sweep(const SweepParams &P, cudaStream_t st) { sweep_kernel<<<P.n_segs * P.split, kSweepThreads, 0, st>>>(P); }

} // namespace katgpu
