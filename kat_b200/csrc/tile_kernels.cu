// tile_kernels.cu — the shared-memory counting path.
//
// What the memory system of a B200 allows (profiles/r1_microbench_memory_system.txt): random
// 8-byte read-modify-writes reach 20 G/s in HBM and 40-70 G/s even when the 32-64 MB they touch
// sit in L2, because every atomic or partial store is a sector RMW at the L2 slice; shared-memory
// atomics run at 1000 G/s chip-wide and full-sector (32-byte) stores at 90-170 G/s.  So the
// table is never updated in place.  Instead:
//
//   B1  bin_bases_kernel   extract + canonicalise + hash (extract.cuh: Cut16, mix_words), then
//       bin the hashed keys by the top Pb1 bits (the table "slice") into per-CTA segments.
//       Keys are staged per bin in a shared-memory ring (one shared atomic hands out the
//       position) and leave the SM only as full 32-byte sectors.
//   B2  bin_keys_kernel    one CTA per (slice, share): bins the slice's keys by the next Pb2
//       bits — the 64 KB tile of the table they fall in — with the same staging ring.
//   T   tile_insert_kernel loads one tile of the table (2048 buckets) into shared memory with
//       TMA bulk copies, applies the tile's keys with shared-memory atomics (hit: add to the
//       count field; miss: CAS into the first free slot of the home bucket), and writes the tile
//       back with one TMA bulk store.  Keys whose home bucket is full (they need a probe
//       outside the tile) or whose count field could carry go to a side list and take the
//       general random-access path afterwards (~1 % of the keys at load 0.3).
//
// The table layout (table.cuh) is unchanged, so every walk kernel works on the result.
// Replaces the same reference loop as count_kernels.cu (lib/src/jellyfish_helper.cc:202-211,
// JFI/hash_counter.hpp:98-130, JFI/large_hash_array.hpp:298-302).
#include "tile_kernels.cuh"

#include <cstdlib>
#include "extract.cuh"

namespace katgpu {

// ---- side list ------------------------------------------------------------------------------
struct SideTally {
    unsigned long long direct_new = 0, direct_total = 0;
};
// From a binning kernel (no tile is held in shared memory anywhere): park the key, or — list
// full — insert it right away the general way.
static __device__ __noinline__ void side_push_or_insert(const SideList &S, const TableGeom &g, uint64_t h, SideTally &t) {
    const unsigned long long idx = atomicAdd(S.n, 1ULL);
    if (idx < S.cap) {
        S.keys[idx] = h;
        atomicAdd(S.reserved, 1ULL);
        return;
    }
    atomicAdd(S.n, ~0ULL); // undo
    Bucket bk = {};
    t.direct_new += insert_hashed(g, h, 1, bk, false) == 1;
    t.direct_total++;
}
__device__ __forceinline__ void tally_flush(const TableGeom &g, SideTally &t) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        t.direct_total += __shfl_down_sync(0xffffffffu, t.direct_total, d);
        t.direct_new += __shfl_down_sync(0xffffffffu, t.direct_new, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (t.direct_total) atomicAdd(&g.ctr->n_total, t.direct_total);
        if (t.direct_new) atomicAdd(&g.ctr->n_distinct, t.direct_new);
    }
}

// ---- staging ring: F bins x R slots, slot-major so that the flush (one thread per bin) reads
// consecutive words -------------------------------------------------------------------------------
template <uint32_t G> // keys per granule written to a segment: 4 (one sector) or 8 (lane pairs write 64 bytes per request)
struct BinCore {
    uint64_t *ring;  // [R][F]
    uint32_t *cnt;   // [F] keys staged (may run past R: those keys went to the side list)
    uint32_t *cur;   // [F] keys already in the destination segment
    uint32_t F, R, shift, mask;
    uint32_t ge;            // effective granule: 8 keys (64 bytes, written by a lane pair in ONE store instruction) when the
                            // CTA has two threads per bin, else 4 keys (one 32-byte sector per thread)
    uint32_t cnt_s, ring_s; // the same arrays as 32-bit shared-window addresses (append: no generic-address arithmetic per key)
    uint64_t *seg0;  // destination segment of bin 0
    uint64_t seg_stride; // keys between the segments of consecutive bins
    uint32_t cap;
    uint32_t *len0;  // &len[segment of bin 0]
    uint32_t len_stride;

    __device__ __forceinline__ void init(const BinDst &d, uint32_t offset, unsigned char *smem) {
        F = d.F, R = d.R, shift = d.shift, mask = d.mask, cap = d.cap;
        ge = (G == 8 && blockDim.x == 2 * F) ? 8u : (G == 8 && (F & 31u) == 0 && KATGPU_OLD_G8 ? 8u : 4u);
        cnt = reinterpret_cast<uint32_t *>(smem);
        cur = cnt + F;
        ring = reinterpret_cast<uint64_t *>(smem + (((size_t)F * 8 + 15) & ~(size_t)15));
        cnt_s = smem_u32(cnt);
        ring_s = smem_u32(ring);
        seg0 = d.keys + (uint64_t)offset * d.cap;
        seg_stride = (uint64_t)d.stride * d.cap;
        len0 = d.len + offset;
        len_stride = d.stride;
        for (uint32_t i = threadIdx.x; i < F; i += blockDim.x) {
            cnt[i] = 0;
            cur[i] = len0[(size_t)i * len_stride];
        }
    }
    __device__ __forceinline__ void append(uint64_t h, const SideList &S, const TableGeom &g, SideTally &t) {
        const uint32_t d = (uint32_t)(h >> shift) & mask;
        uint32_t pos;
        asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(pos) : "r"(cnt_s + d * 4u) : "memory");
        if (pos < R) asm volatile("st.shared.u64 [%0], %1;" ::"r"(ring_s + (pos * F + d) * 8u), "l"(h) : "memory");
        else side_push_or_insert(S, g, h, t); // ring full (sized so that this is rare)
    }
    // One bin, one thread, any state: unaligned segment head (a previous launch ended there),
    // final flush (tails are written too), full segment (keys go to the side list).
    __device__ __forceinline__ void flush_bin_serial(uint32_t d, uint32_t n, uint32_t c, bool final, const SideList &S, const TableGeom &g, SideTally &t) {
        if (c + n > cap) { // segment full (sized by the host so that this is rare)
            for (uint32_t i = 0; i < n; i++) side_push_or_insert(S, g, ring[(size_t)i * F + d], t);
            cnt[d] = 0;
            return;
        }
        uint64_t *seg = seg0 + (uint64_t)d * seg_stride;
        const uint32_t aligned_end = (c + n) & ~(ge - 1);
        const uint32_t w = final ? n : (aligned_end > c ? aligned_end - c : 0u);
        uint32_t i = 0;
        for (; i < w && ((c + i) & 3u); i++) seg[c + i] = ring[(size_t)i * F + d]; // head of a segment a previous launch left unaligned
        for (; i + 4 <= w; i += 4) {
            const uint64_t k0 = ring[(size_t)i * F + d], k1 = ring[(size_t)(i + 1) * F + d], k2 = ring[(size_t)(i + 2) * F + d],
                           k3 = ring[(size_t)(i + 3) * F + d];
            asm volatile("st.global.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(seg + c + i), "l"(k0), "l"(k1), "l"(k2), "l"(k3) : "memory");
        }
        for (; i < w; i++) seg[c + i] = ring[(size_t)i * F + d]; // tail (final flush only)
        const uint32_t r = n - w; // < G
        for (uint32_t j = 0; j < r; j++) ring[(size_t)j * F + d] = ring[(size_t)(w + j) * F + d];
        cnt[d] = r;
        cur[d] = c + w;
    }
    // Move staged keys to the segments in whole granules of G keys (32-byte sector stores);
    // `final` also writes the tails.  Collective: every thread of the CTA calls it.
    // G == 8: the two sectors of a granule are written by ADJACENT LANES in one store instruction
    // (lane 2j and 2j+1 first serve bin 2j, then bin 2j+1; the owner's fill state travels by
    // shuffle), so that HBM sees one 64-byte request instead of two sector requests — it takes
    // ~35 G scattered write requests per second whatever their size.
    __device__ __forceinline__ void flush(bool final, const SideList &S, const TableGeom &g, SideTally &t) {
        __syncthreads();
        if (G == 8 && blockDim.x == 2 * F && !final) {
            // two threads per bin: lanes 2j and 2j+1 write the two sectors of each 64-byte granule of bin j in the
            // same store instruction -> one 64-byte request; no state has to travel between lanes
            const uint32_t d = threadIdx.x >> 1, sub = threadIdx.x & 1u;
            const uint32_t n = min(cnt[d], R), c = cur[d];
            const bool simple = n > 0 && (c & 7u) == 0 && c + n <= cap;
            const uint32_t w = simple ? (n & ~7u) : 0u;
            const uint64_t *col = ring + d;
            uint64_t *dst = seg0 + (uint64_t)d * seg_stride + c;
            for (uint32_t i = 4 * sub; i < w; i += 8) {
                const uint64_t k0 = col[(size_t)i * F], k1 = col[(size_t)(i + 1) * F], k2 = col[(size_t)(i + 2) * F], k3 = col[(size_t)(i + 3) * F];
                asm volatile("st.global.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "l"(k0), "l"(k1), "l"(k2), "l"(k3) : "memory");
            }
            __syncwarp(); // both lanes of the pair have read the bin's rows before its leftover moves
            if (sub == 0) {
                if (simple) {
                    const uint32_t r = n - w; // < 8
                    if (w)
                        for (uint32_t j = 0; j < r; j++) ring[(size_t)j * F + d] = ring[(size_t)(w + j) * F + d];
                    cnt[d] = r;
                    cur[d] = c + w;
                } else if (n)
                    flush_bin_serial(d, n, c, false, S, g, t); // unaligned head (a previous launch ended there) or a full segment
                else
                    cnt[d] = 0;
            }
            __syncthreads();
            return;
        }
        if (ge == 8 && !final) {
            const uint32_t lane = threadIdx.x & 31u, sub = lane & 1u;
            for (uint32_t d = threadIdx.x; d < F; d += blockDim.x) { // warp-uniform trip count (F and blockDim are multiples of 32)
                const uint32_t n = min(cnt[d], R), c = cur[d];
                const bool simple = n > 0 && (c & 7u) == 0 && c + n <= cap;
                const uint32_t w = simple ? (n & ~7u) : 0u;
                if (!simple) {
                    if (n) flush_bin_serial(d, n, c, false, S, g, t);
                    else cnt[d] = 0;
                }
#pragma unroll
                for (uint32_t phase = 0; phase < 2; phase++) {
                    const uint32_t src = (lane & ~1u) | phase;
                    const uint32_t wb = __shfl_sync(0xffffffffu, w, src), cb = __shfl_sync(0xffffffffu, c, src);
                    const uint32_t db = (d & ~1u) | phase;
                    const uint64_t *col = ring + db;
                    uint64_t *dst = seg0 + (uint64_t)db * seg_stride + cb;
                    for (uint32_t i = 4 * sub; i < wb; i += 8) {
                        const uint64_t k0 = col[(size_t)i * F], k1 = col[(size_t)(i + 1) * F], k2 = col[(size_t)(i + 2) * F], k3 = col[(size_t)(i + 3) * F];
                        asm volatile("st.global.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "l"(k0), "l"(k1), "l"(k2), "l"(k3) : "memory");
                    }
                }
                __syncwarp(); // both lanes of a pair have read a bin's rows before its owner moves the leftover
                if (simple) {
                    const uint32_t r = n - w; // < 8
                    if (w)
                        for (uint32_t j = 0; j < r; j++) ring[(size_t)j * F + d] = ring[(size_t)(w + j) * F + d];
                    cnt[d] = r;
                    cur[d] = c + w;
                }
            }
            __syncthreads();
            return;
        }
        for (uint32_t d = threadIdx.x; d < F; d += blockDim.x) {
            const uint32_t n = min(cnt[d], R), c = cur[d];
            if (n == 0) {
                cnt[d] = 0;
                continue;
            }
            flush_bin_serial(d, n, c, final, S, g, t);
        }
        __syncthreads();
    }
    __device__ __forceinline__ void store_lengths() { // after flush(true)
        for (uint32_t i = threadIdx.x; i < F; i += blockDim.x) len0[(size_t)i * len_stride] = cur[i];
    }
};

// ---- HyperLogLog over a 1/1024 sample of the key space (distinct-key estimate for sizing) ----
__device__ __forceinline__ uint64_t hll_hash64(uint64_t h) {
    uint64_t x = h * 0xD6E8FEB86659FD93ULL;
    x ^= x >> 32;
    x *= 0xD6E8FEB86659FD93ULL;
    x ^= x >> 32;
    return x;
}
__device__ __forceinline__ void hll_sample(uint32_t *hll, uint64_t h, uint32_t sshift) {
    if ((((uint32_t)(h >> sshift)) & ((1u << kTileHllSampleShift) - 1)) != 0) return; // rare below
    const uint64_t x = hll_hash64(h);
    atomicMax(&hll[x >> (64 - kTileHllBits)], (uint32_t)__clzll((x << kTileHllBits) | (1ULL << (kTileHllBits - 1))) + 1);
}

// ---- B1: base stream -> level-1 segments -------------------------------------------------------
struct __align__(128) BinBasesHead {
    uint8_t ascii[kBin1Tile + kHalo];
    uint32_t W[2][kBin1Groups + 2];
    uint32_t B[2][kBin1Groups + 2];
    uint64_t mbar;
    uint32_t hll[kTileHllM];
};

template <bool CANON, bool WIDE>
__global__ void __launch_bounds__(kBin1Threads, kBin1MinCtas) bin_bases_kernel(const __grid_constant__ BinBasesParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    BinBasesHead &S = *reinterpret_cast<BinBasesHead *>(smem_raw);
    const TableGeom &g = P.g;
    const int tid = threadIdx.x;
    const uint32_t k = g.k, sh = g.sh;
    const uint32_t kmask_lo = (uint32_t)g.kmask, kmask_hi = (uint32_t)(g.kmask >> 32);
    const uint32_t sshift = P.dst.shift - kTileHllSampleShift; // host: dst.shift >= 10
    BinCore<kBin1Granule> C;
    C.init(P.dst, blockIdx.x, smem_raw + sizeof(BinBasesHead));
    for (uint32_t i = tid; i < (uint32_t)kTileHllM; i += kBin1Threads) S.hll[i] = 0;
    SideTally tally;

    auto tile_tma = [&](int t) { return P.tma_ok && ((uint64_t)(t + 1) * kBin1Tile + kHalo <= P.n_valid); };
    // static striding (per-CTA segment fill is predictable), rotated so that successive launches
    // hand their first tiles to different CTAs
    int tile = (int)((blockIdx.x + gridDim.x - P.cta_rot % gridDim.x) % gridDim.x);
    if (tid == 0) {
        mbar_init(&S.mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (tile < P.n_tiles && tile_tma(tile)) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&S.mbar, kBin1Tile + kHalo);
            tma_load_1d(S.ascii, P.bases + (uint64_t)tile * kBin1Tile, kBin1Tile + kHalo, &S.mbar);
        }
    }
    __syncthreads();
    uint32_t parity = 0, buf = 0;

    for (; tile < P.n_tiles; tile += gridDim.x, buf ^= 1) {
        const uint64_t tile_base = (uint64_t)tile * kBin1Tile;
        if (tile_tma(tile)) {
            while (!mbar_try_wait(&S.mbar, parity)) {
            }
            parity ^= 1;
        } else { // ragged tail / unaligned input: guarded loads, bytes past n_valid read as 0 (= not a base)
            for (int i = tid; i < (kBin1Tile + kHalo); i += kBin1Threads) {
                const uint64_t p = tile_base + i;
                S.ascii[i] = p < P.n_valid ? P.bases[p] : (uint8_t)0;
            }
            __syncthreads();
        }
        uint32_t *Wb = S.W[buf], *Bb = S.B[buf];
        for (int gi = tid; gi < kBin1Groups + 2; gi += kBin1Threads) {
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
                mbar_expect_tx(&S.mbar, kBin1Tile + kHalo);
                tma_load_1d(S.ascii, P.bases + (uint64_t)nt * kBin1Tile, kBin1Tile + kHalo, &S.mbar);
            }
        }
        // kBin1PerThread k-mers per thread: a thread owns (half of) one 16-base group and sees its 48-base window
        const uint32_t grp = kBin1PerThread == 16 ? (uint32_t)tid : ((uint32_t)tid & (uint32_t)(kBin1Groups - 1));
        const bool second_half = kBin1PerThread != 16 && (uint32_t)tid >= (uint32_t)kBin1Groups; // warp-uniform
        const uint64_t p0 = tile_base + (uint64_t)grp * 16;
        const uint64_t bad48 = ((uint64_t)Bb[grp] << 32) | ((uint64_t)Bb[grp + 1] << 16) | Bb[grp + 2];
        const uint32_t nst = p0 >= P.n_starts ? 0u : (uint32_t)min((uint64_t)16, P.n_starts - p0);
        Cut16 cut;
        cut.init(Wb[grp], Wb[grp + 1], Wb[grp + 2], bad48, k, nst);
        auto one = [&](uint32_t r) { // r is a compile-time constant after unrolling
            uint32_t lo, hi;
            cut.key<CANON, WIDE>(r, kmask_lo, kmask_hi, lo, hi);
            const uint64_t h = mix_words<WIDE>(lo, hi, sh, kmask_lo, kmask_hi);
            if (cut.ok(r)) {
                C.append(h, P.side, g, tally);
                hll_sample(S.hll, h, sshift);
            }
        };
        if (!second_half) {
#pragma unroll
            for (uint32_t r = 0; r < (uint32_t)kBin1PerThread; r++) one(r);
        } else {
#pragma unroll
            for (uint32_t r = 16 - kBin1PerThread; r < 16; r++) one(r);
        }
        C.flush(false, P.side, g, tally);
    }
    C.flush(true, P.side, g, tally);
    C.store_lengths();
    for (uint32_t i = tid; i < (uint32_t)kTileHllM; i += kBin1Threads)
        if (S.hll[i]) atomicMax(P.hll + i, S.hll[i]);
    tally_flush(g, tally);
}

// ---- level 1 of keys that are already hashed (received from other ranks): keys -> level-1 segments ---
// Two shapes: 1024 threads (one CTA owns an SM: the default) and 512 threads with at most 64 registers (half an
// SM, so that the route kernel of the NEXT round — NVLink bound from 4 GPUs on — could run beside it).  The half-SM
// pair was measured and lost: 4 GPUs 354 ms per step against 258 with whole-SM kernels that alternate (the small
// ring of the half-SM binning costs more than the idle SMs it fills); it stays selectable for experiments
// (KATGPU_BINH_THREADS=512 KATGPU_DIST_ROUTE_CTAS=1).
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 2048 / THREADS / 2) bin_hashed_kernel(const __grid_constant__ BinKeysParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t *hll = reinterpret_cast<uint32_t *>(smem_raw); // [kTileHllM]
    const TableGeom &g = P.g;
    const uint32_t tid = threadIdx.x;
    SideTally tally;
    BinCore<kBinHGranule> C1; // per-CTA destination segments; one CTA per SM, so its ring may use most of the shared memory
    C1.init(P.dst, blockIdx.x, smem_raw + kTileHllM * 4);
    for (uint32_t i = tid; i < (uint32_t)kTileHllM; i += THREADS) hll[i] = 0;
    __syncthreads();
    const uint32_t sshift = P.dst.shift - kTileHllSampleShift;
    constexpr int kU = kBinTile / THREADS; // keys per thread per tile of kBinTile keys; host: P.tile_keys == kBinTile * P.cpf sizes the ring
    // tiles first, first + step, ... of `n` keys at `keys`
    const uint32_t cpf = P.cpf ? P.cpf : 1u; // tiles binned between two flushes (the ring is sized for them)
    uint32_t since_flush = 0;
    auto bin_tiles = [&](const uint64_t *keys, uint64_t n, uint64_t first, uint64_t step) {
        const uint64_t n_tiles = (n + kBinTile - 1) / kBinTile;
        for (uint64_t tile = first; tile < n_tiles; tile += step) {
            const uint64_t lo = tile * kBinTile;
            constexpr int kSub = 4; // keys per thread whose loads are in flight together
#pragma unroll 1
            for (int u0 = 0; u0 < kU; u0 += kSub) {
                uint64_t kc[kSub];
#pragma unroll
                for (int u = 0; u < kSub; u++) { // every load of the batch in flight before its first key is binned
                    const uint64_t i = lo + (uint64_t)(u0 + u) * THREADS + tid;
                    kc[u] = i < n ? __ldg(keys + i) : 0;
                }
#pragma unroll
                for (int u = 0; u < kSub; u++) {
                    if (lo + (uint64_t)(u0 + u) * THREADS + tid < n) {
                        C1.append(kc[u], P.side, g, tally);
                        hll_sample(hll, kc[u], sshift);
                    }
                }
            }
            if (++since_flush >= cpf) {
                C1.flush(false, P.side, g, tally);
                since_flush = 0;
            }
        }
    };
    if (P.in_keys) // source A: one flat array, tiles strided over the grid
        bin_tiles(P.in_keys, P.n_keys, (blockIdx.x + gridDim.x - P.cta_rot % gridDim.x) % gridDim.x, gridDim.x);
    else // source C: the dense segments other ranks' route_bases_kernel wrote here, whole segments per CTA
        for (uint32_t sgm = blockIdx.x; sgm < P.n_src; sgm += gridDim.x) bin_tiles(P.seg_keys + (size_t)sgm * P.seg_cap, P.seg_len[sgm], 0, 1);
    C1.flush(true, P.side, g, tally);
    C1.store_lengths();
    for (uint32_t i = tid; i < (uint32_t)kTileHllM; i += THREADS)
        if (hll[i]) atomicMax(P.hll + i, hll[i]);
    tally_flush(g, tally);
}

// ---- B2: the level-1 segments of one slice, one share of them per CTA -> per-tile segments --------
// The keys arrive through a ring of TMA bulk copies (one elected thread issues them kBin2Stages - 1 pieces
// ahead; full / empty mbarriers), so no warp ever waits on a global load and no register holds a prefetched key.
struct B2Cursor {
    uint32_t s, base, len;
    const uint64_t *src;
};
__global__ void __launch_bounds__(kBin2Threads, 1) bin_keys_kernel(const __grid_constant__ BinKeysParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t(*kbuf)[kBin2Piece] = reinterpret_cast<uint64_t(*)[kBin2Piece]>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)kBin2Stages * kBin2Piece * 8), *empty = full + kBin2Stages;
    const TableGeom &g = P.g;
    const uint32_t tid = threadIdx.x;
    SideTally tally;
    BinCore<kBin2Granule> C;
    const uint32_t s_local = blockIdx.x / P.parts, q = blockIdx.x % P.parts;
    C.init(P.dst, (s_local * P.dst.F) * P.parts + q, smem_raw + kBin2StageBytes);
    const uint32_t seg_lo = (uint32_t)(((uint64_t)P.src_segs * q) / P.parts), seg_hi = (uint32_t)(((uint64_t)P.src_segs * (q + 1)) / P.parts);
    // the pieces of this CTA: every non-empty source segment in kBin2Piece-key steps (the same walk for the
    // producer, ahead, and for the consumers)
    auto seek = [&](B2Cursor &c) { // position on the next non-empty segment at or after c.s
        for (; c.s < seg_hi; c.s++) {
            const size_t sidx = (size_t)(P.first_slice + s_local) * P.src_segs + c.s;
            c.len = P.src_len[sidx];
            if (c.len) {
                c.src = P.src_keys + sidx * P.src_cap;
                c.base = 0;
                return;
            }
        }
    };
    auto advance = [&](B2Cursor &c) {
        c.base += kBin2Piece;
        if (c.base >= c.len) {
            c.s++;
            seek(c);
        }
    };
    B2Cursor cons{seg_lo, 0, 0, nullptr};
    seek(cons);
    B2Cursor prod = cons;
    uint32_t issued = 0, consumed = 0, since_flush = 0;
    if (tid == 0) {
        for (int i = 0; i < kBin2Stages; i++) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], kBin2Threads);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&]() { // thread 0: piece number `issued` into its stage, once every thread has read the stage's previous piece
        const uint32_t st = issued % kBin2Stages;
        if (issued >= (uint32_t)kBin2Stages)
            while (!mbar_try_wait(&empty[st], ((issued / kBin2Stages) - 1) & 1)) {
            }
        const uint32_t n = min((uint32_t)kBin2Piece, prod.len - prod.base);
        const uint32_t bytes = (n * 8 + 15) & ~15u; // (segment capacities are multiples of 16 keys: the round-up stays inside)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&full[st], bytes);
        tma_load_1d(kbuf[st], prod.src + prod.base, bytes, &full[st]);
        issued++;
        advance(prod);
    };
    if (tid == 0)
        for (int i = 0; i < kBin2Stages - 1 && prod.s < seg_hi; i++) issue();
    while (cons.s < seg_hi) {
        if (tid == 0 && prod.s < seg_hi) issue();
        const uint32_t st = consumed % kBin2Stages;
        while (!mbar_try_wait(&full[st], (consumed / kBin2Stages) & 1)) {
        }
        const uint32_t n = min((uint32_t)kBin2Piece, cons.len - cons.base);
        uint64_t k0 = 0, k1 = 0;
        if (tid < n) k0 = kbuf[st][tid];
        if (tid + kBin2Threads < n) k1 = kbuf[st][tid + kBin2Threads];
        mbar_arrive(&empty[st]); // this thread is done with the stage
        if (tid < n) C.append(k0, P.side, g, tally);
        if (tid + kBin2Threads < n) C.append(k1, P.side, g, tally);
        consumed++;
        advance(cons);
        if (++since_flush >= P.cpf) { // (uniform: every thread walks the same pieces)
            C.flush(false, P.side, g, tally);
            since_flush = 0;
        }
    }
    C.flush(true, P.side, g, tally);
    C.store_lengths();
    tally_flush(g, tally);
}

// ---- T: apply a tile's keys in shared memory ---------------------------------------------------
constexpr uint32_t kParkCap = 1024; // parked keys a CTA collects per tile before it touches the global side list
struct __align__(128) TileSmem {
    uint64_t slots[kTileBytes / 8];
    uint64_t park[kParkCap];
    uint64_t mbar;
    unsigned long long park_base;
    uint32_t total, go, pushed, park_n;
};

template <bool TAG32>
__global__ void __launch_bounds__(kTileThreads, KATGPU_TILE_MIN_CTAS) tile_insert_kernel(const __grid_constant__ TileParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TileSmem &S = *reinterpret_cast<TileSmem *>(smem_raw);
    const TableGeom &g = P.g;
    const uint32_t tid = threadIdx.x;
    const uint64_t tagmask = g.tagmask, rmask = g.rmask, probe1 = 1ULL << g.rbits, one = 1ULL << g.cshift;
    const uint32_t rbits = g.rbits, cshift_hi = g.cshift - 32, inc_hi = 1u << (g.cshift - 32); // host: cshift >= 32
    const uint32_t slots_s = smem_u32(S.slots);
    constexpr int kTileUnroll = KATGPU_TILE_UNROLL;
    unsigned long long my_new = 0, my_total = 0;
    uint32_t parity = 0;
    if (tid == 0) {
        mbar_init(&S.mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        S.pushed = 0;
        S.park_n = 0;
    }
    __syncthreads();
    for (uint32_t t = blockIdx.x; t < P.n_tiles; t += gridDim.x) {
        const uint32_t *len = P.len + (size_t)t * P.parts;
        uint64_t *gtile = g.slots + ((uint64_t)(P.first_tile + t) << (kTileBucketsLog2 + 2));
        if (tid == 0) {
            uint32_t total = 0;
            for (uint32_t q = 0; q < P.parts; q++) total += len[q];
            uint32_t go = total != 0;
            if (go) {
                // start the tile load first: it does not depend on the side-list check below
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&S.mbar, kTileBytes);
#pragma unroll
                for (int c = 0; c < kTileBytes; c += 16384)
                    tma_load_1d(reinterpret_cast<unsigned char *>(S.slots) + c, reinterpret_cast<const unsigned char *>(gtile) + c, 16384, &S.mbar);
                // make sure the side list can take every key of this tile, else leave the tile to the general kernel
                const unsigned long long r0 = atomicAdd(P.side.reserved, (unsigned long long)total);
                if (r0 + total > P.side.cap) {
                    atomicAdd(P.side.reserved, ~(unsigned long long)total + 1ULL);
                    const uint32_t at = atomicAdd(P.side.n_deferred, 1u);
                    if (at < P.side.deferred_cap) P.side.deferred[at] = t;
                    go = 0;
                }
            }
            S.total = total;
            S.go = go;
        }
        __syncthreads();
        const uint64_t total = S.total;
        if (total == 0) {
            __syncthreads(); // S.go / S.total are rewritten by thread 0 in the next iteration
            continue;
        }
        // the first keys travel while the tile does
        uint64_t kn[kTileUnroll];
        auto fetch = [&](const uint64_t *src, uint32_t n, uint32_t base) {
#pragma unroll
            for (int u = 0; u < kTileUnroll; u++) {
                const uint32_t i = base + u * kTileThreads + tid;
                kn[u] = i < n ? __ldg(src + i) : 0;
            }
        };
        const bool go = S.go != 0;
        if (go) fetch(P.keys + (size_t)t * P.parts * P.cap, len[0], 0);
        while (!mbar_try_wait(&S.mbar, parity)) {
        }
        parity ^= 1;
        if (!go) { // deferred: the tile was loaded for nothing
            __syncthreads();
            continue;
        }
        // a count field that could carry while this tile is in shared memory takes the general path
        const uint32_t safe_below = (uint32_t)(g.cmax > total ? g.cmax - total : 0); // cmax < 2^32
        // One key against its home bucket in shared memory.  Hit: a native 32-bit add on the high
        // word (the count field starts at bit cshift >= 32).  Miss: CAS into the first free slot
        // (slots fill in order and never empty, so the key is not further along its probe
        // sequence).  Full bucket, or a count that could carry: side list.
        // Parked keys are collected in shared memory and appended to the global side list with ONE
        // atomic per tile: at table load 0.56 almost 1 % of the keys park, and 66 M increments of
        // one global word serialise (measured: tile kernel 43 -> 77 ms with per-key pushes).
        auto park_key = [&](uint64_t h) {
            const uint32_t at = atomicAdd(&S.park_n, 1u);
            if (at < kParkCap) {
                S.park[at] = h;
                return;
            }
            const unsigned long long idx = atomicAdd(P.side.n, 1ULL); // buffer full; below cap: this tile's keys are reserved
            P.side.keys[idx] = h;
            atomicAdd(&S.pushed, 1u);
        };
        // general form: any tag width, retries when a CAS is lost, and follows the probe sequence while it
        // stays inside the tile (probe j looks at bucket home + j(j+1)/2, table.cuh): at load 0.56 one key in
        // 120 finds its home bucket full, and all but a few of those find room one or three buckets on
        // In the key loop only the home bucket is looked at (j_end = 1): a lane that walked on would hold its
        // whole warp back (one warp iteration in four has such a lane at load 0.56: measured 43 -> 80 ms).
        // Keys that need more are parked in shared memory and finished after the loop, densely, by all threads.
        auto insert_general = [&](uint64_t h, uint32_t j0, uint32_t j_end) {
            const uint32_t home = ((uint32_t)(h >> rbits)) & ((1u << kTileBucketsLog2) - 1);
            bool park = j0 >= j_end, done = false;
            for (uint32_t j = j0; j < j_end && !done && !park; j++) {
                const uint32_t bucket = home + ((j * (j + 1)) >> 1);
                if (bucket >= (1u << kTileBucketsLog2)) { // the probe leaves the tile
                    park = true;
                    break;
                }
                const uint32_t bsm = slots_s + (bucket << 5);
                const uint64_t want = (h & rmask) | ((uint64_t)(j + 1) << rbits);
                for (int attempt = 0;; attempt++) {
                    uint64_t w0, w1, w2, w3;
                    asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "r"(bsm));
                    asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(w2), "=l"(w3) : "r"(bsm + 16));
                    const bool m0 = (w0 & tagmask) == want, m1 = (w1 & tagmask) == want, m2 = (w2 & tagmask) == want, m3 = (w3 & tagmask) == want;
                    if (m0 | m1 | m2 | m3) {
                        const uint32_t hit = m0 ? 0u : m1 ? 8u : m2 ? 16u : 24u;
                        const uint32_t seen_hi = (uint32_t)((m0 ? w0 : m1 ? w1 : m2 ? w2 : w3) >> 32);
                        if ((seen_hi >> cshift_hi) < safe_below) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(bsm + hit + 4), "r"(inc_hi) : "memory");
                        else park = true;
                        done = true;
                        break;
                    }
                    if (w3 != 0) break; // this bucket is full and does not hold the key: next probe
                    if (attempt >= 6) { // a pathological CAS storm
                        park = true;
                        break;
                    }
                    const uint32_t emp = w0 == 0 ? 0u : w1 == 0 ? 8u : w2 == 0 ? 16u : 24u;
                    unsigned long long old;
                    asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(old) : "r"(bsm + emp), "l"(0ULL), "l"(want | one) : "memory");
                    if (old == 0) {
                        my_new++;
                        done = true;
                        break;
                    } // else: somebody took the slot (maybe for this very key); look at the bucket again
                }
            }
            if (park || !done) park_key(h);
        };
        // TAG32: rem + probe live in the low word, so does the "occupied" test (probe >= 1); one
        // straight-line pass covers a hit, a claim and a full bucket; only a lost CAS re-reads.
        auto insert_one = [&](uint64_t h) {
            if (!TAG32) {
                insert_general(h, 0, 1);
                return;
            }
            const uint32_t bsm = slots_s + ((((uint32_t)(h >> rbits)) & ((1u << kTileBucketsLog2) - 1)) << 5);
            const uint32_t want32 = ((uint32_t)h & (uint32_t)rmask) | (uint32_t)probe1, mask32 = (uint32_t)tagmask;
            uint32_t l0, h0, l1, h1, l2, h2, l3, h3;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(l0), "=r"(h0), "=r"(l1), "=r"(h1) : "r"(bsm));
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(l2), "=r"(h2), "=r"(l3), "=r"(h3) : "r"(bsm + 16));
            const bool m0 = (l0 & mask32) == want32, m1 = (l1 & mask32) == want32, m2 = (l2 & mask32) == want32, m3 = (l3 & mask32) == want32;
            if (m0 | m1 | m2 | m3) {
                const uint32_t hit = m0 ? 0u : m1 ? 8u : m2 ? 16u : 24u;
                const uint32_t seen_hi = m0 ? h0 : m1 ? h1 : m2 ? h2 : h3;
                if ((seen_hi >> cshift_hi) < safe_below) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(bsm + hit + 4), "r"(inc_hi) : "memory");
                else park_key(h);
                return;
            }
            if (l3 != 0) { // home bucket full: parked, finished after the loop
                park_key(h);
                return;
            }
            const uint32_t emp = l0 == 0 ? 0u : l1 == 0 ? 8u : l2 == 0 ? 16u : 24u;
            unsigned long long old;
            asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(old) : "r"(bsm + emp), "l"(0ULL), "l"(((uint64_t)inc_hi << 32) | want32) : "memory");
            if (old == 0) my_new++;
            else insert_general(h, 0, 1); // lost the slot: look again
        };
        for (uint32_t q = 0; q < P.parts; q++) {
            const uint32_t n = len[q];
            const uint64_t *src = P.keys + ((size_t)t * P.parts + q) * P.cap;
            if (q > 0) fetch(src, n, 0);
            for (uint32_t base = 0; base < n; base += kTileThreads * kTileUnroll) {
                uint64_t h[kTileUnroll];
#pragma unroll
                for (int u = 0; u < kTileUnroll; u++) h[u] = kn[u];
                if (base + kTileThreads * kTileUnroll < n) fetch(src, n, base + kTileThreads * kTileUnroll); // next round in flight during this one
#pragma unroll
                for (int u = 0; u < kTileUnroll; u++)
                    if (base + u * kTileThreads + tid < n) insert_one(h[u]);
            }
        }
        __syncthreads();
        { // second chance for the parked keys: the probe sequence inside the tile, then the side list
            const uint32_t n_park = min(S.park_n, kParkCap);
            constexpr uint32_t kMine = (kParkCap + kTileThreads - 1) / kTileThreads;
            uint64_t mine[kMine];
#pragma unroll
            for (uint32_t u = 0; u < kMine; u++) {
                const uint32_t i = tid + u * kTileThreads;
                mine[u] = i < n_park ? S.park[i] : 0;
            }
            __syncthreads();
            if (tid == 0) S.park_n = 0; // survivors are collected in the same buffer
            __syncthreads();
#pragma unroll
            for (uint32_t u = 0; u < kMine; u++)
                if (tid + u * kTileThreads < n_park) insert_general(mine[u], 0, (uint32_t)kTileProbes);
            __syncthreads();
        }
        const uint32_t parked = min(S.park_n, kParkCap);
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gtile), "r"(smem_u32(S.slots)), "r"((uint32_t)kTileBytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (parked) S.park_base = atomicAdd(P.side.n, (unsigned long long)parked);
        }
        __syncthreads();
        if (parked) { // the tile travels back to HBM while the parked keys are appended to the side list
            const unsigned long long base = S.park_base;
            for (uint32_t i = tid; i < parked; i += kTileThreads) P.side.keys[base + i] = S.park[i];
        }
        if (tid == 0) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // the tile has been read out of shared memory
            const uint32_t pushed = S.pushed + parked;
            S.pushed = 0;
            S.park_n = 0;
            my_total += total - pushed;
            if (total > pushed) atomicAdd(P.side.reserved, ~(unsigned long long)(total - pushed) + 1ULL); // release the unused reservation
        }
        __syncthreads();
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); // every tile written back before the CTA retires
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        my_total += __shfl_down_sync(0xffffffffu, my_total, d);
        my_new += __shfl_down_sync(0xffffffffu, my_new, d);
    }
    if ((tid & 31) == 0) {
        if (my_total) atomicAdd(&g.ctr->n_total, my_total);
        if (my_new) atomicAdd(&g.ctr->n_distinct, my_new);
    }
}

// Tiles the tile kernel left alone (side list nearly full): insert their keys the general way.
// tiles == nullptr: every segment 0 .. n_cap-1 (the whole pending buffer, general sweep).
__global__ void __launch_bounds__(256) insert_deferred_kernel(const __grid_constant__ TileParams P, const uint32_t *tiles, uint32_t n_cap) {
    const TableGeom &g = P.g;
    const uint32_t n = tiles ? min(*P.side.n_deferred, n_cap) : n_cap;
    unsigned long long my_new = 0, my_total = 0;
    for (uint32_t j = blockIdx.x; j < n; j += gridDim.x) {
        const uint32_t t = tiles ? tiles[j] : j;
        for (uint32_t q = 0; q < P.parts; q++) {
            const uint32_t len = P.len[(size_t)t * P.parts + q];
            const uint64_t *src = P.keys + ((size_t)t * P.parts + q) * P.cap;
            for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) {
                Bucket bk = {};
                my_new += insert_hashed(g, src[i], 1, bk, false) == 1;
                my_total++;
            }
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

// ---- host side -------------------------------------------------------------------------------
static size_t ring_bytes(uint32_t F, uint32_t R) { return (((size_t)F * 8 + 15) & ~(size_t)15) + (size_t)F * R * 8; }
size_t bin_bases_smem(uint32_t F, uint32_t R) { return sizeof(BinBasesHead) + ring_bytes(F, R); }
size_t bin_keys_smem(uint32_t F, uint32_t R) { return (size_t)kTileHllM * 4 + ring_bytes(F, R); }
size_t bin_keys_b_smem(uint32_t F, uint32_t R) { return (size_t)kBin2StageBytes + ring_bytes(F, R); }

cudaError_t tile_configure(int *bps_bases, int *bps_tile) {
    const int big = 200 << 10;
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(bin_bases_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(bin_bases_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(bin_bases_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(bin_bases_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(bin_keys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(bin_hashed_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, big)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(bin_hashed_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, big)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(tile_insert_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileSmem))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(tile_insert_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileSmem))) != cudaSuccess) return e;
    *bps_bases = kBin1MinCtas; // the host sizes the staging ring so that this many CTAs fit an SM
    int a = 0, b = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, tile_insert_kernel<true>, kTileThreads, sizeof(TileSmem))) != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, tile_insert_kernel<false>, kTileThreads, sizeof(TileSmem));
    *bps_tile = a < b ? a : b;
    return e;
}
void launch_bin_bases(const BinBasesParams &P, int grid, cudaStream_t st) {
    const size_t smem = bin_bases_smem(P.dst.F, P.dst.R);
    const bool canon = P.g.canonical != 0, wide = P.g.b >= 32;
    if (canon && wide) bin_bases_kernel<true, true><<<grid, kBin1Threads, smem, st>>>(P);
    else if (canon) bin_bases_kernel<true, false><<<grid, kBin1Threads, smem, st>>>(P);
    else if (wide) bin_bases_kernel<false, true><<<grid, kBin1Threads, smem, st>>>(P);
    else bin_bases_kernel<false, false><<<grid, kBin1Threads, smem, st>>>(P);
}
void launch_bin_keys(const BinKeysParams &P, int grid, cudaStream_t st) {
    // (KATGPU_BINH_THREADS=512: the half-SM shape, measured slower — see the comment at the kernel)
    static const bool half = getenv("KATGPU_BINH_THREADS") && atoi(getenv("KATGPU_BINH_THREADS")) == 512;
    if (P.seg_keys && half) bin_hashed_kernel<512><<<grid, 512, bin_keys_smem(P.dst.F, P.dst.R), st>>>(P);
    else if (P.in_keys || P.seg_keys) bin_hashed_kernel<1024><<<grid, 1024, bin_keys_smem(P.dst.F, P.dst.R), st>>>(P);
    else bin_keys_kernel<<<grid, kBin2Threads, bin_keys_b_smem(P.dst.F, P.dst.R), st>>>(P);
}
void launch_tile_insert(const TileParams &P, int grid, cudaStream_t st) {
    if (P.g.tagmask <= 0xFFFFFFFFULL) tile_insert_kernel<true><<<grid, kTileThreads, sizeof(TileSmem), st>>>(P);
    else tile_insert_kernel<false><<<grid, kTileThreads, sizeof(TileSmem), st>>>(P);
}
void launch_insert_deferred(const TileParams &P, const uint32_t *tiles, uint32_t n, cudaStream_t st) {
    insert_deferred_kernel<<<148 * 4, 256, 0, st>>>(P, tiles, n);
}

} // namespace katgpu
