// count_kernels.cu — K1: extract + canonicalise + insert (the dominant kernel),
// K5: extract + route to owner ranks, and the hashed-key insert / rehash
// kernels.  sm_100a only.
//
// Replaces the reference's HOT LOOP A:
//   for (MerIterator mers(parser, canonical); mers; ++mers) ary.add(*mers, 1);
// (lib/src/jellyfish_helper.cc:202-211; JFI/mer_iterator.hpp:61-89;
//  JFI/hash_counter.hpp:98-130).  The rolling per-thread iterator becomes a
// tile-parallel scheme: a CTA stages 8 KiB (+32 B halo) of the ASCII base
// stream into shared memory with one TMA bulk copy (cp.async.bulk, mbarrier
// completion), packs it to 2 bits/base + 1 invalid bit/base, and every thread
// then cuts 16 consecutive k-mers out of the packed words with funnel shifts.
#include "count_kernels.cuh"
#include "extract.cuh"

namespace katgpu {

struct __align__(128) ExtractSmem {
    uint8_t ascii[kTile + kHalo];
    uint32_t W[kGroups + 2]; // 2-bit codes, 16 bases per word, first base in bits 31:30
    uint32_t B[kGroups + 2]; // invalid bits, 16 per word (low half), first base in bit 15
    uint64_t mbar;
    int next_tile;
    unsigned long long wtot[kThreads / 32][2];
    unsigned long long gbase[kMaxRanks];
};

// Stage tile `t` of the base stream: TMA when the whole tile + halo is in
// bounds and aligned, otherwise a guarded cooperative copy (ragged tail).
__device__ __forceinline__ bool tile_uses_tma(const ExtractParams &P, int t) {
    return P.tma_ok && ((uint64_t)(t + 1) * kTile + kHalo <= P.n_valid);
}

__device__ __forceinline__ void issue_tile_tma(const ExtractParams &P, ExtractSmem &S, int t) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&S.mbar, kTile + kHalo);
    tma_load_1d(S.ascii, P.bases + (uint64_t)t * kTile, kTile + kHalo, &S.mbar);
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, MODE == kModeInsert ? KATGPU_MIN_CTAS : 2)
    extract_kernel(const __grid_constant__ ExtractParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ExtractSmem &S = *reinterpret_cast<ExtractSmem *>(smem_raw);
    uint64_t *stage = reinterpret_cast<uint64_t *>(smem_raw + sizeof(ExtractSmem)); // ROUTE only
    const TableGeom &g = P.g;
    const int tid = threadIdx.x;
    const uint32_t k = g.k;

    if (tid == 0) {
        mbar_init(&S.mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        int t = (int)atomicAdd(P.tile_counter, 1u);
        S.next_tile = t;
        if (t < P.n_tiles && tile_uses_tma(P, t)) issue_tile_tma(P, S, t);
    }
    __syncthreads();
    int tile = S.next_tile;
    uint32_t parity = 0;
    unsigned long long my_total = 0, my_new = 0;

    while (tile < P.n_tiles) {
        const uint64_t tile_base = (uint64_t)tile * kTile;
        if (tile_uses_tma(P, tile)) {
            while (!mbar_try_wait(&S.mbar, parity)) {
            }
            parity ^= 1;
        } else {
            // ragged tail / unaligned input: guarded loads, bytes past n_valid read as 0 (= not a base)
            for (int i = tid; i < (kTile + kHalo); i += kThreads) {
                const uint64_t p = tile_base + i;
                S.ascii[i] = p < P.n_valid ? P.bases[p] : (uint8_t)0;
            }
            __syncthreads();
        }
        // ---- pack ----
        for (int gi = tid; gi < kGroups + 2; gi += kThreads) {
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
        // ascii is free again: fetch the next tile while this one is hashed
        if (tid == 0) {
            int t = (int)atomicAdd(P.tile_counter, 1u);
            S.next_tile = t;
            if (t < P.n_tiles && tile_uses_tma(P, t)) issue_tile_tma(P, S, t);
        }
        // ---- k-mers ----
#pragma unroll 1
        for (int pass = 0; pass < kGroups / kThreads; pass++) {
            const int gi = pass * kThreads + tid;
            const uint64_t p0 = tile_base + (uint64_t)gi * 16;
            const uint64_t hi = ((uint64_t)S.W[gi] << 32) | S.W[gi + 1];
            const uint64_t lo = (uint64_t)S.W[gi + 2] << 32;
            const uint64_t bad48 = ((uint64_t)S.B[gi] << 32) | ((uint64_t)S.B[gi + 1] << 16) | S.B[gi + 2];
            // number of k-mer starts this thread owns (stream end / chunk overlap)
            const uint32_t nst = p0 >= P.n_starts ? 0u : (uint32_t)min((uint64_t)16, P.n_starts - p0);

            Roll16 roll;
            roll.init(hi, lo, bad48, k, nst);
            if (MODE == kModeInsert) {
#pragma unroll 1
                for (uint32_t r0 = 0; r0 < 16; r0 += kUnroll) {
                    uint64_t h[kUnroll];
                    bool ok[kUnroll];
                    Bucket bk[kUnroll];
#pragma unroll
                    for (int u = 0; u < kUnroll; u++) {
                        if (r0 + u > 0) roll.step(r0 + u, g.kmask);
                        ok[u] = roll.ok(r0 + u);
                        h[u] = mix_key(g, roll.key(g.canonical));
                        if (ok[u]) bk[u] = load_bucket(g.slots + (((h[u] >> g.rbits) & g.bmask) << 2));
                    }
#pragma unroll
                    for (int u = 0; u < kUnroll; u++) {
                        if (ok[u]) {
                            const int res = insert_home(g, h[u], bk[u]);
                            my_total++;
                            my_new += (res == 1);
                        }
                    }
                }
            } else {
                // ---- ROUTE: bucket hashed keys by owner rank, CTA-aggregated ----
                uint64_t hh[16];
                uint64_t owners = 0; // 4 bits per k-mer, 0xF = none
                unsigned long long c0 = 0, c1 = 0; // 16-bit fields: owners 0-3, 4-7
#pragma unroll
                for (uint32_t r = 0; r < 16; r++) {
                    if (r > 0) roll.step(r, g.kmask);
                    const bool ok = roll.ok(r);
                    hh[r] = mix_key(g, roll.key(g.canonical));
                    const uint32_t o = ok ? owner_of(hh[r], P.n_ranks) : 0xFu;
                    owners |= (uint64_t)o << (4 * r);
                    if (ok) {
                        if (o < 4) c0 += 1ULL << (16 * o);
                        else c1 += 1ULL << (16 * (o - 4));
                    }
                }
                // block-wide exclusive scan of the packed per-owner counts
                const int lane = tid & 31, warp = tid >> 5;
                unsigned long long i0 = c0, i1 = c1;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const unsigned long long t0 = __shfl_up_sync(0xffffffffu, i0, d);
                    const unsigned long long t1 = __shfl_up_sync(0xffffffffu, i1, d);
                    if (lane >= d) {
                        i0 += t0;
                        i1 += t1;
                    }
                }
                if (lane == 31) {
                    S.wtot[warp][0] = i0;
                    S.wtot[warp][1] = i1;
                }
                __syncthreads(); // also: previous pass finished reading `stage`
                unsigned long long w0 = 0, w1 = 0, tot0 = 0, tot1 = 0;
#pragma unroll
                for (int w = 0; w < kThreads / 32; w++) {
                    if (w < warp) {
                        w0 += S.wtot[w][0];
                        w1 += S.wtot[w][1];
                    }
                    tot0 += S.wtot[w][0];
                    tot1 += S.wtot[w][1];
                }
                unsigned long long ex0 = i0 - c0 + w0, ex1 = i1 - c1 + w1; // exclusive prefix per owner
                uint32_t seg[kMaxRanks + 1];                                // segment starts in `stage`
                seg[0] = 0;
#pragma unroll
                for (int o = 0; o < kMaxRanks; o++) {
                    const uint32_t cnt = (uint32_t)(((o < 4 ? tot0 : tot1) >> (16 * (o & 3))) & 0xFFFFu);
                    seg[o + 1] = seg[o] + cnt;
                }
                if (tid < (int)P.n_ranks) {
                    const uint32_t cnt = seg[tid + 1] - seg[tid];
                    S.gbase[tid] = cnt ? atomicAdd(P.cursors + tid, (unsigned long long)cnt) : 0ULL;
                }
#pragma unroll
                for (uint32_t r = 0; r < 16; r++) {
                    const uint32_t o = (uint32_t)(owners >> (4 * r)) & 0xFu;
                    if (o != 0xFu) {
                        uint32_t off;
                        if (o < 4) {
                            off = (uint32_t)(ex0 >> (16 * o)) & 0xFFFFu;
                            ex0 += 1ULL << (16 * o);
                        } else {
                            off = (uint32_t)(ex1 >> (16 * (o - 4))) & 0xFFFFu;
                            ex1 += 1ULL << (16 * (o - 4));
                        }
                        stage[seg[o] + off] = hh[r];
                    }
                }
                __syncthreads();
                // coalesced copy-out, one contiguous run per owner
                const uint32_t total = seg[kMaxRanks];
                for (uint32_t i = tid; i < total; i += kThreads) {
                    uint32_t o = 0;
#pragma unroll
                    for (int q = 1; q < kMaxRanks; q++) o += (i >= seg[q]);
                    const unsigned long long at = S.gbase[o] + (i - seg[o]);
                    if (at < P.out_cap) P.out[o][at] = stage[i];
                    else atomicExch(P.route_overflow, 1u);
                }
            }
        }
        __syncthreads(); // W/B (and next_tile) settled
        tile = S.next_tile;
    }
    // one atomic pair per warp for the instance / distinct counters
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


// Insert a list of hashed keys (multiplicity 1 each): the owner side of the
// multi-GPU exchange.
__global__ void __launch_bounds__(256) insert_hashed_kernel(const __grid_constant__ TableGeom g,
                                                            const uint64_t *__restrict__ hs, uint64_t n) {
    unsigned long long my_new = 0, my_total = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * kUnroll;
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x * kUnroll; base < n; base += stride) {
        uint64_t h[kUnroll];
        bool ok[kUnroll];
        Bucket bk[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            const uint64_t i = base + (uint64_t)u * blockDim.x + threadIdx.x;
            ok[u] = i < n;
            h[u] = ok[u] ? __ldg(hs + i) : 0;
            if (ok[u]) bk[u] = load_bucket(g.slots + (((h[u] >> g.rbits) & g.bmask) << 2));
        }
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            if (ok[u]) {
                my_new += insert_home(g, h[u], bk[u]) == 1;
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

// Insert (hashed key, count) pairs: spill replay after a grow.
__global__ void __launch_bounds__(256) insert_pairs_kernel(const __grid_constant__ TableGeom g,
                                                           const uint64_t *__restrict__ pairs, uint64_t n) {
    unsigned long long my_new = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        Bucket bk = {};
        my_new += insert_hashed(g, pairs[2 * i], pairs[2 * i + 1], bk, false) == 1;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) my_new += __shfl_down_sync(0xffffffffu, my_new, d);
    if ((threadIdx.x & 31) == 0 && my_new) atomicAdd(&g.ctr->n_distinct, my_new);
}

// Add (k-mer, count) records: the LOAD path (HashLoader::loadHash, lib/src/jellyfish_helper.cc:176-178).
__global__ void __launch_bounds__(256) add_pairs_kernel(const __grid_constant__ TableGeom g, const uint64_t *__restrict__ keys,
                                                        const uint64_t *__restrict__ counts, uint64_t n) {
    unsigned long long my_new = 0, my_total = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t c = counts[i];
        if (c == 0) continue; // add(key, 0) of the reference creates the key with count 0, which no walk reports
        Bucket bk = {};
        my_new += insert_hashed(g, mix_key(g, keys[i] & g.kmask), c, bk, false) == 1;
        my_total += c;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        my_new += __shfl_down_sync(0xffffffffu, my_new, d);
        my_total += __shfl_down_sync(0xffffffffu, my_total, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (my_new) atomicAdd(&g.ctr->n_distinct, my_new);
        if (my_total) atomicAdd(&g.ctr->n_total, my_total);
    }
}

// Table doubling (the role of hash_counter::double_size, JFI/hash_counter.hpp:204-244):
// walk the old table, re-insert every (hashed key, exact count) into the new one.
__global__ void __launch_bounds__(256) rehash_kernel(const __grid_constant__ TableGeom old_g,
                                                     const __grid_constant__ TableGeom new_g, uint64_t n_slots) {
    unsigned long long my_new = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t w = old_g.slots[i];
        if (w == 0) continue;
        uint64_t h, cnt;
        decode_slot(old_g, i, w, h, cnt);
        Bucket bk = {};
        my_new += insert_hashed(new_g, h, cnt, bk, false) == 1;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) my_new += __shfl_down_sync(0xffffffffu, my_new, d);
    if ((threadIdx.x & 31) == 0 && my_new) atomicAdd(&new_g.ctr->n_distinct, my_new);
}

// ---- launchers ----
size_t extract_smem_bytes(int mode) {
    return sizeof(ExtractSmem) + (mode == kModeRoute ? (size_t)kThreads * 16 * sizeof(uint64_t) : 0);
}

cudaError_t extract_configure(int mode, int *blocks_per_sm) {
    const size_t smem = extract_smem_bytes(mode);
    cudaError_t e;
    if (mode == kModeInsert) {
        e = cudaFuncSetAttribute(extract_kernel<kModeInsert>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, extract_kernel<kModeInsert>, kThreads, smem);
    }
    e = cudaFuncSetAttribute(extract_kernel<kModeRoute>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, extract_kernel<kModeRoute>, kThreads, smem);
}

void launch_extract(int mode, const ExtractParams &P, int grid, cudaStream_t st) {
    const size_t smem = extract_smem_bytes(mode);
    if (mode == kModeInsert) extract_kernel<kModeInsert><<<grid, kThreads, smem, st>>>(P);
    else extract_kernel<kModeRoute><<<grid, kThreads, smem, st>>>(P);
}
void launch_insert_hashed(const TableGeom &g, const uint64_t *hs, uint64_t n, int grid, cudaStream_t st) {
    insert_hashed_kernel<<<grid, 256, 0, st>>>(g, hs, n);
}
void launch_insert_pairs(const TableGeom &g, const uint64_t *pairs, uint64_t n, int grid, cudaStream_t st) {
    insert_pairs_kernel<<<grid, 256, 0, st>>>(g, pairs, n);
}
void launch_add_pairs(const TableGeom &g, const uint64_t *keys, const uint64_t *counts, uint64_t n, int grid, cudaStream_t st) {
    add_pairs_kernel<<<grid, 256, 0, st>>>(g, keys, counts, n);
}
void launch_rehash(const TableGeom &old_g, const TableGeom &new_g, uint64_t n_slots, int grid, cudaStream_t st) {
    rehash_kernel<<<grid, 256, 0, st>>>(old_g, new_g, n_slots);
}

} // namespace katgpu
