// wide_kernels.cu — the counting and walk kernels for k-mers of 33..64 bases (two-word keys, wide.cuh).  sm_100a only.
//
// Same roles as the one-word kernels: wide_extract_kernel = HOT LOOP A (lib/src/jellyfish_helper.cc:202-211,
// JFI/mer_iterator.hpp:61-89, JFI/hash_counter.hpp:98-130) — a CTA stages a tile of the base stream in shared
// memory with one TMA bulk copy, every thread rolls a 128-bit k-mer and its reverse complement along its
// stretch of the tile and adds each canonical k-mer to the table (16-byte CAS to claim, RED to count);
// wide_hist / wide_gcp / wide_comp = HOT LOOPS B / C / D (src/histogram.cc:183-199, src/gcp.cc:179-197,
// src/comp.cc:387-484) with the accumulator scheme of reduce_kernels.cu; lookup / export / rehash / add_pairs.
#include "wide_kernels.cuh"
#include "extract.cuh"
#include "walk_common.cuh"

#include <algorithm>

namespace katgpu {

// ------------------------------------------------------------------ counting --
struct __align__(128) WideSmem {
    uint8_t ascii[kWideTile + kWideHalo];
    uint64_t mbar;
    int next_tile;
    unsigned long long s_new, s_tot;
};

__device__ __forceinline__ bool wide_tile_uses_tma(const WideExtractParams &P, int t) {
    return P.tma_ok && ((uint64_t)(t + 1) * kWideTile + kWideHalo <= P.n_valid);
}
__device__ __forceinline__ void wide_issue_tma(const WideExtractParams &P, WideSmem &S, int t) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&S.mbar, kWideTile + kWideHalo);
    tma_load_1d(S.ascii, P.bases + (uint64_t)t * kWideTile, kWideTile + kWideHalo, &S.mbar);
}

__global__ void __launch_bounds__(kWideThreads, 4) wide_extract_kernel(const __grid_constant__ WideExtractParams P) {
    __shared__ WideSmem S;
    const WideGeom &g = P.g;
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(&S.mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        S.s_new = S.s_tot = 0;
        const int t = (int)atomicAdd(P.tile_counter, 1u);
        S.next_tile = t;
        if (t < P.n_tiles && wide_tile_uses_tma(P, t)) wide_issue_tma(P, S, t);
    }
    __syncthreads();
    int tile = S.next_tile;
    uint32_t parity = 0;
    unsigned long long my_new = 0, my_tot = 0;
    while (tile < P.n_tiles) {
        const uint64_t tile_base = (uint64_t)tile * kWideTile;
        if (wide_tile_uses_tma(P, tile)) {
            while (!mbar_try_wait(&S.mbar, parity)) {
            }
            parity ^= 1;
        } else { // ragged tail / unaligned input: guarded loads, bytes past n_valid read as 0 (= not a base)
            for (int i = tid; i < kWideTile + kWideHalo; i += kWideThreads) {
                const uint64_t p = tile_base + i;
                S.ascii[i] = p < P.n_valid ? P.bases[p] : (uint8_t)0;
            }
            __syncthreads();
        }
        // this thread's stretch: kWidePer start positions (9 words: an odd stride, no bank conflicts) + k-1 bases of tail
        const uint64_t p0 = tile_base + (uint64_t)tid * kWidePer;
        const uint32_t n_here = p0 < P.n_starts ? (uint32_t)min((uint64_t)kWidePer, P.n_starts - p0) : 0u;
        const uint32_t *words = reinterpret_cast<const uint32_t *>(S.ascii) + tid * (kWidePer / 4);
        uint32_t w = 0;
        wide_roll_segment(
            g, n_here,
            [&](uint32_t i) {
                if ((i & 3u) == 0) w = words[i >> 2];
                const uint32_t ch = w & 0xFFu;
                w >>= 8;
                return ch;
            },
            [&](const Key128 &key) {
                my_new += wide_insert(g, key, 1ULL) == 1;
                my_tot++;
            });
        __syncthreads(); // everybody is done with S.ascii
        if (tid == 0) {
            const int t = (int)atomicAdd(P.tile_counter, 1u);
            S.next_tile = t;
            if (t < P.n_tiles && wide_tile_uses_tma(P, t)) wide_issue_tma(P, S, t);
        }
        __syncthreads();
        tile = S.next_tile;
    }
    if (my_new) atomicAdd(&S.s_new, my_new);
    if (my_tot) atomicAdd(&S.s_tot, my_tot);
    __syncthreads();
    if (tid == 0) {
        if (S.s_new) atomicAdd(&g.ctr->claimed, S.s_new);
        if (S.s_tot) atomicAdd(&g.ctr->total, S.s_tot);
    }
}

// (key, count) records -> table: HashLoader::loadHash (lib/src/jellyfish_helper.cc:176-178)
__global__ void __launch_bounds__(256) wide_add_pairs_kernel(const __grid_constant__ WideGeom g, const uint64_t *__restrict__ keys_hi,
                                                             const uint64_t *__restrict__ keys_lo, const uint64_t *__restrict__ counts,
                                                             uint64_t n) {
    unsigned long long my_new = 0, my_tot = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        Key128 key;
        key.lo = keys_lo[i];
        key.hi = keys_hi[i] & g.hi_mask;
        const unsigned long long c = counts[i];
        if (c == 0) continue; // a record with count 0 adds nothing; occupancy is count != 0
        my_new += wide_insert(g, key, c) == 1;
        my_tot += c;
    }
    my_new = warp_sum(my_new);
    my_tot = warp_sum(my_tot);
    if ((threadIdx.x & 31) == 0) {
        if (my_new) atomicAdd(&g.ctr->claimed, my_new);
        if (my_tot) atomicAdd(&g.ctr->total, my_tot);
    }
}

// hash_counter::double_size (JFI/hash_counter.hpp:204-244): every entry of the old table into the new one
__global__ void __launch_bounds__(256) wide_rehash_kernel(const __grid_constant__ WideGeom og, const __grid_constant__ WideGeom ng) {
    unsigned long long my_new = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= og.n_slots; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long c = og.counts[i];
        if (c == 0) continue;
        const Key128 key = og.keys[i]; // the extra slot reads as the all-ones key and lands in the new extra slot
        my_new += wide_insert(ng, key, c) == 1;
    }
    my_new = warp_sum(my_new);
    if ((threadIdx.x & 31) == 0 && my_new) atomicAdd(&ng.ctr->claimed, my_new);
}

// distinct keys and k-mer instances of a finished table: out[0] += #(count != 0), out[1] += sum of counts
__global__ void __launch_bounds__(256) wide_stats_kernel(const __grid_constant__ WideGeom g, unsigned long long *out) {
    unsigned long long dis = 0, tot = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= g.n_slots; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long c = g.counts[i];
        dis += c != 0;
        tot += c;
    }
    dis = warp_sum(dis);
    tot = warp_sum(tot);
    if ((threadIdx.x & 31) == 0) {
        if (dis) atomicAdd(out, dis);
        if (tot) atomicAdd(out + 1, tot);
    }
}

// ------------------------------------------------------------------ K2 hist --
template <bool SMEM>
__global__ void __launch_bounds__(512) wide_hist_kernel(const __grid_constant__ WideGeom g, uint64_t base, uint64_t ceil_, uint64_t inc,
                                                        uint32_t nb, unsigned long long *__restrict__ out) {
    extern __shared__ uint32_t sbins[];
    if (SMEM) {
        for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) sbins[i] = 0;
        __syncthreads();
    }
    const uint64_t n = g.n_slots + 1;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (n + stride - 1) / stride; // whole warps iterate together (match.any needs the full warp)
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint64_t it = 0; it < rounds; it++, i += stride) {
        const unsigned long long v = i < n ? __ldg(g.counts + i) : 0ULL;
        const bool occ = v != 0;
        uint32_t bin = 0;
        if (occ) {
            if (v < base) bin = 0;
            else if (v > ceil_) bin = nb - 1;
            else bin = (uint32_t)((v - base) / inc);
        }
        if (SMEM) warp_agg_inc_u32(sbins, bin, occ);
        else warp_agg_inc_u64(out, bin, occ);
    }
    if (SMEM) {
        __syncthreads();
        for (uint32_t b = threadIdx.x; b < nb; b += blockDim.x)
            if (sbins[b]) atomicAdd(out + b, (unsigned long long)sbins[b]);
    }
}

// ------------------------------------------------------------------- K3 gcp --
template <bool SMEM>
__global__ void __launch_bounds__(512) wide_gcp_kernel(const __grid_constant__ WideGeom g, double scale, uint32_t cvg_bins,
                                                       unsigned long long *__restrict__ out) {
    extern __shared__ uint32_t scell[];
    const uint32_t cols = cvg_bins + 1;
    const uint32_t cells = g.k * cols;
    const bool unit = scale == 1.0;
    if (SMEM) {
        for (uint32_t i = threadIdx.x; i < cells; i += blockDim.x) scell[i] = 0;
        __syncthreads();
    }
    const uint64_t n = g.n_slots + 1;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (n + stride - 1) / stride;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint64_t it = 0; it < rounds; it++, i += stride) {
        const unsigned long long cnt = i < n ? __ldg(g.counts + i) : 0ULL;
        bool act = cnt != 0;
        uint32_t cell = 0;
        if (act) {
            const Key128 key = g.keys[i];
            const uint32_t gc = wide_gc(key);
            uint64_t cvg = scale_count(cnt, scale, unit);
            if (cvg > cvg_bins) cvg = cvg_bins;
            act = gc < g.k; // rows 0..k-1 only: all-G/C k-mers are dropped like the reference
            cell = gc * cols + (uint32_t)cvg;
        }
        if (SMEM) warp_agg_inc_u32(scell, cell, act);
        else warp_agg_inc_u64(out, cell, act);
    }
    if (SMEM) {
        __syncthreads();
        for (uint32_t c = threadIdx.x; c < cells; c += blockDim.x)
            if (scell[c]) atomicAdd(out + c, (unsigned long long)scell[c]);
    }
}

// ------------------------------------------------------------------ K4 comp --
// The probe key: the stored key of the walked table, canonicalised when asked and not canonical already
// (JellyfishHelper::getCount, lib/src/jellyfish_helper.cc:189-194; src/comp.cc:400,447).
__device__ __forceinline__ unsigned long long wide_probe_other(const WideGeom &ga, const WideGeom &gb, const Key128 &key, bool canon_probe) {
    return wide_lookup(gb, (canon_probe && !ga.canonical) ? wide_canonical(ga, key) : key);
}

__global__ void __launch_bounds__(512) wide_comp_kernel(const __grid_constant__ WideCompParams P) {
    extern __shared__ uint32_t sm[];
    uint32_t *stile = sm;                       // [kCompSub][kCompSub] corner of main
    uint32_t *sspec = sm + kCompSub * kCompSub; // up to 3 spectra of P.len (if P.spec_smem)
    const WideGeom &ga = P.ga;
    for (uint32_t i = threadIdx.x; i < kCompSub * kCompSub; i += blockDim.x) stile[i] = 0;
    if (P.spec_smem)
        for (uint32_t i = threadIdx.x; i < 3 * P.len; i += blockDim.x) sspec[i] = 0;
    __syncthreads();
    const bool unit1 = P.d1_scale == 1.0, unit2 = P.d2_scale == 1.0;
    unsigned long long tot = 0, dis = 0, only_tot = 0, only_dis = 0, sh1 = 0, sh2 = 0, shd = 0;
    const uint64_t n = ga.n_slots + 1;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (n + stride - 1) / stride;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint64_t it = 0; it < rounds; it++, i += stride) {
        const uint64_t ca = i < n ? __ldg(ga.counts + i) : 0ULL;
        const bool occ = ca != 0;
        uint64_t cb = 0, cc = 0;
        if (occ) {
            const Key128 key = ga.keys[i];
            if (P.pass != 3) cb = wide_probe_other(ga, P.gb, key, P.canon_b);
            if (P.pass == 1 && P.have_c) cc = wide_probe_other(ga, P.gc, key, P.canon_c);
            tot += ca;
            dis++;
        }
        if (P.pass == 3) continue;
        // spectra of the walked table (+ shared spectra in pass 1)
        const uint32_t sa = spectrum_bin(ca, P.len), sb = spectrum_bin(cb, P.len);
        const bool shared = occ && cb != 0;
        if (P.spec_smem) {
            warp_agg_inc_u32(sspec, sa, occ);
            if (P.pass == 1) {
                warp_agg_inc_u32(sspec + P.len, sa, shared);
                warp_agg_inc_u32(sspec + 2 * P.len, sb, shared);
            }
        } else {
            warp_agg_inc_u64(P.spec_walk, sa, occ);
            if (P.pass == 1) {
                warp_agg_inc_u64(P.spec_shared1, sa, shared);
                warp_agg_inc_u64(P.spec_shared2, sb, shared);
            }
        }
        if (occ && cb == 0) {
            only_tot += ca;
            only_dis++;
        }
        if (P.pass == 1 && shared) {
            sh1 += ca;
            sh2 += cb;
            shd++;
        }
        // matrix cell
        uint64_t s1, s2;
        bool act;
        if (P.pass == 1) {
            s1 = scale_count(ca, P.d1_scale, unit1);
            s2 = scale_count(cb, P.d2_scale, unit2);
            act = occ;
        } else {
            s1 = 0;
            s2 = scale_count(ca, P.d2_scale, unit2);
            act = occ && cb == 0; // only k-mers absent from hash1 (src/comp.cc:452-462)
        }
        if (s1 >= P.d1_bins) s1 = P.d1_bins - 1;
        if (s2 >= P.d2_bins) s2 = P.d2_bins - 1;
        const bool in_tile = s1 < kCompSub && s2 < kCompSub;
        warp_agg_inc_u32(stile, (uint32_t)(s1 * kCompSub + s2), act && in_tile);
        warp_agg_inc_u64(P.main_mx, s1 * P.d2_bins + s2, act && !in_tile);
        if (P.pass == 1 && P.have_c) { // src/comp.cc:426-433
            uint64_t s3 = scale_count(cc, P.d2_scale, unit2);
            if (s3 >= P.d2_bins) s3 = P.d2_bins - 1;
            const uint64_t cell = s1 * P.d2_bins + s3;
            warp_agg_inc_u64(P.ends_mx, cell, occ && s2 == s3);
            warp_agg_inc_u64(P.mixed_mx, cell, occ && s2 != s3 && s3 > 0);
            warp_agg_inc_u64(P.middle_mx, cell, occ && s2 != s3 && s3 == 0);
        }
    }
    __syncthreads();
    if (P.pass != 3) {
        for (uint32_t c = threadIdx.x; c < kCompSub * kCompSub; c += blockDim.x) {
            const uint32_t v = stile[c];
            const uint32_t r = c / kCompSub, q = c % kCompSub;
            if (v && r < P.d1_bins && q < P.d2_bins) atomicAdd(P.main_mx + (uint64_t)r * P.d2_bins + q, (unsigned long long)v);
        }
        if (P.spec_smem)
            for (uint32_t c = threadIdx.x; c < P.len; c += blockDim.x) {
                if (sspec[c]) atomicAdd(P.spec_walk + c, (unsigned long long)sspec[c]);
                if (P.pass == 1) {
                    if (sspec[P.len + c]) atomicAdd(P.spec_shared1 + c, (unsigned long long)sspec[P.len + c]);
                    if (sspec[2 * P.len + c]) atomicAdd(P.spec_shared2 + c, (unsigned long long)sspec[2 * P.len + c]);
                }
            }
    }
    tot = warp_sum(tot); dis = warp_sum(dis); only_tot = warp_sum(only_tot); only_dis = warp_sum(only_dis);
    sh1 = warp_sum(sh1); sh2 = warp_sum(sh2); shd = warp_sum(shd);
    if ((threadIdx.x & 31) == 0) {
        unsigned long long *c = P.counters; // order of katgpu_comp_counters (include/katgpu.h)
        if (P.pass == 1) {
            atomicAdd(c + 0, tot); atomicAdd(c + 3, dis); atomicAdd(c + 6, only_tot); atomicAdd(c + 8, only_dis);
            atomicAdd(c + 10, sh1); atomicAdd(c + 11, sh2); atomicAdd(c + 12, shd);
        } else if (P.pass == 2) {
            atomicAdd(c + 1, tot); atomicAdd(c + 4, dis); atomicAdd(c + 7, only_tot); atomicAdd(c + 9, only_dis);
        } else {
            atomicAdd(c + 2, tot); atomicAdd(c + 5, dis);
        }
    }
}

// ---------------------------------------------------------- lookup / export --
__global__ void __launch_bounds__(256) wide_lookup_kernel(const __grid_constant__ WideGeom g, const uint64_t *__restrict__ keys_hi,
                                                          const uint64_t *__restrict__ keys_lo, uint64_t n, int canonicalise,
                                                          uint64_t *__restrict__ out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        Key128 key;
        key.lo = keys_lo[i];
        key.hi = keys_hi[i] & g.hi_mask;
        if (canonicalise) key = wide_canonical(g, key);
        out[i] = wide_lookup(g, key);
    }
}

__global__ void __launch_bounds__(256) wide_export_kernel(const __grid_constant__ WideGeom g, uint64_t *__restrict__ keys_hi,
                                                          uint64_t *__restrict__ keys_lo, uint64_t *__restrict__ counts, uint64_t cap,
                                                          unsigned long long *cursor, const ExportSel sel) {
    const uint64_t n = g.n_slots + 1;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (n + stride - 1) / stride;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint64_t it = 0; it < rounds; it++, i += stride) {
        const unsigned long long c = i < n ? g.counts[i] : 0ULL;
        bool occ = c != 0;
        Key128 key = {0, 0};
        if (occ) {
            key = g.keys[i];
            if (sel.mode) {
                const uint32_t gc = wide_gc(key);
                const bool in_bounds = sel.low_gc <= gc && gc <= sel.high_gc && sel.low_count <= c && c <= sel.high_count;
                occ = in_bounds == (sel.mode > 0);
            }
        }
        const uint32_t m = __ballot_sync(0xffffffffu, occ);
        if (m == 0) continue;
        const int lane = threadIdx.x & 31;
        unsigned long long basei = 0;
        if (lane == 0) basei = atomicAdd(cursor, (unsigned long long)__popc(m));
        basei = __shfl_sync(0xffffffffu, basei, 0);
        if (occ) {
            const uint64_t at = basei + __popc(m & ((1u << lane) - 1));
            if (at < cap) {
                keys_hi[at] = key.hi;
                keys_lo[at] = key.lo;
                counts[at] = c;
            }
        }
    }
}

// ---- launchers ----
void launch_wide_extract(const WideExtractParams &P, int grid, cudaStream_t st) { wide_extract_kernel<<<grid, kWideThreads, 0, st>>>(P); }
void launch_wide_add_pairs(const WideGeom &g, const uint64_t *keys_hi, const uint64_t *keys_lo, const uint64_t *counts, uint64_t n, int grid,
                           cudaStream_t st) {
    wide_add_pairs_kernel<<<grid, 256, 0, st>>>(g, keys_hi, keys_lo, counts, n);
}
void launch_wide_rehash(const WideGeom &og, const WideGeom &ng, int grid, cudaStream_t st) { wide_rehash_kernel<<<grid, 256, 0, st>>>(og, ng); }
void launch_wide_stats(const WideGeom &g, unsigned long long *out, int grid, cudaStream_t st) { wide_stats_kernel<<<grid, 256, 0, st>>>(g, out); }
void launch_wide_hist(const WideGeom &g, uint64_t base, uint64_t ceil_, uint64_t inc, uint32_t nb, unsigned long long *out, int grid,
                      cudaStream_t st) {
    const size_t smem = (size_t)nb * 4;
    if (smem <= kMaxReduceSmem) {
        cudaFuncSetAttribute(wide_hist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxReduceSmem);
        wide_hist_kernel<true><<<grid, 512, smem, st>>>(g, base, ceil_, inc, nb, out);
    } else
        wide_hist_kernel<false><<<grid, 512, 0, st>>>(g, base, ceil_, inc, nb, out);
}
void launch_wide_gcp(const WideGeom &g, double scale, uint32_t cvg_bins, unsigned long long *out, int grid, cudaStream_t st) {
    const size_t smem = (size_t)g.k * (cvg_bins + 1) * 4;
    if (smem <= kMaxReduceSmem) {
        cudaFuncSetAttribute(wide_gcp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxReduceSmem);
        wide_gcp_kernel<true><<<grid, 512, smem, st>>>(g, scale, cvg_bins, out);
    } else
        wide_gcp_kernel<false><<<grid, 512, 0, st>>>(g, scale, cvg_bins, out);
}
void launch_wide_comp(WideCompParams P, int grid, cudaStream_t st) {
    size_t smem = (size_t)kCompSub * kCompSub * 4;
    P.spec_smem = (size_t)3 * P.len * 4 + smem <= kMaxReduceSmem;
    if (P.spec_smem) smem += (size_t)3 * P.len * 4;
    cudaFuncSetAttribute(wide_comp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxReduceSmem);
    wide_comp_kernel<<<grid, 512, smem, st>>>(P);
}
void launch_wide_lookup(const WideGeom &g, const uint64_t *keys_hi, const uint64_t *keys_lo, uint64_t n, int canonicalise, uint64_t *out,
                        int grid, cudaStream_t st) {
    wide_lookup_kernel<<<grid, 256, 0, st>>>(g, keys_hi, keys_lo, n, canonicalise, out);
}
void launch_wide_export(const WideGeom &g, uint64_t *keys_hi, uint64_t *keys_lo, uint64_t *counts, uint64_t cap, unsigned long long *cursor,
                        int grid, cudaStream_t st, const ExportSel &sel) {
    wide_export_kernel<<<grid, 256, 0, st>>>(g, keys_hi, keys_lo, counts, cap, cursor, sel);
}

} // namespace katgpu
