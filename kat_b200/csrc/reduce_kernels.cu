// reduce_kernels.cu — table walks: K2 hist, K3 gcp, K4 comp (dual-hash walk),
// point lookups and export.  sm_100a only.
//
// Replaces the reference's HOT LOOPS B/C/D: Histogram::binSlice
// (src/histogram.cc:183-199), Gcp::analyseSlice (src/gcp.cc:179-197),
// Comp::compareSlice (src/comp.cc:387-484) and their per-thread accumulators +
// merge (sparse_matrix.hpp:324-335, comp_counters.cc:230-254).  The per-thread
// private matrices become per-CTA shared-memory tiles of 32-bit cells, flushed
// once with 64-bit global atomics; increments are warp-aggregated with
// match.any so the hot cells (count 1) cost one shared atomic per warp.
#include "reduce_kernels.cuh"
#include "walk_common.cuh"

#include <algorithm>

namespace katgpu {

// ------------------------------------------------------------------ K2 hist --
template <bool SMEM>
__global__ void __launch_bounds__(512) hist_kernel(const __grid_constant__ TableGeom g, uint64_t n_slots,
                                                   uint64_t base, uint64_t ceil_, uint64_t inc, uint32_t nb,
                                                   unsigned long long *__restrict__ out) {
    extern __shared__ uint32_t sbins[];
    if (SMEM) {
        for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) sbins[i] = 0;
        __syncthreads();
    }
    const uint64_t n_pairs = n_slots >> 1;
    const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(g.slots);
    // whole warps iterate together (match.any needs the full warp); one 16-byte load per thread and
    // round at full occupancy (a 4-deep unrolled variant measured 14 % slower: 5.78 vs 5.08 ms for 17 GB)
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (n_pairs + stride - 1) / stride;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint64_t it = 0; it < rounds; it++, i += stride) {
        ulonglong2 w = make_ulonglong2(0, 0);
        if (i < n_pairs) w = __ldg(src + i);
#pragma unroll
        for (int s = 0; s < 2; s++) {
            const uint64_t word = s ? w.y : w.x;
            const bool occ = word != 0;
            uint32_t bin = 0;
            if (occ) {
                uint64_t v = word >> g.cshift;
                if (word & (1ULL << (g.rbits + kProbeBits))) {
                    uint64_t h, c;
                    decode_slot(g, 2 * i + s, word, h, c);
                    v = c;
                }
                if (v < base) bin = 0;
                else if (v > ceil_) bin = nb - 1;
                else bin = (uint32_t)(inc == 1 ? (v - base) : (v - base) / inc);
            }
            if (SMEM) warp_agg_inc_u32(sbins, bin, occ);
            else warp_agg_inc_u64(out, bin, occ);
        }
    }
    if (SMEM) {
        __syncthreads();
        for (uint32_t b = threadIdx.x; b < nb; b += blockDim.x)
            if (sbins[b]) atomicAdd(out + b, (unsigned long long)sbins[b]);
    }
}

// ------------------------------------------------------------------- K3 gcp --
template <bool SMEM>
__global__ void __launch_bounds__(512, 2) gcp_kernel(const __grid_constant__ TableGeom g, uint64_t n_slots,
                                                  double scale, uint32_t cvg_bins,
                                                  unsigned long long *__restrict__ out) {
    extern __shared__ uint32_t scell[];
    const uint32_t cols = cvg_bins + 1;
    const uint32_t cells = g.k * cols;
    const bool unit = scale == 1.0;
    if (SMEM) {
        for (uint32_t i = threadIdx.x; i < cells; i += blockDim.x) scell[i] = 0;
        __syncthreads();
    }
    const uint64_t n_pairs = n_slots >> 1;
    const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(g.slots);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (n_pairs + stride * kWalkUnroll - 1) / (stride * kWalkUnroll);
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint64_t it = 0; it < rounds; it++, i += stride * kWalkUnroll) {
        ulonglong2 w[kWalkUnroll];
#pragma unroll
        for (int u = 0; u < kWalkUnroll; u++) {
            const uint64_t j = i + (uint64_t)u * stride;
            w[u] = j < n_pairs ? __ldcs(src + j) : make_ulonglong2(0, 0);
        }
#pragma unroll
        for (int u = 0; u < kWalkUnroll; u++) {
#pragma unroll
            for (int s = 0; s < 2; s++) {
                const uint64_t word = s ? w[u].y : w[u].x;
                bool act = word != 0;
                uint32_t cell = 0;
                if (act) {
                    uint64_t h, cnt;
                    decode_slot(g, 2 * (i + (uint64_t)u * stride) + s, word, h, cnt);
                    const uint64_t key = unmix_key(g, h);
                    // gcCount (lib/include/kat/str_utils.hpp:151-161): C = 01, G = 10 -> the two bits differ
                    const uint32_t gc = (uint32_t)__popcll((key ^ (key >> 1)) & 0x5555555555555555ULL);
                    uint64_t cvg = scale_count(cnt, scale, unit);
                    if (cvg > cvg_bins) cvg = cvg_bins;
                    act = gc < g.k; // rows 0..k-1 only: all-G/C k-mers are dropped like the reference
                    cell = gc * cols + (uint32_t)cvg;
                }
                if (SMEM) warp_agg_inc_u32(scell, cell, act);
                else warp_agg_inc_u64(out, cell, act);
            }
        }
    }
    if (SMEM) {
        __syncthreads();
        for (uint32_t c = threadIdx.x; c < cells; c += blockDim.x)
            if (scell[c]) atomicAdd(out + c, (unsigned long long)scell[c]);
    }
}

// ------------------------------------------------------------------ K4 comp --
// pass 1: walk A (hash1), probe B (hash2) and optionally C (hash3).
// pass 2: walk A (hash2), probe B (hash1): only hash2 counters + main[0][s2].
// pass 3: walk A (hash3): totals only.
__device__ __forceinline__ uint64_t probe_other(const TableGeom &ga, const TableGeom &gb, uint64_t h,
                                                bool canon_probe) {
    if (canon_probe && !ga.canonical) { // stored key may be the non-canonical strand
        uint64_t key = unmix_key(ga, h);
        const uint64_t rc = revcomp_kmer(key, ga.k);
        if (rc < key) key = rc;
        h = mix_key(gb, key);
    }
    return lookup_hashed(gb, h);
}
__global__ void __launch_bounds__(512) comp_kernel(const __grid_constant__ CompParams P) {
    extern __shared__ uint32_t sm[];
    uint32_t *stile = sm;                             // [kCompSub][kCompSub] corner of main
    uint32_t *sspec = sm + kCompSub * kCompSub;       // up to 3 spectra of P.len (if P.spec_smem)
    const TableGeom &ga = P.ga;
    for (uint32_t i = threadIdx.x; i < kCompSub * kCompSub; i += blockDim.x) stile[i] = 0;
    if (P.spec_smem)
        for (uint32_t i = threadIdx.x; i < 3 * P.len; i += blockDim.x) sspec[i] = 0;
    __syncthreads();
    const bool unit1 = P.d1_scale == 1.0, unit2 = P.d2_scale == 1.0;
    unsigned long long tot = 0, dis = 0, only_tot = 0, only_dis = 0, sh1 = 0, sh2 = 0, shd = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (P.n_slots + stride - 1) / stride;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint64_t it = 0; it < rounds; it++, i += stride) {
        const uint64_t word = i < P.n_slots ? __ldg(ga.slots + i) : 0;
        const bool occ = word != 0;
        uint64_t ca = 0, cb = 0, cc = 0;
        if (occ) {
            uint64_t h;
            decode_slot(ga, i, word, h, ca);
            if (P.pass != 3) cb = probe_other(ga, P.gb, h, P.canon_b);
            if (P.pass == 1 && P.have_c) cc = probe_other(ga, P.gc, h, P.canon_c);
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
        unsigned long long *c = P.counters;
        // order of katgpu_comp_counters (include/katgpu.h)
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
__global__ void __launch_bounds__(256) lookup_kernel(const __grid_constant__ TableGeom g,
                                                     const uint64_t *__restrict__ keys, uint64_t n,
                                                     int canonicalise, uint64_t *__restrict__ out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t key = keys[i] & g.kmask;
        if (canonicalise) {
            const uint64_t rc = revcomp_kmer(key, g.k);
            if (rc < key) key = rc;
        }
        out[i] = lookup_hashed(g, mix_key(g, key));
    }
}

// sel.mode 0: every (key, count); +1: only k-mers whose count and GC count lie inside the bounds
// (FilterKmer::inBounds, src/filter_kmer.cc:296-309); -1: only those outside.
__global__ void __launch_bounds__(256) export_kernel(const __grid_constant__ TableGeom g, uint64_t n_slots,
                                                     uint64_t *__restrict__ keys, uint64_t *__restrict__ counts,
                                                     uint64_t cap, unsigned long long *cursor, const ExportSel sel) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (n_slots + stride - 1) / stride;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint64_t it = 0; it < rounds; it++, i += stride) {
        const uint64_t word = i < n_slots ? g.slots[i] : 0;
        bool occ = word != 0;
        uint64_t key = 0, c = 0;
        if (occ) {
            uint64_t h;
            decode_slot(g, i, word, h, c);
            key = unmix_key(g, h);
            if (sel.mode) {
                const uint32_t gc = (uint32_t)__popcll((key ^ (key >> 1)) & 0x5555555555555555ULL);
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
                keys[at] = key;
                counts[at] = c;
            }
        }
    }
}

// ---- launchers ----
void launch_hist(const TableGeom &g, uint64_t n_slots, uint64_t base, uint64_t ceil_, uint64_t inc, uint32_t nb,
                 unsigned long long *out, int grid, cudaStream_t st) {
    const size_t smem = (size_t)nb * 4;
    if (smem <= kMaxReduceSmem) {
        cudaFuncSetAttribute(hist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxReduceSmem);
        hist_kernel<true><<<grid, 512, smem, st>>>(g, n_slots, base, ceil_, inc, nb, out);
    } else
        hist_kernel<false><<<grid, 512, 0, st>>>(g, n_slots, base, ceil_, inc, nb, out);
}
void launch_gcp(const TableGeom &g, uint64_t n_slots, double scale, uint32_t cvg_bins, unsigned long long *out,
                int grid, cudaStream_t st) {
    const size_t smem = (size_t)g.k * (cvg_bins + 1) * 4;
    if (smem <= kMaxReduceSmem) {
        cudaFuncSetAttribute(gcp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxReduceSmem);
        // `grid` = SM count: as many CTAs per SM as their matrices fit in shared memory (at most 4)
        const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, ((size_t)220 << 10) / (smem + 1024)));
        gcp_kernel<true><<<grid * per_sm, 512, smem, st>>>(g, n_slots, scale, cvg_bins, out);
    } else
        gcp_kernel<false><<<grid, 512, 0, st>>>(g, n_slots, scale, cvg_bins, out);
}
void launch_comp(CompParams P, int grid, cudaStream_t st) {
    size_t smem = (size_t)kCompSub * kCompSub * 4;
    P.spec_smem = (size_t)3 * P.len * 4 + smem <= kMaxReduceSmem;
    if (P.spec_smem) smem += (size_t)3 * P.len * 4;
    cudaFuncSetAttribute(comp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxReduceSmem);
    comp_kernel<<<grid, 512, smem, st>>>(P);
}
void launch_lookup(const TableGeom &g, const uint64_t *keys, uint64_t n, int canonicalise, uint64_t *out, int grid,
                   cudaStream_t st) {
    lookup_kernel<<<grid, 256, 0, st>>>(g, keys, n, canonicalise, out);
}
void launch_export(const TableGeom &g, uint64_t n_slots, uint64_t *keys, uint64_t *counts, uint64_t cap,
                   unsigned long long *cursor, int grid, cudaStream_t st, const ExportSel &sel) {
    export_kernel<<<grid, 256, 0, st>>>(g, n_slots, keys, counts, cap, cursor, sel);
}

} // namespace katgpu
