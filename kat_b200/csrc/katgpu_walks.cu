// katgpu_walks.cu — the consumer side of the C-ABI: katgpu_hist / katgpu_gcp / katgpu_comp / katgpu_lookup /
// katgpu_table_export over a finished table (reduce_kernels.cu); on a multi-GPU context the local walk is
// followed by one all-reduce of the output (katgpu_dist.cu).
#include "katgpu_internal.cuh"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

using namespace katgpu;

// ---- table walks ----
static katgpu_status finished(katgpu_table *t, const char *who) {
    if (!t) return fail(KATGPU_ERR_INVALID, "%s: NULL table", who);
    if (!t->ended) return fail(KATGPU_ERR_STATE, "%s: call katgpu_count_end first", who);
    return KATGPU_OK;
}

struct DevBuf { // zeroed device scratch, freed on scope exit
    void *p = nullptr;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
};

extern "C" katgpu_status katgpu_hist(katgpu_table *t, uint64_t base, uint64_t ceil_, uint64_t inc, uint64_t nb,
                                     uint64_t *bins_out) {
    TraceScope trace__("katgpu_hist");
    KT(finished(t, "katgpu_hist"));
    if (!bins_out || nb == 0 || inc == 0 || nb > 0x7fffffffULL) return fail(KATGPU_ERR_INVALID, "katgpu_hist: bad arguments");
    katgpu_ctx *c = t->ctx;
    CU(cudaSetDevice(c->device));
    if (t->wide) return wide_hist(t, base, ceil_, inc, nb, bins_out);
    DevBuf d;
    CU(cudaMalloc(&d.p, nb * 8));
    CU(cudaMemsetAsync(d.p, 0, nb * 8, c->stream));
    {
        ProfScope ps(c, "hist", c->stream);
        launch_hist(t->g, t->n_slots, base, ceil_, inc, (uint32_t)nb, (unsigned long long *)d.p, c->sm_count * 4, c->stream);
    }
    c->launches++;
    CU(cudaGetLastError());
    if (c->dist) KT(dist_allreduce_u64(c, d.p, nb)); // ranks own disjoint k-mers: bins add up
    CU(cudaMemcpyAsync(bins_out, d.p, nb * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return KATGPU_OK;
}

extern "C" katgpu_status katgpu_gcp(katgpu_table *t, double cvg_scale, uint32_t cvg_bins, uint64_t *mx_out) {
    KT(finished(t, "katgpu_gcp"));
    if (!mx_out) return fail(KATGPU_ERR_INVALID, "katgpu_gcp: NULL output");
    katgpu_ctx *c = t->ctx;
    CU(cudaSetDevice(c->device));
    if (t->wide) return wide_gcp(t, cvg_scale, cvg_bins, mx_out);
    const size_t cells = (size_t)t->g.k * ((size_t)cvg_bins + 1);
    DevBuf d;
    CU(cudaMalloc(&d.p, cells * 8));
    CU(cudaMemsetAsync(d.p, 0, cells * 8, c->stream));
    {
        ProfScope ps(c, "gcp", c->stream);
        launch_gcp(t->g, t->n_slots, cvg_scale, cvg_bins, (unsigned long long *)d.p, c->sm_count, c->stream);
    }
    c->launches++;
    CU(cudaGetLastError());
    if (c->dist) KT(dist_allreduce_u64(c, d.p, cells));
    CU(cudaMemcpyAsync(mx_out, d.p, cells * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return KATGPU_OK;
}

extern "C" katgpu_status katgpu_comp(katgpu_table *t1, katgpu_table *t2, katgpu_table *t3, double d1_scale,
                                     double d2_scale, uint32_t d1_bins, uint32_t d2_bins, uint64_t *main_mx,
                                     uint64_t *ends_mx, uint64_t *middle_mx, uint64_t *mixed_mx,
                                     katgpu_comp_counters *counters, uint64_t *spectra) {
    KT(finished(t1, "katgpu_comp"));
    KT(finished(t2, "katgpu_comp"));
    if (t3) KT(finished(t3, "katgpu_comp"));
    if (!main_mx || !counters || !spectra || d1_bins == 0 || d2_bins == 0) return fail(KATGPU_ERR_INVALID, "katgpu_comp: bad arguments");
    if (t3 && (!ends_mx || !middle_mx || !mixed_mx)) return fail(KATGPU_ERR_INVALID, "katgpu_comp: 3-input mode needs ends/middle/mixed outputs");
    if (t1->g.k != t2->g.k || (t3 && t3->g.k != t1->g.k)) return fail(KATGPU_ERR_INVALID, "katgpu_comp: tables were counted with different k");
    if (t1->ctx != t2->ctx || (t3 && t3->ctx != t1->ctx)) return fail(KATGPU_ERR_INVALID, "katgpu_comp: tables live on different contexts");
    katgpu_ctx *c = t1->ctx;
    // Multi-GPU: the dual-hash walk is local because both inputs place a k-mer on the same owner rank.
    // That holds for the stored k-mer; a probe that re-canonicalises a non-canonical stored k-mer
    // (src/comp.cc:400,447) may belong to another rank.
    if (c->dist && !(t1->g.canonical && t2->g.canonical && (!t3 || t3->g.canonical)))
        return fail(KATGPU_ERR_INVALID, "katgpu_comp: on a multi-GPU context every input must be counted canonically");
    CU(cudaSetDevice(c->device));
    if (t1->wide) // same k everywhere, so every table is wide
        return wide_comp(t1, t2, t3, d1_scale, d2_scale, d1_bins, d2_bins, main_mx, ends_mx, middle_mx, mixed_mx, counters, spectra);
    const uint32_t len = std::min(d1_bins, d2_bins);
    const size_t cells = (size_t)d1_bins * d2_bins;
    const size_t n_mx = t3 ? 4 : 1;
    const size_t words = n_mx * cells + 16 + 4 * (size_t)len;
    DevBuf d;
    CU(cudaMalloc(&d.p, words * 8));
    CU(cudaMemsetAsync(d.p, 0, words * 8, c->stream));
    unsigned long long *base = (unsigned long long *)d.p;
    CompParams P{};
    P.d1_scale = d1_scale;
    P.d2_scale = d2_scale;
    P.d1_bins = d1_bins;
    P.d2_bins = d2_bins;
    P.len = len;
    P.main_mx = base;
    P.ends_mx = t3 ? base + cells : nullptr;
    P.middle_mx = t3 ? base + 2 * cells : nullptr;
    P.mixed_mx = t3 ? base + 3 * cells : nullptr;
    P.counters = base + n_mx * cells;
    unsigned long long *spec = P.counters + 16;
    const int grid = c->sm_count * 2;
    {
        ProfScope ps(c, "comp", c->stream);
        // pass 1: walk hash1, probe hash2 (canonicalised iff input 2 is canonical, src/comp.cc:400) and hash3
        P.pass = 1;
        P.ga = t1->g;
        P.n_slots = t1->n_slots;
        P.gb = t2->g;
        P.canon_b = t2->g.canonical;
        P.have_c = t3 != nullptr;
        if (t3) {
            P.gc = t3->g;
            P.canon_c = t3->g.canonical;
        }
        P.spec_walk = spec;
        P.spec_shared1 = spec + 2 * (size_t)len;
        P.spec_shared2 = spec + 3 * (size_t)len;
        launch_comp(P, grid, c->stream);
        // pass 2: walk hash2, probe hash1 — always canonicalised (src/comp.cc:447)
        P.pass = 2;
        P.ga = t2->g;
        P.n_slots = t2->n_slots;
        P.gb = t1->g;
        P.canon_b = 1;
        P.have_c = 0;
        P.spec_walk = spec + len;
        launch_comp(P, grid, c->stream);
        c->launches += 2;
        if (t3) { // pass 3: hash3 totals (src/comp.cc:466-479)
            P.pass = 3;
            P.ga = t3->g;
            P.n_slots = t3->n_slots;
            launch_comp(P, grid, c->stream);
            c->launches++;
        }
    }
    CU(cudaGetLastError());
    if (c->dist) KT(dist_allreduce_u64(c, d.p, words)); // matrices, counters and spectra in one all-reduce
    CU(cudaMemcpyAsync(main_mx, P.main_mx, cells * 8, cudaMemcpyDeviceToHost, c->stream));
    if (t3) {
        CU(cudaMemcpyAsync(ends_mx, P.ends_mx, cells * 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(middle_mx, P.middle_mx, cells * 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(mixed_mx, P.mixed_mx, cells * 8, cudaMemcpyDeviceToHost, c->stream));
    }
    static_assert(sizeof(katgpu_comp_counters) == 13 * 8, "counter block layout");
    CU(cudaMemcpyAsync(counters, P.counters, sizeof(katgpu_comp_counters), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(spectra, spec, 4 * (size_t)len * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return KATGPU_OK;
}

extern "C" katgpu_status katgpu_lookup(katgpu_table *t, const uint64_t *keys, size_t n, int canonicalise,
                                       uint64_t *counts_out) {
    KT(finished(t, "katgpu_lookup"));
    if ((!keys || !counts_out) && n) return fail(KATGPU_ERR_INVALID, "katgpu_lookup: NULL argument");
    if (t->wide) return fail(KATGPU_ERR_INVALID, "katgpu_lookup: k > 32 needs two words per key: katgpu_lookup_wide");
    if (n == 0) return KATGPU_OK;
    katgpu_ctx *c = t->ctx;
    CU(cudaSetDevice(c->device));
    DevBuf dk, dc;
    CU(cudaMalloc(&dk.p, n * 8));
    CU(cudaMalloc(&dc.p, n * 8));
    CU(cudaMemcpyAsync(dk.p, keys, n * 8, cudaMemcpyHostToDevice, c->stream));
    const int grid = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)c->sm_count * 8);
    launch_lookup(t->g, (const uint64_t *)dk.p, n, canonicalise, (uint64_t *)dc.p, grid, c->stream);
    c->launches++;
    CU(cudaGetLastError());
    if (c->dist) KT(dist_allreduce_u64(c, dc.p, n)); // the owner rank answers, everybody else holds 0
    CU(cudaMemcpyAsync(counts_out, dc.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return KATGPU_OK;
}

// Two words per key: every k.  keys_hi may be NULL for a one-word table (k <= 32).
extern "C" katgpu_status katgpu_lookup_wide(katgpu_table *t, const uint64_t *keys_hi, const uint64_t *keys_lo, size_t n, int canonicalise,
                                            uint64_t *counts_out) {
    KT(finished(t, "katgpu_lookup_wide"));
    if (!t->wide) return katgpu_lookup(t, keys_lo, n, canonicalise, counts_out);
    if ((!keys_hi || !keys_lo || !counts_out) && n) return fail(KATGPU_ERR_INVALID, "katgpu_lookup_wide: NULL argument");
    if (n == 0) return KATGPU_OK;
    CU(cudaSetDevice(t->ctx->device));
    return wide_lookup_host(t, keys_hi, keys_lo, n, canonicalise, counts_out);
}

static katgpu_status export_selected(katgpu_table *t, const ExportSel &sel, uint64_t *keys_out, uint64_t *counts_out, size_t cap,
                                      uint64_t *n_out, const char *who) {
    KT(finished(t, who));
    if (!n_out || ((!keys_out || !counts_out) && cap)) return fail(KATGPU_ERR_INVALID, "%s: NULL argument", who);
    if (t->wide) return fail(KATGPU_ERR_INVALID, "%s: k > 32 needs two words per key: katgpu_table_export_wide", who);
    katgpu_ctx *c = t->ctx;
    CU(cudaSetDevice(c->device));
    DevBuf dk, dc, dn;
    CU(cudaMalloc(&dk.p, std::max<size_t>(cap, 1) * 8));
    CU(cudaMalloc(&dc.p, std::max<size_t>(cap, 1) * 8));
    CU(cudaMalloc(&dn.p, 8));
    CU(cudaMemsetAsync(dn.p, 0, 8, c->stream));
    launch_export(t->g, t->n_slots, (uint64_t *)dk.p, (uint64_t *)dc.p, cap, (unsigned long long *)dn.p, c->sm_count * 8, c->stream, sel);
    c->launches++;
    CU(cudaGetLastError());
    unsigned long long n = 0;
    CU(cudaMemcpyAsync(&n, dn.p, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *n_out = n;
    const size_t m = std::min<size_t>(cap, n);
    if (m) {
        CU(cudaMemcpy(keys_out, dk.p, m * 8, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(counts_out, dc.p, m * 8, cudaMemcpyDeviceToHost));
    }
    return KATGPU_OK;
}

extern "C" katgpu_status katgpu_table_export(katgpu_table *t, uint64_t *keys_out, uint64_t *counts_out, size_t cap,
                                             uint64_t *n_out) {
    return export_selected(t, ExportSel(), keys_out, counts_out, cap, n_out, "katgpu_table_export");
}

extern "C" katgpu_status katgpu_table_export_filtered(katgpu_table *t, uint64_t low_count, uint64_t high_count, uint32_t low_gc,
                                                      uint32_t high_gc, int inside, uint64_t *keys_out, uint64_t *counts_out,
                                                      size_t cap, uint64_t *n_out) {
    ExportSel sel;
    sel.mode = inside ? 1 : -1;
    sel.low_count = low_count, sel.high_count = high_count, sel.low_gc = low_gc, sel.high_gc = high_gc;
    return export_selected(t, sel, keys_out, counts_out, cap, n_out, "katgpu_table_export_filtered");
}

// Two words per key: every k (keys_hi_out receives zeros for a one-word table and may then be NULL).
static katgpu_status export_selected_wide(katgpu_table *t, const ExportSel &sel, uint64_t *keys_hi_out, uint64_t *keys_lo_out,
                                           uint64_t *counts_out, size_t cap, uint64_t *n_out, const char *who) {
    KT(finished(t, who));
    if (!t->wide) {
        KT(export_selected(t, sel, keys_lo_out, counts_out, cap, n_out, who));
        if (keys_hi_out) memset(keys_hi_out, 0, std::min<size_t>(cap, *n_out) * 8);
        return KATGPU_OK;
    }
    if (!n_out || ((!keys_hi_out || !keys_lo_out || !counts_out) && cap)) return fail(KATGPU_ERR_INVALID, "%s: NULL argument", who);
    CU(cudaSetDevice(t->ctx->device));
    return wide_export(t, sel, keys_hi_out, keys_lo_out, counts_out, cap, n_out);
}
extern "C" katgpu_status katgpu_table_export_wide(katgpu_table *t, uint64_t *keys_hi_out, uint64_t *keys_lo_out, uint64_t *counts_out,
                                                  size_t cap, uint64_t *n_out) {
    return export_selected_wide(t, ExportSel(), keys_hi_out, keys_lo_out, counts_out, cap, n_out, "katgpu_table_export_wide");
}
extern "C" katgpu_status katgpu_table_export_filtered_wide(katgpu_table *t, uint64_t low_count, uint64_t high_count, uint32_t low_gc,
                                                           uint32_t high_gc, int inside, uint64_t *keys_hi_out, uint64_t *keys_lo_out,
                                                           uint64_t *counts_out, size_t cap, uint64_t *n_out) {
    ExportSel sel;
    sel.mode = inside ? 1 : -1;
    sel.low_count = low_count, sel.high_count = high_count, sel.low_gc = low_gc, sel.high_gc = high_gc;
    return export_selected_wide(t, sel, keys_hi_out, keys_lo_out, counts_out, cap, n_out, "katgpu_table_export_filtered_wide");
}
