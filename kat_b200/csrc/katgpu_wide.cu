// katgpu_wide.cu — host side of the two-word-key tables (k = 33..64; wide.cuh, wide_kernels.cu): lifetime, growth,
// the admission of k-mer starts against the load limit, and the walks.  The C-ABI entry points of katgpu_api.cu /
// katgpu_walks.cu hand over to these functions when katgpu_table::wide is set; the host feed pipeline (staging
// ring, copy stream) and the device FASTQ parser are shared with the one-word path.
#include "katgpu_internal.cuh"

#include <algorithm>

using namespace katgpu;

struct WideTable {
    WideGeom g{};
    size_t keys_bytes = 0, counts_bytes = 0;
    WideCounters *h_ctr = nullptr; // pinned mirror
};

static const double kWideSoftLoad = 0.70; // admit k-mer starts up to this load (worst case: every one is a new key)

static katgpu_status wide_alloc(katgpu_ctx *c, int k, int canonical, uint64_t n_slots, WideGeom &g, size_t &keys_bytes, size_t &counts_bytes) {
    g = WideGeom{};
    g.n_slots = n_slots;
    g.mask = n_slots - 1;
    g.k = (uint32_t)k;
    g.canonical = canonical ? 1u : 0u;
    g.hi_mask = k == 64 ? ~0ULL : ((1ULL << (2 * k - 64)) - 1);
    g.hi_top = (uint32_t)(2 * k - 66);
    g.rc_shift = (uint32_t)(128 - 2 * k);
    g.max_probes = (uint32_t)std::min<uint64_t>(n_slots, 1u << 16);
    keys_bytes = (size_t)(n_slots + 1) * sizeof(Key128);
    counts_bytes = (size_t)(n_slots + 1) * 8;
    void *p = nullptr;
    CU(pool_alloc(c, keys_bytes, &p));
    g.keys = (Key128 *)p;
    CU(pool_alloc(c, counts_bytes, &p));
    g.counts = (unsigned long long *)p;
    CU(pool_alloc(c, 256, &p));
    g.ctr = (WideCounters *)p;
    ProfScope ps(c, "memset", c->stream);
    CU(cudaMemsetAsync(g.keys, 0xFF, keys_bytes, c->stream)); // empty marker: all ones
    CU(cudaMemsetAsync(g.counts, 0, counts_bytes, c->stream));
    CU(cudaMemsetAsync(g.ctr, 0, 256, c->stream));
    return KATGPU_OK;
}
static void wide_release(katgpu_ctx *c, WideGeom &g, size_t keys_bytes, size_t counts_bytes) {
    if (g.keys) pool_free(c, g.keys, keys_bytes);
    if (g.counts) pool_free(c, g.counts, counts_bytes);
    if (g.ctr) pool_free(c, g.ctr, 256);
    g.keys = nullptr;
    g.counts = nullptr;
    g.ctr = nullptr;
}

katgpu_status wide_begin(katgpu_table *t, int k, int canonical, uint64_t hash_size) {
    katgpu_ctx *c = t->ctx;
    WideTable *w = new (std::nothrow) WideTable();
    if (!w) return fail(KATGPU_ERR_NOMEM, "katgpu_count_begin: out of host memory");
    uint64_t n_slots = 1024;
    while (n_slots < hash_size && n_slots < (1ULL << 40)) n_slots <<= 1;
    if (cudaMallocHost((void **)&w->h_ctr, sizeof(WideCounters)) != cudaSuccess) {
        delete w;
        return fail(KATGPU_ERR_NOMEM, "katgpu_count_begin: out of pinned host memory");
    }
    const katgpu_status st = wide_alloc(c, k, canonical, n_slots, w->g, w->keys_bytes, w->counts_bytes);
    if (st != KATGPU_OK) {
        wide_release(c, w->g, w->keys_bytes, w->counts_bytes);
        cudaFreeHost(w->h_ctr);
        delete w;
        return st;
    }
    t->wide = w;
    t->n_slots = n_slots;
    t->g.k = (uint32_t)k; // the generic argument checks of the walks read these two
    t->g.canonical = canonical ? 1u : 0u;
    return KATGPU_OK;
}

void wide_free(katgpu_table *t) {
    WideTable *w = t->wide;
    if (!w) return;
    wide_release(t->ctx, w->g, w->keys_bytes, w->counts_bytes);
    if (w->h_ctr) cudaFreeHost(w->h_ctr);
    delete w;
    t->wide = nullptr;
}

static katgpu_status wide_read_counters(katgpu_table *t, WideCounters &out) {
    katgpu_ctx *c = t->ctx;
    WideTable *w = t->wide;
    CU(cudaMemcpyAsync(w->h_ctr, w->g.ctr, sizeof(WideCounters), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    out = *w->h_ctr;
    if (out.failed) return fail(KATGPU_ERR_HASH_FULL, "Hash full: a probe sequence of the wide table ran out");
    return KATGPU_OK;
}

// hash_counter::double_size (JFI/hash_counter.hpp:204-244)
static katgpu_status wide_grow(katgpu_table *t) {
    katgpu_ctx *c = t->ctx;
    WideTable *w = t->wide;
    WideGeom ng;
    size_t kb = 0, cb = 0;
    katgpu_status st = wide_alloc(c, (int)w->g.k, (int)w->g.canonical, w->g.n_slots * 2, ng, kb, cb);
    if (st != KATGPU_OK) {
        wide_release(c, ng, kb, cb);
        return st;
    }
    {
        ProfScope ps(c, "rehash", c->stream);
        launch_wide_rehash(w->g, ng, c->sm_count * 8, c->stream);
        c->launches++;
    }
    CU(cudaMemcpyAsync(&ng.ctr->total, &w->g.ctr->total, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    wide_release(c, w->g, w->keys_bytes, w->counts_bytes);
    w->g = ng;
    w->keys_bytes = kb;
    w->counts_bytes = cb;
    t->n_slots = ng.n_slots;
    t->grow_events++;
    return KATGPU_OK;
}

// How many new keys may be admitted now; grows the table when the answer would be too small.
static katgpu_status wide_admit(katgpu_table *t, uint64_t want, uint64_t &granted) {
    for (;;) {
        WideCounters ctr;
        KT(wide_read_counters(t, ctr));
        const uint64_t soft = (uint64_t)(kWideSoftLoad * (double)t->n_slots);
        const uint64_t room = ctr.claimed < soft ? soft - ctr.claimed : 0;
        const uint64_t min_chunk = std::min<uint64_t>(want, std::max<uint64_t>(t->n_slots / 64, (uint64_t)kWideTile));
        if (room >= min_chunk) {
            granted = std::min(want, room);
            return KATGPU_OK;
        }
        if (!t->allow_grow) {
            if (room > 0) {
                granted = std::min(want, room);
                return KATGPU_OK;
            }
            return fail(KATGPU_ERR_HASH_FULL, "Hash full");
        }
        KT(wide_grow(t));
    }
}

// k-mers starting in [0, n_starts) of dev[0 .. n_valid): extract + canonicalise + insert
katgpu_status wide_feed_range(katgpu_table *t, const uint8_t *dev, uint64_t n_valid, uint64_t n_starts) {
    katgpu_ctx *c = t->ctx;
    WideTable *w = t->wide;
    if (n_valid < (uint64_t)w->g.k) return KATGPU_OK;
    n_starts = std::min(n_starts, n_valid - w->g.k + 1);
    uint64_t done = 0;
    while (done < n_starts) {
        uint64_t granted = 0;
        KT(wide_admit(t, n_starts - done, granted));
        if (granted < n_starts - done && granted >= (uint64_t)kWideTile) granted = granted / kWideTile * kWideTile; // keep the next launch TMA-aligned
        WideExtractParams P{};
        P.g = w->g;
        P.bases = dev + done;
        P.n_valid = n_valid - done;
        P.n_starts = granted;
        P.tile_counter = c->tile_counter;
        P.n_tiles = (int)((granted + kWideTile - 1) / kWideTile);
        P.tma_ok = (reinterpret_cast<uintptr_t>(P.bases) & 15) == 0;
        CU(cudaMemsetAsync(c->tile_counter, 0, sizeof(unsigned int), c->stream));
        const int grid = std::min(P.n_tiles, c->sm_count * 4);
        {
            ProfScope ps(c, "count", c->stream);
            launch_wide_extract(P, grid, c->stream);
        }
        c->launches++;
        CU(cudaGetLastError());
        done += granted;
    }
    return KATGPU_OK;
}

katgpu_status wide_add_pairs(katgpu_table *t, const uint64_t *keys_hi, const uint64_t *keys_lo, const uint64_t *counts, size_t n) {
    katgpu_ctx *c = t->ctx;
    WideTable *w = t->wide;
    const size_t kBatch = 1u << 23; // records per staging round
    const size_t cap = std::min(n, kBatch);
    if (cap == 0) return KATGPU_OK;
    void *d = nullptr;
    CU(cudaMalloc(&d, cap * 24));
    uint64_t *d_hi = (uint64_t *)d, *d_lo = d_hi + cap, *d_cn = d_lo + cap;
    katgpu_status st = KATGPU_OK;
    size_t done = 0;
    while (done < n && st == KATGPU_OK) {
        uint64_t granted = 0;
        st = wide_admit(t, std::min(n - done, cap), granted);
        if (st != KATGPU_OK) break;
        cudaError_t e = cudaMemcpyAsync(d_hi, keys_hi + done, granted * 8, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_lo, keys_lo + done, granted * 8, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_cn, counts + done, granted * 8, cudaMemcpyHostToDevice, c->stream);
        if (e != cudaSuccess) {
            st = fail(KATGPU_ERR_CUDA, "katgpu_count_add_pairs_wide: %s", cudaGetErrorString(e));
            break;
        }
        {
            ProfScope ps(c, "insert", c->stream);
            const int grid = (int)std::min<uint64_t>((granted + 255) / 256, (uint64_t)c->sm_count * 8);
            launch_wide_add_pairs(w->g, d_hi, d_lo, d_cn, granted, grid, c->stream);
        }
        c->launches++;
        done += granted;
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) st = fail(KATGPU_ERR_CUDA, "katgpu_count_add_pairs_wide: %s", cudaGetErrorString(cudaGetLastError()));
    }
    cudaFree(d);
    return st;
}

// distinct keys / k-mer instances of the table as it stands (the extra slot included)
katgpu_status wide_totals(katgpu_table *t, uint64_t &n_distinct, uint64_t &n_total) {
    katgpu_ctx *c = t->ctx;
    WideTable *w = t->wide;
    WideCounters ctr;
    KT(wide_read_counters(t, ctr)); // reports a failed probe sequence
    void *d = nullptr;
    CU(pool_alloc(c, 256, &d));
    unsigned long long out[2] = {0, 0};
    cudaError_t e = cudaMemsetAsync(d, 0, 16, c->stream);
    if (e == cudaSuccess) {
        launch_wide_stats(w->g, (unsigned long long *)d, c->sm_count * 8, c->stream);
        c->launches++;
        e = cudaMemcpyAsync(out, d, 16, cudaMemcpyDeviceToHost, c->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    pool_free(c, d, 256);
    if (e != cudaSuccess) return fail(KATGPU_ERR_CUDA, "wide table totals: %s", cudaGetErrorString(e));
    n_distinct = out[0];
    n_total = out[1];
    return KATGPU_OK;
}

uint64_t wide_bytes(const katgpu_table *t) { return t->wide->keys_bytes + t->wide->counts_bytes + 256; }

// ---- walks -------------------------------------------------------------------------------------------------------
struct WideDevBuf {
    void *p = nullptr;
    ~WideDevBuf() {
        if (p) cudaFree(p);
    }
};

katgpu_status wide_hist(katgpu_table *t, uint64_t base, uint64_t ceil_, uint64_t inc, uint64_t nb, uint64_t *bins_out) {
    katgpu_ctx *c = t->ctx;
    WideDevBuf d;
    CU(cudaMalloc(&d.p, nb * 8));
    CU(cudaMemsetAsync(d.p, 0, nb * 8, c->stream));
    {
        ProfScope ps(c, "hist", c->stream);
        launch_wide_hist(t->wide->g, base, ceil_, inc, (uint32_t)nb, (unsigned long long *)d.p, c->sm_count * 4, c->stream);
    }
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(bins_out, d.p, nb * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return KATGPU_OK;
}

katgpu_status wide_gcp(katgpu_table *t, double cvg_scale, uint32_t cvg_bins, uint64_t *mx_out) {
    katgpu_ctx *c = t->ctx;
    const size_t cells = (size_t)t->wide->g.k * ((size_t)cvg_bins + 1);
    WideDevBuf d;
    CU(cudaMalloc(&d.p, cells * 8));
    CU(cudaMemsetAsync(d.p, 0, cells * 8, c->stream));
    {
        ProfScope ps(c, "gcp", c->stream);
        launch_wide_gcp(t->wide->g, cvg_scale, cvg_bins, (unsigned long long *)d.p, c->sm_count * 2, c->stream);
    }
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(mx_out, d.p, cells * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return KATGPU_OK;
}

katgpu_status wide_comp(katgpu_table *t1, katgpu_table *t2, katgpu_table *t3, double d1_scale, double d2_scale, uint32_t d1_bins,
                        uint32_t d2_bins, uint64_t *main_mx, uint64_t *ends_mx, uint64_t *middle_mx, uint64_t *mixed_mx,
                        katgpu_comp_counters *counters, uint64_t *spectra) {
    katgpu_ctx *c = t1->ctx;
    const uint32_t len = std::min(d1_bins, d2_bins);
    const size_t cells = (size_t)d1_bins * d2_bins;
    const size_t n_mx = t3 ? 4 : 1;
    const size_t words = n_mx * cells + 16 + 4 * (size_t)len;
    WideDevBuf d;
    CU(cudaMalloc(&d.p, words * 8));
    CU(cudaMemsetAsync(d.p, 0, words * 8, c->stream));
    unsigned long long *base = (unsigned long long *)d.p;
    WideCompParams P{};
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
        P.ga = t1->wide->g;
        P.gb = t2->wide->g;
        P.canon_b = (int)t2->wide->g.canonical;
        P.have_c = t3 != nullptr;
        P.gc = t3 ? t3->wide->g : t2->wide->g;
        P.canon_c = t3 ? (int)t3->wide->g.canonical : 0;
        P.spec_walk = spec;
        P.spec_shared1 = spec + 2 * (size_t)len;
        P.spec_shared2 = spec + 3 * (size_t)len;
        launch_wide_comp(P, grid, c->stream);
        // pass 2: walk hash2, probe hash1 — always canonicalised (src/comp.cc:447)
        P.pass = 2;
        P.ga = t2->wide->g;
        P.gb = t1->wide->g;
        P.canon_b = 1;
        P.have_c = 0;
        P.spec_walk = spec + len;
        launch_wide_comp(P, grid, c->stream);
        c->launches += 2;
        if (t3) { // pass 3: hash3 totals (src/comp.cc:466-479)
            P.pass = 3;
            P.ga = t3->wide->g;
            launch_wide_comp(P, grid, c->stream);
            c->launches++;
        }
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(main_mx, P.main_mx, cells * 8, cudaMemcpyDeviceToHost, c->stream));
    if (t3) {
        CU(cudaMemcpyAsync(ends_mx, P.ends_mx, cells * 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(middle_mx, P.middle_mx, cells * 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(mixed_mx, P.mixed_mx, cells * 8, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaMemcpyAsync(counters, P.counters, sizeof(katgpu_comp_counters), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(spectra, spec, 4 * (size_t)len * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return KATGPU_OK;
}

katgpu_status wide_lookup_host(katgpu_table *t, const uint64_t *keys_hi, const uint64_t *keys_lo, size_t n, int canonicalise, uint64_t *counts_out) {
    katgpu_ctx *c = t->ctx;
    WideDevBuf d;
    CU(cudaMalloc(&d.p, n * 24));
    uint64_t *d_hi = (uint64_t *)d.p, *d_lo = d_hi + n, *d_out = d_lo + n;
    CU(cudaMemcpyAsync(d_hi, keys_hi, n * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d_lo, keys_lo, n * 8, cudaMemcpyHostToDevice, c->stream));
    const int grid = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)c->sm_count * 8);
    launch_wide_lookup(t->wide->g, d_hi, d_lo, n, canonicalise, d_out, grid, c->stream);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(counts_out, d_out, n * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return KATGPU_OK;
}

katgpu_status wide_export(katgpu_table *t, const ExportSel &sel, uint64_t *keys_hi_out, uint64_t *keys_lo_out, uint64_t *counts_out, size_t cap,
                          uint64_t *n_out) {
    katgpu_ctx *c = t->ctx;
    const size_t m = std::max<size_t>(cap, 1);
    WideDevBuf d, dn;
    CU(cudaMalloc(&d.p, m * 24));
    CU(cudaMalloc(&dn.p, 8));
    uint64_t *d_hi = (uint64_t *)d.p, *d_lo = d_hi + m, *d_cn = d_lo + m;
    CU(cudaMemsetAsync(dn.p, 0, 8, c->stream));
    launch_wide_export(t->wide->g, d_hi, d_lo, d_cn, cap, (unsigned long long *)dn.p, c->sm_count * 8, c->stream, sel);
    c->launches++;
    CU(cudaGetLastError());
    unsigned long long n = 0;
    CU(cudaMemcpyAsync(&n, dn.p, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *n_out = n;
    const size_t got = std::min<size_t>(cap, n);
    if (got) {
        if (keys_hi_out) CU(cudaMemcpy(keys_hi_out, d_hi, got * 8, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(keys_lo_out, d_lo, got * 8, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(counts_out, d_cn, got * 8, cudaMemcpyDeviceToHost));
    }
    return KATGPU_OK;
}
