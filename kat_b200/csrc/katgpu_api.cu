// katgpu_api.cu — C-ABI of libkatgpu (include/katgpu.h): contexts, the device block cache, table
// lifetime (geometry, allocation, admission control, doubling), the counting paths and the host->device
// feed pipeline.  The raw-FASTQ feed is in katgpu_fastq.cu, the table walks in katgpu_walks.cu, the
// multi-GPU engine in katgpu_dist.cu.
#include "katgpu_internal.cuh"
#include "synth.cuh"

#include <algorithm>
#include <atomic>
#include <thread>
#include <cstdarg>
#include <cstring>
#include <cmath>
#include <new>

using namespace katgpu;

static thread_local char g_err[512] = "";
static int g_debug_count_bits = 0;

katgpu_status katgpu_fail(katgpu_status st, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return st;
}

// ---- device block cache ----
void pool_flush(katgpu_ctx *c) {
    for (auto &kv : c->pool) cudaFree(kv.second);
    c->pool.clear();
    c->pooled_bytes = 0;
}
// Cached blocks are handed out best-fit: the smallest one that holds the request and is at most
// 25 % larger (buffer sizes follow the free memory of the moment, so the second table of a job asks
// for a little less than the first got; cudaMalloc / cudaFree of tens of GB cost tens of ms each).
cudaError_t pool_alloc(katgpu_ctx *c, size_t bytes, void **out) {
    bytes = std::max<size_t>(bytes, 256);
    auto it = c->pool.lower_bound(bytes);
    if (it != c->pool.end() && it->first <= bytes + bytes / 4) {
        *out = it->second;
        c->pooled_bytes -= it->first;
        c->block_size[it->second] = it->first;
        c->pool.erase(it);
        return cudaSuccess;
    }
    TraceScope trace__("      cudaMalloc (block cache miss)");
    cudaError_t e = cudaMalloc(out, bytes);
    if (e == cudaErrorMemoryAllocation && !c->pool.empty()) {
        cudaGetLastError();
        if (getenv("KATGPU_TRACE")) fprintf(stderr, "[katgpu]       out of memory for %.2f GB: releasing %.2f GB of cached blocks\n", bytes / 1073741824.0, c->pooled_bytes / 1073741824.0);
        pool_flush(c);
        e = cudaMalloc(out, bytes);
    }
    if (e == cudaSuccess) c->block_size[*out] = bytes;
    if (getenv("KATGPU_TRACE")) fprintf(stderr, "[katgpu]       new block of %.3f GB (%zu cached blocks, %.2f GB)\n", bytes / 1073741824.0, c->pool.size(), c->pooled_bytes / 1073741824.0);
    return e;
}
void pool_free(katgpu_ctx *c, void *p, size_t bytes) {
    if (!p) return;
    bytes = std::max<size_t>(bytes, 256);
    auto it = c->block_size.find(p); // the block's true size (it may have been handed out for a smaller request)
    if (it != c->block_size.end()) {
        bytes = it->second;
        c->block_size.erase(it);
    }
    c->pool.emplace(bytes, p);
    c->pooled_bytes += bytes;
}

static void prof_collect(katgpu_ctx *c) {
    for (auto &sp : c->spans) {
        cudaEventSynchronize(sp.b);
        float ms = 0;
        cudaEventElapsedTime(&ms, sp.a, sp.b);
        auto &e = c->prof[sp.family];
        e.first += ms;
        e.second += 1;
        cudaEventDestroy(sp.a);
        cudaEventDestroy(sp.b);
    }
    c->spans.clear();
}

// ---- basics ----
extern "C" const char *katgpu_last_error(void) { return g_err; }
extern "C" int katgpu_version(void) { return KATGPU_VERSION; }
extern "C" int katgpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
extern "C" void katgpu_debug_set_count_bits(int bits) { g_debug_count_bits = bits; }

extern "C" katgpu_status katgpu_init(int device, katgpu_ctx **out) {
    TraceScope trace__("katgpu_init");
    if (!out) return fail(KATGPU_ERR_INVALID, "katgpu_init: out is NULL");
    *out = nullptr;
    int n = katgpu_device_count();
    if (n <= 0)
        return fail(KATGPU_ERR_NO_DEVICE, "katgpu_init: no CUDA device visible; libkatgpu has no CPU path");
    if (device < 0 || device >= n) return fail(KATGPU_ERR_INVALID, "katgpu_init: device %d out of range (%d devices)", device, n);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(KATGPU_ERR_NO_DEVICE, "katgpu_init: device %d is sm_%d%d; libkatgpu is built for sm_100a only", device,
                    prop.major, prop.minor);
    // Random 32-byte bucket probes: ask L2 to fetch 32-byte sectors from HBM instead of
    // promoting every miss to 64/128 bytes (a hint; KATGPU_L2_FETCH_BYTES overrides).
    {
        size_t gran = 32;
        if (const char *e = getenv("KATGPU_L2_FETCH_BYTES")) gran = (size_t)atoi(e);
        if (gran) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
        cudaGetLastError();
    }
    katgpu_ctx *c = new (std::nothrow) katgpu_ctx();
    if (!c) return fail(KATGPU_ERR_NOMEM, "katgpu_init: out of host memory");
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CU(extract_configure(kModeInsert, &c->bps[0]));
    CU(extract_configure(kModeRoute, &c->bps[1]));
    CU(partition_configure(&c->bps_part));
    if (c->bps_part < 1) return fail(KATGPU_ERR_CUDA, "katgpu_init: partition kernel does not fit on an SM");
    if (c->bps_part > 2) c->bps_part = 2;
    CU(sweep_occupancy(&c->bps_sweep));
    CU(tile_configure(&c->bps_bin1, &c->bps_tile));
    if (c->bps_tile < 1) return fail(KATGPU_ERR_CUDA, "katgpu_init: tile kernel does not fit on an SM");
    if (c->bps[0] < 1 || c->bps[1] < 1) return fail(KATGPU_ERR_CUDA, "katgpu_init: extract kernel does not fit on an SM");
    CU(cudaMalloc(&c->tile_counter, 256));
    CU(cudaMalloc(&c->route_cursors, 256));
    CU(cudaMallocHost(&c->h_ctr, sizeof(TableCounters)));
    for (int i = 0; i < kRing; i++) {
        CU(cudaEventCreateWithFlags(&c->chunk_free[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->chunk_ready[i], cudaEventDisableTiming));
    }
    *out = c;
    return KATGPU_OK;
}

extern "C" void katgpu_shutdown(katgpu_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->copy_stream);
    dist_shutdown(c);
    prof_collect(c);
    for (int i = 0; i < kRing; i++) {
        if (c->chunk[i]) cudaFree(c->chunk[i]);
        if (c->chunk_free[i]) cudaEventDestroy(c->chunk_free[i]);
        if (c->chunk_ready[i]) cudaEventDestroy(c->chunk_ready[i]);
    }
    for (int i = 0; i < 2; i++) {
        if (c->raw[i]) cudaFree(c->raw[i]);
        if (c->raw_free[i]) cudaEventDestroy(c->raw_free[i]);
    }
    if (c->nl_buf) cudaFree(c->nl_buf);
    if (c->nl_tiles_buf) cudaFree(c->nl_tiles_buf);
    if (c->parse_flags) cudaFree(c->parse_flags);
    for (int i = 0; i < 8; i++) {
        if (c->bounce[i]) cudaFreeHost(c->bounce[i]);
        if (c->bounce_free[i]) cudaEventDestroy(c->bounce_free[i]);
    }
    pool_flush(c);
    if (c->tile_counter) cudaFree(c->tile_counter);
    if (c->route_cursors) cudaFree(c->route_cursors);
    if (c->h_ctr) cudaFreeHost(c->h_ctr);
    cudaStreamDestroy(c->stream);
    cudaStreamDestroy(c->copy_stream);
    delete c;
}

extern "C" katgpu_status katgpu_host_alloc(size_t bytes, void **out) {
    if (!out) return fail(KATGPU_ERR_INVALID, "katgpu_host_alloc: out is NULL");
    CU(cudaMallocHost(out, bytes ? bytes : 1));
    return KATGPU_OK;
}
extern "C" void katgpu_host_free(void *p) {
    if (p) cudaFreeHost(p);
}
extern "C" katgpu_status katgpu_device_alloc(katgpu_ctx *c, size_t bytes, void **out) {
    if (!c || !out) return fail(KATGPU_ERR_INVALID, "katgpu_device_alloc: NULL argument");
    CU(cudaSetDevice(c->device));
    cudaError_t e = cudaMalloc(out, bytes ? bytes : 1);
    if (e == cudaErrorMemoryAllocation && !c->pool.empty()) { // give cached blocks back first
        cudaGetLastError();
        pool_flush(c);
        e = cudaMalloc(out, bytes ? bytes : 1);
    }
    CU(e);
    return KATGPU_OK;
}
extern "C" void katgpu_device_free(katgpu_ctx *c, void *dev) {
    if (c) cudaSetDevice(c->device);
    if (dev) cudaFree(dev);
}
extern "C" katgpu_status katgpu_ipc_export(void *dev, void *handle64) {
    if (!dev || !handle64) return fail(KATGPU_ERR_INVALID, "katgpu_ipc_export: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
    CU(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t *>(handle64), dev));
    return KATGPU_OK;
}
extern "C" katgpu_status katgpu_ipc_import(const void *handle64, void **out) {
    if (!handle64 || !out) return fail(KATGPU_ERR_INVALID, "katgpu_ipc_import: NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    CU(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return KATGPU_OK;
}
extern "C" void katgpu_ipc_close(void *dev) {
    if (dev) cudaIpcCloseMemHandle(dev);
}
extern "C" katgpu_status katgpu_synchronize(katgpu_ctx *c) {
    if (!c) return fail(KATGPU_ERR_INVALID, "katgpu_synchronize: NULL context");
    CU(cudaStreamSynchronize(c->copy_stream));
    CU(cudaStreamSynchronize(c->stream));
    return KATGPU_OK;
}
extern "C" katgpu_status katgpu_release_cached_memory(katgpu_ctx *c) {
    if (!c) return fail(KATGPU_ERR_INVALID, "katgpu_release_cached_memory: NULL context");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->copy_stream));
    pool_flush(c);
    return KATGPU_OK;
}
extern "C" katgpu_status katgpu_memcpy_d2h(katgpu_ctx *c, void *dst, const void *src, size_t n) {
    if (!c) return fail(KATGPU_ERR_INVALID, "katgpu_memcpy_d2h: NULL context");
    CU(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return KATGPU_OK;
}
extern "C" katgpu_status katgpu_memcpy_h2d(katgpu_ctx *c, void *dst, const void *src, size_t n) {
    if (!c) return fail(KATGPU_ERR_INVALID, "katgpu_memcpy_h2d: NULL context");
    CU(cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return KATGPU_OK;
}
extern "C" uint64_t katgpu_launch_count(katgpu_ctx *c) { return c ? c->launches : 0; }
extern "C" katgpu_status katgpu_profile_enable(katgpu_ctx *c, int enable) {
    if (!c) return fail(KATGPU_ERR_INVALID, "katgpu_profile_enable: NULL context");
    CU(cudaStreamSynchronize(c->stream));
    prof_collect(c);
    c->prof.clear();
    c->profiling = enable != 0;
    return KATGPU_OK;
}
extern "C" katgpu_status katgpu_profile_get(katgpu_ctx *c, const char *family, double *ms, uint64_t *launches) {
    if (!c || !family) return fail(KATGPU_ERR_INVALID, "katgpu_profile_get: NULL argument");
    CU(cudaStreamSynchronize(c->stream));
    prof_collect(c);
    auto it = c->prof.find(family);
    if (ms) *ms = it == c->prof.end() ? 0.0 : it->second.first;
    if (launches) *launches = it == c->prof.end() ? 0 : it->second.second;
    return KATGPU_OK;
}

// Per-launch bound on the increments one slot can receive -> when RED is provably safe.
void set_add_bound(TableGeom &g, uint64_t bound) {
    g.add_bound = bound;
    g.red_below = bound <= g.cmax ? g.cmax - bound + 1 : 0;
}

// ---- table geometry ----
static uint64_t inv_mod_2_64(uint64_t a) { // Newton iteration, a odd
    uint64_t x = a;
    for (int i = 0; i < 6; i++) x *= 2 - a * x;
    return x;
}
static int ceil_log2(uint64_t v) {
    int l = 0;
    while (l < 63 && (1ULL << l) < v) l++;
    return l;
}

// Fill every field of g except the pointers.  Lb = log2(buckets).
static katgpu_status make_geom(int k, int canonical, int Lb, int count_bits_cap, TableGeom &g) {
    const int b = 2 * k;
    if (Lb > b) Lb = b;
    if (Lb < 1) Lb = 1;
    const int rbits = b - Lb;
    int cbits = 64 - (rbits + kProbeBits + 1);
    // The count field never reaches into the low word (cshift >= 32): the tile kernel updates it
    // with a native 32-bit shared-memory atomic on the high word.  Counts beyond 2^32 - 1 carry
    // into the exact side table like any other field overflow.
    if (cbits > 32) cbits = 32;
    if (count_bits_cap >= 2 && count_bits_cap < cbits) cbits = count_bits_cap;
    if (cbits < 2) return fail(KATGPU_ERR_INVALID, "table geometry: no room for a count field (k=%d, Lb=%d)", k, Lb);
    g.k = k;
    g.b = b;
    g.Lb = Lb;
    g.rbits = rbits;
    g.cbits = cbits;
    g.cshift = 64 - cbits;
    g.sh = (b + 1) / 2;
    g.canonical = canonical ? 1 : 0;
    g.rmask = rbits ? ((1ULL << rbits) - 1) : 0;
    g.tagmask = (1ULL << (rbits + kProbeBits)) - 1;
    g.kmask = b == 64 ? ~0ULL : ((1ULL << b) - 1);
    g.bmask = (1ULL << Lb) - 1;
    g.cmax = (1ULL << cbits) - 1;
    g.m1inv = inv_mod_2_64(kMixM1);
    g.m2inv = inv_mod_2_64(kMixM2);
    g.spill_cap = kSpillCap;
    g.add_bound = ~0ULL; // set per launch (set_add_bound)
    g.red_below = 0;
    return KATGPU_OK;
}
// Smallest Lb that leaves a count field of at least 16 bits (so that count
// overflow into the side table is rare), and at least 2^10 buckets.
static int min_Lb(int k) {
    const int b = 2 * k;
    int lb = b - (64 - kProbeBits - 1 - 16);
    if (lb < 10) lb = 10;
    if (lb > b) lb = b;
    return lb;
}

static katgpu_status alloc_table_storage(katgpu_ctx *c, TableGeom &g, uint64_t &n_slots) {
    TraceScope trace__("  alloc_table_storage");
    n_slots = (uint64_t)kSlotsPerBucket << g.Lb;
    void *p = nullptr;
    CU(pool_alloc(c, n_slots * sizeof(uint64_t), &p));
    g.slots = (uint64_t *)p;
    CU(pool_alloc(c, ((size_t)kOvfCap + 1) * 2 * sizeof(uint64_t), &p));
    g.ovf = (uint64_t *)p;
    CU(pool_alloc(c, (size_t)kSpillCap * 2 * sizeof(uint64_t), &p));
    g.spill = (uint64_t *)p;
    CU(pool_alloc(c, sizeof(TableCounters), &p));
    g.ctr = (TableCounters *)p;
    {
        ProfScope ps(c, "memset", c->stream);
        CU(cudaMemsetAsync(g.slots, 0, n_slots * sizeof(uint64_t), c->stream));
        CU(cudaMemsetAsync(g.ovf, 0, ((size_t)kOvfCap + 1) * 2 * sizeof(uint64_t), c->stream));
        CU(cudaMemsetAsync(g.ctr, 0, sizeof(TableCounters), c->stream));
    }
    return KATGPU_OK;
}
static void free_table_storage(katgpu_ctx *c, TableGeom &g) {
    if (g.slots) pool_free(c, g.slots, ((uint64_t)kSlotsPerBucket << g.Lb) * sizeof(uint64_t));
    if (g.ovf) pool_free(c, g.ovf, ((size_t)kOvfCap + 1) * 2 * sizeof(uint64_t));
    if (g.spill) pool_free(c, g.spill, (size_t)kSpillCap * 2 * sizeof(uint64_t));
    if (g.ctr) pool_free(c, g.ctr, sizeof(TableCounters));
    g.slots = g.ovf = g.spill = nullptr;
    g.ctr = nullptr;
}

extern "C" katgpu_status katgpu_count_begin(katgpu_ctx *c, int k, int canonical, uint64_t hash_size, int allow_grow,
                                            katgpu_table **out) {
    TraceScope trace__("katgpu_count_begin");
    if (!c || !out) return fail(KATGPU_ERR_INVALID, "katgpu_count_begin: NULL argument");
    *out = nullptr;
    if (k < 1 || k > 64) return fail(KATGPU_ERR_INVALID, "katgpu_count_begin: k=%d outside 1..64 (one or two 64-bit words per k-mer)", k);
    if (k > 32 && c->dist) return fail(KATGPU_ERR_INVALID, "katgpu_count_begin: k=%d > 32 is counted on one GPU (not on a multi-GPU context)", k);
    CU(cudaSetDevice(c->device));
    katgpu_table *t = new (std::nothrow) katgpu_table();
    if (!t) return fail(KATGPU_ERR_NOMEM, "katgpu_count_begin: out of host memory");
    t->ctx = c;
    t->allow_grow = allow_grow != 0;
    t->count_bits_cap = g_debug_count_bits;
    if (k > 32) { // two-word keys: their own table layout and kernels (wide.cuh)
        const katgpu_status stw = wide_begin(t, k, canonical, hash_size);
        if (stw != KATGPU_OK) {
            delete t;
            return stw;
        }
        *out = t;
        return KATGPU_OK;
    }
    int Lb = ceil_log2(std::max<uint64_t>(hash_size, 1)) - 2; // slots -> buckets
    Lb = std::max(Lb, min_Lb(k));
    katgpu_status st = make_geom(k, canonical, Lb, t->count_bits_cap, t->g);
    if (st == KATGPU_OK) st = alloc_table_storage(c, t->g, t->n_slots);
    if (st != KATGPU_OK) {
        free_table_storage(c, t->g);
        delete t;
        return st;
    }
    *out = t;
    return KATGPU_OK;
}

static void pend_release(katgpu_table *t) {
    katgpu_ctx *c = t->ctx;
    if (t->pend_keys) pool_free(c, t->pend_keys, t->pend_bytes);
    if (t->pend_seglen) pool_free(c, t->pend_seglen, t->pend_seglen_bytes);
    if (t->pend_hll) pool_free(c, t->pend_hll, kHllM * 4);
    if (t->pend_done) pool_free(c, t->pend_done, kMaxParts * 4);
    if (t->side_keys) pool_free(c, t->side_keys, t->side_bytes);
    if (t->side_ctr) pool_free(c, t->side_ctr, 64);
    if (t->side_deferred) pool_free(c, t->side_deferred, ((size_t)1 << 20) * 4);
    t->pend_keys = nullptr;
    t->pend_seglen = t->pend_hll = t->pend_done = nullptr;
    t->side_keys = nullptr;
    t->side_ctr = nullptr;
    t->side_deferred = nullptr;
}

extern "C" void katgpu_table_free(katgpu_table *t) {
    TraceScope trace__("katgpu_table_free");
    if (!t) return;
    cudaSetDevice(t->ctx->device);
    cudaStreamSynchronize(t->ctx->stream);
    if (t->wide) wide_free(t);
    else free_table_storage(t->ctx, t->g);
    pend_release(t);
    delete t;
}

// Read the device counter block (synchronises the compute stream).
katgpu_status read_counters(katgpu_table *t, TableCounters &out) {
    katgpu_ctx *c = t->ctx;
    if (t->wide) { // wide tables keep their own counter block (katgpu_wide.cu)
        out = TableCounters{};
        return KATGPU_OK;
    }
    CU(cudaMemcpyAsync(c->h_ctr, t->g.ctr, sizeof(TableCounters), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    out = *c->h_ctr;
    return KATGPU_OK;
}

// Double the table (hash_counter::double_size, JFI/hash_counter.hpp:204-244)
// and replay any parked keys.
static katgpu_status grow_table(katgpu_table *t, const TableCounters &cur, int target_Lb = 0) {
    katgpu_ctx *c = t->ctx;
    if ((int)t->g.Lb >= (int)t->g.b) return fail(KATGPU_ERR_HASH_FULL, "Hash full: table already covers the whole %u-bit key space", t->g.b);
    TableGeom ng{};
    KT(make_geom(t->g.k, t->g.canonical, std::max<int>(t->g.Lb + 1, target_Lb), t->count_bits_cap, ng));
    uint64_t new_slots = 0;
    katgpu_status st = alloc_table_storage(c, ng, new_slots);
    if (st != KATGPU_OK) {
        free_table_storage(c, ng);
        return st;
    }
    const int grid = c->sm_count * 8;
    {
        ProfScope ps(c, "rehash", c->stream);
        launch_rehash(t->g, ng, t->n_slots, grid, c->stream);
        c->launches++;
        const uint64_t nsp = std::min<uint64_t>(cur.spill_n, kSpillCap);
        if (nsp) {
            launch_insert_pairs(ng, t->g.spill, nsp, grid, c->stream);
            c->launches++;
        }
    }
    // carry the instance counter over (n_distinct is recomputed by the rehash)
    CU(cudaMemcpyAsync(&ng.ctr->n_total, &t->g.ctr->n_total, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    free_table_storage(c, t->g);
    t->g = ng;
    t->n_slots = new_slots;
    t->grow_events++;
    return KATGPU_OK;
}

static const double kMaxLoad = 0.80;  // grow beyond this
static const double kHardLoad = 0.90; // never admit more k-mers than this could hold

// After a launch: replay spills (growing), check the overflow side table.
katgpu_status settle(katgpu_table *t, TableCounters &ctr) {
    KT(read_counters(t, ctr));
    if (ctr.ovf_full) return fail(KATGPU_ERR_OVERFLOW, "count-overflow side table exhausted (%u keys)", kOvfCap);
    while (ctr.spill_n > 0) {
        if (ctr.spill_n > kSpillCap)
            return fail(KATGPU_ERR_HASH_FULL, "Hash full: %llu keys could not be placed (spill list holds %llu)", ctr.spill_n, (unsigned long long)kSpillCap);
        if (!t->allow_grow) return fail(KATGPU_ERR_HASH_FULL, "Hash full");
        KT(grow_table(t, ctr));
        KT(read_counters(t, ctr));
    }
    return KATGPU_OK;
}

// How many k-mer starts may be admitted now; grows the table when the answer
// would be too small.  Worst case every admitted k-mer is a new key.
static katgpu_status admit(katgpu_table *t, TableCounters &ctr, uint64_t want, uint64_t &granted) {
    for (;;) {
        const double slots = (double)t->n_slots;
        const bool full_space = (int)t->g.Lb >= (int)t->g.b; // bijective: cannot overflow
        if (full_space) {
            granted = want;
            return KATGPU_OK;
        }
        const uint64_t hard = (uint64_t)(slots * kHardLoad);
        const uint64_t room = ctr.n_distinct < hard ? hard - ctr.n_distinct : 0;
        const bool over_soft = (double)ctr.n_distinct > slots * kMaxLoad;
        const uint64_t min_chunk = std::min<uint64_t>(want, std::max<uint64_t>(t->n_slots / 64, 4096));
        if (room >= min_chunk && !(over_soft && t->allow_grow)) {
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
        KT(grow_table(t, ctr));
        KT(read_counters(t, ctr));
    }
}

static katgpu_status launch_extract_insert(katgpu_table *t, const uint8_t *dev, uint64_t n_valid, uint64_t n_starts) {
    katgpu_ctx *c = t->ctx;
    ExtractParams P{};
    P.g = t->g;
    set_add_bound(P.g, n_starts); // every k-mer of this launch could hit the same slot
    P.bases = dev;
    P.n_valid = n_valid;
    P.n_starts = n_starts;
    P.tile_counter = c->tile_counter;
    P.n_tiles = (int)((n_starts + kTile - 1) / kTile);
    P.tma_ok = (reinterpret_cast<uintptr_t>(dev) & 15) == 0;
    if (P.n_tiles == 0) return KATGPU_OK;
    CU(cudaMemsetAsync(c->tile_counter, 0, sizeof(unsigned int), c->stream));
    const int grid = std::min(P.n_tiles, c->sm_count * c->bps[0]);
    {
        ProfScope ps(c, "count", c->stream);
        launch_extract(kModeInsert, P, grid, c->stream);
    }
    c->launches++;
    CU(cudaGetLastError());
    return KATGPU_OK;
}

// ---- partitioned (L2-resident) counting path ----------------------------------
static int count_path_mode() { // 0 auto, 1 direct, 2 partition (L2 sweep), 3 tiled (shared-memory tiles)
    const char *e = getenv("KATGPU_COUNT_PATH");
    if (!e) return 0;
    if (!strcmp(e, "direct")) return 1;
    if (!strcmp(e, "partition")) return 2;
    if (!strcmp(e, "tiled")) return 3;
    return 0;
}
static const int kAutoPendingPath = 3; // which pending-key path "auto" picks for large jobs (2 = L2 sweep, 3 = tiled)
static bool tiled_possible(const katgpu_table *t) { return (int)t->g.Lb >= kTileBucketsLog2 + 2; }
bool use_partition(const katgpu_table *t, uint64_t kmers_hint) {
    if (t->wide) return false; // two-word keys are inserted in read order
    if (t->pend_keys) return true; // keep the path the pending buffer was built for
    const int m = count_path_mode();
    if (m == 1 || t->g.Lb < 6) return false;
    if (m == 2) return true;
    if (m == 3) return tiled_possible(t);
    // worth it once the table is far larger than L2 and the batch amortises one table pass
    return t->n_slots * 8 >= (512ull << 20) && kmers_hint >= (32ull << 20);
}
static bool want_tiled(const katgpu_table *t) {
    const int m = count_path_mode();
    return tiled_possible(t) && (m == 3 || (m == 0 && kAutoPendingPath == 3));
}

// Staging-ring slots per bin for a binning pass of `tile_keys` keys over F bins: the leftover of
// the previous flush (< granule) + mean + ~4 sigma, within `max_bytes` of shared memory (a ring that
// is too small only sends more keys the slow way).
uint32_t ring_slots(uint32_t F, uint32_t tile_keys, size_t max_bytes, uint32_t granule) {
    const double mean = (double)tile_keys / (double)F;
    uint64_t R = granule + (uint64_t)std::ceil(mean + 4.0 * std::sqrt(mean));
    const uint64_t fit = max_bytes > (size_t)F * 8 + 16 ? (max_bytes - (size_t)F * 8 - 16) / ((size_t)F * 8) : 4;
    R = std::max<uint64_t>(4, std::min(R, fit));
    return (uint32_t)R;
}

katgpu_status pend_alloc(katgpu_table *t, uint64_t kmers_hint, bool &ok) {
    TraceScope trace__("  pend_alloc");
    ok = true;
    if (t->pend_keys) return KATGPU_OK;
    katgpu_ctx *c = t->ctx;
    const bool tiled = want_tiled(t);
    const uint64_t table_bytes = t->n_slots * 8;
    uint64_t slice_mb = 32; // table bytes swept at a time; must stay L2-resident
    if (const char *e = getenv("KATGPU_SLICE_MB")) slice_mb = std::max(1, atoi(e));
    int Pb = ceil_log2(std::max<uint64_t>(1, table_bytes / (slice_mb << 20)));
    Pb = std::max(1, std::min(Pb, 9));
    Pb = std::min<int>(Pb, (int)t->g.Lb - 2);
    if (tiled) // level 1 of two: leave at least one bit for level 2 (slice -> 64 KB tile)
        Pb = std::max(1, std::min(9, (int)t->g.Lb - kTileBucketsLog2 - 1));
    const uint64_t tile_pos = tiled ? (c->dist ? kBinTile : kBin1Tile) : kTile; // positions one CTA bins between two flushes
    // level-1 segment columns = CTAs of the kernel that fills them: bin_bases_kernel (bps_bin1 per SM), or on a
    // multi-GPU job bin_hashed_kernel (one per SM)
    const uint64_t nparts = 1ull << Pb, n_segs = (uint64_t)c->sm_count * (c->dist ? 1 : tiled ? c->bps_bin1 : c->bps_part);
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    const uint64_t budget_keys = (uint64_t)((double)(free_b + c->pooled_bytes) * (tiled ? 0.60 : 0.70) / 8.0);
    // Size every (slice, CTA) segment from the same fill model feed_partitioned() uses, so that a
    // feed of `kmers_hint` positions fits in ONE partition round (each extra round = one more full
    // pass over the table in the sweep).
    const double m_tile = (double)tile_pos / (double)nparts; // keys one tile sends to one slice
    const double pad = tiled ? 0.0 : 2.0;                    // the L2-sweep path pads every run to 4 keys
    const uint64_t hint = std::max<uint64_t>(kmers_hint, t->n_slots / 4);
    const uint64_t cta_tiles = ((hint + tile_pos - 1) / tile_pos + n_segs - 1) / n_segs + 1;
    uint64_t seg_cap = (uint64_t)(((double)cta_tiles * (m_tile + pad) + 4.0 * std::sqrt((double)cta_tiles * m_tile)) * 1.02) + 64;
    seg_cap = (seg_cap + 15) / 16 * 16;
    const uint64_t min_cap = ((uint64_t)(2.2 * (m_tile + 4.0 * std::sqrt(m_tile) + 4.0)) + 255) / 16 * 16;
    seg_cap = std::max(seg_cap, min_cap);
    if (nparts * n_segs * seg_cap > budget_keys) seg_cap = std::max(min_cap, budget_keys / (nparts * n_segs) / 16 * 16);
    if (nparts * n_segs * seg_cap > budget_keys) {
        ok = false; // not enough memory for a useful pending buffer: use the direct path
        return KATGPU_OK;
    }
    seg_cap = std::min<uint64_t>(seg_cap, 0x7fffffffu / 16 * 16);
    void *p = nullptr;
    cudaError_t e = pool_alloc(c, nparts * n_segs * seg_cap * 8, &p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        ok = false;
        return KATGPU_OK;
    }
    t->pend_keys = (uint64_t *)p;
    t->pend_bytes = nparts * n_segs * seg_cap * 8;
    // small per-job buffers also come from the block cache: cudaMalloc/cudaFree synchronise the
    // device and were measured at up to 45 ms per job when HBM is nearly full
    t->pend_seglen_bytes = nparts * n_segs * 4;
    CU(pool_alloc(c, t->pend_seglen_bytes, &p));
    t->pend_seglen = (uint32_t *)p;
    CU(pool_alloc(c, kHllM * 4, &p));
    t->pend_hll = (uint32_t *)p;
    CU(pool_alloc(c, kMaxParts * 4, &p));
    t->pend_done = (uint32_t *)p;
    CU(cudaMemsetAsync(t->pend_seglen, 0, nparts * n_segs * 4, c->stream));
    CU(cudaMemsetAsync(t->pend_hll, 0, kHllM * 4, c->stream));
    t->pend_Pb = Pb;
    t->pend_segs = (uint32_t)n_segs;
    t->pend_seg_cap = (uint32_t)seg_cap;
    t->pend_fill = 0;
    t->pend_bound = 0;
    t->tiled = tiled;
    if (tiled) {
        // bps_bin1 CTAs of bin_bases_kernel per SM (two: <= 111 KB each), 16.5 KB of it fixed
        // (one CTA per SM in the wide shape: ~200 KB, 21 KB of it fixed)
        const size_t ring1_bytes = KATGPU_BIN1_WIDE ? ((size_t)200 << 10) - 21504 : ((c->bps_bin1 >= 3 ? (size_t)74 : (size_t)111) << 10) - 16896;
        t->ring1 = ring_slots((uint32_t)nparts, kBin1Tile, ring1_bytes, kBin1Granule);
        // side list: keys that leave the fast path (~1 % at load 0.3, ~5 % at 0.55)
        t->side_cap = std::min<uint64_t>(std::max<uint64_t>(hint / 8, 4u << 20), 512u << 20);
        if (const char *e = getenv("KATGPU_SIDE_CAP")) t->side_cap = std::max<uint64_t>(16, strtoull(e, nullptr, 10)); // tests: force deferred tiles
        t->side_bytes = t->side_cap * 8;
        if (pool_alloc(c, t->side_bytes, &p) != cudaSuccess) {
            cudaGetLastError();
            t->side_cap = 1u << 20;
            t->side_bytes = t->side_cap * 8;
            CU(pool_alloc(c, t->side_bytes, &p));
        }
        t->side_keys = (uint64_t *)p;
        CU(pool_alloc(c, 64, &p));
        t->side_ctr = (unsigned long long *)p;
        CU(cudaMemsetAsync(t->side_ctr, 0, 64, c->stream));
        CU(pool_alloc(c, ((size_t)1 << 20) * 4, &p));
        t->side_deferred = (uint32_t *)p;
    }
    return KATGPU_OK;
}

bool dist_can_tile(const katgpu_table *t) { return want_tiled(t); }

static double hll_estimate(const uint32_t *reg, int m) {
    double z = 0;
    int zeros = 0;
    for (int i = 0; i < m; i++) {
        z += std::ldexp(1.0, -(int)reg[i]);
        zeros += reg[i] == 0;
    }
    const double alpha = 0.7213 / (1.0 + 1.079 / m);
    double e = alpha * m * (double)m / z;
    if (e <= 2.5 * m && zeros > 0) e = m * std::log((double)m / zeros);
    return e;
}

static katgpu_status admit(katgpu_table *t, TableCounters &ctr, uint64_t want, uint64_t &granted);

// Insert `n` hashed keys held in device memory with the general random-access kernel.
static katgpu_status insert_hashed_direct(katgpu_table *t, TableCounters &ctr, const uint64_t *hs, uint64_t n) {
    katgpu_ctx *c = t->ctx;
    uint64_t done = 0;
    while (done < n) {
        uint64_t granted = 0;
        KT(admit(t, ctr, n - done, granted));
        const int grid = (int)std::min<uint64_t>((granted + 256 * kUnroll - 1) / (256 * kUnroll), (uint64_t)c->sm_count * 8);
        {
            ProfScope ps(c, "insert", c->stream);
            TableGeom gl = t->g;
            set_add_bound(gl, granted);
            launch_insert_hashed(gl, hs + done, granted, grid, c->stream);
        }
        c->launches++;
        CU(cudaGetLastError());
        done += granted;
        KT(settle(t, ctr));
    }
    return KATGPU_OK;
}

SideList side_list_of(const katgpu_table *t) {
    SideList S{};
    S.keys = t->side_keys;
    S.n = t->side_ctr;
    S.reserved = t->side_ctr + 1;
    S.cap = t->side_cap;
    S.deferred = t->side_deferred;
    S.n_deferred = reinterpret_cast<uint32_t *>(t->side_ctr + 2);
    S.deferred_cap = 1u << 20;
    return S;
}

// The tiled sweep (tile_kernels.cu): for groups of slices, bin the slice's keys by table tile
// (bin_keys_kernel), then update every tile in shared memory (tile_insert_kernel); finally the
// side list goes through the general kernel.  Called with the table already sized.
static katgpu_status sweep_tiled(katgpu_table *t, TableCounters &ctr) {
    TraceScope trace__("    sweep_tiled");
    katgpu_ctx *c = t->ctx;
    const uint32_t F1 = 1u << t->pend_Pb, n_segs = t->pend_segs;
    const int total_bits = (int)t->g.Lb - kTileBucketsLog2; // table tiles = 2^total_bits
    const int Pb2 = total_bits - (int)t->pend_Pb;
    TableGeom gchk = t->g;
    set_add_bound(gchk, ~0ULL); // rare general-path inserts inside the kernels: always the checked add
    SideList side = side_list_of(t);
    bool general = Pb2 < 1 || Pb2 > 10; // (a table grown past 2^32 slots after binning started)
    if (const char *e = getenv("KATGPU_TILED_GENERAL")) general = general || atoi(e) != 0;
    // sizes of the level-2 (per-tile) segments for one group of slices; smaller groups when the
    // allocation fails, the general path when even one slice does not fit
    uint32_t F2 = 0, spg = 0, parts = 1;
    uint64_t cap2 = 0;
    uint64_t *keys2 = nullptr;
    uint32_t *len2 = nullptr;
    size_t bytes2 = 0, len2_bytes = 0;
    if (!general) {
        F2 = 1u << Pb2;
        // keys a slice can hold at most: sizing from the capacity (not from the current fill) gives
        // every sweep of a job — and of the next job — the same buffer sizes, so the block cache hits
        const double slice_keys = std::max(64.0, (double)t->pend_seg_cap * (double)n_segs);
        size_t free_b = 0, total_b = 0;
        CU(cudaMemGetInfo(&free_b, &total_b));
        double budget = std::min(0.80 * (double)(free_b + c->pooled_bytes), 24.0 * 1024 * 1024 * 1024);
        if (const char *e = getenv("KATGPU_TILED_GROUP_MB")) budget = std::max(1.0, atof(e)) * 1024 * 1024;
        const uint32_t want_ctas = (uint32_t)c->sm_count; // bin_keys_kernel: one 1024-thread CTA per SM
        for (;;) {
            spg = (uint32_t)std::max(1.0, std::min((double)F1, budget / (slice_keys * 8.0 * 1.3)));
            parts = std::min<uint32_t>(std::max<uint32_t>(1, std::min<uint32_t>(8, (want_ctas + spg - 1) / spg)), n_segs);
            // bin_keys_kernel keeps ONE CTA per SM, all of equal work: a launch of ns * parts CTAs takes
            // ceil(ns * parts / SMs) rounds whether or not the last one is full (124 slices x 2 shares = 248 CTAs on 148
            // SMs: two rounds, a third of the second empty; 9 rounds for the 512 slices where 7 would do).  Among the
            // group sizes the budget allows (down to half of it) and 1..8 shares per slice, take the pair that spends
            // the fewest CTA slots over the whole sweep; ties go to fewer shares, then to larger groups.
            if (!getenv("KATGPU_TILED_NO_WAVE_FIT")) {
                const uint32_t spg_max = spg, sms = want_ctas;
                double best = 1e300;
                for (uint32_t pt = 1; pt <= std::min<uint32_t>(8, n_segs); pt++)
                    for (uint32_t sg = spg_max; sg >= std::max<uint32_t>(1, spg_max / 2); sg--) {
                        const uint32_t full = F1 / sg, rest = F1 % sg;
                        const uint64_t rounds = (uint64_t)full * ((sg * pt + sms - 1) / sms) + ((uint64_t)rest * pt + sms - 1) / sms;
                        const double cost = (double)rounds / pt * (1.0 + 0.03 * (pt - 1)) // a CTA's work goes as 1 / shares (+ its fixed part)
                                            + 0.02 * (double)(full + (rest != 0));        // per-launch overhead, in rounds
                        if (cost < best - 1e-9) best = cost, spg = sg, parts = pt;
                    }
            }
            for (;;) {
                const double mean2 = slice_keys / ((double)F2 * parts);
                cap2 = ((uint64_t)(mean2 + 6.0 * std::sqrt(mean2) + 32.0) + 15) / 16 * 16;
                if ((double)spg * F2 * parts * cap2 * 8.0 <= budget || spg == 1) break;
                spg = std::max<uint32_t>(1, spg * 3 / 4);
            }
            bytes2 = (size_t)spg * F2 * parts * cap2 * 8;
            len2_bytes = (size_t)spg * F2 * parts * 4;
            void *p = nullptr, *q = nullptr;
            if (pool_alloc(c, bytes2, &p) == cudaSuccess) {
                if (pool_alloc(c, len2_bytes, &q) == cudaSuccess) {
                    keys2 = (uint64_t *)p;
                    len2 = (uint32_t *)q;
                    break;
                }
                pool_free(c, p, bytes2);
            }
            cudaGetLastError();
            if (spg == 1) { // not even one slice's tile segments fit
                general = true;
                break;
            }
            budget *= 0.5;
        }
    }
    if (general) { // insert the level-1 segments as they are, the general way
        TileParams P{};
        P.g = gchk;
        P.keys = t->pend_keys;
        P.len = t->pend_seglen;
        P.cap = t->pend_seg_cap;
        P.parts = 1;
        P.n_tiles = F1 * n_segs;
        P.side = side;
        {
            ProfScope ps(c, "insert", c->stream);
            launch_insert_deferred(P, nullptr, P.n_tiles, c->stream);
        }
        c->launches++;
        CU(cudaGetLastError());
    } else {
        // keys binned between two flushes of bin_keys_kernel and its ring (one CTA per SM, <= 200 KB)
        // keys bin_keys_kernel bins between two flushes: as many chunks as its ring (one CTA per SM) can take
        // without cutting into the 4-sigma head room
        const size_t ring2_bytes = (196u << 10) - kBin2StageBytes;
        uint32_t cpf = (uint32_t)kBin2MaxCpf * 2; // in pieces of kBin2Piece keys
        for (; cpf > 1; cpf--) {
            const double mean = (double)(kBin2Piece * cpf) / F2;
            if (((double)kBin2Granule + mean + 3.5 * std::sqrt(mean) + 1.0) * F2 * 8.0 + F2 * 8.0 + 16.0 <= (double)ring2_bytes) break; // (ring_slots then takes what fits: >= 3.5 sigma)
        }
        const uint32_t tile2 = kBin2Piece * cpf;
        const uint32_t ring2 = ring_slots(F2, tile2, ring2_bytes, kBin2Granule);
        if (getenv("KATGPU_TRACE"))
            fprintf(stderr, "[katgpu] sweep_tiled: F1=%u F2=%u slices/group=%u parts=%u cap2=%llu ring1=%u ring2=%u tile2=%u level-2 buffer %.1f MB side cap %llu\n",
                    F1, F2, spg, parts, (unsigned long long)cap2, t->ring1, ring2, tile2, bytes2 / 1048576.0, (unsigned long long)t->side_cap);
        katgpu_status st = KATGPU_OK;
        TraceScope trace_k("      level-2 binning + tiles");
        for (uint32_t s0 = 0; s0 < F1 && st == KATGPU_OK; s0 += spg) {
            const uint32_t ns = std::min(spg, F1 - s0);
            cudaMemsetAsync(len2, 0, (size_t)ns * F2 * parts * 4, c->stream);
            BinKeysParams B{};
            B.g = gchk;
            B.src_keys = t->pend_keys;
            B.src_len = t->pend_seglen;
            B.src_segs = n_segs;
            B.src_cap = t->pend_seg_cap;
            B.first_slice = s0;
            B.parts = parts;
            B.tile_keys = tile2;
            B.cpf = cpf;
            B.dst.keys = keys2;
            B.dst.len = len2;
            B.dst.cap = (uint32_t)cap2;
            B.dst.stride = parts;
            B.dst.shift = t->g.b - t->pend_Pb - Pb2;
            B.dst.mask = F2 - 1;
            B.dst.F = F2;
            B.dst.R = ring2;
            B.side = side;
            {
                ProfScope ps(c, "bin2", c->stream);
                launch_bin_keys(B, (int)(ns * parts), c->stream);
            }
            c->launches++;
            TileParams P{};
            P.g = gchk;
            P.keys = keys2;
            P.len = len2;
            P.cap = (uint32_t)cap2;
            P.parts = parts;
            P.first_tile = s0 * F2;
            P.n_tiles = ns * F2;
            P.side = side;
            {
                ProfScope ps(c, "tile", c->stream);
                launch_tile_insert(P, (int)std::min<uint32_t>(P.n_tiles, (uint32_t)(c->sm_count * c->bps_tile)), c->stream);
            }
            c->launches++;
            // tiles refused for lack of side-list room (pathological inputs): general kernel, then forget them
            {
                ProfScope ps(c, "insert", c->stream);
                launch_insert_deferred(P, side.deferred, side.deferred_cap, c->stream);
            }
            c->launches++;
            cudaMemsetAsync(side.n_deferred, 0, 4, c->stream);
            if (cudaGetLastError() != cudaSuccess) st = fail(KATGPU_ERR_CUDA, "sweep_tiled: launch failed");
        }
        cudaStreamSynchronize(c->stream);
        pool_free(c, keys2, bytes2);
        pool_free(c, len2, len2_bytes);
        KT(st);
    }
    // keys that left the fast path: general insertion (may grow the table)
    TraceScope trace_s("      side list");
    unsigned long long side_n = 0;
    CU(cudaMemcpyAsync(&side_n, t->side_ctr, 8, cudaMemcpyDeviceToHost, c->stream));
    KT(settle(t, ctr)); // synchronises; spills of the general inserts above
    side_n = std::min<unsigned long long>(side_n, t->side_cap);
    if (getenv("KATGPU_TRACE")) fprintf(stderr, "[katgpu] sweep_tiled: %llu keys took the side list\n", side_n);
    if (side_n) KT(insert_hashed_direct(t, ctr, t->side_keys, side_n));
    CU(cudaMemsetAsync(t->side_ctr, 0, 64, c->stream));
    CU(cudaMemsetAsync(t->pend_seglen, 0, ((size_t)1 << t->pend_Pb) * t->pend_segs * 4, c->stream));
    CU(cudaMemsetAsync(t->pend_hll, 0, kHllM * 4, c->stream));
    t->pend_fill = 0;
    t->pend_tiles = 0;
    t->pend_total_tiles = 0;
    t->pend_bound = 0;
    return settle(t, ctr);
}

// Insert every pending key, slice by slice (sweep_kernel), after sizing the table
// for the estimated number of new distinct keys.
katgpu_status sweep_pending(katgpu_table *t, TableCounters &ctr) {
    TraceScope trace__("  sweep_pending");
    if (!t->pend_keys || t->pend_bound == 0) return KATGPU_OK;
    katgpu_ctx *c = t->ctx;
    uint32_t h_hll[kHllM];
    CU(cudaMemcpyAsync(h_hll, t->pend_hll, sizeof h_hll, cudaMemcpyDeviceToHost, c->stream));
    KT(read_counters(t, ctr)); // synchronises
    const double est_new = std::min((double)t->pend_bound, hll_estimate(h_hll, kHllM) * (double)(1u << kHllSampleShift) * 1.25 + 4096.0);
    if ((int)t->g.Lb < (int)t->g.b) {
        const double need = (double)ctr.n_distinct + est_new;
        if (need > kMaxLoad * (double)t->n_slots) {
            if (t->allow_grow) {
                int lb = t->g.Lb;
                while (lb < (int)t->g.b && need > 0.70 * (double)((uint64_t)kSlotsPerBucket << lb)) lb++;
                KT(grow_table(t, ctr, lb));
                KT(read_counters(t, ctr));
            } else if ((double)ctr.n_distinct + est_new / 1.25 * 0.8 > (double)t->n_slots)
                return fail(KATGPU_ERR_HASH_FULL, "Hash full");
        }
    }
    if (t->tiled) return sweep_tiled(t, ctr);
    SweepParams P{};
    P.g = t->g;
    // Keys that can target one slot in this sweep: a slice never holds more than its segments'
    // capacity, and no more than were fed.  (pend_fill is a fill MODEL for uniform hashes — on
    // skewed input one slice really fills its segments, so it is not a bound.)
    set_add_bound(P.g, std::min<uint64_t>(t->pend_bound, (uint64_t)t->pend_seg_cap * t->pend_segs));
    P.keys = t->pend_keys;
    P.seglen = t->pend_seglen;
    P.seg_cap = t->pend_seg_cap;
    P.n_segs = t->pend_segs;
    P.part_lo = 0;
    P.part_hi = 1u << t->pend_Pb;
    // every sweep CTA must be resident at once (they synchronise slice by slice)
    const uint32_t resident = (uint32_t)(c->sm_count * c->bps_sweep);
    P.split = std::max<uint32_t>(1, resident / t->pend_segs);
    P.window = 1;
    if (const char *e = getenv("KATGPU_SWEEP_WINDOW")) P.window = (uint32_t)std::max(0, atoi(e));
    if (P.split * t->pend_segs > resident) P.window = 0; // cannot guarantee co-residency: no barrier
    P.slice_done = t->pend_done;
    P.slice_slots = t->n_slots >> t->pend_Pb;
    P.prefetch = 1;
    P.experiment = 0;
    if (const char *e = getenv("KATGPU_SWEEP_EXPERIMENT")) P.experiment = (uint32_t)atoi(e); // measurement only
    if (const char *e = getenv("KATGPU_SWEEP_PREFETCH")) P.prefetch = (uint32_t)atoi(e);
    CU(cudaMemsetAsync(t->pend_done, 0, kMaxParts * 4, c->stream));
    {
        ProfScope ps(c, "sweep", c->stream);
        launch_sweep(P, c->stream);
    }
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemsetAsync(t->pend_seglen, 0, ((size_t)1 << t->pend_Pb) * t->pend_segs * 4, c->stream));
    CU(cudaMemsetAsync(t->pend_hll, 0, kHllM * 4, c->stream));
    t->pend_fill = 0;
    t->pend_tiles = 0;
    t->pend_total_tiles = 0;
    t->pend_bound = 0;
    return settle(t, ctr);
}

static katgpu_status feed_partitioned(katgpu_table *t, const uint8_t *dev, uint64_t n_valid, uint64_t n_starts,
                                      TableCounters &ctr, const uint64_t *in_keys = nullptr) {
    katgpu_ctx *c = t->ctx;
    // keys one CTA tile sends to one slice: mean m (every position a k-mer) + up to 3 padding keys;
    // after T tiles the fullest segment holds about T*(m+2) + 4*sqrt(T*m) (a segment that still
    // overflows falls back to direct insertion inside the kernel, so this only has to be a good guess)
    const uint64_t tile_pos = t->tiled ? (in_keys ? kBinTile : kBin1Tile) : kTile;
    const double m = (double)tile_pos / (double)(1u << t->pend_Pb), pad = t->tiled ? 0.0 : 2.0;
    auto fill_after = [&](uint64_t tiles) { return (double)tiles * (m + pad) + 4.0 * std::sqrt((double)tiles * m); };
    uint64_t done = 0;
    while (done < n_starts) {
        uint64_t cta_tiles = 0; // largest T with fill_after(pend_tiles + T) <= seg_cap
        {
            const double cap = (double)t->pend_seg_cap;
            uint64_t lo = 0, hi = (uint64_t)(cap / (m + pad)) + 1;
            while (lo < hi) {
                const uint64_t mid = (lo + hi + 1) / 2;
                if (fill_after(mid) <= cap) lo = mid; else hi = mid - 1;
            }
            cta_tiles = lo > t->pend_tiles ? lo - t->pend_tiles : 0;
        }
        if (cta_tiles == 0) {
            if (t->pend_tiles == 0) return fail(KATGPU_ERR_INVALID, "internal: pending segments smaller than one tile");
            KT(sweep_pending(t, ctr));
            continue;
        }
        const uint64_t max_starts = cta_tiles * t->pend_segs * tile_pos;
        const uint64_t take = std::min(n_starts - done, max_starts);
        const uint64_t n_tiles = (take + tile_pos - 1) / tile_pos;
        const uint32_t cta_rot = (uint32_t)(t->pend_total_tiles % t->pend_segs);
        if (t->tiled) {
            BinDst D{};
            D.keys = t->pend_keys;
            D.len = t->pend_seglen;
            D.cap = t->pend_seg_cap;
            D.stride = t->pend_segs; // segment (slice d, CTA j) = d * n_segs + j
            D.shift = t->g.b - t->pend_Pb;
            D.mask = (1u << t->pend_Pb) - 1;
            D.F = 1u << t->pend_Pb;
            D.R = t->ring1;
            TableGeom gchk = t->g;
            set_add_bound(gchk, ~0ULL);
            if (in_keys) {
                BinKeysParams B{};
                B.g = gchk;
                B.in_keys = in_keys + done;
                B.n_keys = take;
                B.cta_rot = cta_rot;
                B.tile_keys = kBinTile;
                B.dst = D;
                B.side = side_list_of(t);
                B.hll = t->pend_hll;
                ProfScope ps(c, "bin1", c->stream);
                launch_bin_keys(B, (int)t->pend_segs, c->stream);
            } else {
                BinBasesParams B{};
                B.g = gchk;
                B.bases = dev + done;
                B.n_valid = n_valid - done;
                B.n_starts = take;
                B.n_tiles = (int)n_tiles;
                B.tma_ok = (reinterpret_cast<uintptr_t>(dev + done) & 15) == 0;
                B.cta_rot = cta_rot;
                B.dst = D;
                B.side = side_list_of(t);
                B.hll = t->pend_hll;
                ProfScope ps(c, "bin1", c->stream);
                launch_bin_bases(B, (int)t->pend_segs, c->stream);
            }
        } else {
            PartitionParams P{};
            P.g = t->g;
            P.bases = in_keys ? nullptr : dev + done;
            P.n_valid = n_valid - done;
            P.n_starts = take;
            P.in_keys = in_keys ? in_keys + done : nullptr;
            P.n_keys = take;
            P.n_tiles = (int)n_tiles;
            P.tma_ok = !in_keys && (reinterpret_cast<uintptr_t>(dev + done) & 15) == 0;
            P.cta_rot = cta_rot;
            P.Pb = t->pend_Pb;
            P.seg_cap = t->pend_seg_cap;
            P.keys = t->pend_keys;
            P.seglen = t->pend_seglen;
            P.hll = t->pend_hll;
            ProfScope ps(c, "partition", c->stream);
            launch_partition(P, (int)t->pend_segs, c->stream);
        }
        c->launches++;
        CU(cudaGetLastError());
        t->pend_total_tiles += n_tiles;
        t->pend_tiles = (t->pend_total_tiles + t->pend_segs - 1) / t->pend_segs;
        t->pend_fill = fill_after(t->pend_tiles);
        t->pend_bound += take;
        done += take;
    }
    return KATGPU_OK;
}

// Count a device-resident stream, in admission-controlled sub-launches.
// Grow (when allowed) until `n` more keys fit even if every one of them is new.
katgpu_status dist_make_room(katgpu_table *t, TableCounters &ctr, uint64_t n) {
    uint64_t granted = 0;
    for (int i = 0; i < 64; i++) {
        KT(admit(t, ctr, n, granted));
        if (granted >= n || !t->allow_grow) return KATGPU_OK; // without growth an overflow surfaces as "Hash full" in settle()
        KT(grow_table(t, ctr));
        KT(read_counters(t, ctr));
    }
    return KATGPU_OK;
}

katgpu_status feed_device_range(katgpu_table *t, const uint8_t *dev, uint64_t n_valid, uint64_t n_starts,
                                       TableCounters &ctr, uint64_t kmers_hint) {
    if (t->wide) return wide_feed_range(t, dev, n_valid, n_starts);
    // a k-mer start needs k bases
    if (n_valid < (uint64_t)t->g.k) return KATGPU_OK;
    n_starts = std::min(n_starts, n_valid - t->g.k + 1);
    if (use_partition(t, kmers_hint)) {
        bool ok = false;
        KT(pend_alloc(t, kmers_hint, ok));
        if (ok) return feed_partitioned(t, dev, n_valid, n_starts, ctr);
    }
    KT(sweep_pending(t, ctr)); // keep insertion order irrelevant but the table consistent
    uint64_t done = 0;
    while (done < n_starts) {
        uint64_t granted = 0;
        KT(admit(t, ctr, n_starts - done, granted));
        if (granted < n_starts - done) granted = std::max<uint64_t>(granted / kTile * kTile, kTile); // keep TMA alignment
        granted = std::min(granted, n_starts - done);
        KT(launch_extract_insert(t, dev + done, n_valid - done, granted));
        done += granted;
        if (done < n_starts || true) KT(settle(t, ctr));
    }
    return KATGPU_OK;
}

extern "C" katgpu_status katgpu_count_feed_device(katgpu_table *t, const void *bases_dev, size_t n) {
    TraceScope trace__("katgpu_count_feed_device");
    if (!t || (!bases_dev && n)) return fail(KATGPU_ERR_INVALID, "katgpu_count_feed_device: NULL argument");
    if (t->ended) return fail(KATGPU_ERR_STATE, "katgpu_count_feed_device: table already finished");
    CU(cudaSetDevice(t->ctx->device));
    if (t->ctx->dist) return dist_feed_device(t, (const uint8_t *)bases_dev, n);
    TableCounters ctr;
    KT(read_counters(t, ctr));
    return feed_device_range(t, (const uint8_t *)bases_dev, n, n, ctr, n);
}

// Device staging buffers of the host feed: `ring` buffers of at least `chunk` + halo bytes.
katgpu_status ensure_feed_staging(katgpu_ctx *c, size_t chunk, int ring) {
    if (c->chunk_cap >= chunk + kHalo && c->chunk_n >= ring) return KATGPU_OK;
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->copy_stream));
    for (int i = 0; i < kRing; i++) {
        if (c->chunk[i]) cudaFree(c->chunk[i]);
        c->chunk[i] = nullptr;
    }
    const size_t cap = std::max(c->chunk_cap, chunk + kHalo);
    ring = std::max(ring, c->chunk_n);
    for (int i = 0; i < ring; i++) {
        void *p = nullptr;
        CU(cudaMalloc(&p, cap + 64));
        c->chunk[i] = (uint8_t *)p;
    }
    c->chunk_cap = cap;
    c->chunk_n = ring;
    return KATGPU_OK;
}

extern "C" katgpu_status katgpu_count_feed(katgpu_table *t, const char *bases, size_t n) {
    TraceScope trace__("katgpu_count_feed");
    if (!t || (!bases && n)) return fail(KATGPU_ERR_INVALID, "katgpu_count_feed: NULL argument");
    if (t->ended) return fail(KATGPU_ERR_STATE, "katgpu_count_feed: table already finished");
    katgpu_ctx *c = t->ctx;
    CU(cudaSetDevice(c->device));
    if (n == 0) return KATGPU_OK;
    if (c->dist) return dist_feed_host(t, bases, n);
    const size_t chunk = std::min(kChunkBytes, (n + 4095) / 4096 * 4096);
    const size_t n_chunks = (n + chunk - 1) / chunk;
    const int ring = (int)std::min<size_t>(kRing, std::max<size_t>(2, n_chunks));
    const size_t halo = t->wide ? (size_t)kWideHalo : (size_t)kHalo; // >= k-1 bases behind the last start of a chunk
    KT(ensure_feed_staging(c, chunk + (halo - kHalo), ring));
    TableCounters ctr;
    KT(read_counters(t, ctr));
    // chunk i holds host bytes [i*chunk, (i+1)*chunk + halo): k-mers starting in
    // the first `chunk` positions are counted from it, the halo feeds their tails.
    auto issue_copy = [&](size_t i) -> katgpu_status {
        const int b = (int)(i % ring);
        const size_t off = i * chunk;
        const size_t len = std::min(n - off, chunk + halo);
        CU(cudaStreamWaitEvent(c->copy_stream, c->chunk_free[b], 0));
        CU(cudaMemcpyAsync(c->chunk[b], bases + off, len, cudaMemcpyHostToDevice, c->copy_stream));
        CU(cudaEventRecord(c->chunk_ready[b], c->copy_stream));
        return KATGPU_OK;
    };
    // the free events start out "complete": record them once on an idle stream
    for (int b = 0; b < ring; b++) CU(cudaEventRecord(c->chunk_free[b], c->stream));
    // With the partitioned path and a long feed, insert the first half of the keys while the
    // second half is still crossing PCIe: the copy engine keeps filling the ring during the sweep.
    // (KATGPU_MID_SWEEPS: how many; default 3 = after each quarter: e2e 269 / 225 / 217 / 238 ms for 1 / 2 / 3 / 4 with a 2^31-slot table)
    int mid_sweeps = 3;
    if (const char *e = getenv("KATGPU_MID_SWEEPS")) mid_sweeps = std::max(0, std::min(atoi(e), 8));
    const bool mid_ok = n_chunks >= 16 && mid_sweeps > 0 && use_partition(t, n);
    size_t issued = 0;
    for (size_t i = 0; i < n_chunks; i++) {
        while (issued < n_chunks && issued < i + (size_t)ring) KT(issue_copy(issued++)); // run ahead of the kernels
        const int b = (int)(i % ring);
        const size_t off = i * chunk;
        const size_t len = std::min(n - off, chunk + halo);
        const size_t starts = std::min(n - off, chunk);
        CU(cudaStreamWaitEvent(c->stream, c->chunk_ready[b], 0));
        KT(feed_device_range(t, c->chunk[b], len, starts, ctr, n));
        CU(cudaEventRecord(c->chunk_free[b], c->stream));
        if (mid_ok && i + 1 < n_chunks && ((i + 1) * (size_t)(mid_sweeps + 1)) / n_chunks != (i * (size_t)(mid_sweeps + 1)) / n_chunks) KT(sweep_pending(t, ctr));
    }
    // `bases` belongs to the caller again when this returns: every host->device copy has completed
    // (the kernels consuming the staged chunks may still be running — they only touch device memory).
    CU(cudaStreamSynchronize(c->copy_stream));
    return KATGPU_OK;
}

extern "C" katgpu_status katgpu_count_reserve(katgpu_table *t, uint64_t expected_kmers) {
    if (!t) return fail(KATGPU_ERR_INVALID, "katgpu_count_reserve: NULL table");
    if (t->ended) return fail(KATGPU_ERR_STATE, "katgpu_count_reserve: table already finished");
    CU(cudaSetDevice(t->ctx->device));
    if (use_partition(t, expected_kmers)) {
        bool ok = false;
        KT(pend_alloc(t, expected_kmers, ok));
    }
    return KATGPU_OK;
}

extern "C" katgpu_status katgpu_count_end(katgpu_table *t, uint64_t *n_distinct, uint64_t *n_total) {
    TraceScope trace__("katgpu_count_end");
    if (!t) return fail(KATGPU_ERR_INVALID, "katgpu_count_end: NULL table");
    CU(cudaSetDevice(t->ctx->device));
    if (t->wide) {
        if (!t->ended) KT(wide_totals(t, t->wide_distinct, t->wide_total));
        t->ended = true;
        if (n_distinct) *n_distinct = t->wide_distinct;
        if (n_total) *n_total = t->wide_total;
        return KATGPU_OK;
    }
    TableCounters ctr;
    KT(read_counters(t, ctr));
    if (t->ctx->dist && !t->ended) KT(dist_count_end(t, ctr)); // every rank's last route round has been binned here
    KT(sweep_pending(t, ctr));
    KT(settle(t, ctr));
    pend_release(t); // the pending-key buffers are only needed while counting
    uint64_t tot[2] = {ctr.n_distinct, ctr.n_total};
    if (t->ctx->dist) { // every k-mer lives on exactly one rank: the job's totals are the sums
        if (!t->ended) {
            katgpu_ctx *c = t->ctx;
            void *d = nullptr;
            CU(pool_alloc(c, 256, &d));
            CU(cudaMemcpyAsync(d, tot, 16, cudaMemcpyHostToDevice, c->stream));
            katgpu_status st = dist_allreduce_u64(c, d, 2);
            if (st == KATGPU_OK && cudaMemcpyAsync(tot, d, 16, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) st = fail(KATGPU_ERR_CUDA, "katgpu_count_end: copy failed");
            cudaStreamSynchronize(c->stream);
            pool_free(c, d, 256);
            KT(st);
            t->dist_global_distinct = tot[0];
            t->dist_global_total = tot[1];
        }
        tot[0] = t->dist_global_distinct;
        tot[1] = t->dist_global_total;
    }
    t->ended = true;
    if (n_distinct) *n_distinct = tot[0];
    if (n_total) *n_total = tot[1];
    return KATGPU_OK;
}

extern "C" katgpu_status katgpu_table_get_info(katgpu_table *t, katgpu_table_info *out) {
    if (!t || !out) return fail(KATGPU_ERR_INVALID, "katgpu_table_get_info: NULL argument");
    CU(cudaSetDevice(t->ctx->device));
    if (t->wide) {
        *out = katgpu_table_info{};
        out->k = (int)t->g.k;
        out->canonical = (int)t->g.canonical;
        out->slots = t->n_slots;
        out->bytes = wide_bytes(t);
        KT(wide_totals(t, out->n_distinct, out->n_total));
        out->count_bits = 64;
        out->grow_events = t->grow_events;
        return KATGPU_OK;
    }
    TableCounters ctr;
    KT(read_counters(t, ctr));
    out->k = t->g.k;
    out->canonical = t->g.canonical;
    out->slots = t->n_slots;
    out->bytes = t->n_slots * 8 + ((uint64_t)kOvfCap + 1) * 16 + kSpillCap * 16 + sizeof(TableCounters);
    out->n_distinct = ctr.n_distinct;
    out->n_total = ctr.n_total;
    out->count_bits = t->g.cbits;
    out->grow_events = t->grow_events;
    out->overflow_keys = ctr.ovf_n;
    return KATGPU_OK;
}

// ---- multi-GPU plumbing ----
extern "C" katgpu_status katgpu_route_device(katgpu_table *t, const void *bases_dev, size_t n, int n_ranks,
                                             void *const *out_dev, uint64_t cap_keys, uint64_t *counts_out) {
    if (!t || !out_dev || !counts_out || (!bases_dev && n)) return fail(KATGPU_ERR_INVALID, "katgpu_route_device: NULL argument");
    if (n_ranks < 1 || n_ranks > kMaxRanks) return fail(KATGPU_ERR_INVALID, "katgpu_route_device: n_ranks=%d outside 1..%d", n_ranks, kMaxRanks);
    if (t->wide) return fail(KATGPU_ERR_INVALID, "katgpu_route_device: one-word keys only (k <= 32)");
    katgpu_ctx *c = t->ctx;
    CU(cudaSetDevice(c->device));
    for (int r = 0; r < n_ranks; r++) counts_out[r] = 0;
    if (n < (size_t)t->g.k) return KATGPU_OK;
    ExtractParams P{};
    P.g = t->g;
    P.bases = (const uint8_t *)bases_dev;
    P.n_valid = n;
    P.n_starts = n - t->g.k + 1;
    P.tile_counter = c->tile_counter;
    P.n_tiles = (int)((P.n_starts + kTile - 1) / kTile);
    P.tma_ok = (reinterpret_cast<uintptr_t>(bases_dev) & 15) == 0;
    P.n_ranks = (uint32_t)n_ranks;
    for (int r = 0; r < n_ranks; r++) P.out[r] = (uint64_t *)out_dev[r];
    P.cursors = c->route_cursors;
    P.out_cap = cap_keys;
    P.route_overflow = reinterpret_cast<unsigned int *>(c->route_cursors + kMaxRanks);
    CU(cudaMemsetAsync(c->tile_counter, 0, sizeof(unsigned int), c->stream));
    CU(cudaMemsetAsync(c->route_cursors, 0, 256, c->stream));
    const int grid = std::min(P.n_tiles, c->sm_count * c->bps[1]);
    {
        ProfScope ps(c, "route", c->stream);
        launch_extract(kModeRoute, P, grid, c->stream);
    }
    c->launches++;
    CU(cudaGetLastError());
    unsigned long long h[kMaxRanks + 1];
    CU(cudaMemcpyAsync(h, c->route_cursors, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int r = 0; r < n_ranks; r++) counts_out[r] = h[r];
    if ((unsigned int)h[kMaxRanks] != 0)
        return fail(KATGPU_ERR_INVALID, "katgpu_route_device: an owner bucket overflowed its capacity of %llu keys", (unsigned long long)cap_keys);
    return KATGPU_OK;
}

extern "C" katgpu_status katgpu_count_add_pairs(katgpu_table *t, const uint64_t *keys, const uint64_t *counts, size_t n) {
    if (!t || ((!keys || !counts) && n)) return fail(KATGPU_ERR_INVALID, "katgpu_count_add_pairs: NULL argument");
    if (t->ended) return fail(KATGPU_ERR_STATE, "katgpu_count_add_pairs: table already finished");
    if (t->wide) return fail(KATGPU_ERR_INVALID, "katgpu_count_add_pairs: k > 32 needs two words per key: katgpu_count_add_pairs_wide");
    katgpu_ctx *c = t->ctx;
    if (c->dist) return fail(KATGPU_ERR_STATE, "katgpu_count_add_pairs: not available on a multi-GPU context (load hashes on one GPU)");
    CU(cudaSetDevice(c->device));
    TableCounters ctr;
    KT(read_counters(t, ctr));
    KT(sweep_pending(t, ctr));
    const size_t kBatch = 1u << 24; // records per staging round (2 x 128 MiB on the device)
    uint64_t *d_keys = nullptr, *d_counts = nullptr;
    const size_t cap = std::min(n, kBatch);
    if (cap == 0) return KATGPU_OK;
    CU(cudaMalloc((void **)&d_keys, cap * 8));
    if (cudaMalloc((void **)&d_counts, cap * 8) != cudaSuccess) {
        cudaFree(d_keys);
        return fail(KATGPU_ERR_NOMEM, "katgpu_count_add_pairs: out of device memory");
    }
    katgpu_status st = KATGPU_OK;
    size_t done = 0;
    while (done < n && st == KATGPU_OK) {
        uint64_t granted = 0;
        st = admit(t, ctr, std::min(n - done, cap), granted);
        if (st != KATGPU_OK) break;
        cudaError_t e = cudaMemcpyAsync(d_keys, keys + done, granted * 8, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_counts, counts + done, granted * 8, cudaMemcpyHostToDevice, c->stream);
        if (e != cudaSuccess) {
            st = fail(KATGPU_ERR_CUDA, "katgpu_count_add_pairs: %s", cudaGetErrorString(e));
            break;
        }
        {
            ProfScope ps(c, "insert", c->stream);
            TableGeom gl = t->g;
            set_add_bound(gl, ~0ULL); // counts are arbitrary: always the checked (ATOM) add
            const int grid = (int)std::min<uint64_t>((granted + 255) / 256, (uint64_t)c->sm_count * 8);
            launch_add_pairs(gl, d_keys, d_counts, granted, grid, c->stream);
        }
        c->launches++;
        done += granted;
        st = settle(t, ctr); // synchronises: the staging buffers are free again
    }
    cudaFree(d_keys);
    cudaFree(d_counts);
    return st;
}

// (key, count) records with two words per key; works for every k (keys_hi is ignored — may be NULL — when k <= 32)
extern "C" katgpu_status katgpu_count_add_pairs_wide(katgpu_table *t, const uint64_t *keys_hi, const uint64_t *keys_lo,
                                                     const uint64_t *counts, size_t n) {
    if (!t || ((!keys_lo || !counts) && n)) return fail(KATGPU_ERR_INVALID, "katgpu_count_add_pairs_wide: NULL argument");
    if (!t->wide) return katgpu_count_add_pairs(t, keys_lo, counts, n);
    if (!keys_hi && n) return fail(KATGPU_ERR_INVALID, "katgpu_count_add_pairs_wide: NULL argument");
    if (t->ended) return fail(KATGPU_ERR_STATE, "katgpu_count_add_pairs_wide: table already finished");
    CU(cudaSetDevice(t->ctx->device));
    return wide_add_pairs(t, keys_hi, keys_lo, counts, n);
}

extern "C" katgpu_status katgpu_insert_hashed_device(katgpu_table *t, const void *hashed_dev, uint64_t n) {
    if (!t || (!hashed_dev && n)) return fail(KATGPU_ERR_INVALID, "katgpu_insert_hashed_device: NULL argument");
    if (t->wide) return fail(KATGPU_ERR_INVALID, "katgpu_insert_hashed_device: one-word keys only (k <= 32)");
    if (t->ended) return fail(KATGPU_ERR_STATE, "katgpu_insert_hashed_device: table already finished");
    katgpu_ctx *c = t->ctx;
    CU(cudaSetDevice(c->device));
    TableCounters ctr;
    KT(read_counters(t, ctr));
    const uint64_t *hs = (const uint64_t *)hashed_dev;
    if (use_partition(t, n)) { // partition the received keys by table slice; inserted at the next sweep
        bool ok = false;
        KT(pend_alloc(t, n, ok));
        if (ok) return feed_partitioned(t, nullptr, n, n, ctr, hs);
    }
    KT(sweep_pending(t, ctr));
    uint64_t done = 0;
    while (done < n) {
        uint64_t granted = 0;
        KT(admit(t, ctr, n - done, granted));
        const int grid = (int)std::min<uint64_t>((granted + 256 * kUnroll - 1) / (256 * kUnroll), (uint64_t)c->sm_count * 8);
        {
            ProfScope ps(c, "insert", c->stream);
            TableGeom gl = t->g;
            set_add_bound(gl, granted);
            launch_insert_hashed(gl, hs + done, granted, grid, c->stream);
        }
        c->launches++;
        CU(cudaGetLastError());
        done += granted;
        KT(settle(t, ctr));
    }
    return KATGPU_OK;
}

// ---- synthetic reads ----
__global__ void synth_kernel(SynthParams p, char *out) {
    const uint64_t per = (uint64_t)p.read_len + 1;
    const uint64_t total = p.n_reads * per;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = i / per;
        out[i] = synth_byte(p, p.first_read + r, (uint32_t)(i - r * per));
    }
}
static katgpu_status synth_params(uint64_t genome_len, uint64_t seed_g, uint64_t seed_r, uint64_t first_read,
                                  uint64_t n_reads, uint32_t read_len, double sub_rate, double n_rate, SynthParams &p) {
    if (read_len == 0 || genome_len < read_len || sub_rate < 0 || sub_rate > 1 || n_rate < 0 || n_rate > 1)
        return fail(KATGPU_ERR_INVALID, "katgpu_synth_reads: bad parameters");
    p.genome_len = genome_len;
    p.seed_g = seed_g;
    p.seed_r = seed_r;
    p.first_read = first_read;
    p.n_reads = n_reads;
    p.read_len = read_len;
    p.thr_sub = (uint32_t)(sub_rate * 16777216.0);
    p.thr_n = (uint32_t)(n_rate * 16777216.0);
    return KATGPU_OK;
}
extern "C" katgpu_status katgpu_synth_reads_device(katgpu_ctx *c, uint64_t genome_len, uint64_t seed_g, uint64_t seed_r,
                                                   uint64_t first_read, uint64_t n_reads, uint32_t read_len,
                                                   double sub_rate, double n_rate, void *out_dev) {
    if (!c || !out_dev) return fail(KATGPU_ERR_INVALID, "katgpu_synth_reads_device: NULL argument");
    SynthParams p;
    KT(synth_params(genome_len, seed_g, seed_r, first_read, n_reads, read_len, sub_rate, n_rate, p));
    CU(cudaSetDevice(c->device));
    if (n_reads) synth_kernel<<<c->sm_count * 16, 256, 0, c->stream>>>(p, (char *)out_dev);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    return KATGPU_OK;
}
extern "C" katgpu_status katgpu_synth_reads_host(uint64_t genome_len, uint64_t seed_g, uint64_t seed_r, uint64_t first_read,
                                                 uint64_t n_reads, uint32_t read_len, double sub_rate, double n_rate,
                                                 char *out) {
    if (!out) return fail(KATGPU_ERR_INVALID, "katgpu_synth_reads_host: NULL argument");
    SynthParams p;
    KT(synth_params(genome_len, seed_g, seed_r, first_read, n_reads, read_len, sub_rate, n_rate, p));
    const uint64_t per = (uint64_t)read_len + 1;
    for (uint64_t r = 0; r < n_reads; r++)
        for (uint32_t j = 0; j <= read_len; j++) out[r * per + j] = synth_byte(p, first_read + r, j);
    return KATGPU_OK;
}
