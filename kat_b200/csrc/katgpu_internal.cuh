// katgpu_internal.cuh — what the translation units of libkatgpu share: error plumbing, the
// context and table objects behind the opaque C-ABI handles, the device block cache, kernel
// timing scopes, and the entry points one file calls in another (katgpu_api.cu: lifetime, feed
// pipeline, walks; katgpu_dist.cu: multi-GPU route rounds and collectives).
#pragma once
#include "../../include/katgpu.h"
#include "count_kernels.cuh"
#include "partition_kernels.cuh"
#include "tile_kernels.cuh"
#include "reduce_kernels.cuh"
#include "parse_kernels.cuh"
#include "wide_kernels.cuh"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

katgpu_status katgpu_fail(katgpu_status st, const char *fmt, ...);
#define fail katgpu_fail
#define CU(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess)                                                                               \
            return fail(e__ == cudaErrorMemoryAllocation ? KATGPU_ERR_NOMEM : KATGPU_ERR_CUDA, "%s: %s (%s:%d)", \
                        #call, cudaGetErrorString(e__), __FILE__, __LINE__);                                  \
    } while (0)
#define KT(call)                          \
    do {                                  \
        katgpu_status s__ = (call);       \
        if (s__ != KATGPU_OK) return s__; \
    } while (0)

struct ProfSpan {
    std::string family;
    cudaEvent_t a, b;
};

constexpr int kRing = 24; // host-feed staging ring: lets the copy engine run ahead of a mid-feed sweep
struct DistState;         // katgpu_dist.cu
struct WideTable;         // katgpu_wide.cu

struct katgpu_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    int bps[2] = {1, 1}; // resident CTAs per SM of the extract kernels
    int bps_part = 1;    // ... of the partition kernel
    int bps_sweep = 1;   // ... of the sweep kernel
    int bps_bin1 = 2, bps_tile = 1; // ... of the tiled path's bin_bases / tile_insert kernels
    unsigned int *tile_counter = nullptr; // device
    unsigned long long *route_cursors = nullptr; // device [kMaxRanks] + overflow flag after
    // host feed staging
    uint8_t *chunk[kRing] = {};
    size_t chunk_cap = 0;
    int chunk_n = 0;          // staging buffers currently allocated (2 .. kRing)
    cudaEvent_t chunk_free[kRing] = {}, chunk_ready[kRing] = {};
    katgpu::TableCounters *h_ctr = nullptr; // pinned mirror
    uint64_t launches = 0;
    bool profiling = false;
    std::vector<ProfSpan> spans;
    std::map<std::string, std::pair<double, uint64_t>> prof;
    // freed device blocks kept for reuse: a `kat hist` job after the first one does not pay
    // cudaMalloc/cudaFree of tens of GB again (the reference likewise keeps its mmap'd arena)
    std::multimap<size_t, void *> pool;
    std::map<void *, size_t> block_size; // true size of every block handed out by pool_alloc
    size_t pooled_bytes = 0;
    DistState *dist = nullptr; // set by katgpu_dist_init: this context is one rank of a multi-GPU job
    // raw-file feed (katgpu_count_feed_fastq): two device buffers for file bytes, pinned bounce
    // buffers filled by several host threads, newline index scratch
    uint8_t *raw[2] = {};
    size_t raw_cap = 0;
    cudaEvent_t raw_free[2] = {};
    int raw_next = 0;
    uint32_t *nl_buf = nullptr, *nl_tiles_buf = nullptr, *parse_flags = nullptr; // device
    size_t nl_cap = 0, nl_tiles_cap = 0;
    uint8_t *bounce[8] = {};
    cudaEvent_t bounce_free[8] = {};
    size_t bounce_bytes = 0;
};

struct katgpu_table {
    katgpu_ctx *ctx = nullptr;
    katgpu::TableGeom g{};
    uint64_t n_slots = 0;
    bool allow_grow = true, ended = false;
    uint32_t grow_events = 0;
    int count_bits_cap = 0;
    // keys partitioned by table slice, waiting for the next sweep (partitioned path)
    uint64_t *pend_keys = nullptr;
    uint32_t *pend_seglen = nullptr, *pend_hll = nullptr, *pend_done = nullptr;
    uint32_t pend_Pb = 0, pend_segs = 0, pend_seg_cap = 0;
    double pend_fill = 0;      // conservative estimate of the fullest segment (keys)
    uint64_t pend_tiles = 0;   // tiles the busiest CTA has partitioned since the last sweep
    uint64_t pend_total_tiles = 0; // all tiles since the last sweep (also the CTA rotation)
    uint64_t pend_bound = 0;   // upper bound on pending k-mer instances
    size_t pend_bytes = 0, pend_seglen_bytes = 0;
    // tiled path (tile_kernels.cu): the pending keys are level-1 segments of the same layout
    bool tiled = false;
    uint32_t ring1 = 0;           // staging-ring slots per bin of the level-1 kernels
    uint64_t *side_keys = nullptr; // side list + its counters {n, reserved} and the deferred-tile list
    unsigned long long *side_ctr = nullptr;
    uint32_t *side_deferred = nullptr;
    uint64_t side_cap = 0;
    size_t side_bytes = 0;
    // multi-GPU (katgpu_dist.cu): keys this rank has binned per level-1 CTA since the last sweep
    uint64_t dist_cta_keys = 0;
    uint64_t dist_global_distinct = 0, dist_global_total = 0; // after katgpu_count_end
    // k = 33..64: the table lives in the two-word layout of wide.cuh (katgpu_wide.cu); `g` then only carries k and
    // the canonical flag, and every entry point hands over to the wide_* functions below
    WideTable *wide = nullptr;
    uint64_t wide_distinct = 0, wide_total = 0; // after katgpu_count_end
};

constexpr size_t kChunkBytes = 128u << 20; // host feed granularity
constexpr uint64_t kSpillCap = 1u << 20;   // (hashed key, count) pairs

// KATGPU_TRACE=1: wall-clock of every host-side phase to stderr (debugging aid).
struct TraceScope {
    const char *name;
    std::chrono::steady_clock::time_point t0;
    bool on;
    explicit TraceScope(const char *n) : name(n), t0(std::chrono::steady_clock::now()), on(getenv("KATGPU_TRACE") != nullptr) {}
    ~TraceScope() {
        if (on) fprintf(stderr, "[katgpu] %-28s %8.3f ms\n", name, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
};

// ---- device block cache (katgpu_api.cu) ----
void pool_flush(katgpu_ctx *c);
cudaError_t pool_alloc(katgpu_ctx *c, size_t bytes, void **out);
void pool_free(katgpu_ctx *c, void *p, size_t bytes);

// ---- kernel timing: CUDA events on the stream the kernel is launched on ----
struct ProfScope {
    katgpu_ctx *c;
    cudaStream_t st;
    ProfSpan sp;
    bool on;
    ProfScope(katgpu_ctx *ctx, const char *family, cudaStream_t s) : c(ctx), st(s), on(ctx->profiling) {
        if (on) {
            sp.family = family;
            on = cudaEventCreate(&sp.a) == cudaSuccess;
            if (on && cudaEventCreate(&sp.b) != cudaSuccess) {
                cudaEventDestroy(sp.a);
                on = false;
            }
            if (on) cudaEventRecord(sp.a, st);
            else cudaGetLastError(); // no span for this launch; the launch itself is unaffected
        }
    }
    ~ProfScope() {
        if (on) {
            cudaEventRecord(sp.b, st);
            c->spans.push_back(sp);
        }
    }
};

// ---- shared between katgpu_api.cu and katgpu_dist.cu ----
katgpu_status read_counters(katgpu_table *t, katgpu::TableCounters &out);
katgpu_status settle(katgpu_table *t, katgpu::TableCounters &ctr);
katgpu_status sweep_pending(katgpu_table *t, katgpu::TableCounters &ctr);
katgpu_status pend_alloc(katgpu_table *t, uint64_t kmers_hint, bool &ok);
bool use_partition(const katgpu_table *t, uint64_t kmers_hint);
katgpu::SideList side_list_of(const katgpu_table *t);
void set_add_bound(katgpu::TableGeom &g, uint64_t bound);
katgpu_status feed_device_range(katgpu_table *t, const uint8_t *dev, uint64_t n_valid, uint64_t n_starts,
                                katgpu::TableCounters &ctr, uint64_t kmers_hint);
katgpu_status ensure_feed_staging(katgpu_ctx *c, size_t chunk, int ring);
uint32_t ring_slots(uint32_t F, uint32_t tile_keys, size_t max_bytes, uint32_t granule);
bool dist_can_tile(const katgpu_table *t);                                              // the tiled pending-key path is possible for this table
katgpu_status dist_make_room(katgpu_table *t, katgpu::TableCounters &ctr, uint64_t n); // grow until n more distinct keys fit
// record `ev` once everything queued so far that reads feed input has run (compute stream and, on a job, the route stream)
katgpu_status record_feed_consumed(katgpu_ctx *c, cudaEvent_t ev);

// ---- two-word keys, k = 33..64 (katgpu_wide.cu) ----
katgpu_status wide_begin(katgpu_table *t, int k, int canonical, uint64_t hash_size);
void wide_free(katgpu_table *t);
katgpu_status wide_feed_range(katgpu_table *t, const uint8_t *dev, uint64_t n_valid, uint64_t n_starts);
katgpu_status wide_add_pairs(katgpu_table *t, const uint64_t *keys_hi, const uint64_t *keys_lo, const uint64_t *counts, size_t n);
katgpu_status wide_totals(katgpu_table *t, uint64_t &n_distinct, uint64_t &n_total);
uint64_t wide_bytes(const katgpu_table *t);
katgpu_status wide_hist(katgpu_table *t, uint64_t base, uint64_t ceil_, uint64_t inc, uint64_t nb, uint64_t *bins_out);
katgpu_status wide_gcp(katgpu_table *t, double cvg_scale, uint32_t cvg_bins, uint64_t *mx_out);
katgpu_status wide_comp(katgpu_table *t1, katgpu_table *t2, katgpu_table *t3, double d1_scale, double d2_scale, uint32_t d1_bins,
                        uint32_t d2_bins, uint64_t *main_mx, uint64_t *ends_mx, uint64_t *middle_mx, uint64_t *mixed_mx,
                        katgpu_comp_counters *counters, uint64_t *spectra);
katgpu_status wide_lookup_host(katgpu_table *t, const uint64_t *keys_hi, const uint64_t *keys_lo, size_t n, int canonicalise, uint64_t *counts_out);
katgpu_status wide_export(katgpu_table *t, const katgpu::ExportSel &sel, uint64_t *keys_hi_out, uint64_t *keys_lo_out, uint64_t *counts_out,
                          size_t cap, uint64_t *n_out);

// ---- multi-GPU hooks (katgpu_dist.cu); all are no-ops / never called when ctx->dist is null ----
void dist_shutdown(katgpu_ctx *c);
katgpu_status dist_feed_device(katgpu_table *t, const uint8_t *dev, uint64_t n);
katgpu_status dist_feed_host(katgpu_table *t, const char *bases, size_t n);
katgpu_status dist_count_end(katgpu_table *t, katgpu::TableCounters &ctr);
katgpu_status dist_allreduce_u64(katgpu_ctx *c, void *dev_buf, size_t count); // sum, in place, on c->stream
int dist_n_ranks(const katgpu_ctx *c);
int dist_rank(const katgpu_ctx *c);
