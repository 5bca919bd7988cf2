// katgpu_fastq.cu — katgpu_count_feed_fastq: raw 4-line FASTQ bytes from host memory to the GPU (pinned bounce
// buffers filled by several host threads), parsed there (parse_kernels.cu), then counted like a device feed.
#include "katgpu_internal.cuh"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

using namespace katgpu;

// ---- raw FASTQ feed: file bytes in, parsed on the device (parse_kernels.cu) ------------------------
// Host -> device copy of pageable memory at more than one core's memcpy rate: `threads` host threads
// copy 8 MiB pieces into pinned bounce buffers and queue the DMA of each piece themselves.
static katgpu_status staged_h2d(katgpu_ctx *c, uint8_t *dst, const char *src, size_t n) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type == cudaMemoryTypeHost) { // already pinned: one DMA
        CU(cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, c->copy_stream));
        return KATGPU_OK;
    }
    cudaGetLastError();
    const size_t piece = (size_t)8 << 20;
    const int n_bounce = 8;
    if (!c->bounce[0]) {
        for (int i = 0; i < n_bounce; i++) {
            CU(cudaMallocHost((void **)&c->bounce[i], piece));
            CU(cudaEventCreateWithFlags(&c->bounce_free[i], cudaEventDisableTiming));
        }
        c->bounce_bytes = piece;
    }
    const size_t n_pieces = (n + piece - 1) / piece;
    int threads = 4;
    if (const char *e = getenv("KATGPU_COPY_THREADS")) threads = std::max(1, std::min(atoi(e), n_bounce));
    threads = (int)std::min<size_t>(threads, n_pieces);
    std::atomic<int> failed{0};
    auto work = [&](int tix) {
        cudaSetDevice(c->device);
        // thread tix owns bounce buffers tix, tix + threads, ... and pieces tix, tix + threads, ...
        int slot = tix;
        for (size_t i = tix; i < n_pieces && !failed.load(); i += threads) {
            const size_t off = i * piece, len = std::min(piece, n - off);
            if (cudaEventSynchronize(c->bounce_free[slot]) != cudaSuccess) failed = 1;
            memcpy(c->bounce[slot], src + off, len);
            if (cudaMemcpyAsync(dst + off, c->bounce[slot], len, cudaMemcpyHostToDevice, c->copy_stream) != cudaSuccess) failed = 1;
            if (cudaEventRecord(c->bounce_free[slot], c->copy_stream) != cudaSuccess) failed = 1;
            slot += threads;
            if (slot >= n_bounce) slot = tix;
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < threads; t++) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
    if (failed.load()) return fail(KATGPU_ERR_CUDA, "staged host->device copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    return KATGPU_OK;
}

extern "C" katgpu_status katgpu_count_feed_fastq(katgpu_table *t, const char *bytes, size_t n, unsigned trim5p, int *regular) {
    TraceScope trace__("katgpu_count_feed_fastq");
    if (!t || !regular || (!bytes && n)) return fail(KATGPU_ERR_INVALID, "katgpu_count_feed_fastq: NULL argument");
    if (t->ended) return fail(KATGPU_ERR_STATE, "katgpu_count_feed_fastq: table already finished");
    if (n > ((size_t)1 << 30)) return fail(KATGPU_ERR_INVALID, "katgpu_count_feed_fastq: at most 1 GiB of whole records per call");
    katgpu_ctx *c = t->ctx;
    CU(cudaSetDevice(c->device));
    *regular = 0;
    // a job's feeds are collective: a rank without data still takes part (with an empty stream)
    if (n == 0 || bytes[n - 1] != '\n') {
        if (n == 0) *regular = 1;
        if (c->dist) return dist_feed_device(t, nullptr, 0);
        return KATGPU_OK;
    }
    if (c->raw_cap < n + 64) {
        CU(cudaDeviceSynchronize()); // nothing reads the old buffers any more (compute, copy and route streams)
        const size_t cap = std::max<size_t>((n + 64 + ((size_t)64 << 20) - 1) / ((size_t)64 << 20) * ((size_t)64 << 20), c->raw_cap);
        for (int i = 0; i < 2; i++) {
            if (c->raw[i]) pool_free(c, c->raw[i], c->raw_cap);
            c->raw[i] = nullptr;
        }
        for (int i = 0; i < 2; i++) {
            void *p = nullptr;
            CU(pool_alloc(c, cap, &p));
            c->raw[i] = (uint8_t *)p;
            if (!c->raw_free[i]) CU(cudaEventCreateWithFlags(&c->raw_free[i], cudaEventDisableTiming));
            CU(cudaEventRecord(c->raw_free[i], c->stream));
        }
        c->raw_cap = cap;
    }
    const uint32_t n_tiles = nl_tiles(n);
    const size_t nl_need = n / 2 + 16; // the shortest regular record, "@\nA\n+\nI\n", has a newline every 2 bytes
    if (c->nl_cap < nl_need || c->nl_tiles_cap < n_tiles) {
        CU(cudaStreamSynchronize(c->stream));
        if (c->nl_buf) pool_free(c, c->nl_buf, c->nl_cap * 4);
        if (c->nl_tiles_buf) pool_free(c, c->nl_tiles_buf, c->nl_tiles_cap * 4);
        void *p = nullptr;
        c->nl_cap = std::max(nl_need, c->nl_cap);
        c->nl_tiles_cap = std::max<size_t>(n_tiles, c->nl_tiles_cap);
        CU(pool_alloc(c, c->nl_cap * 4, &p));
        c->nl_buf = (uint32_t *)p;
        CU(pool_alloc(c, c->nl_tiles_cap * 4, &p));
        c->nl_tiles_buf = (uint32_t *)p;
        if (!c->parse_flags) {
            CU(pool_alloc(c, 256, &p));
            c->parse_flags = (uint32_t *)p;
        }
    }
    const int b = c->raw_next;
    c->raw_next ^= 1;
    uint8_t *raw = c->raw[b];
    // the buffer is free once the kernels of the feed two calls back have read it
    CU(cudaStreamWaitEvent(c->copy_stream, c->raw_free[b], 0));
    KT(staged_h2d(c, raw, bytes, n));
    CU(cudaEventRecord(c->chunk_ready[0], c->copy_stream));
    CU(cudaStreamWaitEvent(c->stream, c->chunk_ready[0], 0));
    CU(cudaMemsetAsync(c->parse_flags, 0, 8, c->stream));
    {
        ProfScope ps(c, "parse", c->stream);
        launch_nl_count(raw, n, c->nl_tiles_buf, n_tiles, c->stream);
        launch_nl_scan(c->nl_tiles_buf, n_tiles, c->parse_flags + 1, c->stream);
    }
    c->launches += 2;
    uint32_t h_flags[2] = {0, 0};
    CU(cudaMemcpyAsync(h_flags, c->parse_flags, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    const uint32_t n_nl = h_flags[1];
    if (n_nl == 0 || (n_nl & 3u) || n_nl > nl_need) { // not four lines per record (or lines too short to be one)
        CU(cudaEventRecord(c->raw_free[b], c->stream));
        if (c->dist) return dist_feed_device(t, nullptr, 0); // the other ranks' feeds are collective: take part, with nothing
        return KATGPU_OK;
    }
    {
        ProfScope ps(c, "parse", c->stream);
        launch_nl_scatter(raw, n, c->nl_tiles_buf, n_tiles, c->nl_buf, (uint32_t)std::min<size_t>(c->nl_cap, 0xFFFFFFFFu), c->stream);
        launch_fastq_mask(raw, c->nl_buf, n_nl / 4, trim5p, c->parse_flags, c->sm_count * 8, c->stream);
    }
    c->launches += 2;
    CU(cudaMemcpyAsync(h_flags, c->parse_flags, 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    if (h_flags[0] != 0) { // some record is not in the regular layout: nothing was counted
        CU(cudaEventRecord(c->raw_free[b], c->stream));
        if (c->dist) return dist_feed_device(t, nullptr, 0);
        return KATGPU_OK;
    }
    *regular = 1;
    katgpu_status st;
    if (c->dist) st = dist_feed_device(t, raw, n);
    else {
        TableCounters ctr;
        KT(read_counters(t, ctr));
        st = feed_device_range(t, raw, n, n, ctr, n / 2);
    }
    KT(st);
    return record_feed_consumed(c, c->raw_free[b]);
}
