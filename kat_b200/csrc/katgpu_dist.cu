// katgpu_dist.cu — one context per GPU, N contexts = one multi-GPU job.
//
//   katgpu_dist_init(ctx, rank, n, id)   joins the job: an NCCL communicator (called directly, the
//                                        library is dlopen'ed so single-GPU users need none), two
//                                        receive areas per rank, peer-mapped on every other rank
//                                        (cudaIpc across processes, plain peer access inside one).
//   katgpu_count_feed*                   become collective: the stream is processed in ROUNDS; in a
//                                        round every rank runs route_bases_kernel (dist_kernels.cu),
//                                        which stores each k-mer's hashed key into its owner's receive
//                                        area over NVLink, then every owner bins what it received into
//                                        its level-1 segments (bin_keys_kernel, segment source).  The
//                                        two receive areas alternate, so round r+1 is routed while
//                                        round r is binned.  The only cross-rank synchronisation is
//                                        two tiny all-reduces per round (stream-ordered barriers).
//   katgpu_count_end                     local tiled sweep (level-2 binning + tile_insert), totals summed.
//   katgpu_hist / gcp / comp / lookup    local walk + ONE ncclAllReduce(sum, u64) of the output.
//
// Ownership is owner_of(h) (table.cuh): a multiplicative hash of the hashed key scaled to N, so any
// N in 2..8 works and the local tables use every bit of h exactly as on one GPU.
#include "katgpu_internal.cuh"
#include "dist_kernels.cuh"

#include <dlfcn.h>
#include <nccl.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstring>

using namespace katgpu;

// ---- NCCL, loaded on first use ----------------------------------------------------------------
namespace {
struct NcclApi {
    void *so = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
};
NcclApi g_nccl;

katgpu_status nccl_load() {
    if (g_nccl.so) return KATGPU_OK;
    // RTLD_NOLOAD first: inside a process that already mapped an NCCL (torch) use that one
    void *so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!so) so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!so) so = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!so) return fail(KATGPU_ERR_INVALID, "multi-GPU needs NCCL: %s", dlerror());
    NcclApi a;
    a.so = so;
#define SYM(field, name)                                                                   \
    a.field = reinterpret_cast<decltype(a.field)>(dlsym(so, name));                        \
    if (!a.field) return fail(KATGPU_ERR_INVALID, "libnccl: symbol %s missing", name)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(AllReduce, "ncclAllReduce");
    SYM(AllGather, "ncclAllGather");
    SYM(GetErrorString, "ncclGetErrorString");
    SYM(GetVersion, "ncclGetVersion");
#undef SYM
    g_nccl = a;
    return KATGPU_OK;
}
} // namespace
#define NC(call)                                                                                          \
    do {                                                                                                  \
        ncclResult_t r__ = (call);                                                                        \
        if (r__ != ncclSuccess) return fail(KATGPU_ERR_CUDA, "%s: %s", #call, g_nccl.GetErrorString(r__)); \
    } while (0)

// ---- per-context state --------------------------------------------------------------------------
struct PeerInfo { // what every rank tells the others about its receive block
    uint64_t pid, ptr;
    int32_t device, pad;
    cudaIpcMemHandle_t handle;
};

struct DistState {
    int rank = 0, n = 1;
    ncclComm_t comm = nullptr;
    cudaStream_t comm_stream = nullptr, route_stream = nullptr;
    int *d_token = nullptr;            // barrier payload
    int bps_route = 2;
    // receive block of this rank: {keys[2][n][grid][cap], len[2][n][grid]}; peers map it whole
    uint8_t *block = nullptr;
    size_t block_bytes = 0;
    uint8_t *peer_block[kMaxRanks] = {}; // rank o's block as addressable from this device
    bool peer_ipc[kMaxRanks] = {};
    uint32_t grid = 0, cap = 0;          // CTAs of route_bases_kernel, keys per (sender, CTA) segment
    uint64_t round_positions = 0;        // k-mer starts a rank routes per round
    cudaEvent_t ev_routed[2] = {}, ev_x[2] = {}, ev_binned[2] = {}, ev_y[2] = {}, ev_tmp = nullptr;
    uint64_t round = 0;
    uint64_t keys_sent = 0, keys_received = 0; // totals, for the NVLink line of the bench
    size_t keys_off(int buf) const { return (size_t)buf * n * grid * cap * 8; }
    size_t len_off(int buf) const { return 2 * (size_t)n * grid * cap * 8 + (size_t)buf * n * grid * 4; }
};

int dist_n_ranks(const katgpu_ctx *c) { return c && c->dist ? c->dist->n : 1; }
int dist_rank(const katgpu_ctx *c) { return c && c->dist ? c->dist->rank : 0; }

extern "C" katgpu_status katgpu_dist_unique_id(void *id128) {
    if (!id128) return fail(KATGPU_ERR_INVALID, "katgpu_dist_unique_id: NULL argument");
    KT(nccl_load());
    static_assert(sizeof(ncclUniqueId) == 128, "NCCL unique id is 128 bytes");
    ncclUniqueId id;
    NC(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
    return KATGPU_OK;
}

// Stream-ordered barrier across the job: when it completes on `comm_stream` here, every rank has
// reached the same point of ITS comm stream (kernel completion makes peer stores visible).
static katgpu_status barrier_on_comm_stream(DistState *d) {
    NC(g_nccl.AllReduce(d->d_token, d->d_token, 1, ncclInt32, ncclSum, d->comm, d->comm_stream));
    return KATGPU_OK;
}

extern "C" katgpu_status katgpu_dist_init(katgpu_ctx *c, int rank, int n_ranks, const void *id128) {
    if (!c || !id128) return fail(KATGPU_ERR_INVALID, "katgpu_dist_init: NULL argument");
    if (c->dist) return fail(KATGPU_ERR_STATE, "katgpu_dist_init: context already belongs to a job");
    if (n_ranks < 2 || n_ranks > kMaxRanks || rank < 0 || rank >= n_ranks)
        return fail(KATGPU_ERR_INVALID, "katgpu_dist_init: rank %d of %d (2..%d ranks, one NVSwitch box)", rank, n_ranks, kMaxRanks);
    KT(nccl_load());
    CU(cudaSetDevice(c->device));
    DistState *d = new (std::nothrow) DistState();
    if (!d) return fail(KATGPU_ERR_NOMEM, "katgpu_dist_init: out of host memory");
    d->rank = rank;
    d->n = n_ranks;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    ncclResult_t r = g_nccl.CommInitRank(&d->comm, n_ranks, id, rank);
    if (r != ncclSuccess) {
        delete d;
        return fail(KATGPU_ERR_CUDA, "ncclCommInitRank: %s", g_nccl.GetErrorString(r));
    }
    c->dist = d;
    CU(cudaStreamCreateWithFlags(&d->comm_stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&d->route_stream, cudaStreamNonBlocking));
    CU(cudaMalloc((void **)&d->d_token, 512));
    CU(cudaMemset(d->d_token, 0, 512));
    for (int i = 0; i < 2; i++) {
        CU(cudaEventCreateWithFlags(&d->ev_routed[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&d->ev_x[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&d->ev_binned[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&d->ev_y[i], cudaEventDisableTiming));
    }
    CU(cudaEventCreateWithFlags(&d->ev_tmp, cudaEventDisableTiming));
    CU(route_configure(&d->bps_route));
    if (d->bps_route < 1) return fail(KATGPU_ERR_CUDA, "katgpu_dist_init: route kernel does not fit on an SM");

    // receive block: sized once, for rounds of KATGPU_DIST_ROUND_MB megabytes of base stream per rank
    // A round is two kernels per rank (route, owner-side binning) and two barriers: long rounds amortise
    // their tails.  The receive block is 2 areas x N senders x 8 bytes per k-mer start of a round: 1 GiB / N
    // of base stream per round keeps it at 16 GiB whatever N.
    uint64_t round_mb = std::max(128, std::min(512, 1024 / n_ranks));
    if (const char *e = getenv("KATGPU_DIST_ROUND_MB")) round_mb = std::max(1, atoi(e));
    // two route CTAs per SM (512 threads, 58 registers, 47 KB each); KATGPU_DIST_ROUTE_CTAS=1 leaves half of every SM free
    int route_ctas = d->bps_route;
    if (const char *e = getenv("KATGPU_DIST_ROUTE_CTAS")) route_ctas = std::max(1, std::min(atoi(e), d->bps_route));
    d->grid = (uint32_t)(c->sm_count * route_ctas);
    const uint64_t tiles = ((round_mb << 20) + kBinTile - 1) / kBinTile;
    const uint64_t tiles_per_cta = (tiles + d->grid - 1) / d->grid;
    d->round_positions = tiles_per_cta * d->grid * (uint64_t)kBinTile;
    d->cap = (uint32_t)(tiles_per_cta * kBinTile); // every start of a CTA's tiles could go to one owner (poly-A input)
    d->block_bytes = 2 * (size_t)n_ranks * d->grid * d->cap * 8 + 2 * (size_t)n_ranks * d->grid * 4;
    d->block_bytes = (d->block_bytes + 255) / 256 * 256;
    {
        void *p = nullptr;
        cudaError_t e = cudaMalloc(&p, d->block_bytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            pool_flush(c);
            CU(cudaMalloc(&p, d->block_bytes));
        }
        d->block = (uint8_t *)p;
        CU(cudaMemset(d->block + d->len_off(0), 0, 2 * (size_t)n_ranks * d->grid * 4));
    }

    // tell everybody where it is; map everybody else's
    PeerInfo mine{};
    mine.pid = (uint64_t)getpid();
    mine.ptr = (uint64_t)(uintptr_t)d->block;
    mine.device = c->device;
    CU(cudaIpcGetMemHandle(&mine.handle, d->block));
    PeerInfo *d_all = nullptr, h_all[kMaxRanks];
    CU(cudaMalloc((void **)&d_all, sizeof(PeerInfo) * kMaxRanks));
    CU(cudaMemcpy(d_all + rank, &mine, sizeof mine, cudaMemcpyHostToDevice));
    NC(g_nccl.AllGather(d_all + rank, d_all, sizeof(PeerInfo), ncclChar, d->comm, d->comm_stream));
    CU(cudaStreamSynchronize(d->comm_stream));
    CU(cudaMemcpy(h_all, d_all, sizeof(PeerInfo) * n_ranks, cudaMemcpyDeviceToHost));
    cudaFree(d_all);
    for (int o = 0; o < n_ranks; o++) {
        if (o == rank) {
            d->peer_block[o] = d->block;
            continue;
        }
        if (h_all[o].pid == mine.pid) { // same process (one host thread per GPU): plain peer access
            int can = 0;
            CU(cudaDeviceCanAccessPeer(&can, c->device, h_all[o].device));
            if (!can) return fail(KATGPU_ERR_CUDA, "katgpu_dist_init: device %d cannot access device %d", c->device, h_all[o].device);
            cudaError_t e = cudaDeviceEnablePeerAccess(h_all[o].device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU(e);
            cudaGetLastError();
            d->peer_block[o] = (uint8_t *)(uintptr_t)h_all[o].ptr;
        } else {
            void *p = nullptr;
            CU(cudaIpcOpenMemHandle(&p, h_all[o].handle, cudaIpcMemLazyEnablePeerAccess));
            d->peer_block[o] = (uint8_t *)p;
            d->peer_ipc[o] = true;
        }
    }
    KT(barrier_on_comm_stream(d)); // nobody proceeds (or frees) before all maps exist
    CU(cudaStreamSynchronize(d->comm_stream));
    return KATGPU_OK;
}

void dist_shutdown(katgpu_ctx *c) {
    DistState *d = c->dist;
    if (!d) return;
    cudaStreamSynchronize(d->route_stream);
    cudaStreamSynchronize(d->comm_stream);
    for (int o = 0; o < d->n; o++)
        if (d->peer_ipc[o] && d->peer_block[o]) cudaIpcCloseMemHandle(d->peer_block[o]);
    if (d->comm) g_nccl.CommDestroy(d->comm);
    if (d->block) cudaFree(d->block);
    if (d->d_token) cudaFree(d->d_token);
    for (int i = 0; i < 2; i++) {
        cudaEventDestroy(d->ev_routed[i]);
        cudaEventDestroy(d->ev_x[i]);
        cudaEventDestroy(d->ev_binned[i]);
        cudaEventDestroy(d->ev_y[i]);
    }
    cudaEventDestroy(d->ev_tmp);
    cudaStreamDestroy(d->comm_stream);
    cudaStreamDestroy(d->route_stream);
    delete d;
    c->dist = nullptr;
}

extern "C" int katgpu_dist_rank(katgpu_ctx *c) { return dist_rank(c); }
extern "C" int katgpu_dist_n_ranks(katgpu_ctx *c) { return dist_n_ranks(c); }

extern "C" katgpu_status katgpu_dist_barrier(katgpu_ctx *c) {
    if (!c || !c->dist) return fail(KATGPU_ERR_STATE, "katgpu_dist_barrier: not a multi-GPU context");
    DistState *d = c->dist;
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(d->route_stream));
    KT(barrier_on_comm_stream(d));
    CU(cudaStreamSynchronize(d->comm_stream));
    return KATGPU_OK;
}

extern "C" katgpu_status katgpu_dist_stats(katgpu_ctx *c, uint64_t *keys_sent, uint64_t *keys_received, uint64_t *rounds) {
    if (!c || !c->dist) return fail(KATGPU_ERR_STATE, "katgpu_dist_stats: not a multi-GPU context");
    if (keys_sent) *keys_sent = c->dist->keys_sent;
    if (keys_received) *keys_received = c->dist->keys_received;
    if (rounds) *rounds = c->dist->round;
    return KATGPU_OK;
}

// Sum `count` uint64 values in place across the job; ordered after everything queued on c->stream,
// and c->stream continues after it.  All NCCL work of a context goes through its one comm stream.
katgpu_status dist_allreduce_u64(katgpu_ctx *c, void *dev_buf, size_t count) {
    DistState *d = c->dist;
    CU(cudaEventRecord(d->ev_tmp, c->stream));
    CU(cudaStreamWaitEvent(d->comm_stream, d->ev_tmp, 0));
    NC(g_nccl.AllReduce(dev_buf, dev_buf, count, ncclUint64, ncclSum, d->comm, d->comm_stream));
    CU(cudaEventRecord(d->ev_tmp, d->comm_stream));
    CU(cudaStreamWaitEvent(c->stream, d->ev_tmp, 0));
    return KATGPU_OK;
}

// Every rank's feed size, gathered once per feed call (host-synchronous): from it each rank knows
// how many rounds the job needs and about how many keys every round will bring — no host
// synchronisation is needed inside the rounds.
static katgpu_status allgather_sizes(katgpu_ctx *c, uint64_t mine, uint64_t *all) {
    DistState *d = c->dist;
    unsigned long long *buf = reinterpret_cast<unsigned long long *>(d->d_token + 16); // [kMaxRanks]
    CU(cudaMemcpyAsync(buf + d->rank, &mine, 8, cudaMemcpyHostToDevice, d->comm_stream));
    NC(g_nccl.AllGather(buf + d->rank, buf, 1, ncclUint64, d->comm, d->comm_stream));
    CU(cudaMemcpyAsync(all, buf, 8 * (size_t)d->n, cudaMemcpyDeviceToHost, d->comm_stream));
    CU(cudaStreamSynchronize(d->comm_stream));
    return KATGPU_OK;
}

// ---- one round -------------------------------------------------------------------------------------
// Route [dev, dev + n_valid) (k-mers starting in the first n_starts positions; 0 = this rank has
// nothing left but must still take part), then bin what this rank received.  Everything is queued
// on streams — the host does not wait for the GPU here, so route(r+1) follows route(r) immediately
// and overlaps the owner-side binning of round r.
//   ready_ev     the route kernel waits for it (the H2D copy of a staged chunk), may be null
//   consumed_ev  recorded when the route kernel has read its input, may be null
//   job_starts   k-mer starts all ranks route in this round (an upper bound on the keys this rank receives)
static katgpu_status route_round(katgpu_table *t, const uint8_t *dev, uint64_t n_valid, uint64_t n_starts, TableCounters &ctr,
                                 cudaEvent_t ready_ev, cudaEvent_t consumed_ev, uint64_t kmers_hint, uint64_t job_starts) {
    katgpu_ctx *c = t->ctx;
    DistState *d = c->dist;
    const int b = (int)(d->round & 1);
    if (n_valid < (uint64_t)t->g.k) n_starts = 0;
    else n_starts = std::min(n_starts, n_valid - t->g.k + 1);
    if (n_starts > d->round_positions) return fail(KATGPU_ERR_INVALID, "internal: route round larger than the receive area");
    RouteParams P{};
    P.g = t->g;
    P.bases = dev;
    P.n_valid = n_valid;
    P.n_starts = n_starts;
    P.n_tiles = (int)((n_starts + kBinTile - 1) / kBinTile);
    P.tma_ok = (reinterpret_cast<uintptr_t>(dev) & 15) == 0;
    P.n_ranks = (uint32_t)d->n;
    for (int o = 0; o < d->n; o++) {
        P.seg[o] = reinterpret_cast<uint64_t *>(d->peer_block[o] + d->keys_off(b)) + (size_t)d->rank * d->grid * d->cap;
        P.len[o] = reinterpret_cast<uint32_t *>(d->peer_block[o] + d->len_off(b)) + (size_t)d->rank * d->grid;
    }
    P.cap = d->cap;
    // the receive areas `b` of every owner are free again once everybody binned round (r - 2)
    if (d->round >= 2) CU(cudaStreamWaitEvent(d->route_stream, d->ev_y[b], 0));
    if (ready_ev) CU(cudaStreamWaitEvent(d->route_stream, ready_ev, 0));
    {
        ProfScope ps(c, "route", d->route_stream);
        launch_route_bases(P, (int)d->grid, d->route_stream);
    }
    c->launches++;
    CU(cudaGetLastError());
    if (consumed_ev) CU(cudaEventRecord(consumed_ev, d->route_stream));
    CU(cudaEventRecord(d->ev_routed[b], d->route_stream));
    // barrier X: every rank's route kernel of this round has finished -> my receive area is complete
    CU(cudaStreamWaitEvent(d->comm_stream, d->ev_routed[b], 0));
    KT(barrier_on_comm_stream(d));
    CU(cudaEventRecord(d->ev_x[b], d->comm_stream));
    // owner side, on the compute stream
    CU(cudaStreamWaitEvent(c->stream, d->ev_x[b], 0));
    const size_t n_src = (size_t)d->n * d->grid;
    const uint32_t *len_dev = reinterpret_cast<const uint32_t *>(d->block + d->len_off(b));
    // with a uniform owner hash this rank receives its 1/N of the round; the binning kernels cope
    // with any excess exactly (full segments spill to the side list), so an estimate is enough here
    const uint64_t expected = job_starts / (uint64_t)d->n + (job_starts ? 1 : 0);
    d->keys_received += expected;
    d->keys_sent += n_starts;
    if (job_starts) {
        // The pending-key (tiled) path when the table is large enough for it; its buffers are sized once,
        // for the whole feed.
        bool tiled = t->pend_keys != nullptr && t->tiled;
        if (!t->pend_keys && dist_can_tile(t)) {
            bool ok = false;
            KT(pend_alloc(t, std::max<uint64_t>(std::max(expected, kmers_hint), t->n_slots / 4), ok));
            tiled = ok && t->tiled;
        }
        if (tiled) {
            // segments s = j, j + G, ... go to level-1 CTA j
            const uint32_t G = t->pend_segs;
            // (`expected` counts k-mer STARTS, as the hint the segments were sized from does: reads lose k-1 starts each,
            // so this already over-estimates real read sets by ~20 %; an under-estimate only sends keys to the side list)
            const uint64_t per_cta = expected / G + kBinTile / 4;
            const double F = (double)(1u << t->pend_Pb);
            auto fill_of = [&](uint64_t keys) { return (double)keys / F + 4.5 * std::sqrt((double)keys / F) + 16.0; };
            if (t->dist_cta_keys && fill_of(t->dist_cta_keys + per_cta) > (double)t->pend_seg_cap) {
                KT(sweep_pending(t, ctr)); // local: level-2 binning + tile insertion of what is pending
                t->dist_cta_keys = 0;
            }
            BinDst D{};
            D.keys = t->pend_keys;
            D.len = t->pend_seglen;
            D.cap = t->pend_seg_cap;
            D.stride = t->pend_segs;
            D.shift = t->g.b - t->pend_Pb;
            D.mask = (1u << t->pend_Pb) - 1;
            D.F = 1u << t->pend_Pb;
            // bin_hashed_kernel: one 1024-thread CTA per SM, its staging ring may take most of the shared memory;
            // as many tiles between two flushes as the ring takes with 3.5 sigma of head room
            const bool half_sm = getenv("KATGPU_BINH_THREADS") && atoi(getenv("KATGPU_BINH_THREADS")) == 512;
            const size_t ringh_bytes = ((size_t)(half_sm ? 100 : 190) << 10) - kTileHllM * 4;
            uint32_t cpf = half_sm ? 1 : 3;
            for (; cpf > 1; cpf--) {
                const double mean = (double)(kBinTile * cpf) / D.F;
                if (((double)kBinHGranule + mean + 3.5 * std::sqrt(mean) + 1.0) * D.F * 8.0 + D.F * 8.0 + 16.0 <= (double)ringh_bytes) break;
            }
            D.R = ring_slots(D.F, kBinTile * cpf, ringh_bytes, kBinHGranule);
            BinKeysParams B{};
            B.g = t->g;
            set_add_bound(B.g, ~0ULL);
            B.seg_keys = reinterpret_cast<const uint64_t *>(d->block + d->keys_off(b));
            B.seg_len = len_dev;
            B.n_src = (uint32_t)n_src;
            B.seg_cap = d->cap;
            B.tile_keys = kBinTile * cpf;
            B.cpf = cpf;
            B.dst = D;
            B.side = side_list_of(t);
            B.hll = t->pend_hll;
            {
                ProfScope ps(c, "bin1", c->stream);
                launch_bin_keys(B, (int)G, c->stream);
            }
            c->launches++;
            CU(cudaGetLastError());
            t->dist_cta_keys += per_cta;
            t->pend_bound += job_starts; // upper bound: every key of the round could have come here
            t->pend_total_tiles += (expected + kBinTile - 1) / kBinTile;
        } else {
            // small tables (small k, small -H): the received segments go through the general
            // random-access insertion; the table is grown first so that they fit in any case
            KT(sweep_pending(t, ctr));
            KT(dist_make_room(t, ctr, job_starts));
            TileParams P2{};
            P2.g = t->g;
            set_add_bound(P2.g, ~0ULL);
            P2.keys = reinterpret_cast<const uint64_t *>(d->block + d->keys_off(b));
            P2.len = len_dev;
            P2.cap = d->cap;
            P2.parts = 1;
            P2.n_tiles = (uint32_t)n_src;
            {
                ProfScope ps(c, "insert", c->stream);
                launch_insert_deferred(P2, nullptr, P2.n_tiles, c->stream);
            }
            c->launches++;
            CU(cudaGetLastError());
            KT(settle(t, ctr));
        }
    }
    CU(cudaEventRecord(d->ev_binned[b], c->stream));
    // barrier Y: every owner has consumed receive area `b`
    CU(cudaStreamWaitEvent(d->comm_stream, d->ev_binned[b], 0));
    KT(barrier_on_comm_stream(d));
    CU(cudaEventRecord(d->ev_y[b], d->comm_stream));
    d->round++;
    return KATGPU_OK;
}

// k-mer starts all ranks route in round r when rank q feeds sizes[q] bytes in pieces of `piece`
static uint64_t job_starts_of_round(const uint64_t *sizes, int n, uint64_t piece, uint64_t r) {
    uint64_t tot = 0;
    for (int q = 0; q < n; q++) tot += sizes[q] > r * piece ? std::min(piece, sizes[q] - r * piece) : 0;
    return tot;
}

katgpu_status dist_feed_device(katgpu_table *t, const uint8_t *dev, uint64_t n) {
    katgpu_ctx *c = t->ctx;
    DistState *d = c->dist;
    TableCounters ctr;
    KT(read_counters(t, ctr));
    const uint64_t R = d->round_positions;
    uint64_t sizes[kMaxRanks];
    KT(allgather_sizes(c, n, sizes));
    uint64_t rounds = 0;
    for (int q = 0; q < d->n; q++) rounds = std::max(rounds, (sizes[q] + R - 1) / R);
    for (uint64_t r = 0; r < rounds; r++) {
        const uint64_t off = std::min<uint64_t>(n, r * R);
        KT(route_round(t, dev + off, n - off, std::min<uint64_t>(n - off, R), ctr, nullptr, nullptr, n, job_starts_of_round(sizes, d->n, R, r)));
    }
    return KATGPU_OK;
}

katgpu_status dist_feed_host(katgpu_table *t, const char *bases, size_t n) {
    katgpu_ctx *c = t->ctx;
    DistState *d = c->dist;
    TableCounters ctr;
    KT(read_counters(t, ctr));
    const size_t chunk = (size_t)std::min<uint64_t>(kChunkBytes, d->round_positions);
    const size_t n_chunks = (n + chunk - 1) / chunk;
    uint64_t sizes[kMaxRanks];
    KT(allgather_sizes(c, n, sizes));
    uint64_t rounds = 0;
    for (int q = 0; q < d->n; q++) rounds = std::max<uint64_t>(rounds, (sizes[q] + chunk - 1) / chunk);
    const int ring = (int)std::min<size_t>(kRing, std::max<size_t>(2, n_chunks));
    if (n_chunks) KT(ensure_feed_staging(c, chunk, ring));
    auto issue_copy = [&](size_t i) -> katgpu_status {
        const int b = (int)(i % ring);
        const size_t off = i * chunk;
        const size_t len = std::min(n - off, chunk + kHalo);
        CU(cudaStreamWaitEvent(c->copy_stream, c->chunk_free[b], 0));
        CU(cudaMemcpyAsync(c->chunk[b], bases + off, len, cudaMemcpyHostToDevice, c->copy_stream));
        CU(cudaEventRecord(c->chunk_ready[b], c->copy_stream));
        return KATGPU_OK;
    };
    for (int b = 0; b < ring && n_chunks; b++) CU(cudaEventRecord(c->chunk_free[b], d->route_stream));
    size_t issued = 0;
    for (uint64_t i = 0; i < rounds; i++) {
        const uint64_t js = job_starts_of_round(sizes, d->n, chunk, i);
        if (i < n_chunks) {
            while (issued < n_chunks && issued < i + (size_t)ring) KT(issue_copy(issued++));
            const int b = (int)(i % ring);
            const size_t off = i * chunk;
            KT(route_round(t, c->chunk[b], std::min(n - off, chunk + kHalo), std::min(n - off, chunk), ctr, c->chunk_ready[b], c->chunk_free[b], n, js));
        } else
            KT(route_round(t, nullptr, 0, 0, ctr, nullptr, nullptr, n, js));
    }
    CU(cudaStreamSynchronize(c->copy_stream)); // `bases` belongs to the caller again
    return KATGPU_OK;
}

// Called by katgpu_count_end before the local sweep: all rounds of every feed have completed their
// barriers already; drain the streams so that the sweep sees every binned key.
katgpu_status dist_count_end(katgpu_table *t, TableCounters &ctr) {
    katgpu_ctx *c = t->ctx;
    DistState *d = c->dist;
    CU(cudaStreamSynchronize(d->route_stream));
    CU(cudaStreamSynchronize(d->comm_stream));
    CU(cudaStreamSynchronize(c->stream));
    t->dist_cta_keys = 0;
    return read_counters(t, ctr);
}

// Measurement aid (tools/route_probe.py, ncu captures): route_bases_kernel on ONE GPU with `n_ranks`
// owners whose receive areas are all local memory — the kernel's own cost without NVLink.  The keys
// are written and thrown away; *ms = kernel time (CUDA events on the compute stream), *keys = keys written.
extern "C" katgpu_status katgpu_debug_route_local(katgpu_table *t, const void *bases_dev, size_t n, int n_ranks, double *ms, uint64_t *keys) {
    if (!t || !bases_dev || !ms || n_ranks < 1 || n_ranks > kMaxRanks || t->wide) return fail(KATGPU_ERR_INVALID, "katgpu_debug_route_local: bad arguments");
    katgpu_ctx *c = t->ctx;
    CU(cudaSetDevice(c->device));
    int bps = 2;
    CU(route_configure(&bps));
    const uint32_t grid = (uint32_t)(c->sm_count * std::max(1, bps));
    const uint64_t n_starts = n >= (size_t)t->g.k ? n - t->g.k + 1 : 0;
    const uint64_t tiles = (n_starts + kBinTile - 1) / kBinTile, tiles_per_cta = (tiles + grid - 1) / grid;
    const uint32_t cap = (uint32_t)std::max<uint64_t>(kBinTile, tiles_per_cta * kBinTile);
    // each owner's share is ~1/n_ranks of a CTA's keys: size the areas for 2x that (uniform hashes; this is a probe)
    const uint64_t per_owner = (uint64_t)grid * cap;
    uint64_t *keys_buf = nullptr;
    uint32_t *len_buf = nullptr;
    CU(cudaMalloc((void **)&keys_buf, (size_t)n_ranks * per_owner * 8));
    CU(cudaMalloc((void **)&len_buf, (size_t)n_ranks * grid * 4));
    RouteParams P{};
    P.g = t->g;
    P.bases = (const uint8_t *)bases_dev;
    P.n_valid = n;
    P.n_starts = n_starts;
    P.n_tiles = (int)tiles;
    P.tma_ok = (reinterpret_cast<uintptr_t>(bases_dev) & 15) == 0;
    P.n_ranks = (uint32_t)n_ranks;
    for (int o = 0; o < n_ranks; o++) {
        P.seg[o] = keys_buf + (size_t)o * per_owner;
        P.len[o] = len_buf + (size_t)o * grid;
    }
    P.cap = cap;
    cudaEvent_t a, b;
    CU(cudaEventCreate(&a));
    CU(cudaEventCreate(&b));
    launch_route_bases(P, (int)grid, c->stream); // warm
    CU(cudaEventRecord(a, c->stream));
    launch_route_bases(P, (int)grid, c->stream);
    CU(cudaEventRecord(b, c->stream));
    CU(cudaEventSynchronize(b));
    c->launches += 2;
    float f = 0;
    CU(cudaEventElapsedTime(&f, a, b));
    *ms = f;
    if (keys) {
        std::vector<uint32_t> h((size_t)n_ranks * grid);
        CU(cudaMemcpy(h.data(), len_buf, h.size() * 4, cudaMemcpyDeviceToHost));
        uint64_t tot = 0;
        for (uint32_t v : h) tot += v;
        *keys = tot;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(keys_buf);
    cudaFree(len_buf);
    CU(cudaGetLastError());
    return KATGPU_OK;
}

katgpu_status record_feed_consumed(katgpu_ctx *c, cudaEvent_t ev) {
    if (c->dist) { // the route kernel reads the input on its own stream
        CU(cudaEventRecord(c->dist->ev_tmp, c->dist->route_stream));
        CU(cudaStreamWaitEvent(c->stream, c->dist->ev_tmp, 0));
    }
    CU(cudaEventRecord(ev, c->stream));
    return KATGPU_OK;
}
