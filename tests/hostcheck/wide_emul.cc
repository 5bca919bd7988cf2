// wide_emul.cc — CPU emulation of the two-word-key (k = 33..64) device code of kat_b200/csrc/wide.cuh.
// TEST INFRASTRUCTURE ONLY (tests/test_wide_host_emulation.py): wide.cuh is written __host__ __device__ so that the
// rolling extraction, canonicalisation, hashing, the insert / lookup protocol over the table layout, the reverse
// complement and the GC count can be executed here, thread by thread in the kernel's own tiling
// (wide_extract_kernel: tiles of 9216 start positions, 256 threads x 36 starts, 64 bytes of halo, bytes past the end
// read as 0), and compared with the oracle.  Atomics are replaced by their single-threaded stand-ins.
//
//   wide_emul <k> <canonical> <slots_log2> <stream file> <chunk bytes> [<probe file> <canonicalise>]
// prints "hi lo count" for every table entry (unordered), then "#stats claimed total failed grow_events";
// with a probe file ("hi lo" per line) prints "? count gc" per probe instead of the table.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "wide.cuh"
using namespace katgpu;

static const int kWideThreads = 256, kWidePer = 36, kWideTile = kWideThreads * kWidePer, kWideHalo = 64;

static WideGeom make_geom(int k, int canonical, int slots_log2, std::vector<Key128> &keys, std::vector<unsigned long long> &counts, WideCounters &ctr) {
    WideGeom g;
    g.n_slots = 1ULL << slots_log2;
    g.mask = g.n_slots - 1;
    keys.assign(g.n_slots + 1, Key128{~0ULL, ~0ULL});
    counts.assign(g.n_slots + 1, 0ULL);
    g.keys = keys.data();
    g.counts = counts.data();
    g.ctr = &ctr;
    g.k = (uint32_t)k;
    g.canonical = (uint32_t)canonical;
    g.hi_mask = k == 64 ? ~0ULL : ((1ULL << (2 * k - 64)) - 1);
    g.hi_top = (uint32_t)(2 * k - 66);
    g.rc_shift = (uint32_t)(128 - 2 * k);
    g.max_probes = (uint32_t)std::min<uint64_t>(g.n_slots, 1u << 16);
    return g;
}

// the body of wide_extract_kernel for one launch, thread by thread
static void emulate_launch(const WideGeom &g, const uint8_t *bases, uint64_t n_valid, uint64_t n_starts) {
    const int n_tiles = (int)((n_starts + kWideTile - 1) / kWideTile);
    std::vector<uint8_t> ascii(kWideTile + kWideHalo);
    for (int tile = 0; tile < n_tiles; tile++) {
        const uint64_t tile_base = (uint64_t)tile * kWideTile;
        for (int i = 0; i < kWideTile + kWideHalo; i++) ascii[i] = tile_base + i < n_valid ? bases[tile_base + i] : 0;
        for (int tid = 0; tid < kWideThreads; tid++) {
            const uint64_t p0 = tile_base + (uint64_t)tid * kWidePer;
            const uint32_t n_here = p0 < n_starts ? (uint32_t)std::min<uint64_t>(kWidePer, n_starts - p0) : 0u;
            const uint32_t *words = reinterpret_cast<const uint32_t *>(ascii.data()) + tid * (kWidePer / 4);
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
                    g.ctr->claimed += wide_insert(g, key, 1ULL) == 1;
                    g.ctr->total++;
                });
        }
    }
}

int main(int argc, char **argv) {
    if (argc < 6) return 2;
    const int k = atoi(argv[1]), canonical = atoi(argv[2]);
    int slots_log2 = atoi(argv[3]);
    const size_t chunk = (size_t)strtoull(argv[5], nullptr, 10);
    FILE *f = fopen(argv[4], "rb");
    if (!f) return 3;
    std::vector<uint8_t> bases;
    uint8_t buf[1 << 16];
    size_t r;
    while ((r = fread(buf, 1, sizeof buf, f)) > 0) bases.insert(bases.end(), buf, buf + r);
    fclose(f);
    std::vector<Key128> keys;
    std::vector<unsigned long long> counts;
    WideCounters ctr;
    memset(&ctr, 0, sizeof ctr);
    WideGeom g = make_geom(k, canonical, slots_log2, keys, counts, ctr);
    unsigned grow_events = 0;
    // katgpu_count_feed for wide tables: chunks of `chunk` start positions + (k-1 <= 63) bytes of halo, and inside a
    // chunk as many starts per launch as the load limit admits (worst case every k-mer is new), growing when none fit
    const size_t n = bases.size();
    for (size_t off = 0; off < n; off += chunk) {
        const size_t len = std::min(n - off, chunk + (size_t)kWideHalo - 1), starts_all = std::min(n - off, chunk);
        if (len < (size_t)k) continue;
        const size_t starts = std::min(starts_all, len - k + 1);
        size_t done = 0;
        while (done < starts) {
            const uint64_t soft = (uint64_t)(0.70 * (double)g.n_slots);
            uint64_t room = ctr.claimed < soft ? soft - ctr.claimed : 0;
            if (room < std::min<uint64_t>(starts - done, std::max<uint64_t>(g.n_slots / 64, (uint64_t)kWideTile))) { // wide_admit -> wide_grow
                std::vector<Key128> nk;
                std::vector<unsigned long long> nc;
                WideCounters nctr;
                memset(&nctr, 0, sizeof nctr);
                WideGeom ng = make_geom(k, canonical, ++slots_log2, nk, nc, nctr);
                for (uint64_t i = 0; i <= g.n_slots; i++)
                    if (g.counts[i]) nctr.claimed += wide_insert(ng, g.keys[i], g.counts[i]) == 1;
                nctr.total = ctr.total;
                nctr.failed |= ctr.failed;
                keys.swap(nk);
                counts.swap(nc);
                ctr = nctr;
                g = ng;
                g.keys = keys.data();
                g.counts = counts.data();
                g.ctr = &ctr;
                grow_events++;
                continue;
            }
            uint64_t granted = std::min<uint64_t>(starts - done, room);
            if (granted < starts - done && granted >= (uint64_t)kWideTile) granted = granted / kWideTile * kWideTile;
            emulate_launch(g, bases.data() + off + done, len - done, granted);
            done += granted;
        }
    }
    if (argc >= 8) {
        FILE *p = fopen(argv[6], "r");
        const int canon = atoi(argv[7]);
        unsigned long long hi, lo;
        while (p && fscanf(p, "%llu %llu", &hi, &lo) == 2) {
            Key128 key{lo, hi & g.hi_mask};
            const uint32_t gc = wide_gc(key);
            if (canon) key = wide_canonical(g, key);
            printf("? %llu %u\n", wide_lookup(g, key), gc);
        }
        if (p) fclose(p);
    } else
        for (uint64_t i = 0; i <= g.n_slots; i++)
            if (g.counts[i]) printf("%llu %llu %llu\n", (unsigned long long)g.keys[i].hi, (unsigned long long)g.keys[i].lo, g.counts[i]);
    printf("#stats %llu %llu %u %u\n", ctr.claimed, ctr.total, ctr.failed, grow_events);
    return 0;
}
