// table.cuh — the HBM-resident k-mer count table of libkatgpu and its device
// primitives.  Replaces the role of jellyfish's large_hash::array
// (JFI/large_hash_array.hpp) + RectangularBinaryMatrix hash
// (JFI/rectangular_binary_matrix.hpp); not a port: outputs of KAT do not depend
// on table layout, so the layout is chosen for B200.
//
// Layout: capacity = 4 << Lb slots of ONE 64-bit word, grouped
// in 32-byte buckets (= one HBM/L2 sector) of 4 slots.  A k-mer key (2k = b
// bits) is passed through a bijection `mix` on b-bit integers; the top Lb bits
// of h = mix(key) pick the home bucket, so only the low rbits = b - Lb bits
// ("rem") need storing — the bucket address carries the rest (quotienting).
//
//   63            cshift cshift-1   rbits+6  rbits+5 .. rbits  rbits-1 .. 0
//   [ count (cbits) ][ unused ][ovf flag][   probe j + 1     ][     rem     ]
//
// An empty slot is the all-zero word (probe field of an occupied slot is >= 1).
// One insertion = one 32-byte bucket load + one 64-bit atomic (add on a hit,
// CAS on a claim) on the same sector.  The count field sits at the TOP of the
// word: an atomicAdd that carries out of bit 63 simply wraps, and exactly one
// thread observes old-count + add >= 2^cbits; that thread records the carry in
// the exact 64-bit overflow side table and sets the ovf flag.  Counts are
// therefore exact (the reference chains overflow to 64 bits,
// JFI/large_hash_array.hpp:668-686).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace katgpu {

constexpr int kSlotsPerBucket = 4;
constexpr int kProbeBits = 6;
constexpr int kMaxProbes = (1 << kProbeBits) - 2; // j in [0, 61]; stored j+1 in [1, 62]
constexpr uint32_t kOvfCap = 1u << 16;            // overflow side-table entries
constexpr uint64_t kMixM1 = 0xff51afd7ed558ccdULL;
constexpr uint64_t kMixM2 = 0xc4ceb9fe1a85ec53ULL;
constexpr uint64_t kOwnerMul = 0x9E3779B97F4A7C15ULL;

// Device-side counters of one table (one 256-byte block in HBM).
struct TableCounters {
    unsigned long long n_distinct; // keys claimed
    unsigned long long n_total;    // k-mer instances inserted
    unsigned long long spill_n;    // keys that exhausted kMaxProbes (table full)
    unsigned long long ovf_n;      // entries used in the overflow side table
    unsigned int ovf_full;         // overflow side table exhausted
    unsigned int pad[7];
};

struct TableGeom {
    uint64_t *slots;   // 4 << Lb words
    uint64_t *ovf;     // kOvfCap x {h + 1, carries}
    uint64_t *spill;   // spill_cap x {h, count}
    TableCounters *ctr;
    uint64_t rmask;    // (1 << rbits) - 1
    uint64_t tagmask;  // (1 << (rbits + kProbeBits)) - 1
    uint64_t kmask;    // (1 << b) - 1
    uint64_t bmask;    // buckets - 1
    uint64_t cmax;     // (1 << cbits) - 1
    uint64_t m1inv, m2inv;
    uint64_t spill_cap;
    uint64_t add_bound; // upper bound on the increments any one slot can receive in this launch
    uint64_t red_below; // derived: a count field below this cannot carry in this launch (0 = never)
    uint32_t k, b, Lb, rbits, cshift, cbits, sh, canonical;
};

// Bijection on b-bit integers: xorshift (folds the first bases into the last), then an odd
// multiply mod 2^b.  Only the TOP bits of h choose the slice and the bucket, and those depend
// on every bit of the key; the low bits are merely stored (rem), so no second xorshift is
// needed (bucket occupancy and slice balance measured equal to the two-round finaliser on
// E. coli, random and tandem-repeat k-mers).  sh = ceil(b / 2) <= 32, so x >> sh fits the low
// word: the device fast path (extract.cuh: mix_words) is one funnel shift, one xor, a 64-bit
// multiply and a mask.
__host__ __device__ __forceinline__ uint64_t mix_key(const TableGeom &g, uint64_t x) {
    x ^= x >> g.sh;
    return (x * kMixM1) & g.kmask;
}
__host__ __device__ __forceinline__ uint64_t unmix_key(const TableGeom &g, uint64_t x) {
    x = (x * g.m1inv) & g.kmask;
    return x ^ (x >> g.sh); // 2*sh >= b: the xorshift is an involution
}
// Owner rank of a hashed key: independent of the bucket bits.
__host__ __device__ __forceinline__ uint32_t owner_of(uint64_t h, uint32_t n_ranks) {
    uint64_t y = (h * kOwnerMul) >> 32;
    return (uint32_t)((y * n_ranks) >> 32);
}
__host__ __device__ __forceinline__ uint64_t tri(uint32_t j) { return ((uint64_t)j * (j + 1)) >> 1; }

#ifdef __CUDACC__

// reverse complement of a k-mer held in the low 2k bits (first base most
// significant), mer_dna::reverse_complement (JFI/mer_dna.hpp:100-108,409-439).
__device__ __forceinline__ uint64_t revcomp_kmer(uint64_t fwd, uint32_t k) {
    uint64_t x = __brevll(~fwd);
    x = ((x & 0x5555555555555555ULL) << 1) | ((x >> 1) & 0x5555555555555555ULL);
    return x >> (64 - 2 * k);
}

struct Bucket {
    uint64_t s[kSlotsPerBucket];
};
// One 32-byte sector, L2-coherent (other SMs update it with atomics).
__device__ __forceinline__ Bucket load_bucket(const uint64_t *p) {
    Bucket r; // one 256-bit request (LDG.E.256, sm_100+): a single sector transaction per probe
    asm volatile("ld.global.cg.v4.u64 {%0, %1, %2, %3}, [%4];"
                 : "=l"(r.s[0]), "=l"(r.s[1]), "=l"(r.s[2]), "=l"(r.s[3])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void prefetch_bucket_l2(const uint64_t *p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// Record `carries` overflows of the count field for hashed key h.
static __device__ __noinline__ void ovf_add(const TableGeom &g, uint64_t h, uint64_t carries) {
    uint32_t i = (uint32_t)((h * kOwnerMul) >> 48) & (kOvfCap - 1);
    const unsigned long long want = h + 1;
    if (want == 0) { // h == 2^64-1 (only possible for k = 32): reserved last entry
        atomicAdd(reinterpret_cast<unsigned long long *>(g.ovf + 2 * (size_t)kOvfCap + 1), (unsigned long long)carries);
        return;
    }
    for (uint32_t n = 0; n < kOvfCap; n++, i = (i + 1) & (kOvfCap - 1)) {
        unsigned long long *e = reinterpret_cast<unsigned long long *>(g.ovf + 2 * (size_t)i);
        unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(e);
        if (cur == 0) {
            cur = atomicCAS(e, 0ULL, want);
            if (cur == 0) {
                atomicAdd(&g.ctr->ovf_n, 1ULL);
                cur = want;
            }
        }
        if (cur == want) {
            atomicAdd(e + 1, (unsigned long long)carries);
            return;
        }
    }
    atomicExch(&g.ctr->ovf_full, 1u);
}
static __device__ __noinline__ uint64_t ovf_get(const TableGeom &g, uint64_t h) {
    uint32_t i = (uint32_t)((h * kOwnerMul) >> 48) & (kOvfCap - 1);
    const uint64_t want = h + 1;
    if (want == 0) return g.ovf[2 * (size_t)kOvfCap + 1];
    for (uint32_t n = 0; n < kOvfCap; n++, i = (i + 1) & (kOvfCap - 1)) {
        uint64_t cur = g.ovf[2 * (size_t)i];
        if (cur == want) return g.ovf[2 * (size_t)i + 1];
        if (cur == 0) return 0;
    }
    return 0;
}

// Add `cnt` (< 2^cbits) to an occupied slot; handles the wrap exactly.
// `seen` is a value of the slot word read during this launch.  Counts only
// grow, so the field can end the launch no higher than field(seen) + add_bound:
// when that cannot carry, a fire-and-forget RED replaces the ATOM round trip.
__device__ __forceinline__ void slot_add(const TableGeom &g, uint64_t *slot, uint64_t h, uint64_t cnt,
                                         uint64_t seen) {
    if ((seen >> g.cshift) < g.red_below) {
        asm volatile("red.global.add.u64 [%0], %1;" ::"l"(slot), "l"(cnt << g.cshift) : "memory");
        return;
    }
    unsigned long long old = atomicAdd(reinterpret_cast<unsigned long long *>(slot),
                                       (unsigned long long)(cnt << g.cshift));
    if (((old >> g.cshift) + cnt) > g.cmax) { // this add carried out of the field
        atomicOr(reinterpret_cast<unsigned long long *>(slot), 1ULL << (g.rbits + kProbeBits));
        ovf_add(g, h, 1);
    }
}

// Insert hashed key h with multiplicity cnt.  Returns 1 if a new slot was
// claimed, 0 if the key existed, -1 if kMaxProbes buckets were full (spilled).
// `bk0` is the already-loaded home bucket (software pipelining by the caller).
__device__ __forceinline__ int insert_hashed(const TableGeom &g, uint64_t h, uint64_t cnt, Bucket bk,
                                             bool have_bk0) {
    const uint64_t b0 = h >> g.rbits;
    const uint64_t rem = h & g.rmask;
    uint64_t carries = cnt >> g.cbits; // only rehash/merge paths pass cnt > cmax
    const uint64_t low = cnt & g.cmax;
    for (uint32_t j = 0; j < (uint32_t)kMaxProbes; j++) {
        uint64_t *bp = g.slots + (((b0 + tri(j)) & g.bmask) << 2);
        if (j > 0 || !have_bk0) bk = load_bucket(bp);
        const uint64_t want = rem | ((uint64_t)(j + 1) << g.rbits);
#pragma unroll
        for (int s = 0; s < kSlotsPerBucket; s++) {
            uint64_t w = bk.s[s];
            if (w == 0) { // claim: slots fill in order, never empty again
                const uint64_t fresh = want | (low << g.cshift) |
                                       (carries ? (1ULL << (g.rbits + kProbeBits)) : 0ULL);
                unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long *>(bp + s), 0ULL,
                                                   (unsigned long long)fresh);
                if (old == 0) {
                    if (carries) ovf_add(g, h, carries);
                    return 1;
                }
                w = old;
            }
            if ((w & g.tagmask) == want) {
                if (low) slot_add(g, bp + s, h, low, w);
                if (carries) {
                    atomicOr(reinterpret_cast<unsigned long long *>(bp + s), 1ULL << (g.rbits + kProbeBits));
                    ovf_add(g, h, carries);
                }
                return 0;
            }
        }
    }
    // table full for this key: park it, the host grows the table and re-inserts
    unsigned long long at = atomicAdd(&g.ctr->spill_n, 1ULL);
    if (at < g.spill_cap) {
        g.spill[2 * at] = h;
        g.spill[2 * at + 1] = cnt;
    }
    return -1;
}

// Insertion with the home bucket already in registers — the hot path of every
// counting kernel.  Three tiers, the first two branch-free up to one predicated
// memory instruction so that a warp does not serialise over which slot matched:
//   1. key present in the home bucket  -> one RED (or ATOM when the count field
//      could carry in this launch);
//   2. key absent, home bucket has a free slot -> one CAS into the first free
//      slot (slots fill in order and never empty, so the key cannot live in a
//      later probe bucket);
//   3. anything else (lost CAS race, full home bucket) -> the general loop.
// Returns 1 if a new slot was claimed, 0 if the key existed, -1 if it was parked.
__device__ __forceinline__ int insert_home(const TableGeom &g, uint64_t h, const Bucket &bk) {
    const uint64_t want = (h & g.rmask) | (1ULL << g.rbits); // probe 0
    uint64_t *bp = g.slots + (((h >> g.rbits) & g.bmask) << 2);
    int hit = -1, emp = -1;
    uint64_t wsel = 0;
    if (g.tagmask <= 0xFFFFFFFFULL) { // tag (rem + probe) fits 32 bits: half the compare work (uniform branch)
        const uint32_t want32 = (uint32_t)want, mask32 = (uint32_t)g.tagmask;
#pragma unroll
        for (int s = kSlotsPerBucket - 1; s >= 0; s--) { // descending: the lowest index wins
            const uint64_t w = bk.s[s];
            const bool m = ((uint32_t)w & mask32) == want32;
            hit = m ? s : hit;
            wsel = m ? w : wsel;
            emp = (w == 0) ? s : emp;
        }
    } else {
#pragma unroll
        for (int s = kSlotsPerBucket - 1; s >= 0; s--) {
            const uint64_t w = bk.s[s];
            const bool m = (w & g.tagmask) == want;
            hit = m ? s : hit;
            wsel = m ? w : wsel;
            emp = (w == 0) ? s : emp;
        }
    }
    if (hit >= 0) {
        slot_add(g, bp + hit, h, 1, wsel);
        return 0;
    }
    if (emp >= 0) {
        const unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long *>(bp + emp), 0ULL,
                                                 (unsigned long long)(want | (1ULL << g.cshift)));
        if (old == 0) return 1;
        if ((old & g.tagmask) == want) {
            slot_add(g, bp + emp, h, 1, old);
            return 0;
        }
    }
    Bucket none = {};
    return insert_hashed(g, h, 1, none, false);
}

// Two-stage form of insert_home for software pipelining: stage 1 issues the one memory
// operation of the common cases (RED on a hit, CAS on a claim) without waiting for it;
// stage 2 looks at the CAS result once other keys' operations are in flight as well.
struct InsertTicket {
    unsigned long long old; // CAS result (state 1)
    uint32_t state;         // 0 done (existing key), 1 CAS issued, 2 needs the general path
    uint32_t slot;
};
__device__ __forceinline__ InsertTicket insert_issue(const TableGeom &g, uint64_t h, const Bucket &bk) {
    const uint64_t want = (h & g.rmask) | (1ULL << g.rbits);
    uint64_t *bp = g.slots + (((h >> g.rbits) & g.bmask) << 2);
    int hit = -1, emp = -1;
    uint64_t wsel = 0;
#pragma unroll
    for (int s = kSlotsPerBucket - 1; s >= 0; s--) {
        const uint64_t w = bk.s[s];
        const bool m = (w & g.tagmask) == want;
        hit = m ? s : hit;
        wsel = m ? w : wsel;
        emp = (w == 0) ? s : emp;
    }
    InsertTicket t;
    t.old = 0;
    t.slot = 0;
    t.state = 2;
    if (hit >= 0) {
        slot_add(g, bp + hit, h, 1, wsel);
        t.state = 0;
    } else if (emp >= 0) {
        t.old = atomicCAS(reinterpret_cast<unsigned long long *>(bp + emp), 0ULL,
                          (unsigned long long)(want | (1ULL << g.cshift)));
        t.state = 1;
        t.slot = (uint32_t)emp;
    }
    return t;
}
// Returns 1 if a new slot was claimed, 0 if the key existed, -1 if it was parked.
__device__ __forceinline__ int insert_finish(const TableGeom &g, uint64_t h, const InsertTicket &t) {
    if (t.state == 0) return 0;
    if (t.state == 1) {
        if (t.old == 0) return 1;
        const uint64_t want = (h & g.rmask) | (1ULL << g.rbits);
        if ((t.old & g.tagmask) == want) {
            slot_add(g, g.slots + (((h >> g.rbits) & g.bmask) << 2) + t.slot, h, 1, t.old);
            return 0;
        }
    }
    Bucket none = {};
    return insert_hashed(g, h, 1, none, false);
}

// Exact count of hashed key h (0 if absent).
__device__ __forceinline__ uint64_t lookup_hashed(const TableGeom &g, uint64_t h) {
    const uint64_t b0 = h >> g.rbits;
    const uint64_t rem = h & g.rmask;
    for (uint32_t j = 0; j < (uint32_t)kMaxProbes; j++) {
        const uint64_t *bp = g.slots + (((b0 + tri(j)) & g.bmask) << 2);
        const ulonglong2 a = __ldg(reinterpret_cast<const ulonglong2 *>(bp));
        const ulonglong2 c = __ldg(reinterpret_cast<const ulonglong2 *>(bp) + 1);
        const uint64_t w4[4] = {a.x, a.y, c.x, c.y};
        const uint64_t want = rem | ((uint64_t)(j + 1) << g.rbits);
#pragma unroll
        for (int s = 0; s < kSlotsPerBucket; s++) {
            if (w4[s] == 0) return 0;
            if ((w4[s] & g.tagmask) == want) {
                uint64_t cnt = w4[s] >> g.cshift;
                if (w4[s] & (1ULL << (g.rbits + kProbeBits))) cnt += ovf_get(g, h) << g.cbits;
                return cnt;
            }
        }
    }
    // a key parked in the spill list is invisible until the host re-inserts it;
    // the host never leaves spilled keys behind (count_end grows the table).
    return 0;
}

// Decode an occupied slot word found at slot index `idx` into (h, count).
__device__ __forceinline__ void decode_slot(const TableGeom &g, uint64_t idx, uint64_t w, uint64_t &h,
                                            uint64_t &cnt) {
    const uint32_t j = (uint32_t)((w >> g.rbits) & ((1u << kProbeBits) - 1)) - 1;
    const uint64_t b0 = ((idx >> 2) - tri(j)) & g.bmask;
    h = (b0 << g.rbits) | (w & g.rmask);
    cnt = w >> g.cshift;
    if (w & (1ULL << (g.rbits + kProbeBits))) cnt += ovf_get(g, h) << g.cbits;
}

#endif // __CUDACC__
} // namespace katgpu
