// wide.cuh — k-mers of 33..64 bases: two 64-bit words per key (the reference keeps ceil(2k/64) words per
// mer_dna, JFI/mer_dna.hpp:513-531, and its hash stores keys of any width; KAT itself has no k limit).
//
// The one-word table of table.cuh packs remainder, probe index and count into 64 bits; that does not scale to
// 128-bit keys, so wide tables have their own, simpler layout, chosen for 16-byte atomics (sm_90+):
//
//   keys  [n_slots + 1]  16 bytes each: {lo = last 32 bases, hi = the first k-32 bases, 2 bits per base, first base
//                        most significant — mer_dna's packing}.  Empty = all ones.  A slot is claimed with ONE
//                        128-bit compare-and-swap (atom.cas.b128), so a key is never visible half-written.
//   counts[n_slots + 1]  exact 64-bit counts, red.global.add.u64 — no overflow side table needed.
//
// Open addressing with triangular probing over a power-of-two table (visits every slot).  The all-ones word pair is
// also a legal key when k = 64 (T x 64): that one k-mer lives in the extra slot n_slots, whose key field is never
// written and therefore reads as exactly that key.  Occupancy is `count != 0` everywhere, so the walks treat the extra
// slot like any other.
//
// Everything in this header is __host__ __device__ and free of warp-level primitives, so that the rolling
// extraction, canonicalisation, hashing and the insert / lookup protocol can be run on the CPU against the oracle
// (tests/hostcheck/wide_emul.cc, single-threaded stand-ins for the atomics) before they ever run on a GPU.
#pragma once
#include <cstdint>
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define WD_HD __host__ __device__ __forceinline__
#else
#define WD_HD inline
#endif

namespace katgpu {

struct __attribute__((aligned(16))) Key128 {
    uint64_t lo, hi;
};
WD_HD bool operator==(const Key128 &a, const Key128 &b) { return a.lo == b.lo && a.hi == b.hi; }
WD_HD bool key_less(const Key128 &a, const Key128 &b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); } // mer_dna::operator<, JFI/mer_dna.hpp:235-258
WD_HD bool key_is_empty_marker(const Key128 &a) { return (a.lo & a.hi) == ~0ULL; }

struct WideCounters {
    unsigned long long claimed; // slots claimed (distinct keys, not counting the extra slot)
    unsigned long long total;   // k-mer instances added
    unsigned int failed;        // a probe sequence ran out (cannot happen below the admitted load)
    unsigned int pad[11];
};

struct WideGeom {
    Key128 *keys;
    unsigned long long *counts;
    WideCounters *ctr;
    uint64_t n_slots, mask; // power of two, mask = n_slots - 1
    uint64_t hi_mask;       // the 2k-64 bits of the hi word
    uint32_t k, canonical;
    uint32_t hi_top;        // 2k - 66: bit position in hi of the first base
    uint32_t rc_shift;      // 128 - 2k
    uint32_t max_probes;
};

WD_HD uint64_t wide_tri(uint64_t j) { return (j * (j + 1)) >> 1; }
WD_HD uint64_t wide_mix64(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}
WD_HD uint64_t wide_hash(const Key128 &a) { return wide_mix64(a.lo ^ wide_mix64(a.hi + 0x9E3779B97F4A7C15ULL)); }

// Rolling k-mer and its reverse complement (mer_iterator::operator++, JFI/mer_iterator.hpp:61-89;
// mer_dna::shift_left / shift_right, JFI/mer_dna.hpp:330-378; code table :46-63).
struct WideRoll {
    Key128 f, r;
    uint32_t filled;
};
WD_HD void wide_push(const WideGeom &g, WideRoll &s, uint32_t ch) {
    const uint32_t u = ch & 0xDFu; // fold lower case
    if (u == 0x41u || u == 0x43u || u == 0x47u || u == 0x54u) {
        const uint64_t c = ((ch >> 1) ^ (ch >> 2)) & 3u; // A 0, C 1, G 2, T 3
        s.f.hi = ((s.f.hi << 2) | (s.f.lo >> 62)) & g.hi_mask;
        s.f.lo = (s.f.lo << 2) | c;
        s.r.lo = (s.r.lo >> 2) | (s.r.hi << 62);
        s.r.hi = (s.r.hi >> 2) | ((3u - c) << g.hi_top);
        if (s.filled < g.k) s.filled++;
    } else
        s.filled = 0; // N, IUPAC codes, '\r', anything else: the run starts over
}
WD_HD Key128 wide_current(const WideGeom &g, const WideRoll &s) { return (g.canonical && key_less(s.r, s.f)) ? s.r : s.f; }

// Every k-mer that STARTS in [0, n_starts) of the segment b[0 .. n_starts + k - 2]; bytes that are not bases
// (including the zero padding past the end of the stream) break the run.  byte_at(i) fetches byte i.
template <class ByteAt, class Emit>
WD_HD void wide_roll_segment(const WideGeom &g, uint32_t n_starts, ByteAt byte_at, Emit emit) {
    if (n_starts == 0) return;
    WideRoll s;
    s.f.lo = s.f.hi = s.r.lo = s.r.hi = 0;
    s.filled = 0;
    const uint32_t end = n_starts + g.k - 1;
    for (uint32_t i = 0; i < end; i++) {
        wide_push(g, s, byte_at(i));
        if (s.filled >= g.k) emit(wide_current(g, s)); // starts at i - k + 1, which is >= 0 and < n_starts
    }
}

WD_HD uint64_t wide_rev_groups(uint64_t x) { // reverse the order of the 32 two-bit groups of a word
#if defined(__CUDA_ARCH__)
    x = __brevll(x);
#else
    x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
    x = ((x >> 8) & 0x00FF00FF00FF00FFULL) | ((x & 0x00FF00FF00FF00FFULL) << 8);
    x = ((x >> 16) & 0x0000FFFF0000FFFFULL) | ((x & 0x0000FFFF0000FFFFULL) << 16);
    x = (x >> 32) | (x << 32);
#endif
    return ((x & 0x5555555555555555ULL) << 1) | ((x >> 1) & 0x5555555555555555ULL); // bit reversal swapped the two bits of every group
}
// mer_dna::reverse_complement (JFI/mer_dna.hpp:409-439) of a key held in the low 2k bits of (hi:lo)
WD_HD Key128 wide_revcomp(const WideGeom &g, const Key128 &a) {
    const uint64_t rh = wide_rev_groups(~a.lo), rl = wide_rev_groups(~a.hi);
    const uint32_t s = g.rc_shift; // 0 .. 62
    Key128 o;
    o.lo = s ? ((rl >> s) | (rh << (64 - s))) : rl;
    o.hi = rh >> s;
    return o;
}
WD_HD Key128 wide_canonical(const WideGeom &g, const Key128 &a) {
    const Key128 r = wide_revcomp(g, a);
    return key_less(r, a) ? r : a;
}
// gcCount (lib/include/kat/str_utils.hpp:151-161): C = 01, G = 10 — the two bits of the code differ
WD_HD uint32_t wide_gc(const Key128 &a) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)(__popcll((a.lo ^ (a.lo >> 1)) & 0x5555555555555555ULL) + __popcll((a.hi ^ (a.hi >> 1)) & 0x5555555555555555ULL));
#else
    return (uint32_t)(__builtin_popcountll((a.lo ^ (a.lo >> 1)) & 0x5555555555555555ULL) + __builtin_popcountll((a.hi ^ (a.hi >> 1)) & 0x5555555555555555ULL));
#endif
}

// ---- memory primitives: the GPU's, or single-threaded stand-ins for the CPU emulation --------------------------
WD_HD Key128 wd_load_key(const Key128 *p) {
#if defined(__CUDA_ARCH__)
    Key128 r; // one 16-byte request, L2-coherent (other SMs claim slots with atomics)
    asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(r.lo), "=l"(r.hi) : "l"(p));
    return r;
#else
    return *p;
#endif
}
// 128-bit compare-and-swap: returns what the slot held; it holds `desired` afterwards iff that was `expected`.
WD_HD Key128 wd_cas_key(Key128 *p, const Key128 &expected, const Key128 &desired) {
#if defined(__CUDA_ARCH__)
    Key128 old;
    asm volatile("{\n\t.reg .b128 c, v, d;\n\t"
                 "mov.b128 c, {%3, %4};\n\t"
                 "mov.b128 v, {%5, %6};\n\t"
                 "atom.global.cas.b128 d, [%2], c, v;\n\t"
                 "mov.b128 {%0, %1}, d;\n\t}"
                 : "=l"(old.lo), "=l"(old.hi)
                 : "l"(p), "l"(expected.lo), "l"(expected.hi), "l"(desired.lo), "l"(desired.hi)
                 : "memory");
    return old;
#else
    const Key128 old = *p;
    if (old == expected) *p = desired;
    return old;
#endif
}
WD_HD void wd_add(unsigned long long *p, unsigned long long v) {
#if defined(__CUDA_ARCH__)
    asm volatile("red.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#else
    *p += v;
#endif
}
WD_HD void wd_flag_failure(const WideGeom &g) {
#if defined(__CUDA_ARCH__)
    atomicExch(&g.ctr->failed, 1u);
#else
    g.ctr->failed = 1u;
#endif
}

// count[key] += cnt, inserting the key if absent (hash_counter::add, JFI/hash_counter.hpp:98-130).
// Returns 1 when a slot was claimed, 0 when the key existed (or is the extra slot's), -1 on failure.
WD_HD int wide_insert(const WideGeom &g, const Key128 &key, unsigned long long cnt) {
    if (key_is_empty_marker(key)) {
        wd_add(g.counts + g.n_slots, cnt);
        return 0;
    }
    const Key128 empty = {~0ULL, ~0ULL};
    const uint64_t h = wide_hash(key);
    for (uint64_t j = 0; j < g.max_probes; j++) {
        const uint64_t i = (h + wide_tri(j)) & g.mask;
        Key128 cur = wd_load_key(g.keys + i);
        if (!(cur == key)) {
            // a word of all ones: the slot is empty — or a 16-byte load saw a claim half-way; either way let the
            // atomic decide (it claims the slot when empty and returns what is really there otherwise)
            if (cur.lo != ~0ULL && cur.hi != ~0ULL) continue;
            cur = wd_cas_key(g.keys + i, empty, key);
            if (cur == empty) {
                wd_add(g.counts + i, cnt);
                return 1;
            }
            if (!(cur == key)) continue;
        }
        wd_add(g.counts + i, cnt);
        return 0;
    }
    wd_flag_failure(g);
    return -1;
}

// Exact count of `key`, 0 if absent (array_base::get_val_for_key, JFI/large_hash_array.hpp:358-376).  Only used on
// finished tables, so plain loads suffice.
WD_HD unsigned long long wide_lookup(const WideGeom &g, const Key128 &key) {
    if (key_is_empty_marker(key)) return g.counts[g.n_slots];
    const uint64_t h = wide_hash(key);
    for (uint64_t j = 0; j < g.max_probes; j++) {
        const uint64_t i = (h + wide_tri(j)) & g.mask;
        const Key128 cur = g.keys[i];
        if (cur == key) return g.counts[i];
        if (key_is_empty_marker(cur)) return 0;
    }
    return 0;
}

} // namespace katgpu
