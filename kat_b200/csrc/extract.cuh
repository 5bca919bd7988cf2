// extract.cuh — device helpers shared by the extraction kernels: TMA bulk copy
// + mbarrier wrappers, ASCII -> 2-bit packing, k-mer cutting.
#pragma once
#include "table.cuh"

namespace katgpu {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier
// (SASS: UBLKCP).  dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---- pack phase: 16 ASCII bytes -> 32 bits of 2-bit codes + 16 invalid bits ----
// Code table of JFI/mer_dna.hpp:46-63: A/a=0 C/c=1 G/g=2 T/t=3, every other
// byte is "not a base".  First base of the group lands in the top bits, so a
// k-mer (first base most significant, mer_dna.hpp:330-353) is a plain bit
// field of the concatenated words.
__device__ __forceinline__ void pack4(uint32_t w, uint32_t &code8, uint32_t &bad4) {
    const uint32_t u = w & 0xDFDFDFDFu; // fold lower case
    const uint32_t valid = __vcmpeq4(u, 0x41414141u) | __vcmpeq4(u, 0x43434343u) |
                           __vcmpeq4(u, 0x47474747u) | __vcmpeq4(u, 0x54545454u);
    const uint32_t c = ((w >> 1) ^ (w >> 2)) & 0x03030303u;
    code8 = (c * 0x40100401u) >> 24;                               // c0<<6 | c1<<4 | c2<<2 | c3
    bad4 = (((~valid) & 0x01010101u) * 0x08040201u) >> 24 & 0xFu; // b0<<3 | b1<<2 | b2<<1 | b3
}

// Cut the k-mer starting r bases into the 48-base window (hi:lo) and test the
// invalid bits of its k bases.
__device__ __forceinline__ bool cut_kmer(uint64_t hi, uint64_t lo, uint64_t bad48, uint32_t r, uint32_t k,
                                         uint64_t &fwd) {
    const uint64_t win = r ? ((hi << (2 * r)) | (lo >> (64 - 2 * r))) : hi;
    fwd = win >> (64 - 2 * k);
    const uint64_t badk = (bad48 << (16 + r)) >> (64 - k);
    return badk == 0;
}



// 16 valid-start flags of one thread at once: start r is a k-mer iff none of the k
// invalid bits [r, r+k) is set.  OR over a window of k by doubling (bit 47-t of bad48 =
// base t).  Returns bit (15 - r) = start r is valid.
__device__ __forceinline__ uint32_t valid_starts16(uint64_t bad48, uint32_t k) {
    uint64_t w = bad48; // window 1
    uint32_t len = 1;
#pragma unroll
    for (int d = 0; d < 5; d++) {
        if (2 * len <= k) {
            w |= w << len;
            len *= 2;
        }
    }
    w |= w << (k - len); // two windows of `len` cover a window of k (len <= k < 2 len)
    return (uint32_t)(~w >> 32) & 0xFFFFu;
}

// Rolling extraction state of one thread: the 16 k-mers starting at the 16 bases of its
// group, forward and reverse complement (JFI/mer_dna.hpp:330-378 shift_left/shift_right).
struct Roll16 {
    uint64_t fwd, rc, nextw;
    uint32_t okm, topsh;
    __device__ __forceinline__ void init(uint64_t hi, uint64_t lo, uint64_t bad48, uint32_t k, uint32_t nst) {
        okm = valid_starts16(bad48, k);
        okm &= nst >= 16 ? 0xFFFFu : (0xFFFFu << (16 - nst)) & 0xFFFFu;
        nextw = k < 32 ? ((hi << (2 * k)) | (lo >> (64 - 2 * k))) : lo; // bases k .. k+15, base k on top
        fwd = hi >> (64 - 2 * k);
        rc = revcomp_kmer(fwd, k);
        topsh = 2 * k - 2;
    }
    // advance from start r-1 to start r (r >= 1): base r+k-1 enters, base r-1 leaves
    __device__ __forceinline__ void step(uint32_t r, uint64_t kmask) {
        const uint64_t c = (nextw >> (64 - 2 * r)) & 3ULL;
        fwd = ((fwd << 2) | c) & kmask;
        rc = (rc >> 2) | ((c ^ 3ULL) << topsh);
    }
    __device__ __forceinline__ bool ok(uint32_t r) const { return (okm >> (15 - r)) & 1u; }
    __device__ __forceinline__ uint64_t key(bool canonical) const { return (canonical && rc < fwd) ? rc : fwd; }
};

// ---- window extraction (partition / route kernels) ------------------------------------
// A thread owns the 16 k-mer starts of one 16-base group and sees the 48-base window
// w2:w1:w0 (w2 = bases 0-15, first base in the top bits).  Instead of rolling one k-mer into
// the next (a serial chain of 64-bit shifts by run-time amounts), both strands are prepared
// once as 96-bit strings from which k-mer r is a CONSTANT funnel shift away:
//   F = W >> 2(33-k)            fwd(r) = (F >> 2(15-r)) & kmask
//   R = revcomp(W) (base 0 low) rc(r)  = (R >> 2r)      & kmask
// (JFI/mer_dna.hpp:330-378 shift_left/shift_right define the two strands; 3 instructions per
// strand per k-mer instead of ~9, and the 16 k-mers are independent.)
__device__ __forceinline__ uint32_t revcomp16(uint32_t w) { // 16 bases: reverse the order, complement
    const uint32_t x = __brev(~w);                          // bit pairs now in place but bit-swapped
    return ((x << 1) & 0xAAAAAAAAu) | ((x >> 1) & 0x55555555u);
}
struct Cut16 {
    uint32_t f0, f1, f2, r0, r1, r2, okm;
    __device__ __forceinline__ void init(uint32_t w2, uint32_t w1, uint32_t w0, uint64_t bad48, uint32_t k, uint32_t nst) {
        okm = valid_starts16(bad48, k);
        okm &= nst >= 16 ? 0xFFFFu : (0xFFFFu << (16 - nst)) & 0xFFFFu;
        r0 = revcomp16(w2);
        r1 = revcomp16(w1);
        r2 = revcomp16(w0);
        const uint32_t s = 2 * (33 - k); // 2 .. 64 (uniform)
        uint32_t a0 = w0, a1 = w1, a2 = w2;
        if (s >= 64) {
            a0 = w2, a1 = 0, a2 = 0;
        } else if (s >= 32) {
            a0 = w1, a1 = w2, a2 = 0;
        }
        f0 = __funnelshift_r(a0, a1, s);
        f1 = __funnelshift_r(a1, a2, s);
        f2 = a2 >> (s & 31);
    }
    __device__ __forceinline__ bool ok(uint32_t r) const { return (okm >> (15 - r)) & 1u; }
    // canonical (or forward) k-mer r as two words; r is a compile-time constant after unrolling
    template <bool CANON, bool WIDE>
    __device__ __forceinline__ void key(uint32_t r, uint32_t kmask_lo, uint32_t kmask_hi, uint32_t &lo, uint32_t &hi) const {
        uint32_t flo = __funnelshift_r(f0, f1, 2 * (15 - r));
        const uint32_t fhi = __funnelshift_r(f1, f2, 2 * (15 - r)) & kmask_hi;
        if (!WIDE) flo &= kmask_lo;
        lo = flo, hi = fhi;
        if (CANON) {
            uint32_t rlo = __funnelshift_r(r0, r1, 2 * r);
            const uint32_t rhi = __funnelshift_r(r1, r2, 2 * r) & kmask_hi;
            if (!WIDE) rlo &= kmask_lo;
            const bool rc_smaller = rhi < fhi || (rhi == fhi && rlo < flo); // mer_dna::operator< (JFI/mer_dna.hpp:235-258)
            lo = rc_smaller ? rlo : flo;
            hi = rc_smaller ? rhi : fhi;
        }
    }
};
// mix_key (table.cuh) on a key held as two words; sh = g.sh in [1, 32].
template <bool WIDE>
__device__ __forceinline__ uint64_t mix_words(uint32_t lo, uint32_t hi, uint32_t sh, uint32_t kmask_lo, uint32_t kmask_hi) {
    lo ^= __funnelshift_rc(lo, hi, sh); // low word of x >> sh (the high word of x >> sh is 0: b - sh <= 32)
    const uint64_t x = (((uint64_t)hi << 32) | lo) * kMixM1;
    uint32_t xl = (uint32_t)x;
    const uint32_t xh = (uint32_t)(x >> 32) & kmask_hi;
    if (!WIDE) xl &= kmask_lo;
    return ((uint64_t)xh << 32) | xl;
}

} // namespace katgpu
