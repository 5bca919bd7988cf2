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

} // namespace katgpu
