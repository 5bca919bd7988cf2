// parse_kernels.cu — FASTQ record parsing on the GPU: raw file bytes in
// HBM become a base stream IN PLACE, so that the host never touches the bulk of a read file.
//
// The reference parser (JFI/mer_overlap_sequence_parser.hpp:218-289, read_fastq / read_sequence /
// skip_quals) hands mer_iterator the sequence lines with one 'N' between records; every byte that
// is not ACGTacgt breaks a k-mer (JFI/mer_dna.hpp:46-63).  For the regular layout — four
// '\n'-terminated lines per record: '@' header, sequence, '+' line, quality of the same length — the
// k-mers of that stream are exactly the k-mers of the file image with every header, '+' and quality
// byte (and the first `trim5p` sequence bytes, :228,248-253) overwritten by a non-base byte.  That
// is what these kernels do; the preconditions are those of the host fast path
// (host/seq_parser.cc: fastq_piece) and EVERY one is checked on the device — any violation raises a
// flag, nothing is counted, and the caller falls back to the host's exact emulation of the
// reference parser.
//
//   nl_count_kernel    newlines per 4 KiB tile (SIMD byte compares)
//   nl_scan_kernel     exclusive prefix over the tiles (one CTA)
//   nl_scatter_kernel  position of every newline, in order
//   fastq_mask_kernel  one warp per record: validate the four lines, blank everything but the bases
#include "parse_kernels.cuh"

namespace katgpu {

constexpr int kNlTile = 4096, kNlThreads = 256; // 16 bytes per thread

__device__ __forceinline__ uint32_t count_nl4(uint32_t w) { return (uint32_t)__popc(__vcmpeq4(w, 0x0A0A0A0Au)) >> 3; }

// the 16 bytes of thread `tid` of tile `tile`; bytes at or past n read as 0
__device__ __forceinline__ uint4 load16_guarded(const uint8_t *data, uint64_t n, uint64_t off) {
    if (off + 16 <= n) return *reinterpret_cast<const uint4 *>(data + off);
    uint32_t w[4] = {0, 0, 0, 0};
    for (int j = 0; j < 16; j++)
        if (off + j < n) w[j >> 2] |= (uint32_t)data[off + j] << (8 * (j & 3));
    return make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *warp_sums, uint32_t &block_total) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if ((int)lane >= d) incl += o;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    uint32_t base = 0, tot = 0;
    for (uint32_t w = 0; w < blockDim.x / 32; w++) {
        const uint32_t s = warp_sums[w];
        if (w < warp) base += s;
        tot += s;
    }
    block_total = tot;
    __syncthreads();
    return base + incl - v;
}

__global__ void __launch_bounds__(kNlThreads) nl_count_kernel(const uint8_t *__restrict__ data, uint64_t n, uint32_t *__restrict__ tile_counts) {
    __shared__ uint32_t ws[kNlThreads / 32];
    const uint64_t off = (uint64_t)blockIdx.x * kNlTile + (uint64_t)threadIdx.x * 16;
    const uint4 v = load16_guarded(data, n, off);
    uint32_t c = count_nl4(v.x) + count_nl4(v.y) + count_nl4(v.z) + count_nl4(v.w), tot;
    block_exclusive_scan(c, ws, tot);
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = tot;
}

// in place: tile_counts[i] <- sum of tile_counts[0 .. i); *total <- everything
__global__ void __launch_bounds__(1024) nl_scan_kernel(uint32_t *tile_counts, uint32_t n_tiles, uint32_t *total) {
    __shared__ uint32_t ws[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_tiles; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n_tiles ? tile_counts[i] : 0;
        uint32_t tot;
        const uint32_t ex = block_exclusive_scan(v, ws, tot);
        if (i < n_tiles) tile_counts[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(kNlThreads) nl_scatter_kernel(const uint8_t *__restrict__ data, uint64_t n, const uint32_t *__restrict__ tile_prefix,
                                                                uint32_t *__restrict__ nl, uint32_t nl_cap) {
    __shared__ uint32_t ws[kNlThreads / 32];
    const uint64_t off = (uint64_t)blockIdx.x * kNlTile + (uint64_t)threadIdx.x * 16;
    const uint4 v = load16_guarded(data, n, off);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    const uint32_t c = count_nl4(v.x) + count_nl4(v.y) + count_nl4(v.z) + count_nl4(v.w);
    uint32_t tot;
    uint32_t at = tile_prefix[blockIdx.x] + block_exclusive_scan(c, ws, tot);
    if (c == 0) return;
#pragma unroll
    for (int j = 0; j < 16; j++)
        if (((w[j >> 2] >> (8 * (j & 3))) & 0xFFu) == 0x0Au) {
            if (at < nl_cap) nl[at] = (uint32_t)(off + j);
            at++;
        }
}

// One warp per record r: lines 4r .. 4r+3 end at nl[4r] .. nl[4r+3].
__global__ void __launch_bounds__(256) fastq_mask_kernel(uint8_t *__restrict__ data, const uint32_t *__restrict__ nl, uint32_t n_records,
                                                         uint32_t trim, uint32_t *__restrict__ flags) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    bool bad = false;
    for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_records; r += warps) {
        uint32_t p = 0;
        if (lane < 4) p = nl[4 * r + lane];
        else if (lane == 4) p = r ? nl[4 * r - 1] + 1 : 0;
        const uint32_t e0 = __shfl_sync(0xffffffffu, p, 0), e1 = __shfl_sync(0xffffffffu, p, 1), e2 = __shfl_sync(0xffffffffu, p, 2),
                       e3 = __shfl_sync(0xffffffffu, p, 3), s0 = __shfl_sync(0xffffffffu, p, 4);
        const uint32_t seq = e0 + 1, slen = e1 - seq, plus = e1 + 1, qual = e2 + 1, qlen = e3 - qual;
        // host/seq_parser.cc fastq_piece: '@' header, sequence longer than the trim, not starting with '@' or '+',
        // no '+' inside, a '+' line, quality of the sequence's length
        bool ok = data[s0] == '@' && data[plus] == '+' && slen > trim && slen == qlen;
        if (ok) {
            const uint8_t c0 = data[seq];
            ok = c0 != '@' && c0 != '+';
        }
        bool plus_inside = false;
        for (uint32_t i = seq + lane; i < e1; i += 32) plus_inside |= data[i] == '+';
        ok = ok && !__any_sync(0xffffffffu, plus_inside);
        bad |= !ok;
        // blank the header (its '\n' stays), the trimmed bases, the '+' line and the quality line
        const uint32_t head_end = ok ? seq + trim : seq;
        for (uint32_t i = s0 + lane; i < head_end; i += 32) data[i] = '\n';
        for (uint32_t i = plus + lane; i < e3; i += 32) data[i] = '\n';
    }
    if (bad) atomicOr(flags, 1u);
}

void launch_nl_count(const uint8_t *data, uint64_t n, uint32_t *tile_counts, uint32_t n_tiles, cudaStream_t st) {
    nl_count_kernel<<<n_tiles, kNlThreads, 0, st>>>(data, n, tile_counts);
}
void launch_nl_scan(uint32_t *tile_counts, uint32_t n_tiles, uint32_t *total, cudaStream_t st) {
    nl_scan_kernel<<<1, 1024, 0, st>>>(tile_counts, n_tiles, total);
}
void launch_nl_scatter(const uint8_t *data, uint64_t n, const uint32_t *tile_prefix, uint32_t n_tiles, uint32_t *nl, uint32_t nl_cap, cudaStream_t st) {
    nl_scatter_kernel<<<n_tiles, kNlThreads, 0, st>>>(data, n, tile_prefix, nl, nl_cap);
}
void launch_fastq_mask(uint8_t *data, const uint32_t *nl, uint32_t n_records, uint32_t trim, uint32_t *flags, int grid, cudaStream_t st) {
    fastq_mask_kernel<<<grid, 256, 0, st>>>(data, nl, n_records, trim, flags);
}
uint32_t nl_tiles(uint64_t n) { return (uint32_t)((n + kNlTile - 1) / kNlTile); }

} // namespace katgpu
