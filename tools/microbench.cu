// microbench.cu — primitive rates of the B200 memory system that bound the counting path
// (DESIGN.md §3): random 64/32-bit REDs, random 32-byte bucket loads, 8-byte vs 32-byte
// scattered stores, shared-memory atomics.  Standalone (no torch): nvcc -arch=sm_100a.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint64_t rng(uint64_t &s) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; uint64_t x = s; x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 29; return x; }

// mode 0: red.u64  1: red.u32  2: ld.v4.u64 (32 B)  3: ld 32 B + red.u64 (the sweep's pair)  4: atom.cas.u64
// 5: st.u64 scattered  6: st.v4.u64 (32 B) scattered  7: ld 32 B + plain st.u64 (non-atomic RMW)
__global__ void __launch_bounds__(256) gmem_kernel(uint64_t *buf, uint64_t window_words, uint64_t n_windows, int iters, int mode, uint64_t *sink) {
    uint64_t s = (blockIdx.x * 256ull + threadIdx.x) * 0x9E3779B97F4A7C15ULL + 12345;
    uint64_t acc = 0;
    for (uint64_t w = 0; w < n_windows; w++) {
        uint64_t *base = buf + w * window_words;
        for (int i = 0; i < iters; i += 4) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint64_t r = rng(s);
                uint64_t *p = base + __umul64hi(r << 2 | 3, window_words / 4) * 4; // low 2 bits of r pick the slot below // 32-byte aligned bucket
                if (mode == 0) asm volatile("red.global.add.u64 [%0], %1;" ::"l"(p + (r & 3)), "l"(1ULL << 33) : "memory");
                else if (mode == 1) asm volatile("red.global.add.u32 [%0], %1;" ::"l"(reinterpret_cast<uint32_t *>(p + (r & 3)) + 1), "r"(1u) : "memory");
                else if (mode == 2 || mode == 3 || mode == 7) {
                    uint64_t a, b, c, d;
                    asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
                    acc += a ^ b ^ c ^ d;
                    if (mode == 3) asm volatile("red.global.add.u64 [%0], %1;" ::"l"(p + (r & 3)), "l"(1ULL << 33) : "memory");
                    if (mode == 7) asm volatile("st.global.cg.u64 [%0], %1;" ::"l"(p + (r & 3)), "l"(a + (1ULL << 33)) : "memory");
                } else if (mode == 4) acc += atomicCAS(reinterpret_cast<unsigned long long *>(p + (r & 3)), 0ULL, (unsigned long long)r | 1);
                else if (mode == 5) asm volatile("st.global.cg.u64 [%0], %1;" ::"l"(p + (r & 3)), "l"(r) : "memory");
                else if (mode == 6) asm volatile("st.global.cg.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(r), "l"(r + 1), "l"(r + 2), "l"(r + 3) : "memory");
            }
        }
    }
    if (acc == 0x1234567) *sink = acc;
}

// Scattered writes of whole pieces: `lanes` adjacent lanes (1, 2, 4) write the consecutive 32-byte
// sectors of one random 32/64/128-byte piece in ONE store instruction (coalesced by the LSU);
// mode 1: one thread writes the 2 or 4 sectors of its piece in consecutive instructions instead.
__global__ void __launch_bounds__(256) piece_kernel(uint64_t *buf, uint64_t window_words, int iters, int lanes, int mode, int hint) {
    const uint32_t lane = threadIdx.x & 31, sub = lane % lanes;
    uint64_t s = ((blockIdx.x * 256ull + threadIdx.x) / lanes) * 0x9E3779B97F4A7C15ULL + 999; // same stream for the lanes of a group
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    const uint64_t n_pieces = window_words / (4ull * (mode ? lanes : lanes));
    for (int i = 0; i < iters; i++) {
        const uint64_t r = rng(s);
        uint64_t *p = buf + __umul64hi(r, n_pieces) * 4ull * lanes;
        if (mode == 0) {
            uint64_t *q = p + 4 * sub;
            if (hint) asm volatile("st.global.L2::cache_hint.v4.u64 [%0], {%1,%2,%3,%4}, %5;" ::"l"(q), "l"(r), "l"(r + 1), "l"(r + 2), "l"(r + 3), "l"(pol) : "memory");
            else asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(q), "l"(r), "l"(r + 1), "l"(r + 2), "l"(r + 3) : "memory");
        } else if (sub == 0) {
            for (int j = 0; j < lanes; j++) {
                uint64_t *q = p + 4 * j;
                if (hint) asm volatile("st.global.L2::cache_hint.v4.u64 [%0], {%1,%2,%3,%4}, %5;" ::"l"(q), "l"(r), "l"(r + 1), "l"(r + 2), "l"(r + 3), "l"(pol) : "memory");
                else asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(q), "l"(r), "l"(r + 1), "l"(r + 2), "l"(r + 3) : "memory");
            }
        }
    }
}

// shared-memory atomics: mode 0 atomicAdd u32 random, 1 atomicAdd u64 random, 2 ticket (st + ld + ld + st), 3 plain ld+st u32
__global__ void __launch_bounds__(512) smem_kernel(int iters, int mode, int words, uint64_t *sink) {
    extern __shared__ uint32_t sm[];
    for (int i = threadIdx.x; i < words; i += 512) sm[i] = 0;
    __syncthreads();
    uint64_t s = (blockIdx.x * 512ull + threadIdx.x) * 0x9E3779B97F4A7C15ULL + 777;
    uint32_t acc = 0;
    for (int i = 0; i < iters; i++) {
        const uint32_t r = (uint32_t)(rng(s) >> 20);
#define PICK(n) __umulhi(r, (uint32_t)(n))
        if (mode == 0) acc += atomicAdd(&sm[PICK(words)], 1u);
        else if (mode == 1) acc += (uint32_t)atomicAdd(reinterpret_cast<unsigned long long *>(sm) + PICK(words / 2), 1ULL);
        else if (mode == 2) {
            volatile uint32_t *t = sm; const uint32_t a = PICK(words / 2);
            t[a] = threadIdx.x; __syncwarp();
            const uint32_t win = t[a]; const uint32_t c = t[words / 2 + a];
            if (win == threadIdx.x) t[words / 2 + a] = c + 1;
            acc += c;
        } else { volatile uint32_t *t = sm; const uint32_t a = PICK(words); const uint32_t c = t[a]; t[a] = c + 1; acc += c; }
    }
    if (acc == 0x1234567) *sink = acc;
}

int main() {
    int dev = 0; cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, dev));
    printf("device %s, %d SMs, L2 %d MB\n", pr.name, pr.multiProcessorCount, pr.l2CacheSize >> 20);
    const uint64_t total_bytes = 8ull << 30; uint64_t *buf, *sink; CK(cudaMalloc(&buf, total_bytes)); CK(cudaMalloc(&sink, 8)); CK(cudaMemset(buf, 0, total_bytes));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const char *names[] = {"red.u64", "red.u32", "ld 32B", "ld 32B + red.u64", "atom.cas.u64", "st.u64 (8 B scattered)", "st.v4.u64 (32 B scattered)", "ld 32B + st.u64"};
    const int grid = pr.multiProcessorCount * 6;
    for (uint64_t win_mb : {16ull, 32ull, 64ull, 8192ull}) {
        const uint64_t window_words = (win_mb << 20) / 8, n_windows = win_mb >= 8192 ? 1 : 8;
        for (int mode = 0; mode < 8; mode++) {
            const int iters = win_mb >= 8192 ? 512 : 128;
            gmem_kernel<<<grid, 256>>>(buf, window_words, n_windows, iters, mode, sink); // warm
            cudaEventRecord(a);
            gmem_kernel<<<grid, 256>>>(buf, window_words, n_windows, iters, mode, sink);
            cudaEventRecord(b); CK(cudaEventSynchronize(b));
            float ms; cudaEventElapsedTime(&ms, a, b);
            const double ops = (double)grid * 256 * iters * n_windows;
            printf("window %5llu MB  %-28s %8.2f G ops/s  (%.2f ms)\n", (unsigned long long)win_mb, names[mode], ops / ms / 1e6, ms);
        }
    }
    for (int hint = 0; hint < 2; hint++)
        for (int mode = 0; mode < 2; mode++)
            for (int lanes : {1, 2, 4}) {
                if (mode == 1 && lanes == 1) continue;
                const uint64_t window_words = total_bytes / 8;
                const int iters = 256;
                piece_kernel<<<grid, 256>>>(buf, window_words, iters, lanes, mode, hint);
                cudaEventRecord(a);
                piece_kernel<<<grid, 256>>>(buf, window_words, iters, lanes, mode, hint);
                cudaEventRecord(b); CK(cudaEventSynchronize(b));
                float ms; cudaEventElapsedTime(&ms, a, b);
                const double sectors = (double)grid * 256 * iters * (mode ? 1.0 : 1.0) * (mode ? (double)lanes / lanes : 1.0);
                const double bytes = (mode == 0 ? (double)grid * 256 * iters * 32.0 : (double)grid * 256 / lanes * iters * 32.0 * lanes);
                printf("scattered %3d-byte pieces over 8 GB, %s%s: %8.2f G sectors/s = %6.0f GB/s (%.2f ms)\n", 32 * lanes,
                       mode ? "one thread, consecutive stores" : "one coalesced store by adjacent lanes", hint ? ", L2::evict_last" : "", bytes / 32.0 / ms / 1e6, bytes / ms / 1e6, ms);
                (void)sectors;
            }
    const char *sn[] = {"atomicAdd u32 smem", "atomicAdd u64 smem", "ticket st+ld+ld+st", "plain ld+st u32"};
    CK(cudaFuncSetAttribute(smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 << 10));
    for (int words : {1024, 24576}) for (int mode = 0; mode < 4; mode++) {
        const int iters = 4096, g2 = pr.multiProcessorCount * 2;
        smem_kernel<<<g2, 512, 96 << 10>>>(iters, mode, words, sink);
        cudaEventRecord(a);
        smem_kernel<<<g2, 512, 96 << 10>>>(iters, mode, words, sink);
        cudaEventRecord(b); CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b);
        const double ops = (double)g2 * 512 * iters;
        printf("smem %6d words  %-22s %8.2f G ops/s chip-wide, %.3f ops/clk/SM @1.965GHz (%.2f ms)\n", words, sn[mode], ops / ms / 1e6, ops / ms / 1e6 / pr.multiProcessorCount / 1.965, ms);
    }
    return 0;
}
