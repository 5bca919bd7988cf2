// p2p_bench.cu — what NVLink 5 / NVSwitch gives a kernel that scatters small pieces into a PEER's
// memory (the multi-GPU level-1 binning writes its segments straight into the owner rank's HBM):
// 32/64/128-byte pieces written by 1/2/4 adjacent lanes in one store instruction, to random
// piece-aligned places of an 8 GB peer buffer, one direction and both directions at once, next to
// the copy-engine rate (cudaMemcpyPeerAsync) and the same kernel on local memory.
// Standalone: nvcc -arch=sm_100a tools/p2p_bench.cu -o tools/p2p_bench ; needs >= 2 GPUs.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint64_t rng(uint64_t &s) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; uint64_t x = s; x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 29; return x; }

// remote_num/8 of the pieces go to `remote`, the rest to `local`
__global__ void __launch_bounds__(256) piece_kernel(uint64_t *local, uint64_t *remote, uint64_t words, int iters, int lanes, int remote_num) {
    const uint32_t lane = threadIdx.x & 31, sub = lane % lanes;
    uint64_t s = ((blockIdx.x * 256ull + threadIdx.x) / lanes) * 0x9E3779B97F4A7C15ULL + 999;
    const uint64_t n_pieces = words / (4ull * lanes);
    for (int i = 0; i < iters; i++) {
        const uint64_t r = rng(s);
        uint64_t *buf = (int)((r >> 7) & 7) < remote_num ? remote : local;
        uint64_t *q = buf + __umul64hi(r, n_pieces) * 4ull * lanes + 4 * sub;
        asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(q), "l"(r), "l"(r + 1), "l"(r + 2), "l"(r + 3) : "memory");
    }
}

int main() {
    int n = 0; CK(cudaGetDeviceCount(&n));
    if (n < 2) { printf("need 2 GPUs, have %d\n", n); return 0; }
    const uint64_t bytes = 8ull << 30, words = bytes / 8;
    uint64_t *buf[2]; cudaStream_t st[2]; cudaEvent_t a[2], b[2];
    for (int d = 0; d < 2; d++) {
        CK(cudaSetDevice(d)); int can = 0; CK(cudaDeviceCanAccessPeer(&can, d, 1 - d));
        if (!can) { printf("no peer access %d -> %d\n", d, 1 - d); return 0; }
        CK(cudaDeviceEnablePeerAccess(1 - d, 0));
        CK(cudaMalloc(&buf[d], bytes)); CK(cudaMemset(buf[d], 0, bytes)); CK(cudaStreamCreate(&st[d]));
        cudaEventCreate(&a[d]); cudaEventCreate(&b[d]);
    }
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    printf("%s x%d, %d SMs\n", pr.name, n, pr.multiProcessorCount);
    // copy engine
    for (int rep = 0; rep < 2; rep++) {
        CK(cudaSetDevice(0)); cudaEventRecord(a[0], st[0]);
        CK(cudaMemcpyPeerAsync(buf[1], 1, buf[0], 0, bytes / 2, st[0])); cudaEventRecord(b[0], st[0]); CK(cudaEventSynchronize(b[0]));
        float ms; cudaEventElapsedTime(&ms, a[0], b[0]);
        if (rep) printf("cudaMemcpyPeerAsync 0->1, 4 GB: %.0f GB/s\n", bytes / 2 / ms / 1e6);
    }
    for (int both = 0; both < 2; both++)
        for (int remote_num : {8, 4, 7, 0})
            for (int lanes : {1, 2, 4})
                for (int cps : {2, 6}) {
                    const int grid = pr.multiProcessorCount * cps, iters = 512;
                    float ms[2] = {0, 0};
                    for (int rep = 0; rep < 2; rep++) {
                        for (int d = 0; d < (both ? 2 : 1); d++) {
                            CK(cudaSetDevice(d)); cudaEventRecord(a[d], st[d]);
                            piece_kernel<<<grid, 256, 0, st[d]>>>(buf[d], buf[1 - d], words, iters, lanes, remote_num);
                            cudaEventRecord(b[d], st[d]);
                        }
                        for (int d = 0; d < (both ? 2 : 1); d++) { CK(cudaSetDevice(d)); CK(cudaEventSynchronize(b[d])); cudaEventElapsedTime(&ms[d], a[d], b[d]); }
                    }
                    const double total = (double)grid * 256 * iters * 32.0, rem = total * remote_num / 8.0;
                    printf("%s  %d/8 remote  %3d-byte pieces  %d CTAs/SM: %7.0f GB/s written per GPU, %7.0f GB/s of it over NVLink, %6.2f G requests/s (%.2f ms%s)\n",
                           both ? "both directions" : "one direction  ", remote_num, 32 * lanes, cps, total / ms[0] / 1e6, rem / ms[0] / 1e6,
                           total / (32.0 * lanes) / ms[0] / 1e6, ms[0], both ? "" : "");
                }
    return 0;
}
