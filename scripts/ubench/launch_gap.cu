// launch_gap.cu -- how long after the end of a small-shared-memory kernel does a 218 KB-shared-memory kernel start on
// the same stream (and the other way round)?  %globaltimer stamps: last exit of the previous grid, first instruction of
// the next.  Variants of the small kernel: default carve-out, PreferredSharedMemoryCarveout = max, and a dynamic
// shared-memory request that forces the large carve-out.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o launch_gap launch_gap.cu && ./launch_gap
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#include <cuda.h>
__device__ unsigned long long g_t[4];   // 0: small first instr (min), 1: small last exit (max), 2: big first instr (min), 3: big last exit (max)
__device__ __forceinline__ unsigned long long now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void spin(unsigned long long ns) { const unsigned long long t0 = now(); while (now() - t0 < ns) {} }
extern __shared__ unsigned char dsm[];
__global__ void small_k(unsigned long long ns) {
    if (threadIdx.x == 0) atomicMin(&g_t[0], now());
    spin(ns);
    if (threadIdx.x == 0) { dsm[0] = 1; atomicMax(&g_t[1], now()); }
}
__global__ void __launch_bounds__(384, 1) big_k(unsigned long long ns) {
    if (threadIdx.x == 0) atomicMin(&g_t[2], now());
    spin(ns);
    if (threadIdx.x == 0) { dsm[0] = 1; atomicMax(&g_t[3], now()); }
}
struct Fat { char b[336]; };
__global__ void __launch_bounds__(384, 1) big_tm_k(const __grid_constant__ CUtensorMap tm, const __grid_constant__ Fat f, unsigned long long ns) {
    if (threadIdx.x == 0) atomicMin(&g_t[2], now());
    spin(ns);
    if (threadIdx.x == 0) { dsm[0] = (unsigned char)(f.b[5] + reinterpret_cast<const unsigned char *>(&tm)[3]); atomicMax(&g_t[3], now()); }
}
int main() {
    const int big_smem = 218 * 1024 + 448;
    cudaFuncSetAttribute(big_k, cudaFuncAttributeMaxDynamicSharedMemorySize, big_smem);
    cudaFuncSetAttribute(small_k, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    cudaFuncSetAttribute(big_tm_k, cudaFuncAttributeMaxDynamicSharedMemorySize, 218 * 1024 + 448);
    CUtensorMap tm; memset(&tm, 0, sizeof tm); Fat fat; memset(&fat, 1, sizeof fat);
    for (int variant = 0; variant < 6; ++variant) {
        size_t small_smem = 16;
        int small_blocks = 256;
        cudaFuncSetAttribute(small_k, cudaFuncAttributePreferredSharedMemoryCarveout, variant == 1 ? (int)cudaSharedmemCarveoutMaxShared : (int)cudaSharedmemCarveoutDefault);
        if (variant == 2) small_smem = 110 * 1024;          // two CTAs per SM need the 228 KB carve-out
        if (variant == 3) small_smem = 48 * 1024;           // like the generic front end's exception launch
        double gap_sb = 0, gap_bs = 0;
        const int N = 20;
        for (int it = 0; it < N + 2; ++it) {
            unsigned long long init[4] = {~0ull, 0ull, ~0ull, 0ull}, t[4];
            cudaMemcpyToSymbol(g_t, init, sizeof init);
            cudaDeviceSynchronize();
            big_k<<<148, 384, big_smem>>>(20000);            // warm: the SMs are in the big configuration
            cudaMemcpyToSymbol(g_t, init, sizeof init);       // (stream-ordered on the default stream)
            small_k<<<small_blocks, 256, small_smem>>>(15000);
            if (variant == 4) cudaFuncSetAttribute(big_k, cudaFuncAttributeMaxDynamicSharedMemorySize, big_smem);
            if (variant == 5) big_tm_k<<<148, 384, big_smem>>>(tm, fat, 20000); else
            big_k<<<148, 384, big_smem>>>(20000);
            cudaDeviceSynchronize();
            cudaMemcpyFromSymbol(t, g_t, sizeof t);
            const double sb = (double)(long long)(t[2] - t[1]) / 1e3;
            // and big -> small
            cudaMemcpyToSymbol(g_t, init, sizeof init);
            big_k<<<148, 384, big_smem>>>(20000);
            small_k<<<small_blocks, 256, small_smem>>>(15000);
            cudaDeviceSynchronize();
            cudaMemcpyFromSymbol(t, g_t, sizeof t);
            const double bs = (double)(long long)(t[0] - t[3]) / 1e3;
            if (it >= 2) { gap_sb += sb; gap_bs += bs; }
        }
        const char *names[6] = {"small kernel: 16 B smem, default carve-out", "small kernel: carve-out preference = max shared",
                                "small kernel: 110 KB dynamic smem (forces the 228 KB carve-out)", "small kernel: 48 KB dynamic smem",
                                "default small kernel; cudaFuncSetAttribute(max dyn smem) before every big launch", "default small kernel; big kernel takes a CUtensorMap + 336-byte struct"};
        printf("%-66s small end -> big (218 KB) start %6.2f us | big end -> small start %6.2f us\n", names[variant], gap_sb / N, gap_bs / N);
    }
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
