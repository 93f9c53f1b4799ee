// pipes.cu -- dispatch cost of the epilogue's instruction mix on one SM sub-partition (sm_100a):
// cycles per warp-instruction for independent streams of FFMA, FFMA2, FMNMX3, MUFU.EX2 and mixtures, with 1 and 2
// warps per scheduler.  nvcc -arch=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>
#define CHAINS 8
template <int MODE>
__global__ void k(float *out, long long *cyc, int iters) {
    float2 a[CHAINS]; float m[CHAINS];
    for (int i = 0; i < CHAINS; ++i) { a[i] = make_float2(threadIdx.x * 1e-3f + i, 1.0f + i); m[i] = a[i].x; }
    const float2 w = make_float2(1.0001f, 0.9999f), c = make_float2(1e-7f, -1e-7f);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (MODE == 0 || MODE == 4 || MODE == 5 || MODE == 7) {   // FFMA2
                unsigned long long d, A, W, C;
                asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a[i].x), "f"(a[i].y));
                asm("mov.b64 %0, {%1, %2};" : "=l"(W) : "f"(w.x), "f"(w.y));
                asm("mov.b64 %0, {%1, %2};" : "=l"(C) : "f"(c.x), "f"(c.y));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(A), "l"(W), "l"(C));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(a[i].x), "=f"(a[i].y) : "l"(d));
            }
            if (MODE == 1) {                         // FFMA x2 (same math as one FFMA2)
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i].x) : "f"(w.x), "f"(c.x));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i].y) : "f"(w.y), "f"(c.y));
            }
            if (MODE == 2 || MODE == 4 || MODE == 6 || MODE == 7) {   // FMNMX3
                asm volatile("max.NaN.f32 %0, %0, %1, %2;" : "+f"(m[i]) : "f"(w.x), "f"(c.y));
            }
            if (MODE == 3 || MODE == 5 || MODE == 6 || MODE == 7) {   // MUFU.EX2
                if (MODE != 7 || (i & 1) == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[(i + 4) % CHAINS]));
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < CHAINS; ++i) s += a[i].x + a[i].y + m[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int MODE>
void run(const char *name, int per_iter_instr, int threads) {
    float *out; long long *cyc;
    cudaMalloc(&out, 4 * 1024 * 148); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    k<MODE><<<148, threads>>>(out, cyc, iters); cudaDeviceSynchronize();
    k<MODE><<<148, threads>>>(out, cyc, iters); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_sched_warps = threads / 128.0;
    printf("%-34s warps/scheduler %.0f: %6.2f cycles per iteration-chain-step per scheduler  (%d instr: %.2f cycles/instr/scheduler)\n", name, per_sched_warps,
           (double)h / iters / CHAINS, per_iter_instr, (double)h / iters / CHAINS / (per_iter_instr * per_sched_warps) * per_sched_warps);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int threads : {128, 256, 384}) {
        run<0>("FFMA2", 1, threads);
        run<1>("FFMA + FFMA", 2, threads);
        run<2>("FMNMX3", 1, threads);
        run<3>("MUFU.EX2", 1, threads);
        run<4>("FFMA2 + FMNMX3", 2, threads);
        run<5>("FFMA2 + MUFU.EX2", 2, threads);
        run<6>("FMNMX3 + MUFU.EX2", 2, threads);
        run<7>("FFMA2 + FMNMX3 + 0.5 MUFU.EX2", 3, threads);
    }
    return 0;
}
