// b2cnn_small.cu -- the whole forward pass of one short window in ONE launch.
//
// The production call of the reference is model(x[1,10,120], age[1]) (bin/predictStream.py:157):
// 44 k MAC -- pure launch latency on a GPU.  For windows whose staging fits one CTA's shared
// memory this kernel runs conv1+act+pool -> conv2+act+pool -> LSTM layer 0/1 from the zero state
// -> Linear -> age scale (bin/models.py:23-34) inside a single CTA per window, replacing the four
// launches of the general path.  Any conv/pool geometry (runtime loops), f32 or bf16 input,
// independent-window semantics only (a batch-as-sequence scan goes through the general path).
#include "b2cnn_internal.cuh"

namespace b2cnn {

template <typename T>
__device__ __forceinline__ float ld_small(const T *p);
template <>
__device__ __forceinline__ float ld_small<float>(const float *p) { return __ldg(p); }
template <>
__device__ __forceinline__ float ld_small<__nv_bfloat16>(const __nv_bfloat16 *p) { return __bfloat162float(__ldg(p)); }

struct SmallParams {
    const void *x;
    const float *age;
    float *out;
    int64_t n_age;
    int B, apply_sigmoid;
    Dims d;
    HeadWeights hw;
    ConvWeights cw;
};

__device__ __forceinline__ float age_scale_small(float age, float coef) {
    float s = __fadd_rn(__fmul_rn(age, coef), 1.0f);      // models.py:32: separate multiply and add
    return (s > 0.f || s != s) ? s : 0.f;
}

template <typename Tin>
__global__ void __launch_bounds__(256) small_forward_kernel(const __grid_constant__ SmallParams p) {
    extern __shared__ float sm[];
    const Dims &d = p.d;
    const int C = d.C, K1 = d.K1, K2 = d.K2, PK = d.PK, PS = d.PS, W = d.W, P1 = d.P1, L = d.L;
    float *xs = sm;                     // [C][W]
    float *a1 = xs + C * W;             // [4][P1]
    float *fs = a1 + kCMid * P1;        // [L]
    float *g = fs + L;                  // [64] gate pre-activations
    float *h = g + kGates;              // [16]
    const int tid = threadIdx.x;
    const int b = blockIdx.x;
    const Tin *xb = reinterpret_cast<const Tin *>(p.x) + (int64_t)b * C * d.XP;
    if (d.XP == W) {
        for (int i = tid; i < C * W; i += 256) xs[i] = ld_small<Tin>(xb + i);
    } else {
        for (int c = 0; c < C; ++c)
            for (int i = tid; i < W; i += 256) xs[c * W + i] = ld_small<Tin>(xb + (int64_t)c * d.XP + i);
    }
    __syncthreads();
    // conv1 + act + pool (models.py:23-24); activation before pooling exactly as written
    for (int e = tid; e < P1 * kCMid; e += 256) {
        const int j = e >> 2, o = e & 3;
        float best = 0.f;
        for (int u = 0; u < PK; ++u) {
            const int t = PS * j + u;
            float s = 0.f;
            for (int c = 0; c < C; ++c)
                for (int k = 0; k < K1; ++k) s = fmaf(p.cw.w1[(c * K1 + k) * kCMid + o], xs[c * W + t + k], s);
            float v = s + p.cw.b1[o];
            if (d.has_affine) v = fmaf(v, p.cw.s1[o], p.cw.t1[o]);
            v = apply_act(v, d.act);
            best = (u == 0) ? v : max_nan(best, v);
        }
        a1[o * P1 + j] = best;
    }
    __syncthreads();
    // conv2 + act + pool (models.py:26-27) -> features (models.py:29)
    for (int pl = tid; pl < L; pl += 256) {
        float best = 0.f;
        for (int u = 0; u < PK; ++u) {
            const int q = PS * pl + u;
            float s = 0.f;
            for (int c = 0; c < kCMid; ++c)
                for (int k = 0; k < K2; ++k) s = fmaf(p.cw.w2[c * K2 + k], a1[c * P1 + q + k], s);
            float v = s + p.cw.b2;
            if (d.has_affine) v = fmaf(v, p.cw.s2, p.cw.t2);
            v = apply_act(v, d.act);
            best = (u == 0) ? v : max_nan(best, v);
        }
        fs[pl] = best;
    }
    __syncthreads();
    // LSTM layer 0 from the zero state (models.py:30): gates = (W_ih f + b_ih) + (0 + b_hh)
    if (tid < kGates) {
        float s = 0.f;
        for (int pp = 0; pp < L; ++pp) s = fmaf(fs[pp], __ldg(p.hw.wih0T + (int64_t)pp * kGates + tid), s);
        g[tid] = (s + __ldg(p.hw.bih0 + tid)) + __ldg(p.hw.bhh0 + tid);
    }
    __syncthreads();
    if (tid < kHidden) {
        const float ig = sigmoid_acc(g[tid]), fg = sigmoid_acc(g[kHidden + tid]);
        const float gg = tanhf(g[2 * kHidden + tid]), og = sigmoid_acc(g[3 * kHidden + tid]);
        const float c = fg * 0.f + ig * gg;
        h[tid] = og * tanhf(c);
    }
    __syncthreads();
    float g1 = 0.f;
    if (tid < kGates) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < kHidden; ++k) s = fmaf(__ldg(p.hw.wih1 + tid * kHidden + k), h[k], s);
        g1 = (s + __ldg(p.hw.bih1 + tid)) + __ldg(p.hw.bhh1 + tid);
    }
    __syncthreads();
    if (tid < kGates) g[tid] = g1;
    __syncthreads();
    if (tid < 32) {
        float y = 0.f;
        if (tid < kHidden) {
            const float ig = sigmoid_acc(g[tid]), fg = sigmoid_acc(g[kHidden + tid]);
            const float gg = tanhf(g[2 * kHidden + tid]), og = sigmoid_acc(g[3 * kHidden + tid]);
            const float c = fg * 0.f + ig * gg;
            y = __ldg(p.hw.wo + tid) * (og * tanhf(c));       // Linear(16 -> 1), models.py:31
        }
#pragma unroll
        for (int off = 8; off >= 1; off >>= 1) y += __shfl_xor_sync(0xffffffffu, y, off);
        if (tid == 0) {
            y = (y + __ldg(p.hw.bo)) * age_scale_small(p.age[p.n_age == 1 ? 0 : b], d.age_coef);
            p.out[b] = p.apply_sigmoid ? sigmoid_acc(y) : y;
        }
    }
}

size_t small_smem_bytes(const Dims &d) {
    return sizeof(float) * ((size_t)d.C * d.W + (size_t)kCMid * d.P1 + d.L + kGates + kHidden + 8);
}

bool small_supported(const Dims &d) { return small_smem_bytes(d) <= 96 * 1024 && d.L <= 2048; }

int launch_small_forward(const Dims &d, const ConvWeights &cw, const HeadWeights &hw, const void *x, int dtype, int64_t B,
                         const float *age, int64_t n_age, int apply_sigmoid, float *out, cudaStream_t st, const char **err) {
    SmallParams p;
    p.x = x; p.age = age; p.out = out; p.n_age = n_age; p.B = (int)B; p.apply_sigmoid = apply_sigmoid;
    p.d = d; p.hw = hw; p.cw = cw;
    const size_t smem = small_smem_bytes(d);
    cudaError_t e = cudaSuccess;
    if (dtype == B2CNN_DTYPE_F32) {
        if (smem > 48 * 1024) e = cudaFuncSetAttribute(small_forward_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) small_forward_kernel<float><<<(unsigned)B, 256, smem, st>>>(p);
    } else {
        if (smem > 48 * 1024) e = cudaFuncSetAttribute(small_forward_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) small_forward_kernel<__nv_bfloat16><<<(unsigned)B, 256, smem, st>>>(p);
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return -1; }
    return 1;
}

}  // namespace b2cnn
