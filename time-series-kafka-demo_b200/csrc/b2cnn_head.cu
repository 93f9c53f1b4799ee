// b2cnn_head.cu -- everything after x.view(-1, MAGICNUM) (bin/models.py:29):
//   LSTM layer-0 input projection (features x weight_ih_l0^T), the 2-layer LSTM cell
//   (models.py:30), Linear(16->1) (models.py:31) and the age scale (models.py:32-34).
//
//   proj_kernel          [B x L] x [L x 64] fp32 tiled GEMM with split-K partials
//   reduce_gates_kernel  sums the split-K partials in fixed order (deterministic) + biases
//   head_independent     one thread per window, LSTM from the zero state (predictStream.py:157)
//   head_sequence        one warp scans the batch axis carrying (h, c) -- the reference's
//                        model(x_batch) semantics for B > 1 (models.py:29-30, utils.py:249)
#include "b2cnn_internal.cuh"
#include "b2cnn_head_dev.cuh"

#ifndef B2CNN_HEAD_INFLIGHT
#define B2CNN_HEAD_INFLIGHT 8                    // 16-byte loads of range partials in flight per lane
#endif

namespace b2cnn {

// ---------------------------------------------------------------------------------------
__global__ void transpose_wih_kernel(const float *__restrict__ w, float *__restrict__ wT, int L) {
    // w: [64][L] -> wT: [L][64]
    __shared__ float tile[32][33];
    const int p0 = blockIdx.x * 32, g0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
    for (int i = ty; i < 32; i += 8) {
        const int g = g0 + i, pp = p0 + tx;
        tile[i][tx] = (pp < L) ? w[(int64_t)g * L + pp] : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int pp = p0 + i, g = g0 + tx;
        if (pp < L) wT[(int64_t)pp * kGates + g] = tile[tx][i];
    }
}

void launch_transpose_wih(const float *wih0, float *wih0T, int L, cudaStream_t st) {
    dim3 grid((L + 31) / 32, kGates / 32), block(32, 8);
    transpose_wih_kernel<<<grid, block, 0, st>>>(wih0, wih0T, L);
}

// ---------------------------------------------------------------------------------------
// gates0 partial[ks][b][g] = sum_{p in split ks} F[b][p] * WT[p][g]
// CTA tile: 64 windows x 64 gates, K-chunks of 32, 4x4 register tile per thread.
// ---------------------------------------------------------------------------------------
constexpr int kPM = 64, kPK = 32, kFsStride = 68;

__global__ void __launch_bounds__(256)
proj_kernel(const float *__restrict__ F, int64_t sB, int64_t sP, const float *__restrict__ WT,
            float *__restrict__ part, int B, int L, int k_per_split) {
    __shared__ __align__(16) float Fs[kPK][kFsStride];
    __shared__ __align__(16) float Ws[kPK][kGates];
    const int tid = threadIdx.x;
    const int tm = tid >> 4, tn = tid & 15;
    const int b0 = blockIdx.x * kPM;
    const int ks = blockIdx.y;
    const int kbeg = ks * k_per_split;
    const int kend = min(L, kbeg + k_per_split);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = kbeg; k0 < kend; k0 += kPK) {
#pragma unroll
        for (int it = 0; it < (kPM * kPK) / 256; ++it) {
            const int e = tid + it * 256;
            int m, kk;
            if (sP == 1) { m = e >> 5; kk = e & 31; } else { m = e & 63; kk = e >> 6; }
            const int b = b0 + m, k = k0 + kk;
            Fs[kk][m] = (b < B && k < kend) ? __ldg(F + (int64_t)b * sB + (int64_t)k * sP) : 0.f;
        }
#pragma unroll
        for (int it = 0; it < (kPK * kGates) / 256; ++it) {
            const int e = tid + it * 256;
            const int kk = e >> 6, g = e & 63;
            const int k = k0 + kk;
            Ws[kk][g] = (k < kend) ? __ldg(WT + (int64_t)k * kGates + g) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kPK; ++kk) {
            const float4 a = *reinterpret_cast<const float4 *>(&Fs[kk][4 * tm]);
            const float4 w = *reinterpret_cast<const float4 *>(&Ws[kk][4 * tn]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int b = b0 + 4 * tm + i;
        if (b < B) {
            float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            *reinterpret_cast<float4 *>(part + ((int64_t)ks * B + b) * kGates + 4 * tn) = v;
        }
    }
}

// gates[b][g] = (sum_ks partial[ks][b][g] + b_ih[g]) + b_hh[g]
__global__ void reduce_gates_kernel(const float *__restrict__ part, int ksplit, int64_t B,
                                    const float *__restrict__ bih, const float *__restrict__ bhh,
                                    float *__restrict__ gates) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B * kGates) return;
    const int g = (int)(e & 63);
    float s = 0.f;
    for (int k = 0; k < ksplit; ++k) s += part[(int64_t)k * B * kGates + e];
    gates[e] = (s + __ldg(bih + g)) + __ldg(bhh + g);
}

// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float age_scale(float age, float coef) { return head_age_scale(age, coef); }

// One thread per window; zero initial state so W_hh * h and f * c vanish (kept as written).
__global__ void __launch_bounds__(128)
head_independent_kernel(const float *__restrict__ gates0, HeadWeights hw, const float *__restrict__ age,
                        int64_t n_age, float coef, int apply_sigmoid, float *__restrict__ out, int64_t B) {
    __shared__ float s_wih1[kGates * kHidden];
    __shared__ float s_b1a[kGates], s_b1b[kGates], s_wo[kHidden + 1];
    for (int i = threadIdx.x; i < kGates * kHidden; i += blockDim.x) s_wih1[i] = hw.wih1[i];
    for (int i = threadIdx.x; i < kGates; i += blockDim.x) { s_b1a[i] = hw.bih1[i]; s_b1b[i] = hw.bhh1[i]; }
    for (int i = threadIdx.x; i < kHidden; i += blockDim.x) s_wo[i] = hw.wo[i];
    if (threadIdx.x == 0) s_wo[kHidden] = hw.bo[0];
    __syncthreads();
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float *g = gates0 + b * kGates;
    float h0[kHidden];
#pragma unroll
    for (int u = 0; u < kHidden; ++u) {
        const float ig = sigmoid_acc(g[u]), fg = sigmoid_acc(g[kHidden + u]);
        const float gg = tanhf(g[2 * kHidden + u]), og = sigmoid_acc(g[3 * kHidden + u]);
        const float c = fg * 0.f + ig * gg;
        h0[u] = og * tanhf(c);
    }
    float y = 0.f;
#pragma unroll 1
    for (int u = 0; u < kHidden; ++u) {
        float gi[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int r = q * kHidden + u;
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < kHidden; ++k) s = fmaf(s_wih1[r * kHidden + k], h0[k], s);
            gi[q] = (s + s_b1a[r]) + s_b1b[r];
        }
        const float c = sigmoid_acc(gi[1]) * 0.f + sigmoid_acc(gi[0]) * tanhf(gi[2]);
        const float h1 = sigmoid_acc(gi[3]) * tanhf(c);
        y = fmaf(s_wo[u], h1, y);
    }
    y += s_wo[kHidden];
    y *= age_scale(age[n_age == 1 ? 0 : b], coef);
    out[b] = apply_sigmoid ? sigmoid_acc(y) : y;
}

// Independent windows, reduction fused in: 16 lanes per window (lane u = hidden unit u), 16 windows per CTA.
// gates0[b][g] = (sum_k partial[k][b][g] + b_ih[g]) + b_hh[g] in the same fixed order as reduce_gates_kernel
// (windows a tensor-core front end flagged had their partial rows overwritten by the exact re-computation before
// this kernel runs).  The arithmetic per window is head_window16 (b2cnn_head_dev.cuh), which the fused streaming
// kernel runs too; it equals head_independent_kernel term for term.
// win_list / win_count (device memory, may be null): only the listed windows -- the exception path of the fused kernel.
__global__ void __launch_bounds__(256)
head_reduce_independent_kernel(const float *__restrict__ part, int slices, HeadWeights hw, const float *__restrict__ age,
                               int64_t n_age, float coef, int apply_sigmoid, float *__restrict__ out, int64_t B,
                               const int *__restrict__ win_list, const int *__restrict__ win_count,
                               int *__restrict__ clean_count, int *__restrict__ clean_flags, const int *__restrict__ clean_list) {
    __shared__ float s_w1[kHidden * kGates];               // W_ih_l1 transposed: s_w1[k * 64 + row], conflict-free per k
    // Last consumer of the call's exception list (the re-computation that read it ran before this kernel): put the
    // handle's flag state back to all-zero -- exactly the flags the streaming kernel set, and the count -- so the next
    // call needs no memset.  Nothing in this kernel reads them.
    if (clean_count != nullptr && blockIdx.x == 0) {
        const int n = *clean_count;
        for (int i = threadIdx.x; i < n; i += blockDim.x) clean_flags[clean_list[i]] = 0;
        __syncthreads();
        if (threadIdx.x == 0) *clean_count = 0;
    }
    for (int i = threadIdx.x; i < kGates * kHidden; i += blockDim.x) s_w1[(i & 15) * kGates + (i >> 4)] = __ldg(hw.wih1 + i);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t nwin = win_count ? *win_count : B;
    const int64_t per_pass = (int64_t)gridDim.x * (blockDim.x >> 4);
    const int64_t passes = (nwin + per_pass - 1) / per_pass;                 // the same for every thread: shuffles stay full-warp
    for (int64_t it = 0; it < passes; ++it) {
        const int64_t wi = it * per_pass + (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4);
        const bool live = wi < nwin;
        const int64_t wl = live ? wi : nwin - 1;            // dead lanes shadow the last window
        const int64_t b = win_list ? win_list[wl] : wl;
        const float y = head_window16<true, B2CNN_HEAD_INFLIGHT>(reinterpret_cast<const float4 *>(part + b * kGates) + (lane & 15), B * (kGates / 4), slices, hw,
                                            s_w1, age[n_age == 1 ? 0 : b], coef, apply_sigmoid, lane);
        if (live && (lane & 15) == 0) out[b] = y;
    }
}

// One warp scans the B rows sequentially.  Lane l owns gate rows l and l+32 of every weight
// matrix (registers); units' (h, c) are held twice, by lanes u and u+16.
__global__ void __launch_bounds__(32)
head_sequence_kernel(const float *__restrict__ gates0, HeadWeights hw, const float *__restrict__ age,
                     int64_t n_age, float coef, int apply_sigmoid, float *__restrict__ out, int64_t B) {
    const int l = threadIdx.x, u = l & 15;
    const bool lo = l < 16;
    float whh0a[kHidden], whh0b[kHidden], wih1a[kHidden], wih1b[kHidden], whh1a[kHidden], whh1b[kHidden];
#pragma unroll
    for (int k = 0; k < kHidden; ++k) {
        whh0a[k] = hw.whh0[l * kHidden + k]; whh0b[k] = hw.whh0[(l + 32) * kHidden + k];
        wih1a[k] = hw.wih1[l * kHidden + k]; wih1b[k] = hw.wih1[(l + 32) * kHidden + k];
        whh1a[k] = hw.whh1[l * kHidden + k]; whh1b[k] = hw.whh1[(l + 32) * kHidden + k];
    }
    const float bih1a = hw.bih1[l], bih1b = hw.bih1[l + 32], bhh1a = hw.bhh1[l], bhh1b = hw.bhh1[l + 32];
    const float wo = hw.wo[u], bo = hw.bo[0];
    float h0 = 0.f, c0 = 0.f, h1 = 0.f, c1 = 0.f;
    float na = gates0[l], nb = gates0[l + 32];
    for (int64_t t = 0; t < B; ++t) {
        float ga = na, gb = nb;
        if (t + 1 < B) { na = gates0[(t + 1) * kGates + l]; nb = gates0[(t + 1) * kGates + l + 32]; }
        // ---- layer 0: gates0 already holds (W_ih x + b_ih) + b_hh; add W_hh h_{t-1}
        float ra = 0.f, rb = 0.f;
#pragma unroll
        for (int k = 0; k < kHidden; ++k) {
            const float hk = __shfl_sync(0xffffffffu, h0, k);
            ra = fmaf(whh0a[k], hk, ra); rb = fmaf(whh0b[k], hk, rb);
        }
        ga += ra; gb += rb;
        float A = sigmoid_acc(ga);                         // lanes <16: i ; lanes >=16: f
        float Bv = lo ? tanhf(gb) : sigmoid_acc(gb);       // lanes <16: g ; lanes >=16: o
        float ig = __shfl_sync(0xffffffffu, A, u), fg = __shfl_sync(0xffffffffu, A, u + 16);
        float gg = __shfl_sync(0xffffffffu, Bv, u), og = __shfl_sync(0xffffffffu, Bv, u + 16);
        c0 = fg * c0 + ig * gg;
        h0 = og * tanhf(c0);
        // ---- layer 1: input h0 (new), recurrent h1 (old)
        float sa = 0.f, sb = 0.f; ra = 0.f; rb = 0.f;
#pragma unroll
        for (int k = 0; k < kHidden; ++k) {
            const float xk = __shfl_sync(0xffffffffu, h0, k);
            const float hk = __shfl_sync(0xffffffffu, h1, k);
            sa = fmaf(wih1a[k], xk, sa); sb = fmaf(wih1b[k], xk, sb);
            ra = fmaf(whh1a[k], hk, ra); rb = fmaf(whh1b[k], hk, rb);
        }
        ga = (sa + bih1a) + (ra + bhh1a);
        gb = (sb + bih1b) + (rb + bhh1b);
        A = sigmoid_acc(ga);
        Bv = lo ? tanhf(gb) : sigmoid_acc(gb);
        ig = __shfl_sync(0xffffffffu, A, u); fg = __shfl_sync(0xffffffffu, A, u + 16);
        gg = __shfl_sync(0xffffffffu, Bv, u); og = __shfl_sync(0xffffffffu, Bv, u + 16);
        c1 = fg * c1 + ig * gg;
        h1 = og * tanhf(c1);
        // ---- Linear(16->1) + age scale
        float y = lo ? wo * h1 : 0.f;
#pragma unroll
        for (int off = 8; off >= 1; off >>= 1) y += __shfl_xor_sync(0xffffffffu, y, off);
        if (l == 0) {
            y = (y + bo) * age_scale(age[n_age == 1 ? 0 : t], coef);
            out[t] = apply_sigmoid ? sigmoid_acc(y) : y;
        }
    }
}

// ---------------------------------------------------------------------------------------
int choose_ksplit(int64_t B, int L, int num_sms) {
    // Depends on L only: the summation order of a window's projection must not change with the
    // batch it arrives in (prefix / chunking consistency is tested bit-for-bit).
    (void)B; (void)num_sms;
    int ks = (L + 2047) / 2048;
    return ks < 1 ? 1 : ks;
}

int launch_head(const Dims &d, const HeadWeights &hw, const float *feats, int64_t sB, int64_t sP,
                int64_t B, const float *age, int64_t n_age, int mode, int apply_sigmoid,
                float *out, float *gates_ws, float *partial_ws, int ksplit, cudaStream_t st,
                const char **err) {
    int launches = 0;
    int kps = (d.L + ksplit - 1) / ksplit;
    kps = ((kps + kPK - 1) / kPK) * kPK;           // whole K-chunks per split
    const int ks_eff = (d.L + kps - 1) / kps;
    dim3 grid((unsigned)((B + kPM - 1) / kPM), ks_eff);
    proj_kernel<<<grid, 256, 0, st>>>(feats, sB, sP, hw.wih0T, partial_ws, (int)B, d.L, kps);
    ++launches;
    int n = launch_reduce_gates(partial_ws, ks_eff, B, hw, gates_ws, st, err);
    if (n < 0) return -1;
    launches += n;
    n = launch_lstm_head(d, hw, gates_ws, B, age, n_age, mode, apply_sigmoid, out, st, err);
    if (n < 0) return -1;
    return launches + n;
}

// gates[b][g] = (sum over `slices` partial[s][b][g] + b_ih[g]) + b_hh[g], fixed summation order
int launch_reduce_gates(const float *partial, int slices, int64_t B, const HeadWeights &hw, float *gates,
                        cudaStream_t st, const char **err) {
    const int64_t n = B * kGates;
    reduce_gates_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(partial, slices, B, hw.bih0, hw.bhh0, gates);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return -1; }
    return 1;
}

// independent windows: slice reduction + LSTM cells + Linear + age scale in one launch
int launch_reduce_lstm_head(const Dims &d, const HeadWeights &hw, const float *partial, int slices, int64_t B,
                            const float *age, int64_t n_age, int apply_sigmoid, float *out, cudaStream_t st, const char **err,
                            int *clean_count, int *clean_flags, const int *clean_list) {
    head_reduce_independent_kernel<<<(unsigned)((B * 16 + 255) / 256), 256, 0, st>>>(partial, slices, hw, age, n_age, d.age_coef,
                                                                                  apply_sigmoid, out, B, nullptr, nullptr,
                                                                                  clean_count, clean_flags, clean_list);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return -1; }
    return 1;
}

// LSTM cells + Linear + age scale from layer-0 gate pre-activations (bin/models.py:30-34)
int launch_lstm_head(const Dims &d, const HeadWeights &hw, const float *gates, int64_t B, const float *age,
                     int64_t n_age, int mode, int apply_sigmoid, float *out, cudaStream_t st, const char **err) {
    if (mode == B2CNN_MODE_INDEPENDENT) {
        head_independent_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(gates, hw, age, n_age, d.age_coef,
                                                                             apply_sigmoid, out, B);
    } else {
        head_sequence_kernel<<<1, 32, 0, st>>>(gates, hw, age, n_age, d.age_coef, apply_sigmoid, out, B);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return -1; }
    return 1;
}

}  // namespace b2cnn
