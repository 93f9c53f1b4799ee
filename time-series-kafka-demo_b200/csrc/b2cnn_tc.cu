// b2cnn_tc.cu -- tcgen05 / TMEM / TMA front end for bf16 windows (sm_100a).
//
// conv1 (bin/models.py:23) is 82 % of the path's FLOPs at the headline shape [4096,3,75000]:
// 9.0 M MAC per window, which on the FP32 CUDA cores alone costs ~4x the time the HBM roofline
// allows.  Here it runs on the 5th-gen tensor cores as a banded-Toeplitz GEMM whose M axis is
// the WINDOW index:
//
//   D[w, (s,o)] = sum_{c} sum_{k<16}  X_c[w, 8n+k] * T_c[k, (s,o)],   T_c[k,(s,o)] = w1[o][c][k-s]
//
//   * A = X_c: 128 windows x 64 consecutive samples of channel c, brought by ONE 3-D TMA box
//     {64 samples, 1 channel, 128 windows} into the canonical K-major SWIZZLE_128B layout; the
//     8-position block n of the tile uses the K=16 slice starting at 16-byte chunk n (descriptor
//     start address + 16n bytes), so one landed tile feeds 7 blocks = 56 conv1 positions
//     (tiles advance by 56 samples; the 8 overlapping samples are re-read from L2, not HBM).
//   * B = T_c: the fp32 conv1 weights expanded to a 16 x 32 band matrix (8 output shifts s x 4
//     output channels o) and split into 2 or 3 bf16 pieces (hi/mid/lo) so that, the inputs
//     being exactly bf16, every product is exact and the fp32 accumulation in TMEM carries the
//     full fp32 weight precision.  One MMA (M=128,N=32,K=16) per (block, channel, piece).
//   * The one tap that does not fit a 16-sample slice (s=7, k=9 -> sample 8n+16) is added by
//     the epilogue on the CUDA cores from the same shared-memory tile (12 FMA per block).
//   * Epilogue: thread == window.  Each thread streams through its window's positions in
//     order, so pool1 -> tanh -> conv2 -> pool2 -> tanh are register-local sliding windows with
//     no shuffles, seams or shared-memory exchange; tanh is 1 - 2/(1+2^(2x log2 e)) on
//     MUFU.EX2 + MUFU.RCP with the bias folded into the exponent FMA; pooling runs BEFORE the
//     activation (monotone) with FMNMX3.NAN so NaNs propagate exactly like ATen's max_pool1d.
//   * Zero band entries turn an inf/NaN sample into NaN for its whole 8-position block, a
//     superset of the reference's NaNs: a window whose features contain a NaN is flagged and
//     recomputed by the exact generic kernel (b2cnn_generic.cu), still on the GPU.
//
// Warp roles per CTA (192 threads, 2 CTAs/SM): warp 0 TMA producer, warp 1 TMEM allocator +
// single-thread MMA issuer, warps 2-5 epilogue (TMEM lane quadrant = warp % 4).
// Pipelines: 2 smem stages (full/empty mbarriers), 2 TMEM accumulator stages of 4+3 blocks.
#include <cuda.h>

#include <cstdlib>
#include <cstring>
#include <vector>

#include "b2cnn_tc.cuh"
#include "b2cnn_tc_ptx.cuh"

namespace b2cnn {

// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=32 (conv1 bands)
constexpr uint32_t kIdesc = make_idesc_bf16(128, 32);

static thread_local const char *g_tc_err = "";
const char *tc_error() { return g_tc_err; }

#ifndef B2CNN_OWN_FLAGS
#define B2CNN_OWN_FLAGS 1          // NaN-exception flag state owned by the handle and cleaned by the head kernel: no per-call memset (TcState)
#endif
constexpr int64_t kOwnFlagCap = 65536;   // windows per call served by the handle's own flag state (512 KB); larger batches use the workspace copy
constexpr int kTcM = 128;          // windows per CTA == UMMA M
constexpr int kTcAdv = 56;         // conv1 positions (= samples) a tile advances
constexpr int kTcBlocks = 7;       // 8-position blocks per 64-sample tile
constexpr int kTcABytes = 128 * 128;
constexpr int kTcBBytes = 32 * 16 * 2;   // one band matrix piece: N=32 x K=16 bf16
constexpr int kTcMaxC = 4;
constexpr int kTcThreads = 192;

struct TcParams {
    float *feats;
    int64_t sB, sP;
    int *nanflag;
    const uint8_t *bmats;     // [C][SPLITS][1024] bytes, UMMA K-major no-swizzle core-matrix order
    int B, W, L;
    int tiles_per_cta, feats_per_cta;
    float w9[kCMid][kTcMaxC];   // tap k=9 of conv1: w1[o][c][9]
    float b1s[kCMid];           // conv1 bias * 2 log2 e
    float w2[kCMid][5];
    float b2s;                  // conv2 bias * 2 log2 e
};

// ------------------------------------------------------------------------------------------
// The kernel (MyCNN5 architecture: K1=10, pool(3,2), K2=5)
// ------------------------------------------------------------------------------------------
template <int C, int SPLITS>
__global__ void __launch_bounds__(kTcThreads, 2)
tc_frontend_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;                                    // [2 stages][C][16 KB]
    uint8_t *sBm = smem + 2 * C * kTcABytes;               // [C][SPLITS][1 KB]
    uint64_t *bars = reinterpret_cast<uint64_t *>(sBm + C * SPLITS * kTcBBytes);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 8);
    const uint32_t bar_full = smem_u32(bars + 0), bar_empty = smem_u32(bars + 2);
    const uint32_t bar_tfull = smem_u32(bars + 4), bar_tempty = smem_u32(bars + 6);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b0 = blockIdx.x * kTcM;
    const int p0 = blockIdx.y * p.feats_per_cta;
    const int nfeat = min(p.feats_per_cta, p.L - p0);
    const int nsteps = (nfeat + 4) / 2;                    // features 2j-3, 2j-2 leave at step j
    const int ntiles = (nsteps + kTcBlocks - 1) / kTcBlocks;
    const int T0 = p0 * 4;                                 // first conv1 position == first sample

    // band matrices -> smem (generic proxy), barriers, TMEM
    for (int i = threadIdx.x; i < C * SPLITS * kTcBBytes / 16; i += kTcThreads)
        reinterpret_cast<uint4 *>(sBm)[i] = reinterpret_cast<const uint4 *>(p.bmats)[i];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_empty + 8 * i, 4);
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // band matrices visible to the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            for (int i = 0; i < ntiles; ++i) {
                const int s = i & 1, ph = (i >> 1) & 1;
                mbar_wait(bar_empty + 8 * s, ph ^ 1);
                mbar_expect_tx(bar_full + 8 * s, C * kTcABytes);
#pragma unroll
                for (int c = 0; c < C; ++c)
                    tma_load_3d(smem_u32(sA + (s * C + c) * kTcABytes), &tmap, T0 + kTcAdv * i, c, b0, bar_full + 8 * s);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // whole warp runs the loop (warp-uniform descriptor arithmetic); one elected lane issues
        const uint64_t a_base = desc_sw128_kmajor(smem_u32(sA));
        const uint64_t b_base = desc_none_kmajor(smem_u32(sBm), 128, 256);
        const uint32_t a_lo0 = (uint32_t)a_base, a_hi = (uint32_t)(a_base >> 32);
        const uint32_t b_lo0 = (uint32_t)b_base, b_hi = (uint32_t)(b_base >> 32);
        for (int i = 0; i < ntiles; ++i) {
            const int s = i & 1, ph = (i >> 1) & 1;
            mbar_wait(bar_full + 8 * s, ph);
            tc_fence_after();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                mbar_wait(bar_tempty + 8 * h, (i & 1) ^ 1);
                tc_fence_after();
                if (elect_one()) {
                    const int nb0 = h ? 4 : 0, nb1 = h ? kTcBlocks : 4;
                    for (int n = nb0; n < nb1; ++n) {
                        const uint32_t d = tmem_base + h * 128 + (n - nb0) * 32;
                        const uint32_t a_s = a_lo0 + (uint32_t)(s * C) * (kTcABytes >> 4) + n;
#pragma unroll
                        for (int c = 0; c < C; ++c)
#pragma unroll
                            for (int sp = 0; sp < SPLITS; ++sp)
                                umma_ss(d, a_s + c * (kTcABytes >> 4), a_hi, b_lo0 + (c * SPLITS + sp) * (kTcBBytes >> 4), b_hi, kIdesc,
                                        (c | sp) != 0);
                    }
                    umma_commit(bar_tfull + 8 * h);
                }
                __syncwarp();
            }
        }
    } else {
        // ===================== epilogue: thread == window =====================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int b = b0 + row;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t swz = (uint32_t)(row & 7);
        float pm6[kCMid], pm7[kCMid], ah[4][kCMid], c2c = 0.f, nan_probe = 0.f;
#pragma unroll
        for (int o = 0; o < kCMid; ++o) {
            pm6[o] = 0.f; pm7[o] = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) ah[i][o] = 0.f;
        }
        const bool row_ok = b < p.B;
        float *fout = p.feats + (int64_t)b * p.sB + (int64_t)p0 * p.sP;

        for (int i = 0; i < ntiles; ++i) {
            const int s = i & 1;
            mbar_wait(bar_full + 8 * s, (i >> 1) & 1);     // TMA bytes visible to this thread too
            const uint8_t *tile = sA + (size_t)s * C * kTcABytes + row * 128;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                mbar_wait(bar_tfull + 8 * h, i & 1);
                tc_fence_after();
                const int nb0 = h ? 4 : 0, nb1 = h ? kTcBlocks : 4;
#pragma unroll
                for (int n = nb0; n < nb1; ++n) {
                    float D[32];
                    tmem_ld32(taddr + h * 128 + (n - nb0) * 32, D);
                    if (n == nb1 - 1) {                    // accumulator stage drained -> MMA may refill it
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_tempty + 8 * h);
                    }
                    const int j = i * kTcBlocks + n;       // global step index
                    // ---- the tap that does not fit the 16-sample slice: previous block's s=7, k=9
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const uint16_t raw = *reinterpret_cast<const uint16_t *>(
                            tile + c * kTcABytes + ((uint32_t)((n + 1) ^ swz) << 4));
                        const float xv = __uint_as_float((uint32_t)raw << 16);
#pragma unroll
                        for (int o = 0; o < kCMid; ++o) pm7[o] = fmaf(p.w9[o][c], xv, pm7[o]);
                    }
                    // ---- pool1 (3,2) on pre-activations, then tanh(+bias): a1 positions 4j-1 .. 4j+2
                    float an[4][kCMid];
#pragma unroll
                    for (int o = 0; o < kCMid; ++o) {
                        an[0][o] = tanh_fold(max3_nan(pm6[o], pm7[o], D[0 * 4 + o]), p.b1s[o]);
                        an[1][o] = tanh_fold(max3_nan(D[0 * 4 + o], D[1 * 4 + o], D[2 * 4 + o]), p.b1s[o]);
                        an[2][o] = tanh_fold(max3_nan(D[2 * 4 + o], D[3 * 4 + o], D[4 * 4 + o]), p.b1s[o]);
                        an[3][o] = tanh_fold(max3_nan(D[4 * 4 + o], D[5 * 4 + o], D[6 * 4 + o]), p.b1s[o]);
                        pm6[o] = D[6 * 4 + o];
                        pm7[o] = D[7 * 4 + o];
                    }
                    // ---- conv2 (no bias yet): outputs r = 4j-5 .. 4j-2 from a1[r .. r+4]
                    float c2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int c = 0; c < kCMid; ++c) {
                        const float A8[8] = {ah[0][c], ah[1][c], ah[2][c], ah[3][c], an[0][c], an[1][c], an[2][c], an[3][c]};
#pragma unroll
                        for (int k = 0; k < 5; ++k)
#pragma unroll
                            for (int r = 0; r < 4; ++r) c2[r] = fmaf(p.w2[c][k], A8[r + k], c2[r]);
                    }
#pragma unroll
                    for (int c = 0; c < kCMid; ++c)
#pragma unroll
                        for (int r = 0; r < 4; ++r) ah[r][c] = an[r][c];
                    // ---- pool2 (3,2) + tanh(+bias): features 2j-3 and 2j-2
                    const float f0 = tanh_fold(max3_nan(c2c, c2[0], c2[1]), p.b2s);
                    const float f1 = tanh_fold(max3_nan(c2[1], c2[2], c2[3]), p.b2s);
                    c2c = c2[3];
                    nan_probe = fmaf(f0, 0.f, nan_probe);
                    nan_probe = fmaf(f1, 0.f, nan_probe);
                    const int pr0 = 2 * j - 3;
                    if (row_ok) {
                        if (pr0 >= 0 && pr0 < nfeat) fout[(int64_t)pr0 * p.sP] = f0;
                        if (pr0 + 1 >= 0 && pr0 + 1 < nfeat) fout[(int64_t)(pr0 + 1) * p.sP] = f1;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + 8 * s);   // smem stage fully consumed
        }
        if (row_ok && nan_probe != nan_probe) p.nanflag[b] = 1;
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
    }
}

// windows flagged by the tensor-core kernel -> compact index list for the exact re-computation
__global__ void tc_compact_flags_kernel(int *flags, int B, int *list, int *count) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B && flags[b]) list[atomicAdd(count, 1)] = b;      // flags stay set: the head kernel picks the recomputed rows by them
}

}  // namespace b2cnn
#include "b2cnn_tc_fused.cuh"
#include "b2cnn_stream_f32.cuh"
namespace b2cnn {

// ------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

static uint16_t bf16_rn(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static float bf16_to_f(uint16_t h) {
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

// MyCNN2/3/4 geometry: fused kernel only (ARCH 1)
static bool arch1_ok(const Dims &d) {
    return d.K1 == 5 && d.K2 == 5 && d.PK == 2 && d.PS == 2 && d.C >= 1 && d.C <= 3 && d.act == B2CNN_ACT_TANH &&
           !d.has_affine && d.L >= 32;
}
static bool arch_ok(const Dims &d) {
    return d.K1 == 10 && d.K2 == 5 && d.PK == 3 && d.PS == 2 && d.C >= 1 && d.C <= kTcMaxC &&
           d.act == B2CNN_ACT_TANH && !d.has_affine && d.L >= 32;
}

static int tiles_per_cta_for(const Dims &d);
static int tiles_per_cta_s_for(const Dims &d);

int tc_prepare(TcState &s, const Dims &d, const ConvWeights &cw, const float *d_wih0, const HeadWeights &, int splits, int,
               cudaStream_t st) {
    s.ready = false;
    s.splits = splits;
    s.fused_ready = false;
    if (!arch_ok(d) && !arch1_ok(d)) return 0;   // not an error: this shape takes the generic path
    if (!get_encode()) return 0;
#if B2CNN_OWN_FLAGS
    if (!s.d_flagstate) {
        const int64_t cap = kOwnFlagCap;
        if (cudaMalloc(reinterpret_cast<void **>(&s.d_flagstate), sizeof(int) * (2 * cap + 16)) == cudaSuccess &&
            cudaMemset(s.d_flagstate, 0, sizeof(int) * (2 * cap + 16)) == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess) {
            s.flag_cap = cap; s.flags_clean = true;
        } else {
            cudaFree(s.d_flagstate); s.d_flagstate = nullptr; s.flag_cap = 0;    // not an error: the workspace copy is used
            (void)cudaGetLastError();
        }
    }
#endif
    // band matrices: piece sp of T_c[k][(s,o)] = w1[o][c][k-s], stored as UMMA K-major
    // no-swizzle core matrices: byte = (n/8)*256 + (k/8)*128 + (n%8)*16 + (k%8)*2, n = s*4+o.
    //   d_bmats   [C][3][1 KB]  three bf16 pieces (hi/mid/lo: the full 24-bit fp32 mantissa)
    //   d_bmats2  [C][2][1 KB]  the first two pieces only (16 mantissa bits; tc_splits=2, an option:
    //                           6 instead of 9 MMAs per block, weights rounded to 2^-17 relative)
    // (A mixed-format MMA -- bf16 samples x fp16 weight pieces, 22 bits in two pieces -- was tried:
    //  kind::f16 with a_format != b_format raises "illegal instruction" on sm_100a.)
    std::vector<uint16_t> host((size_t)d.C * 3 * 512, 0), host2((size_t)d.C * 2 * 512, 0);
    for (int c = 0; c < d.C; ++c)
        for (int sft = 0; sft < 8; ++sft)
            for (int o = 0; o < kCMid; ++o)
                for (int k = 0; k < 16; ++k) {
                    const int tap = k - sft;
                    if (tap < 0 || tap >= d.K1) continue;
                    float w = cw.w1[(c * d.K1 + tap) * kCMid + o];
                    const int n = sft * 4 + o;
                    const size_t off = (size_t)(n / 8) * 128 + (k / 8) * 64 + (n % 8) * 8 + (k % 8);   // in bf16 units
                    for (int sp = 0; sp < 3; ++sp) {
                        const uint16_t piece = bf16_rn(w);
                        host[((size_t)c * 3 + sp) * 512 + off] = piece;
                        if (sp < 2) host2[((size_t)c * 2 + sp) * 512 + off] = piece;
                        w -= bf16_to_f(piece);
                    }
                }
    if (!s.d_bmats && cudaMalloc(&s.d_bmats, host.size() * 2 + 16) != cudaSuccess) { g_tc_err = "cudaMalloc(band matrices)"; return -1; }
    if (!s.d_bmats2 && cudaMalloc(&s.d_bmats2, host2.size() * 2 + 16) != cudaSuccess) { g_tc_err = "cudaMalloc(band matrices)"; return -1; }
    if (cudaMemcpyAsync(s.d_bmats, host.data(), host.size() * 2, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(s.d_bmats2, host2.data(), host2.size() * 2, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) { g_tc_err = "upload band matrices"; return -1; }
    s.ready = true;
    s.has_v1 = arch_ok(d);
    // ---- fused kernel: W_ih_l0 packed per (range, chunk)
    s.fused_ready = false;
    s.tiles_per_cta = tiles_per_cta_for(d);
    s.feats_per_cta = 14 * s.tiles_per_cta - 4;      // even: every range starts 16-byte aligned (TMA)
    s.chunks_per_cta = (7 * s.tiles_per_cta + 7) / 8;
    s.n_ranges = (d.L + s.feats_per_cta - 1) / s.feats_per_cta;
    if (d.C <= 3) {
        const size_t bytes = (size_t)s.n_ranges * s.chunks_per_cta * kFuWChunkBytes;
        cudaFree(s.d_wpack); s.d_wpack = nullptr;
        if (cudaMalloc(&s.d_wpack, bytes) != cudaSuccess) { g_tc_err = "cudaMalloc(packed W_ih)"; return -1; }
        const int64_t total = (int64_t)s.n_ranges * s.chunks_per_cta * 1024;
        tc_pack_wih_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_wih0, reinterpret_cast<uint8_t *>(s.d_wpack), d.L,
                                                                          s.feats_per_cta, s.chunks_per_cta, s.n_ranges, d.K1 == 10 ? 3 : 2);
        if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) { g_tc_err = "pack W_ih"; return -1; }
        s.fused_ready = true;
        // stream_f32_kernel: 3-block tiles, its own (L-only) position ranges
        s.stream_ready = false;
        s.tiles_per_cta_s = tiles_per_cta_s_for(d);
        s.feats_per_cta_s = 2 * kSfBlocks * s.tiles_per_cta_s - 4;
        s.chunks_per_cta_s = (kSfBlocks * s.tiles_per_cta_s + 7) / 8;
        s.n_ranges_s = (d.L + s.feats_per_cta_s - 1) / s.feats_per_cta_s;
        const size_t bytes_s = (size_t)s.n_ranges_s * s.chunks_per_cta_s * kFuWChunkBytes;
        cudaFree(s.d_wpack_s); s.d_wpack_s = nullptr;
        if (cudaMalloc(&s.d_wpack_s, bytes_s) != cudaSuccess) { g_tc_err = "cudaMalloc(packed W_ih, fp32 stream)"; return -1; }
        const int64_t total_s = (int64_t)s.n_ranges_s * s.chunks_per_cta_s * 1024;
        tc_pack_wih_kernel<<<(unsigned)((total_s + 255) / 256), 256, 0, st>>>(d_wih0, reinterpret_cast<uint8_t *>(s.d_wpack_s), d.L,
                                                                            s.feats_per_cta_s, s.chunks_per_cta_s, s.n_ranges_s, d.K1 == 10 ? 3 : 2);
        if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) { g_tc_err = "pack W_ih (fp32 stream)"; return -1; }
        s.stream_ready = true;
    }
    return 0;
}

void tc_release(TcState &s) {
    cudaFree(s.d_flagstate);
    s.d_flagstate = nullptr; s.flag_cap = 0; s.flags_clean = false; s.owner_set = false;
    cudaFree(s.d_bmats);
    cudaFree(s.d_bmats2);
    cudaFree(s.d_wpack);
    cudaFree(s.d_wpack_s);
    s.d_wpack_s = nullptr;
    s.stream_ready = false;
    s.d_bmats = nullptr;
    s.d_bmats2 = nullptr;
    s.d_wpack = nullptr;
    s.ready = s.fused_ready = false;
}

bool tc_supported(const TcState &s, const Dims &d, int dtype, int64_t B, int mode) {
    (void)mode; (void)B;
    return s.ready && dtype == B2CNN_DTYPE_BF16 && (arch_ok(d) || (arch1_ok(d) && s.fused_ready && s.opt_fused));
}
bool tc_can_emit_features(const TcState &s) { return s.ready && s.has_v1; }

// A TMA tensor map needs a row pitch that is a multiple of 16 bytes.  Windows whose length is not a
// multiple of 8 samples (7500, 37500 ...) are copied once into a pitch-aligned scratch (costs one
// extra read + write of the input; W % 8 == 0, e.g. the headline 75000, streams straight from x).
static int64_t padded_w(const Dims &d) { return (d.W + 7) & ~7; }
static int64_t flags_bytes(int64_t B) { return ((2 * B + 64) * 4 + 255) / 256 * 256; }
static int64_t xpad_bytes(const Dims &d, int64_t B) { return (d.W % 8) ? (B * d.C * padded_w(d) * 2 + 255) / 256 * 256 : 0; }

// tc workspace: nan flags [B] + list [B] + count, then the optional pitch-aligned copy of x
int64_t tc_workspace_bytes(const TcState &s, const Dims &d, int64_t B) {
    if (!s.ready) return 0;
    return flags_bytes(B) + xpad_bytes(d, B);
}

// rows of W samples, `sp` elements apart -> rows Wp (multiple of 8) apart, tail zero-filled.
// VEC 8: W % 4 == 0 and sp % 4 == 0: every row starts 8-byte aligned; one thread moves 16 output bytes with two
//        8-byte loads (a 16-byte load would be misaligned on every other row) and ONE 16-byte store.
// VEC 1: any W (2-byte accesses).
template <int VEC>
__global__ void tc_repack_rows_kernel(const uint16_t *__restrict__ src, uint16_t *__restrict__ dst, int64_t rows, int W, int64_t sp, int Wp) {
    const int per_row = Wp / VEC;
    for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
        const uint16_t *in = src + r * sp;
        uint16_t *out = dst + r * Wp;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per_row; i += gridDim.x * blockDim.x) {
            if (VEC == 8) {
                uint2 a = make_uint2(0u, 0u), b = make_uint2(0u, 0u);
                if (i * 8 < W) a = __ldg(reinterpret_cast<const uint2 *>(in + i * 8));
                if (i * 8 + 4 < W) b = __ldg(reinterpret_cast<const uint2 *>(in + i * 8 + 4));
                *reinterpret_cast<uint4 *>(out + i * 8) = make_uint4(a.x, a.y, b.x, b.y);
            } else {
                out[i] = i < W ? in[i] : (uint16_t)0;
            }
        }
    }
}

// returns the pointer / pitch the tensor map must describe (x itself when its rows are 16-byte aligned: W % 8 == 0
// for a contiguous tensor, or a producer that padded the row pitch -- b2cnn_forward_pitched)
static const void *tc_stage_input(const Dims &d, const void *x, int64_t B, void *scratch_after_flags, int64_t *pitch,
                                  int *launches, cudaStream_t st) {
    *pitch = d.XP;
    if ((d.XP % 8) == 0) return x;
    const int64_t rows = B * d.C;
    const int Wp = (int)padded_w(d);
    dim3 grid(4, (unsigned)(rows < 32768 ? rows : 32768));
    if (d.W % 4 == 0 && d.XP % 4 == 0)
        tc_repack_rows_kernel<8><<<grid, 256, 0, st>>>(reinterpret_cast<const uint16_t *>(x), reinterpret_cast<uint16_t *>(scratch_after_flags), rows, d.W, d.XP, Wp);
    else
        tc_repack_rows_kernel<1><<<grid, 256, 0, st>>>(reinterpret_cast<const uint16_t *>(x), reinterpret_cast<uint16_t *>(scratch_after_flags), rows, d.W, d.XP, Wp);
    *pitch = Wp;
    ++*launches;
    return scratch_after_flags;
}

static int tiles_per_cta_for(const Dims &d) {
    // About 37 position ranges per window whatever its length (depends on L only, so a window's
    // summation order never depends on the batch): L=18745 -> 37 tiles (514 features) per CTA and
    // 37 ranges x 16 window-tile pairs = 592 CTAs = 4 waves of 148 at B=4096; shorter windows get
    // proportionally shorter ranges so that small batches still fill the SMs.
    if (const char *e = getenv("B2CNN_TC_TILES")) { const int v = atoi(e); if (v >= 1 && v <= 4096) return v; }
    int nt = ((d.L + 36) / 37 + 4 + 13) / 14;
    if (nt < 2) nt = 2;
    if (nt > 37) nt = 37;
    return nt;
}

static int tiles_per_cta_s_for(const Dims &d) {
    // stream_f32_kernel: about 37 position ranges (L only) like the bf16 kernel; a CTA's stream emits
    // 6*tiles - 4 features (even: every range starts 16-byte aligned for TMA)
    if (const char *e = getenv("B2CNN_SF_TILES")) { const int v = atoi(e); if (v >= 1 && v <= 16384) return v; }
    int nt = ((d.L + 36) / 37 + 4 + 5) / 6;
    if (nt < 4) nt = 4;
    return nt;
}

static int launch_tc_kernel(const TcState &s, const Dims &d, const ConvWeights &cw, const void *x, int64_t pitch, int64_t B,
                            float *feats, int64_t sB, int64_t sP, int *nanflag, cudaStream_t st, const char **err) {
    if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) { *err = "x must be 16-byte aligned for TMA"; return -1; }
    CUtensorMap tm;
    cuuint64_t gdim[3] = {(cuuint64_t)d.W, (cuuint64_t)d.C, (cuuint64_t)B};
    cuuint64_t gstr[2] = {(cuuint64_t)pitch * 2, (cuuint64_t)d.C * pitch * 2};
    cuuint32_t box[3] = {64, 1, kTcM};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = get_encode()(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(x), gdim, gstr, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { *err = "cuTensorMapEncodeTiled failed"; return -1; }
    TcParams p;
    memset(&p, 0, sizeof p);
    p.feats = feats; p.sB = sB; p.sP = sP; p.nanflag = nanflag;
    p.bmats = reinterpret_cast<const uint8_t *>(s.d_bmats);
    p.B = (int)B; p.W = d.W; p.L = d.L;
    // a CTA's stream can emit 14*tiles - 3 features; an EVEN count keeps every range's first
    // sample (4 * p0 elements) 16-byte aligned: an unaligned TMA box start faults (measured).
    p.tiles_per_cta = s.tiles_per_cta;
    p.feats_per_cta = s.feats_per_cta;
    for (int o = 0; o < kCMid; ++o) {
        for (int c = 0; c < d.C; ++c) p.w9[o][c] = cw.w1[(c * d.K1 + 9) * kCMid + o];
        p.b1s[o] = cw.b1[o] * k2Log2e;
        for (int k = 0; k < 5; ++k) p.w2[o][k] = cw.w2[o * d.K2 + k];
    }
    p.b2s = cw.b2 * k2Log2e;
    const int n_pr = (d.L + p.feats_per_cta - 1) / p.feats_per_cta;
    dim3 grid((unsigned)((B + kTcM - 1) / kTcM), n_pr);
    const size_t smem = (size_t)2 * d.C * kTcABytes + (size_t)d.C * 3 * kTcBBytes + 256 + 1024;
#define TC_LAUNCH(CC, SS)                                                                              \
    if (d.C == CC) {                                                                 \
        cudaError_t e = cudaFuncSetAttribute(tc_frontend_kernel<CC, SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) { *err = cudaGetErrorString(e); return -1; }                             \
        tc_frontend_kernel<CC, SS><<<grid, kTcThreads, smem, st>>>(tm, p);                             \
    } else
    TC_LAUNCH(3, 3) TC_LAUNCH(1, 3) TC_LAUNCH(2, 3) TC_LAUNCH(4, 3)
    { *err = "no tensor-core instantiation for this channel count"; return -1; }
#undef TC_LAUNCH
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return -1; }
    return 1;
}

// TMA/tcgen05 front end + NaN-flag compaction + exact re-computation of flagged windows.
// `ws`: 2*B+64 ints of scratch (flags, list, count); pass nullptr to allocate stream-ordered.
int tc_frontend(TcState &s, const Dims &d, const ConvWeights &cw, const void *x, int64_t B, float *feats,
                int64_t sB, int64_t sP, void *ws, int num_sms, cudaStream_t st, const char **err) {
    int *flags = reinterpret_cast<int *>(ws);
    const bool own = flags == nullptr;
    if (own && cudaMallocAsync(&flags, (size_t)tc_workspace_bytes(s, d, B), st) != cudaSuccess) { *err = "cudaMallocAsync"; return -1; }
    int *list = flags + B, *count = list + B;
    int launches = -1;
    if (cudaMemsetAsync(flags, 0, sizeof(int) * (2 * B + 1), st) != cudaSuccess) {
        *err = "memset flags";
    } else {
        int staged = 0;
        int64_t pitch = d.XP;
        const void *xin = tc_stage_input(d, x, B, reinterpret_cast<char *>(flags) + flags_bytes(B), &pitch, &staged, st);
        int n = launch_tc_kernel(s, d, cw, xin, pitch, B, feats, sB, sP, flags, st, err);
        if (n >= 0) {
            tc_compact_flags_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(flags, (int)B, list, count);
            int m = launch_frontend_generic_listed(d, cw, x, B2CNN_DTYPE_BF16, B, feats, sB, sP, list, count, st, num_sms, err);
            if (m >= 0) launches = staged + n + 1 + m;
        }
    }
    if (own) cudaFreeAsync(flags, st);
    return launches;
}

int tc_features(TcState &s, const Dims &d, const ConvWeights &cw, const void *x, int64_t B, float *feats,
                int num_sms, cudaStream_t st, const char **err) {
    // parity-test entry: row-major [B][L] features (uncoalesced stores; not a timed path)
    return tc_frontend(s, d, cw, x, B, feats, d.L, 1, nullptr, num_sms, st, err);
}


bool tc_fused_supported(const TcState &s, const Dims &d, int dtype) {
    return s.ready && s.fused_ready && s.opt_fused && dtype == B2CNN_DTYPE_BF16 && (arch_ok(d) || arch1_ok(d)) && d.C <= 3;
}
int tc_partial_slices(const TcState &s) {
    if (!s.fused_ready) return 0;
    return s.stream_ready && s.n_ranges_s > s.n_ranges ? s.n_ranges_s : s.n_ranges;
}

// Where this call keeps count | flags | list: the handle's own, already-zero copy (see TcState) or the head of the
// workspace, zeroed here.  count and flags are adjacent in both, the list needs no zeroing.
static int flag_bufs(TcState &s, void *ws, int64_t B, cudaStream_t st, bool cleaning_head_follows, const char **err) {
    bool own = false;
#if B2CNN_OWN_FLAGS
    if (s.d_flagstate && B <= s.flag_cap && cleaning_head_follows) {
        if (!s.owner_set) { s.owner_set = true; s.owner_stream = st; }
        own = s.owner_stream == st;
    }
#endif
    if (own) {
        s.cur_count = s.d_flagstate; s.cur_flags = s.cur_count + 16; s.cur_list = s.cur_flags + s.flag_cap;
        if (!s.flags_clean && cudaMemsetAsync(s.cur_count, 0, sizeof(int) * (s.flag_cap + 16), st) != cudaSuccess) { *err = "memset flags"; return -1; }
        s.flags_clean = false;                    // until the cleaning head kernel of this call has been launched
    } else {
        s.cur_count = reinterpret_cast<int *>(ws); s.cur_flags = s.cur_count + 16; s.cur_list = s.cur_flags + B;
        if (cudaMemsetAsync(s.cur_count, 0, sizeof(int) * (B + 16), st) != cudaSuccess) { *err = "memset flags"; return -1; }
    }
    s.cur_own = own;
    return 0;
}

static int make_tmap(const Dims &d, const void *x, int64_t pitch, int64_t B, CUtensorMap *tm, const char **err) {
    if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) { *err = "x must be 16-byte aligned for TMA"; return -1; }
    cuuint64_t gdim[3] = {(cuuint64_t)d.W, (cuuint64_t)d.C, (cuuint64_t)B};
    cuuint64_t gstr[2] = {(cuuint64_t)pitch * 2, (cuuint64_t)d.C * pitch * 2};
    cuuint32_t box[3] = {64, 1, kTcM};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = get_encode()(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(x), gdim, gstr, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { *err = "cuTensorMapEncodeTiled failed"; return -1; }
    return 0;
}

// fused front end + projection -> gates[B][64]; flagged (NaN) windows are recomputed exactly.
int tc_fused_gates(TcState &s, const Dims &d, const ConvWeights &cw, const HeadWeights &hw, const void *x, int64_t B,
                   float *feats, float *partial, float *gates, void *ws, int num_sms, cudaStream_t st, const char **err,
                   bool reduce_here, int *slices_out) {
    (void)feats;
    // scratch ints: count | flags | list  (the handle's own zero-between-calls copy, or the head of the workspace)
    if (flag_bufs(s, ws, B, st, !reduce_here, err) != 0) return -1;
    int *count = s.cur_count, *flags = s.cur_flags, *list = s.cur_list;
    int staged = 0;
    int64_t pitch = d.XP;
    const void *xin = tc_stage_input(d, x, B, reinterpret_cast<char *>(ws) + flags_bytes(B), &pitch, &staged, st);
    CUtensorMap tm;
    if (make_tmap(d, xin, pitch, B, &tm, err) != 0) return -1;
    TcFusedParams p;
    memset(&p, 0, sizeof p);
    p.partial = partial; p.nanflag = flags; p.list = list; p.count = count;
    p.wpack = reinterpret_cast<const uint8_t *>(s.d_wpack);
    p.B = (int)B; p.W = d.W; p.L = d.L;
    p.tiles_per_cta = s.tiles_per_cta; p.feats_per_cta = s.feats_per_cta; p.chunks_per_cta = s.chunks_per_cta;
    for (int q2 = 0; q2 < 2; ++q2) {
        for (int c = 0; c < d.C; ++c)
            p.w9p[c][q2] = d.K1 == 10 ? make_float2(cw.w1[(c * d.K1 + 9) * kCMid + 2 * q2], cw.w1[(c * d.K1 + 9) * kCMid + 2 * q2 + 1])
                                      : make_float2(0.f, 0.f);
        p.b1sp[q2] = make_float2(cw.b1[2 * q2] * k2Log2e, cw.b1[2 * q2 + 1] * k2Log2e);
        // conv2 consumes r = (1 - tanh)/2 of conv1: sum w*(1 - 2r) = sum(w) + sum (-2w)*r
        for (int k = 0; k < 5; ++k) p.w2p[q2][k] = make_float2(-2.f * cw.w2[(2 * q2) * d.K2 + k], -2.f * cw.w2[(2 * q2 + 1) * d.K2 + k]);
    }
    {
        double sw = 0.0;
        for (int i = 0; i < kCMid * d.K2; ++i) sw += cw.w2[i];
        p.b2s = (float)((cw.b2 + sw) * (double)k2Log2e);
    }
    const int arch_id = d.K1 == 10 ? 0 : 1;
    const int sp = s.splits;                          // bf16 pieces per conv1 weight: 3 (fp32-equivalent) or 2
    p.bmats = reinterpret_cast<const uint8_t *>(sp == 2 ? s.d_bmats2 : s.d_bmats);
    cudaError_t le = cudaSuccess;
    bool launched = false;
    {
        // persistent grid: one CTA per SM (218 KB of shared memory and all 512 TMEM columns each), static round-robin over
        // the (window-tile pair, position range) items
        const int64_t n_items = ((B + 2 * kTcM - 1) / (2 * kTcM)) * s.n_ranges;
        if (n_items > 0x7fffffff) { *err = "batch too large for the fused kernel's item index"; return -1; }
        p.n_ranges = s.n_ranges; p.n_items = (int)n_items;
        dim3 grid((unsigned)(n_items < num_sms ? n_items : num_sms));
        const size_t smem = (size_t)4 * d.C * kTcABytes + (size_t)d.C * sp * kTcBBytes + 2 * kFuWChunkBytes + FuBars::kTotal * 8 + 16;
#define FU_LAUNCH(CC, SS, AA)                                                                          \
        if (!launched && d.C == CC && sp == SS && arch_id == AA) {                                     \
            le = cudaFuncSetAttribute(tc_fused_kernel<CC, SS, AA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (le == cudaSuccess) tc_fused_kernel<CC, SS, AA><<<grid, kFuThreads, smem, st>>>(tm, p); \
            launched = true;                                                                           \
        }
        FU_LAUNCH(3, 2, 0) FU_LAUNCH(3, 3, 0) FU_LAUNCH(2, 2, 0) FU_LAUNCH(1, 2, 0) FU_LAUNCH(2, 3, 0) FU_LAUNCH(1, 3, 0)
        FU_LAUNCH(3, 2, 1) FU_LAUNCH(3, 3, 1) FU_LAUNCH(2, 2, 1) FU_LAUNCH(1, 2, 1) FU_LAUNCH(2, 3, 1) FU_LAUNCH(1, 3, 1)
#undef FU_LAUNCH
    }
    if (!launched) { *err = "no fused instantiation for this channel count / split"; return -1; }
    if (le != cudaSuccess) { *err = cudaGetErrorString(le); return -1; }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return -1; }
    int launches = 1 + staged;
    if (slices_out) *slices_out = s.n_ranges;
    // the exception path: exact gate partials of the flagged windows, written over their rows of `partial`
    // (one launch; the list was compacted by the kernel itself, an empty list costs one almost-empty launch)
    int n = launch_frontend_generic_gates_listed(d, cw, x, B2CNN_DTYPE_BF16, B, hw.wih0T, partial, s.n_ranges, list, count, st, num_sms, err);
    if (n < 0) return -1;
    launches += n;
    // reduce_here == false: the caller's head kernel sums the range partials itself (independent windows)
    n = reduce_here ? launch_reduce_gates(partial, s.n_ranges, B, hw, gates, st, err) : 0;
    if (n < 0) return -1;
    return launches + n;
}

#ifdef B2CNN_TIMING
// experiments only: read and clear the fused kernel's wait counters (cycles summed over warps and CTAs)
extern "C" int b2cnn_debug_timing(unsigned long long *out16) {
    if (cudaMemcpyFromSymbol(out16, g_fu_timing, sizeof(unsigned long long) * 16) != cudaSuccess) return -1;
    unsigned long long z[16] = {0};
    return cudaMemcpyToSymbol(g_fu_timing, z, sizeof z) == cudaSuccess ? 0 : -1;
}
#endif

bool tc_stream_supported(const TcState &s, const Dims &d, int dtype) {
    return s.ready && s.stream_ready && dtype == B2CNN_DTYPE_F32 && (arch_ok(d) || arch1_ok(d)) && d.C <= 3 && (d.XP % 4) == 0;
}

// fp32 windows: streaming front end + projection -> gates[B][64]; flagged (NaN) windows are recomputed exactly.
int tc_stream_gates(TcState &s, const Dims &d, const ConvWeights &cw, const HeadWeights &hw, const void *x, int64_t B,
                    float *feats, float *partial, float *gates, void *ws, int num_sms, cudaStream_t st, const char **err,
                    bool reduce_here, int *slices_out) {
    (void)feats;
    if (flag_bufs(s, ws, B, st, !reduce_here, err) != 0) return -1;
    int *count = s.cur_count, *flags = s.cur_flags, *list = s.cur_list;
    if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) { *err = "x must be 16-byte aligned for TMA"; return -1; }
    CUtensorMap tm;
    {
        cuuint64_t gdim[3] = {(cuuint64_t)d.W, (cuuint64_t)d.C, (cuuint64_t)B};
        cuuint64_t gstr[2] = {(cuuint64_t)d.XP * 4, (cuuint64_t)d.C * d.XP * 4};
        cuuint32_t box[3] = {32, 1, kTcM};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = get_encode()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void *>(x), gdim, gstr, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { *err = "cuTensorMapEncodeTiled (fp32) failed"; return -1; }
    }
    StreamF32Params pp;
    memset(&pp, 0, sizeof pp);
    TcFusedParams &p = pp.f;
    p.partial = partial; p.nanflag = flags; p.list = list; p.count = count;
    p.wpack = reinterpret_cast<const uint8_t *>(s.d_wpack_s);
    p.B = (int)B; p.W = d.W; p.L = d.L;
    p.tiles_per_cta = s.tiles_per_cta_s; p.feats_per_cta = s.feats_per_cta_s; p.chunks_per_cta = s.chunks_per_cta_s;
    for (int q2 = 0; q2 < 2; ++q2) {
        for (int c = 0; c < d.C; ++c)
            for (int k = 0; k < d.K1; ++k)
                pp.w1p[c][k][q2] = make_float2(cw.w1[(c * d.K1 + k) * kCMid + 2 * q2], cw.w1[(c * d.K1 + k) * kCMid + 2 * q2 + 1]);
        p.b1sp[q2] = make_float2(cw.b1[2 * q2] * k2Log2e, cw.b1[2 * q2 + 1] * k2Log2e);
        for (int k = 0; k < 5; ++k) p.w2p[q2][k] = make_float2(-2.f * cw.w2[(2 * q2) * d.K2 + k], -2.f * cw.w2[(2 * q2 + 1) * d.K2 + k]);
    }
    {
        double sw = 0.0;
        for (int i = 0; i < kCMid * d.K2; ++i) sw += cw.w2[i];
        p.b2s = (float)((cw.b2 + sw) * (double)k2Log2e);
    }
    dim3 grid((unsigned)((B + 2 * kTcM - 1) / (2 * kTcM)), s.n_ranges_s);
    const size_t smem = (size_t)4 * d.C * kSfABytes + 2 * kFuWChunkBytes + SfBars::kTotal * 8 + 16;
    const int arch_id = d.K1 == 10 ? 0 : 1;
    cudaError_t le = cudaSuccess;
    bool launched = false;
#define SF_LAUNCH(CC, AA)                                                                              \
    if (!launched && d.C == CC && arch_id == AA) {                                                     \
        le = cudaFuncSetAttribute(stream_f32_kernel<CC, AA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (le == cudaSuccess) stream_f32_kernel<CC, AA><<<grid, kSfThreads, smem, st>>>(tm, pp);      \
        launched = true;                                                                               \
    }
    SF_LAUNCH(3, 0) SF_LAUNCH(2, 0) SF_LAUNCH(1, 0) SF_LAUNCH(3, 1) SF_LAUNCH(2, 1) SF_LAUNCH(1, 1)
#undef SF_LAUNCH
    if (!launched) { *err = "no fp32 stream instantiation for this channel count"; return -1; }
    if (le != cudaSuccess) { *err = cudaGetErrorString(le); return -1; }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { *err = cudaGetErrorString(e); return -1; }
    int launches = 1;
    if (slices_out) *slices_out = s.n_ranges_s;
    int n = launch_frontend_generic_gates_listed(d, cw, x, B2CNN_DTYPE_F32, B, hw.wih0T, partial, s.n_ranges_s, list, count, st, num_sms, err);
    if (n < 0) return -1;
    launches += n;
    n = reduce_here ? launch_reduce_gates(partial, s.n_ranges_s, B, hw, gates, st, err) : 0;
    if (n < 0) return -1;
    return launches + n;
}

}  // namespace b2cnn
