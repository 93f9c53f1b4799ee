// b2cnn_stream_f32.cuh -- the fused front end + projection for FLOAT32 windows (included by b2cnn_tc.cu
// after b2cnn_tc_fused.cuh; shares its parameters, W_ih packing format and projection).
//
// The reference's tensors are fp32 (bin/predictStream.py:155).  bf16 windows run conv1 on the tensor cores
// (tc_fused_kernel); an fp32 window would need its samples split into three bf16 pieces first, so this
// kernel keeps the streaming structure -- TMA tiles, thread == window, register-local sliding windows, features
// never leaving the SM, projection on tcgen05 with A in TMEM -- and evaluates conv1 itself on the CUDA cores,
// in exact fp32 FMAs straight from the shared-memory tile:
//
//   stage  = 128 windows x 32 samples of one channel (fp32, 128-byte SWIZZLE_128B rows), 2 stages per tile
//   block  = 8 conv1 positions from the 16 samples at tile offset 8n (n = 0,1,2; tiles advance 24 samples);
//            position 8n+7 misses its tap 9 (sample 8n+16) exactly as in the tensor-core kernel and gets it
//            one block later from sample 8(n+1)+8 of the same stream
//   D[s][o] = sum_c sum_k w1[o][c][k] * x_c[8n+s+k]  for the taps inside the 16 samples (948 FMA per block
//            and window as 474 packed FFMA2 over the channel pairs (o, o+1)); everything after D is the
//            epilogue of tc_fused_kernel, unchanged.
// Only real taps are multiplied, so NaN/inf samples propagate exactly like the reference's conv (no band
// zeros here).  Bound: the FP32 pipe -- 9.0 M conv1 MACs per window is ~4x what the HBM roofline of an fp32
// window allows; see DESIGN.md for the measured fraction.
//
// CTA = 2 window tiles x 1 position range, 384 threads:
//   warp 0 TMA producer | warps 1/2 projection MMA issuer of tile 0/1 | warp 3 TMEM allocator + W_ih producer
//   warps 4-7 / 8-11 epilogue of tile 0 / 1
// TMEM columns per window tile (128): gates 64 | projection pieces 2 x 24.
#pragma once

namespace b2cnn {

constexpr int kSfThreads = 384;
constexpr int kSfBlocks = 3;                 // 8-position blocks per 32-sample tile
constexpr int kSfAdv = 24;                   // samples a tile advances
constexpr int kSfABytes = 128 * 128;         // one channel of one stage: 128 windows x 32 fp32

struct StreamF32Params {
    TcFusedParams f;                         // projection / epilogue constants (w9p, b1sp, w2p, b2s, ranges, pointers)
    float2 w1p[kTcMaxC][10][2];              // conv1 weights paired over out-channels: (w1[2q][c][k], w1[2q+1][c][k])
};

struct SfBars {   // uint64_t slots; per window tile t (stride kPerTile)
    static constexpr int kFull = 0, kEmpty = 2, kPFull = 4, kPEmpty = 6, kGFull = 8, kPerTile = 9;
    static constexpr int kWFull = 2 * kPerTile, kWEmpty = kWFull + 2, kTotal = kWEmpty + 2;
};

template <int C, int ARCH>
__global__ void __launch_bounds__(kSfThreads, 1)
stream_f32_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ StreamF32Params pp) {
    const TcFusedParams &p = pp.f;
    constexpr int K1 = ARCH == 0 ? 10 : 5;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sA = smem;                                        // [2 tiles][2 stages][C][16 KB]
    uint8_t *sW = sA + 2 * 2 * C * kSfABytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sW + 2 * kFuWChunkBytes);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + SfBars::kTotal);
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int idx) -> uint32_t { return bar0 + 8u * (uint32_t)idx; };

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int b_cta = blockIdx.x * 2 * kTcM;
    const int p0 = blockIdx.y * p.feats_per_cta;
    const int nfeat = min(p.feats_per_cta, p.L - p0);
    constexpr int FOFF = ARCH == 0 ? 3 : 2;
    const int nsteps_needed = (nfeat + FOFF - 1) / 2 + 1;
    const int ntiles = (nsteps_needed + kSfBlocks - 1) / kSfBlocks;
    const int J = ntiles * kSfBlocks;
    const int nchunks = (J + 7) / 8;
    const int T0 = p0 * 4;

    if ((smem_u32(smem) & 1023u) != 0) __trap();
    if (threadIdx.x == 0) {
        for (int t = 0; t < 2; ++t) {
            const int o = t * SfBars::kPerTile;
            for (int i = 0; i < 2; ++i) {
                mbar_init(BAR(o + SfBars::kFull + i), 1);
                mbar_init(BAR(o + SfBars::kEmpty + i), 4);
                mbar_init(BAR(o + SfBars::kPFull + i), 4);
                mbar_init(BAR(o + SfBars::kPEmpty + i), 1);
            }
            mbar_init(BAR(o + SfBars::kGFull), 1);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(BAR(SfBars::kWFull + i), 1); mbar_init(BAR(SfBars::kWEmpty + i), 2); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 3) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    auto sA_of = [&](int t, int s, int c) -> uint8_t * { return sA + ((size_t)((t * 2 + s) * C + c)) * kSfABytes; };
    auto col_gates = [&](int t) -> uint32_t { return 128u * t; };
    auto col_pieces = [&](int t, int u) -> uint32_t { return 128u * t + 64u + 24u * u; };

    if (warp == 0) {
        // ===================== producer: fp32 window tiles of both window tiles =====================
        if (lane == 0) {
            for (int i = 0; i < ntiles; ++i) {
                const int s = i & 1, ph = (i >> 1) & 1;
                for (int t = 0; t < 2; ++t) {
                    const int o = t * SfBars::kPerTile;
                    mbar_wait_parked(BAR(o + SfBars::kEmpty + s), ph ^ 1);
                    mbar_expect_tx(BAR(o + SfBars::kFull + s), C * kSfABytes);
#pragma unroll
                    for (int c = 0; c < C; ++c)
                        tma_load_3d(smem_u32(sA_of(t, s, c)), &tmap, T0 + kSfAdv * i, c, b_cta + t * kTcM, BAR(o + SfBars::kFull + s));
                }
            }
        }
    } else if (warp == 1 || warp == 2) {
        // ===================== projection MMA issuer of window tile t =====================
        const int t = warp - 1;
        const int o = t * SfBars::kPerTile;
        const uint64_t w_base = desc_none_kmajor(smem_u32(sW), 128, 256);
        const uint32_t w_lo0 = (uint32_t)w_base, w_hi = (uint32_t)(w_base >> 32);
        for (int m = 0; m < nchunks; ++m) {
            const int u = m & 1, ph = (m >> 1) & 1;
            mbar_wait_parked(BAR(SfBars::kWFull + u), ph);
            mbar_wait_parked(BAR(o + SfBars::kPFull + u), ph);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t d = tmem_base + col_gates(t);
                const uint32_t a0 = tmem_base + col_pieces(t, u);
                const uint32_t w0 = w_lo0 + u * (kFuWChunkBytes >> 4);
                // piece pairs (feature piece, weight piece) with fp + wp <= 2: hh hm mh hl lh mm
                umma_ts(d, a0 + 0, w0 + 0 * 128, w_hi, kIdescProj, m != 0);
                umma_ts(d, a0 + 0, w0 + 1 * 128, w_hi, kIdescProj, 1);
                umma_ts(d, a0 + 8, w0 + 0 * 128, w_hi, kIdescProj, 1);
                umma_ts(d, a0 + 0, w0 + 2 * 128, w_hi, kIdescProj, 1);
                umma_ts(d, a0 + 16, w0 + 0 * 128, w_hi, kIdescProj, 1);
                umma_ts(d, a0 + 8, w0 + 1 * 128, w_hi, kIdescProj, 1);
                umma_commit(BAR(o + SfBars::kPEmpty + u));
                umma_commit(BAR(SfBars::kWEmpty + u));
                if (m + 1 == nchunks) umma_commit(BAR(o + SfBars::kGFull));
            }
            __syncwarp();
        }
    } else if (warp == 3) {
        // ===================== W_ih chunk producer (and TMEM allocator) =====================
        if (lane == 0) {
            const uint8_t *wsrc = p.wpack + (size_t)blockIdx.y * p.chunks_per_cta * kFuWChunkBytes;
            for (int m = 0; m < nchunks; ++m) {
                const int u = m & 1;
                mbar_wait_parked(BAR(SfBars::kWEmpty + u), ((m >> 1) & 1) ^ 1);
                mbar_expect_tx(BAR(SfBars::kWFull + u), kFuWChunkBytes);
                bulk_load_1d(smem_u32(sW + u * kFuWChunkBytes), wsrc + (size_t)m * kFuWChunkBytes, kFuWChunkBytes,
                             BAR(SfBars::kWFull + u));
            }
        }
    } else {
        // ===================== epilogue: thread == window =====================
        // iteration jj: stage A of block jj (conv1 from the smem tile -> pool1 -> activation) and stage B of
        // step jj-1 (conv2 -> pool2 -> tanh -> bf16 pieces -> TMEM), as in tc_fused_kernel.
        const int t = (warp - 4) >> 2;
        const int o_bar = t * SfBars::kPerTile;
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int b = b_cta + t * kTcM + row;
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t swz = (uint32_t)(row & 7);                 // SWIZZLE_128B: 16-byte chunk ^= row % 8
        const bool row_ok = b < p.B;
        float2 pm6[2], pm7[2], abuf[2][4][2], nan_probe = make_float2(0.f, 0.f);
        float c2c = 0.f;
#pragma unroll
        for (int q2 = 0; q2 < 2; ++q2) {
            pm6[q2] = make_float2(0.f, 0.f); pm7[q2] = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 4; ++i) { abuf[0][i][q2] = make_float2(0.f, 0.f); abuf[1][i][q2] = make_float2(0.f, 0.f); }
        }
        int n = 0, ti = 0;

        auto iteration = [&](int jj, auto doA_, auto doB_, auto par_) {
            constexpr bool doA = decltype(doA_)::value, doB = decltype(doB_)::value;
            constexpr int PAR = decltype(par_)::value;
            const int jb = jj - 1, m = jb >> 3, kk = jb & 7, u = m & 1;
            const int s = ti & 1;
            // ---------------- top: barriers ----------------
            if constexpr (doA) {
                if (n == 0) mbar_wait_parked(BAR(o_bar + SfBars::kFull + s), (ti >> 1) & 1);
            }
            if constexpr (doB) {
                if (kk == 0) {                              // first store of chunk m into piece buffer u
                    mbar_wait_parked(BAR(o_bar + SfBars::kPEmpty + u), ((m >> 1) & 1) ^ 1);
                    tc_fence_after();
                }
            }
            // ---------------- middle: straight-line math ----------------
            float2 an[4][2];
            if constexpr (doA) {
                // conv1 of block jj on the CUDA cores: D[s][q2] over the 16 samples at tile offset 8n
                float2 D[8][2];
#pragma unroll
                for (int sft = 0; sft < 8; ++sft) { D[sft][0] = make_float2(0.f, 0.f); D[sft][1] = make_float2(0.f, 0.f); }
                const uint8_t *tile = sA_of(t, s, 0) + row * 128;
#pragma unroll 1
                for (int c = 0; c < C; ++c) {
                    float xs[16];
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        const float4 v = *reinterpret_cast<const float4 *>(tile + c * kSfABytes + ((uint32_t)((2 * n + j4) ^ swz) << 4));
                        xs[4 * j4 + 0] = v.x; xs[4 * j4 + 1] = v.y; xs[4 * j4 + 2] = v.z; xs[4 * j4 + 3] = v.w;
                    }
                    if constexpr (ARCH == 0) {
                        // tap 9 of the PREVIOUS block's position 7: sample 8(n-1)+16 of that block == sample 8 of this one
#pragma unroll
                        for (int q2 = 0; q2 < 2; ++q2) pm7[q2] = fma2(pp.w1p[c][9][q2], make_float2(xs[8], xs[8]), pm7[q2]);
                    }
#pragma unroll
                    for (int k = 0; k < K1; ++k)
#pragma unroll
                        for (int sft = 0; sft < 8; ++sft)
                            if (sft + k < 16) {
#pragma unroll
                                for (int q2 = 0; q2 < 2; ++q2)
                                    D[sft][q2] = fma2(pp.w1p[c][k][q2], make_float2(xs[sft + k], xs[sft + k]), D[sft][q2]);
                            }
                }
                if constexpr (ARCH == 0) {
#pragma unroll
                    for (int q2 = 0; q2 < 2; ++q2) {
                        an[0][q2] = sig_fold2(make_float2(max3_nan(pm6[q2].x, pm7[q2].x, D[0][q2].x),
                                                          max3_nan(pm6[q2].y, pm7[q2].y, D[0][q2].y)), p.b1sp[q2]);
                        an[1][q2] = sig_fold2(make_float2(max3_nan(D[0][q2].x, D[1][q2].x, D[2][q2].x),
                                                          max3_nan(D[0][q2].y, D[1][q2].y, D[2][q2].y)), p.b1sp[q2]);
                        an[2][q2] = sig_fold2(make_float2(max3_nan(D[2][q2].x, D[3][q2].x, D[4][q2].x),
                                                          max3_nan(D[2][q2].y, D[3][q2].y, D[4][q2].y)), p.b1sp[q2]);
                        an[3][q2] = sig_fold2(make_float2(max3_nan(D[4][q2].x, D[5][q2].x, D[6][q2].x),
                                                          max3_nan(D[4][q2].y, D[5][q2].y, D[6][q2].y)), p.b1sp[q2]);
                        pm6[q2] = D[6][q2];
                        pm7[q2] = D[7][q2];
                    }
                } else {
#pragma unroll
                    for (int q2 = 0; q2 < 2; ++q2)
#pragma unroll
                        for (int i2 = 0; i2 < 4; ++i2)
                            an[i2][q2] = sig_fold2(make_float2(max_nan(D[2 * i2][q2].x, D[2 * i2 + 1][q2].x),
                                                               max_nan(D[2 * i2][q2].y, D[2 * i2 + 1][q2].y)), p.b1sp[q2]);
                }
            }
            if constexpr (doB) {
                // one accumulator per output position over both channel pairs: c2 = acc.x + acc.y (no pair-sum step)
                float2 acc[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[r] = make_float2(0.f, 0.f);
#pragma unroll
                for (int q2 = 0; q2 < 2; ++q2) {
                    const float2 A8[8] = {abuf[PAR][0][q2], abuf[PAR][1][q2], abuf[PAR][2][q2], abuf[PAR][3][q2],
                                          abuf[PAR ^ 1][0][q2], abuf[PAR ^ 1][1][q2], abuf[PAR ^ 1][2][q2], abuf[PAR ^ 1][3][q2]};
#pragma unroll
                    for (int k = 0; k < 5; ++k)
#pragma unroll
                        for (int r = 0; r < 4; ++r) acc[r] = fma2(p.w2p[q2][k], A8[r + k], acc[r]);
                }
                float c2[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) c2[r] = acc[r].x + acc[r].y;
                float2 f;
                if constexpr (ARCH == 0) {
                    f = tanh_fold2(make_float2(max3_nan(c2c, c2[0], c2[1]), max3_nan(c2[1], c2[2], c2[3])), make_float2(p.b2s, p.b2s));
                    c2c = c2[3];
                } else {
                    f = tanh_fold2(make_float2(max_nan(c2[0], c2[1]), max_nan(c2[2], c2[3])), make_float2(p.b2s, p.b2s));
                }
                nan_probe = fma2(f, make_float2(0.f, 0.f), nan_probe);
                const uint32_t h = pack_bf16x2(f.x, f.y);
                const float2 r1 = sub2(f, make_float2(__uint_as_float(h << 16), __uint_as_float(h & 0xffff0000u)));
                const uint32_t md = pack_bf16x2(r1.x, r1.y);
                const float2 r2 = sub2(r1, make_float2(__uint_as_float(md << 16), __uint_as_float(md & 0xffff0000u)));
                const uint32_t lo = pack_bf16x2(r2.x, r2.y);
                const uint32_t acol = tlane + col_pieces(t, u) + kk;
                tmem_st1(acol, h);
                tmem_st1(acol + 8, md);
                tmem_st1(acol + 16, lo);
            }
            if constexpr (doA) {
#pragma unroll
                for (int q2 = 0; q2 < 2; ++q2)
#pragma unroll
                    for (int r = 0; r < 4; ++r) abuf[PAR][r][q2] = an[r][q2];
            }
            // ---------------- bottom: arrivals ----------------
            if constexpr (doA) {
                if (n == kSfBlocks - 1) {                  // last read of this smem stage
                    __syncwarp();
                    if (lane == 0) mbar_arrive(BAR(o_bar + SfBars::kEmpty + s));
                    n = 0; ++ti;
                } else {
                    ++n;
                }
            }
            if constexpr (doB) {
                if (kk == 7 || jb == J - 1) {
                    const uint32_t abase = tlane + col_pieces(t, u);
                    for (int z = kk + 1; z < 8; ++z) { tmem_st1(abase + z, 0u); tmem_st1(abase + z + 8, 0u); tmem_st1(abase + z + 16, 0u); }
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(BAR(o_bar + SfBars::kPFull + u));
                }
            }
        };
        using T_ = std::integral_constant<bool, true>;
        using F_ = std::integral_constant<bool, false>;
        using P0 = std::integral_constant<int, 0>;
        using P1 = std::integral_constant<int, 1>;
        iteration(0, T_{}, F_{}, P0{});
        int jj = 1;
#pragma unroll 1
        for (; jj + 1 < J; jj += 2) {
            iteration(jj, T_{}, T_{}, P1{});
            iteration(jj + 1, T_{}, T_{}, P0{});
        }
        if (jj < J) { iteration(jj, T_{}, T_{}, P1{}); ++jj; }
        if (J & 1) iteration(J, F_{}, T_{}, P1{}); else iteration(J, F_{}, T_{}, P0{});

        mbar_wait_parked(BAR(o_bar + SfBars::kGFull), 0);
        tc_fence_after();
        float *dst = p.partial + ((int64_t)blockIdx.y * p.B + b) * kGates;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t G[32];
            tmem_ld32_issue(tlane + col_gates(t) + half * 32, G);
            tmem_ld32_wait(G);
            if (row_ok) {
#pragma unroll
                for (int k = 0; k < 32; k += 4)
                    *reinterpret_cast<uint4 *>(dst + half * 32 + k) = make_uint4(G[k], G[k + 1], G[k + 2], G[k + 3]);
            }
        }
        if (row_ok && (nan_probe.x != nan_probe.x || nan_probe.y != nan_probe.y) && atomicExch(&p.nanflag[b], 1) == 0)
            p.list[atomicAdd(p.count, 1)] = b;
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 3) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
    }
}

}  // namespace b2cnn
