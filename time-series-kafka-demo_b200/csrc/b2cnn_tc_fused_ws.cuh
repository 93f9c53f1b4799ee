// b2cnn_tc_fused_ws.cuh -- the fused kernel with a warp-specialised epilogue (included by
// b2cnn_tc.cu after b2cnn_tc_fused.cuh, whose parameters, W_ih packing and barrier ideas it shares).
//
// tc_fused_kernel runs the whole per-window stream in one thread, two epilogue warps per SM
// sub-partition: measured issue-slot utilisation 49 %, MUFU pipe 54 % -- latency-bound, not
// throughput-bound.  Here the stream is cut at the conv1 activations:
//
//   A-warps (one thread per window):  TMEM accumulators -> tap-9 patch -> pool1 -> tanh (32 MUFU)
//                                     -> 16 activations handed over through 16 TMEM columns
//   B-warps (one thread per window):  activations -> conv2 -> pool2 -> tanh (4 MUFU) -> bf16
//                                     pieces -> TMEM A operand of the projection GEMM
//
// so every sub-partition holds 2 MUFU-bound and 2 FMA-bound warps (plus one control warp) whose
// phases interleave.  640 threads, one CTA per SM:
//   warp 0 producer (window tiles)   warp 1/2 MMA issuers   warp 3 TMEM allocator + W_ih producer
//   warps 4-7 / 8-11   A / B warps of window tile 0        warps 12-15 / 16-19  of window tile 1
// TMEM columns per window tile (256): conv1 ring 3 x 32 | activations 2 x 16 | A pieces 2 x 24 | gates 64.
#pragma once

namespace b2cnn {

constexpr int kWsThreads = 640;
constexpr int kWsRing = 3;

struct WsBars {   // uint64_t slots; per window tile t (stride kPerTile)
    static constexpr int kFull = 0, kEmpty = 2, kTFull = 4, kTEmpty = 7, kAFull = 10, kAEmpty = 12, kPFull = 14, kPEmpty = 16,
                         kGFull = 18, kPerTile = 19;
    static constexpr int kWFull = 2 * kPerTile, kWEmpty = kWFull + 2, kTotal = kWEmpty + 2;
};

template <int C, int SPLITS, int ARCH>
__global__ void __launch_bounds__(kWsThreads, 1)
tc_fused_ws_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TcFusedParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sA = smem;                                        // [2 tiles][2 stages][C][16 KB]
    uint8_t *sBm = sA + 2 * 2 * C * kTcABytes;                 // conv1 band matrices
    uint8_t *sW = sBm + C * SPLITS * kTcBBytes;                // W_ih ring [2][6 KB]
    uint64_t *bars = reinterpret_cast<uint64_t *>(sW + 2 * kFuWChunkBytes);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + WsBars::kTotal);
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int idx) -> uint32_t { return bar0 + 8u * (uint32_t)idx; };

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int b_cta = blockIdx.x * 2 * kTcM;
    const int p0 = blockIdx.y * p.feats_per_cta;
    const int nfeat = min(p.feats_per_cta, p.L - p0);
    constexpr int FOFF = ARCH == 0 ? 3 : 2;                    // step j emits features 2j-FOFF, 2j-FOFF+1
    const int nsteps_needed = (nfeat + FOFF - 1) / 2 + 1;
    const int ntiles = (nsteps_needed + kTcBlocks - 1) / kTcBlocks;
    const int J = ntiles * kTcBlocks;
    const int nchunks = (J + 7) / 8;
    const int T0 = p0 * 4;

    if ((smem_u32(smem) & 1023u) != 0) __trap();
    for (int i = threadIdx.x; i < C * SPLITS * kTcBBytes / 16; i += kWsThreads)
        reinterpret_cast<uint4 *>(sBm)[i] = reinterpret_cast<const uint4 *>(p.bmats)[i];
    if (threadIdx.x == 0) {
        for (int t = 0; t < 2; ++t) {
            const int o = t * WsBars::kPerTile;
            for (int i = 0; i < 2; ++i) { mbar_init(BAR(o + WsBars::kFull + i), 1); mbar_init(BAR(o + WsBars::kEmpty + i), 4); }
            for (int i = 0; i < kWsRing; ++i) { mbar_init(BAR(o + WsBars::kTFull + i), 1); mbar_init(BAR(o + WsBars::kTEmpty + i), 4); }
            for (int i = 0; i < 2; ++i) { mbar_init(BAR(o + WsBars::kAFull + i), 4); mbar_init(BAR(o + WsBars::kAEmpty + i), 4); }
            for (int i = 0; i < 2; ++i) { mbar_init(BAR(o + WsBars::kPFull + i), 4); mbar_init(BAR(o + WsBars::kPEmpty + i), 1); }
            mbar_init(BAR(o + WsBars::kGFull), 1);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(BAR(WsBars::kWFull + i), 1); mbar_init(BAR(WsBars::kWEmpty + i), 2); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 3) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    auto sA_of = [&](int t, int s, int c) -> uint8_t * { return sA + ((size_t)((t * 2 + s) * C + c)) * kTcABytes; };
    // TMEM column map of a window tile
    constexpr uint32_t kColAct = kWsRing * 32, kColPieces = 128, kColGates = 192;

    if (warp == 0) {
        // ===================== producer: window tiles =====================
        if (lane == 0) {
            for (int i = 0; i < ntiles; ++i) {
                const int s = i & 1, ph = (i >> 1) & 1;
                for (int t = 0; t < 2; ++t) {
                    const int o = t * WsBars::kPerTile;
                    mbar_wait_parked(BAR(o + WsBars::kEmpty + s), ph ^ 1);
                    mbar_expect_tx(BAR(o + WsBars::kFull + s), C * kTcABytes);
#pragma unroll
                    for (int c = 0; c < C; ++c)
                        tma_load_3d(smem_u32(sA_of(t, s, c)), &tmap, T0 + kTcAdv * i, c, b_cta + t * kTcM, BAR(o + WsBars::kFull + s));
                }
            }
        }
    } else if (warp == 1 || warp == 2) {
        // ===================== MMA issuer of window tile t =====================
        const int t = warp - 1;
        const int o = t * WsBars::kPerTile;
        const uint64_t a_base = desc_sw128_kmajor(smem_u32(sA_of(t, 0, 0)));
        const uint64_t b_base = desc_none_kmajor(smem_u32(sBm), 128, 256);
        const uint64_t w_base = desc_none_kmajor(smem_u32(sW), 128, 256);
        const uint32_t a_lo0 = (uint32_t)a_base, a_hi = (uint32_t)(a_base >> 32);
        const uint32_t b_lo0 = (uint32_t)b_base, b_hi = (uint32_t)(b_base >> 32);
        const uint32_t w_lo0 = (uint32_t)w_base, w_hi = (uint32_t)(w_base >> 32);
        const uint32_t tcol = tmem_base + t * 256;
        auto issue_proj = [&](int m) {
            const int u = m & 1, ph = (m >> 1) & 1;
            mbar_wait_parked(BAR(WsBars::kWFull + u), ph);
            mbar_wait_parked(BAR(o + WsBars::kPFull + u), ph);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t d = tcol + kColGates;
                const uint32_t a0 = tcol + kColPieces + u * 24;
                const uint32_t w0 = w_lo0 + u * (kFuWChunkBytes >> 4);
                umma_ts(d, a0 + 0, w0 + 0 * 128, w_hi, kIdescProj, m != 0);   // hh hm mh hl lh mm
                umma_ts(d, a0 + 0, w0 + 1 * 128, w_hi, kIdescProj, 1);
                umma_ts(d, a0 + 8, w0 + 0 * 128, w_hi, kIdescProj, 1);
                umma_ts(d, a0 + 0, w0 + 2 * 128, w_hi, kIdescProj, 1);
                umma_ts(d, a0 + 16, w0 + 0 * 128, w_hi, kIdescProj, 1);
                umma_ts(d, a0 + 8, w0 + 1 * 128, w_hi, kIdescProj, 1);
                umma_commit(BAR(o + WsBars::kPEmpty + u));
                umma_commit(BAR(WsBars::kWEmpty + u));
            }
            __syncwarp();
        };
        int m_done = 0, n = 0, i = 0, slot = 0, sph = 0;
        for (int j = 0; j < J; ++j) {
            const int s = i & 1;
            if (j >= 8 + kFuLag && ((j - kFuLag) & 7) == 0) { issue_proj(m_done); ++m_done; }
            if (n == 0) mbar_wait_parked(BAR(o + WsBars::kFull + s), (i >> 1) & 1);
            mbar_wait_parked(BAR(o + WsBars::kTEmpty + slot), sph ^ 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t d = tcol + slot * 32;
                const uint32_t a_s = a_lo0 + (uint32_t)(s * C) * (kTcABytes >> 4) + n;
#pragma unroll
                for (int c = 0; c < C; ++c)
#pragma unroll
                    for (int sp = 0; sp < SPLITS; ++sp)
                        umma_ss(d, a_s + c * (kTcABytes >> 4), a_hi, b_lo0 + (c * SPLITS + sp) * (kTcBBytes >> 4), b_hi, kIdesc,
                                (c | sp) != 0);
                umma_commit(BAR(o + WsBars::kTFull + slot));
            }
            __syncwarp();
            if (++n == kTcBlocks) { n = 0; ++i; }
            if (++slot == kWsRing) { slot = 0; sph ^= 1; }
        }
        for (; m_done < nchunks; ++m_done) issue_proj(m_done);
        if (elect_one()) umma_commit(BAR(o + WsBars::kGFull));
        __syncwarp();
    } else if (warp == 3) {
        // ===================== W_ih chunk producer (and TMEM allocator) =====================
        if (lane == 0) {
            const uint8_t *wsrc = p.wpack + (size_t)blockIdx.y * p.chunks_per_cta * kFuWChunkBytes;
            for (int m = 0; m < nchunks; ++m) {
                const int u = m & 1;
                mbar_wait_parked(BAR(WsBars::kWEmpty + u), ((m >> 1) & 1) ^ 1);
                mbar_expect_tx(BAR(WsBars::kWFull + u), kFuWChunkBytes);
                bulk_load_1d(smem_u32(sW + u * kFuWChunkBytes), wsrc + (size_t)m * kFuWChunkBytes, kFuWChunkBytes,
                             BAR(WsBars::kWFull + u));
            }
        }
    } else {
        const int e = warp - 4;                    // 0..15
        const int t = e >> 3;                      // window tile
        const bool stage_a = ((e >> 2) & 1) == 0;  // warps 4-7 / 12-15: A;  8-11 / 16-19: B
        const int o_bar = t * WsBars::kPerTile;
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + t * 256;
        if (stage_a) {
            // ===================== A-warps: accumulators -> pool1 -> tanh -> activations =====================
            const uint32_t swz = (uint32_t)(row & 7);
            float2 pm6[2], pm7[2];
#pragma unroll
            for (int q2 = 0; q2 < 2; ++q2) { pm6[q2] = make_float2(0.f, 0.f); pm7[q2] = make_float2(0.f, 0.f); }
            uint32_t Dn[32];
            mbar_wait_parked(BAR(o_bar + WsBars::kTFull + 0), 0);
            tc_fence_after();
            tmem_ld32_issue(tlane + 0, Dn);
            int n = 0, ti = 0, slot = 0, sph = 0;
#pragma unroll 1
            for (int j = 0; j < J; ++j) {
                const int s = ti & 1;
                float D[32];
                tmem_ld32_wait(Dn);
#pragma unroll
                for (int k = 0; k < 32; ++k) D[k] = __uint_as_float(Dn[k]);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(o_bar + WsBars::kTEmpty + slot));
                if (++slot == kWsRing) { slot = 0; sph ^= 1; }
                if (j + 1 < J) {                                   // prefetch the next block's accumulators
                    mbar_wait_parked(BAR(o_bar + WsBars::kTFull + slot), sph);
                    tc_fence_after();
                    tmem_ld32_issue(tlane + slot * 32, Dn);
                }
                float2 an[4][2];
                if constexpr (ARCH == 0) {
                    if (n == 0) mbar_wait_parked(BAR(o_bar + WsBars::kFull + s), (ti >> 1) & 1);   // TMA bytes visible for the tap-9 reads
                    const uint8_t *tile = sA_of(t, s, 0) + row * 128;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const uint16_t raw = *reinterpret_cast<const uint16_t *>(tile + c * kTcABytes + ((uint32_t)((n + 1) ^ swz) << 4));
                        const float xv = __uint_as_float((uint32_t)raw << 16);
#pragma unroll
                        for (int q2 = 0; q2 < 2; ++q2) pm7[q2] = fma2(p.w9p[c][q2], make_float2(xv, xv), pm7[q2]);
                    }
#pragma unroll
                    for (int q2 = 0; q2 < 2; ++q2) {
                        const int o0 = 2 * q2, o1 = 2 * q2 + 1;
                        an[0][q2] = sig_fold2(make_float2(max3_nan(pm6[q2].x, pm7[q2].x, D[0 * 4 + o0]),
                                                           max3_nan(pm6[q2].y, pm7[q2].y, D[0 * 4 + o1])), p.b1sp[q2]);
                        an[1][q2] = sig_fold2(make_float2(max3_nan(D[0 * 4 + o0], D[1 * 4 + o0], D[2 * 4 + o0]),
                                                           max3_nan(D[0 * 4 + o1], D[1 * 4 + o1], D[2 * 4 + o1])), p.b1sp[q2]);
                        an[2][q2] = sig_fold2(make_float2(max3_nan(D[2 * 4 + o0], D[3 * 4 + o0], D[4 * 4 + o0]),
                                                           max3_nan(D[2 * 4 + o1], D[3 * 4 + o1], D[4 * 4 + o1])), p.b1sp[q2]);
                        an[3][q2] = sig_fold2(make_float2(max3_nan(D[4 * 4 + o0], D[5 * 4 + o0], D[6 * 4 + o0]),
                                                           max3_nan(D[4 * 4 + o1], D[5 * 4 + o1], D[6 * 4 + o1])), p.b1sp[q2]);
                        pm6[q2] = make_float2(D[6 * 4 + o0], D[6 * 4 + o1]);
                        pm7[q2] = make_float2(D[7 * 4 + o0], D[7 * 4 + o1]);
                    }
                } else {
#pragma unroll
                    for (int q2 = 0; q2 < 2; ++q2) {
                        const int o0 = 2 * q2, o1 = 2 * q2 + 1;
#pragma unroll
                        for (int i2 = 0; i2 < 4; ++i2)
                            an[i2][q2] = sig_fold2(make_float2(max_nan(D[(2 * i2) * 4 + o0], D[(2 * i2 + 1) * 4 + o0]),
                                                                max_nan(D[(2 * i2) * 4 + o1], D[(2 * i2 + 1) * 4 + o1])), p.b1sp[q2]);
                    }
                }
                if (++n == kTcBlocks) {                            // last read of this smem stage
                    __syncwarp();
                    if (lane == 0) mbar_arrive(BAR(o_bar + WsBars::kEmpty + s));
                    n = 0; ++ti;
                }
                // ---- hand the 16 activations of this step to the B-warp thread of the same window
                const int buf = j & 1;
                mbar_wait_parked(BAR(o_bar + WsBars::kAEmpty + buf), ((j >> 1) & 1) ^ 1);
                tc_fence_after();
                uint32_t av[16];
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int q2 = 0; q2 < 2; ++q2) {
                        av[r * 4 + q2 * 2 + 0] = __float_as_uint(an[r][q2].x);
                        av[r * 4 + q2 * 2 + 1] = __float_as_uint(an[r][q2].y);
                    }
                tmem_st16(tlane + kColAct + buf * 16, av);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(o_bar + WsBars::kAFull + buf));
            }
        } else {
            // ===================== B-warps: activations -> conv2 -> pool2 -> tanh -> projection operand =====================
            const int b = b_cta + t * kTcM + row;
            const bool row_ok = b < p.B;
            float2 ah[4][2], nan_probe = make_float2(0.f, 0.f);
            float c2c = 0.f;
#pragma unroll
            for (int r = 0; r < 4; ++r) { ah[r][0] = make_float2(0.f, 0.f); ah[r][1] = make_float2(0.f, 0.f); }
#pragma unroll 1
            for (int j = 0; j < J; ++j) {
                const int buf = j & 1;
                const int m = j >> 3, kk = j & 7, u = m & 1;
                mbar_wait_parked(BAR(o_bar + WsBars::kAFull + buf), (j >> 1) & 1);
                tc_fence_after();
                uint32_t av[16];
                tmem_ld16(tlane + kColAct + buf * 16, av);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(o_bar + WsBars::kAEmpty + buf));
                if (kk == 0) {                                     // first store of chunk m into piece buffer u
                    mbar_wait_parked(BAR(o_bar + WsBars::kPEmpty + u), ((m >> 1) & 1) ^ 1);
                    tc_fence_after();
                }
                float2 an[4][2];
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int q2 = 0; q2 < 2; ++q2)
                        an[r][q2] = make_float2(__uint_as_float(av[r * 4 + q2 * 2]), __uint_as_float(av[r * 4 + q2 * 2 + 1]));
                float2 acc[4][2];
#pragma unroll
                for (int r = 0; r < 4; ++r) { acc[r][0] = make_float2(0.f, 0.f); acc[r][1] = make_float2(0.f, 0.f); }
#pragma unroll
                for (int q2 = 0; q2 < 2; ++q2) {
                    const float2 A8[8] = {ah[0][q2], ah[1][q2], ah[2][q2], ah[3][q2], an[0][q2], an[1][q2], an[2][q2], an[3][q2]};
#pragma unroll
                    for (int k = 0; k < 5; ++k)
#pragma unroll
                        for (int r = 0; r < 4; ++r) acc[r][q2] = fma2(p.w2p[q2][k], A8[r + k], acc[r][q2]);
                }
                float c2[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float2 sacc = add2(acc[r][0], acc[r][1]);
                    c2[r] = sacc.x + sacc.y;
                }
                float2 f;                                          // features 2j-FOFF, 2j-FOFF+1
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
                const uint32_t acol = tlane + kColPieces + u * 24 + kk;
                tmem_st1(acol, h);
                tmem_st1(acol + 8, md);
                tmem_st1(acol + 16, lo);
#pragma unroll
                for (int r = 0; r < 4; ++r) { ah[r][0] = an[r][0]; ah[r][1] = an[r][1]; }
                if (kk == 7 || j == J - 1) {
                    const uint32_t abase = tlane + kColPieces + u * 24;
                    for (int z = kk + 1; z < 8; ++z) { tmem_st1(abase + z, 0u); tmem_st1(abase + z + 8, 0u); tmem_st1(abase + z + 16, 0u); }
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(BAR(o_bar + WsBars::kPFull + u));
                }
            }
            // ---- gate pre-activations of this CTA's position range -> partial[range][window][64]
            mbar_wait_parked(BAR(o_bar + WsBars::kGFull), 0);
            tc_fence_after();
            float *dst = p.partial + ((int64_t)blockIdx.y * p.B + b) * kGates;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t G[32];
                tmem_ld32_issue(tlane + kColGates + half * 32, G);
                tmem_ld32_wait(G);
                if (row_ok) {
#pragma unroll
                    for (int k = 0; k < 32; k += 4)
                        *reinterpret_cast<uint4 *>(dst + half * 32 + k) = make_uint4(G[k], G[k + 1], G[k + 2], G[k + 3]);
                }
            }
            if (row_ok && (nan_probe.x != nan_probe.x || nan_probe.y != nan_probe.y)) p.nanflag[b] = 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 3) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

}  // namespace b2cnn
