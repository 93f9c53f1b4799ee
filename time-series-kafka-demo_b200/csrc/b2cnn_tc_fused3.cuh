// b2cnn_tc_fused3.cuh -- the fused kernel with THREE window tiles per SM (included by b2cnn_tc.cu
// after b2cnn_tc_fused.cuh; same arithmetic, parameters and W_ih packing).
//
// tc_fused_kernel holds two window tiles per SM, i.e. two epilogue warps per scheduler, and is
// latency-bound (issue slots ~50 %, MUFU pipe ~60 %).  Shared memory is what limits the tile
// count (a 128-window x 64-sample x 3-channel stage is 48 KB), so this variant gives every
// window tile ONE smem stage, a 2-slot accumulator ring and one projection-piece buffer, and
// relies on the three tiles being de-phased in time to hide each other's refill bubbles:
//   544 threads: warp 0 producer | warps 1-3 MMA issuers | warp 4 TMEM allocator + W_ih producer
//                warps 5-16 epilogue (3 per scheduler; TMEM lane quadrant = warp % 4)
//   smem  : 3 x 48 KB window tiles + 9 KB band matrices + 2 x 6 KB W_ih chunks
//   TMEM  : gates 3 x 64 | conv1 rings 3 x (2 x 32) | projection pieces 3 x 24
#pragma once

namespace b2cnn {

constexpr int kF3Threads = 544;
constexpr int kF3Tiles = 3;

struct F3Bars {   // uint64_t slots; per window tile t (stride kPerTile)
    static constexpr int kFull = 0, kEmpty = 1, kTFull = 2, kTEmpty = 4, kPFull = 6, kPEmpty = 7, kGFull = 8, kPerTile = 9;
    static constexpr int kWFull = kF3Tiles * kPerTile, kWEmpty = kWFull + 2, kTotal = kWEmpty + 2;
};

template <int C, int SPLITS, int ARCH>
__global__ void __launch_bounds__(kF3Threads, 1)
tc_fused3_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TcFusedParams p, const uint32_t stagger_ns) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sA = smem;                                        // [3 tiles][C][16 KB], one stage each
    uint8_t *sBm = sA + kF3Tiles * C * kTcABytes;
    uint8_t *sW = sBm + C * SPLITS * kTcBBytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sW + 2 * kFuWChunkBytes);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + F3Bars::kTotal);
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int idx) -> uint32_t { return bar0 + 8u * (uint32_t)idx; };

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int b_cta = blockIdx.x * kF3Tiles * kTcM;
    const int p0 = blockIdx.y * p.feats_per_cta;
    const int nfeat = min(p.feats_per_cta, p.L - p0);
    constexpr int FOFF = ARCH == 0 ? 3 : 2;
    const int nsteps_needed = (nfeat + FOFF - 1) / 2 + 1;
    const int ntiles = (nsteps_needed + kTcBlocks - 1) / kTcBlocks;
    const int J = ntiles * kTcBlocks;
    const int nchunks = (J + 7) / 8;
    const int T0 = p0 * 4;

    if ((smem_u32(smem) & 1023u) != 0) __trap();
    for (int i = threadIdx.x; i < C * SPLITS * kTcBBytes / 16; i += kF3Threads)
        reinterpret_cast<uint4 *>(sBm)[i] = reinterpret_cast<const uint4 *>(p.bmats)[i];
    if (threadIdx.x == 0) {
        for (int t = 0; t < kF3Tiles; ++t) {
            const int o = t * F3Bars::kPerTile;
            mbar_init(BAR(o + F3Bars::kFull), 1);
            mbar_init(BAR(o + F3Bars::kEmpty), 4);
            for (int i = 0; i < 2; ++i) { mbar_init(BAR(o + F3Bars::kTFull + i), 1); mbar_init(BAR(o + F3Bars::kTEmpty + i), 4); }
            mbar_init(BAR(o + F3Bars::kPFull), 4);
            mbar_init(BAR(o + F3Bars::kPEmpty), 1);
            mbar_init(BAR(o + F3Bars::kGFull), 1);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(BAR(F3Bars::kWFull + i), 1); mbar_init(BAR(F3Bars::kWEmpty + i), kF3Tiles); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    auto sA_of = [&](int t, int c) -> uint8_t * { return sA + ((size_t)(t * C + c)) * kTcABytes; };
    // TMEM column map (all N=64 accumulators on multiples of 64, N=32 on multiples of 32)
    auto col_gates = [&](int t) -> uint32_t { return 64u * t; };
    auto col_ring = [&](int t, int slot) -> uint32_t { return 192u + 64u * t + 32u * slot; };
    auto col_pieces = [&](int t) -> uint32_t { return 384u + 24u * t; };

    if (warp == 0) {
        // ===================== producer: one smem stage per window tile =====================
        if (lane == 0) {
            for (int i = 0; i < ntiles; ++i) {
                for (int t = 0; t < kF3Tiles; ++t) {
                    const int o = t * F3Bars::kPerTile;
                    if (i == 0 && t > 0 && stagger_ns) __nanosleep(stagger_ns);     // de-phase the tiles once
                    mbar_wait_parked(BAR(o + F3Bars::kEmpty), (i & 1) ^ 1);
                    mbar_expect_tx(BAR(o + F3Bars::kFull), C * kTcABytes);
#pragma unroll
                    for (int c = 0; c < C; ++c)
                        tma_load_3d(smem_u32(sA_of(t, c)), &tmap, T0 + kTcAdv * i, c, b_cta + t * kTcM, BAR(o + F3Bars::kFull));
                }
            }
        }
    } else if (warp >= 1 && warp <= 3) {
        // ===================== MMA issuer of window tile t =====================
        const int t = warp - 1;
        const int o = t * F3Bars::kPerTile;
        const uint64_t a_base = desc_sw128_kmajor(smem_u32(sA_of(t, 0)));
        const uint64_t b_base = desc_none_kmajor(smem_u32(sBm), 128, 256);
        const uint64_t w_base = desc_none_kmajor(smem_u32(sW), 128, 256);
        const uint32_t a_lo0 = (uint32_t)a_base, a_hi = (uint32_t)(a_base >> 32);
        const uint32_t b_lo0 = (uint32_t)b_base, b_hi = (uint32_t)(b_base >> 32);
        const uint32_t w_lo0 = (uint32_t)w_base, w_hi = (uint32_t)(w_base >> 32);
        auto issue_proj = [&](int m) {
            const int u = m & 1;
            mbar_wait_parked(BAR(F3Bars::kWFull + u), (m >> 1) & 1);
            mbar_wait_parked(BAR(o + F3Bars::kPFull), m & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t d = tmem_base + col_gates(t);
                const uint32_t a0 = tmem_base + col_pieces(t);
                const uint32_t w0 = w_lo0 + u * (kFuWChunkBytes >> 4);
                umma_ts(d, a0 + 0, w0 + 0 * 128, w_hi, kIdescProj, m != 0);   // hh hm mh hl lh mm
                umma_ts(d, a0 + 0, w0 + 1 * 128, w_hi, kIdescProj, 1);
                umma_ts(d, a0 + 8, w0 + 0 * 128, w_hi, kIdescProj, 1);
                umma_ts(d, a0 + 0, w0 + 2 * 128, w_hi, kIdescProj, 1);
                umma_ts(d, a0 + 16, w0 + 0 * 128, w_hi, kIdescProj, 1);
                umma_ts(d, a0 + 8, w0 + 1 * 128, w_hi, kIdescProj, 1);
                umma_commit(BAR(o + F3Bars::kPEmpty));
                umma_commit(BAR(F3Bars::kWEmpty + u));
            }
            __syncwarp();
        };
        int m_done = 0, n = 0, i = 0;
        for (int j = 0; j < J; ++j) {
            const int slot = j & 1;
            if (n == 0) mbar_wait_parked(BAR(o + F3Bars::kFull), i & 1);
            mbar_wait_parked(BAR(o + F3Bars::kTEmpty + slot), ((j >> 1) & 1) ^ 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t d = tmem_base + col_ring(t, slot);
                const uint32_t a_s = a_lo0 + n;
#pragma unroll
                for (int c = 0; c < C; ++c)
#pragma unroll
                    for (int sp = 0; sp < SPLITS; ++sp)
                        umma_ss(d, a_s + c * (kTcABytes >> 4), a_hi, b_lo0 + (c * SPLITS + sp) * (kTcBBytes >> 4), b_hi, kIdesc,
                                (c | sp) != 0);
                umma_commit(BAR(o + F3Bars::kTFull + slot));
            }
            __syncwarp();
            if (++n == kTcBlocks) { n = 0; ++i; }
            // Project chunk m right after conv1 block 8m+8 was issued: the epilogue iteration that
            // finishes the chunk (stage B of step 8m+7) is the one that first needs block 8m+8, so
            // the projection must not be waited for BEFORE that block is issued (single piece buffer).
            if (j >= 8 && (j & 7) == 0) { issue_proj(m_done); ++m_done; }
        }
        for (; m_done < nchunks; ++m_done) issue_proj(m_done);
        if (elect_one()) umma_commit(BAR(o + F3Bars::kGFull));
        __syncwarp();
    } else if (warp == 4) {
        // ===================== W_ih chunk producer (and TMEM allocator) =====================
        if (lane == 0) {
            const uint8_t *wsrc = p.wpack + (size_t)blockIdx.y * p.chunks_per_cta * kFuWChunkBytes;
            for (int m = 0; m < nchunks; ++m) {
                const int u = m & 1;
                mbar_wait_parked(BAR(F3Bars::kWEmpty + u), ((m >> 1) & 1) ^ 1);
                mbar_expect_tx(BAR(F3Bars::kWFull + u), kFuWChunkBytes);
                bulk_load_1d(smem_u32(sW + u * kFuWChunkBytes), wsrc + (size_t)m * kFuWChunkBytes, kFuWChunkBytes,
                             BAR(F3Bars::kWFull + u));
            }
        }
    } else {
        // ===================== epilogue: thread == window (stage A of block jj + stage B of step jj-1) =====================
        const int t = (warp - 5) >> 2;
        const int o_bar = t * F3Bars::kPerTile;
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int b = b_cta + t * kTcM + row;
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t swz = (uint32_t)(row & 7);
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
            const int jb = jj - 1, m = jb >> 3, kk = jb & 7;
            uint32_t Dr[32];
            auto Dv = [&](int idx) -> float { return __uint_as_float(Dr[idx]); };
            // ---------------- top: barriers and the accumulator load ----------------
            if constexpr (doA) {
                mbar_wait_parked(BAR(o_bar + F3Bars::kTFull + (jj & 1)), (jj >> 1) & 1);
                tc_fence_after();
                tmem_ld32_issue(tlane + col_ring(t, jj & 1), Dr);
                tmem_ld32_wait(Dr);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(o_bar + F3Bars::kTEmpty + (jj & 1)));
                if (ARCH == 0 && n == 0) mbar_wait_parked(BAR(o_bar + F3Bars::kFull), ti & 1);   // TMA bytes visible for the tap-9 reads
            }
            if constexpr (doB) {
                if (kk == 0) {                              // the single piece buffer must have been projected
                    mbar_wait_parked(BAR(o_bar + F3Bars::kPEmpty), (m & 1) ^ 1);
                    tc_fence_after();
                }
            }
            // ---------------- middle: straight-line math ----------------
            float2 an[4][2];
            if constexpr (doA) {
                if constexpr (ARCH == 0) {
                    const uint8_t *tile = sA_of(t, 0) + row * 128;
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
                        an[0][q2] = sig_fold2(make_float2(max3_nan(pm6[q2].x, pm7[q2].x, Dv(0 * 4 + o0)),
                                                          max3_nan(pm6[q2].y, pm7[q2].y, Dv(0 * 4 + o1))), p.b1sp[q2]);
                        an[1][q2] = sig_fold2(make_float2(max3_nan(Dv(0 * 4 + o0), Dv(1 * 4 + o0), Dv(2 * 4 + o0)),
                                                          max3_nan(Dv(0 * 4 + o1), Dv(1 * 4 + o1), Dv(2 * 4 + o1))), p.b1sp[q2]);
                        an[2][q2] = sig_fold2(make_float2(max3_nan(Dv(2 * 4 + o0), Dv(3 * 4 + o0), Dv(4 * 4 + o0)),
                                                          max3_nan(Dv(2 * 4 + o1), Dv(3 * 4 + o1), Dv(4 * 4 + o1))), p.b1sp[q2]);
                        an[3][q2] = sig_fold2(make_float2(max3_nan(Dv(4 * 4 + o0), Dv(5 * 4 + o0), Dv(6 * 4 + o0)),
                                                          max3_nan(Dv(4 * 4 + o1), Dv(5 * 4 + o1), Dv(6 * 4 + o1))), p.b1sp[q2]);
                        pm6[q2] = make_float2(Dv(6 * 4 + o0), Dv(6 * 4 + o1));
                        pm7[q2] = make_float2(Dv(7 * 4 + o0), Dv(7 * 4 + o1));
                    }
                } else {
#pragma unroll
                    for (int q2 = 0; q2 < 2; ++q2) {
                        const int o0 = 2 * q2, o1 = 2 * q2 + 1;
#pragma unroll
                        for (int i2 = 0; i2 < 4; ++i2)
                            an[i2][q2] = sig_fold2(make_float2(max_nan(Dv((2 * i2) * 4 + o0), Dv((2 * i2 + 1) * 4 + o0)),
                                                               max_nan(Dv((2 * i2) * 4 + o1), Dv((2 * i2 + 1) * 4 + o1))), p.b1sp[q2]);
                    }
                }
            }
            if constexpr (doB) {
                float2 acc[4][2];
#pragma unroll
                for (int r = 0; r < 4; ++r) { acc[r][0] = make_float2(0.f, 0.f); acc[r][1] = make_float2(0.f, 0.f); }
#pragma unroll
                for (int q2 = 0; q2 < 2; ++q2) {
                    const float2 A8[8] = {abuf[PAR][0][q2], abuf[PAR][1][q2], abuf[PAR][2][q2], abuf[PAR][3][q2],
                                          abuf[PAR ^ 1][0][q2], abuf[PAR ^ 1][1][q2], abuf[PAR ^ 1][2][q2], abuf[PAR ^ 1][3][q2]};
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
                const uint32_t acol = tlane + col_pieces(t) + kk;
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
                if (n == kTcBlocks - 1) {                  // last read of this window tile's smem stage
                    __syncwarp();
                    if (lane == 0) mbar_arrive(BAR(o_bar + F3Bars::kEmpty));
                    n = 0; ++ti;
                } else {
                    ++n;
                }
            }
            if constexpr (doB) {
                if (kk == 7 || jb == J - 1) {
                    const uint32_t abase = tlane + col_pieces(t);
                    for (int z = kk + 1; z < 8; ++z) { tmem_st1(abase + z, 0u); tmem_st1(abase + z + 8, 0u); tmem_st1(abase + z + 16, 0u); }
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(BAR(o_bar + F3Bars::kPFull));
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

        mbar_wait_parked(BAR(o_bar + F3Bars::kGFull), 0);
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
        if (row_ok && (nan_probe.x != nan_probe.x || nan_probe.y != nan_probe.y)) p.nanflag[b] = 1;
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

}  // namespace b2cnn
