// CTA-pair GEMM kernel (tcgen05 cta_group::2).  Included by gemm_tc.cu inside namespace ecgvit::{anonymous} after the
// tile constants (BM, BK, UMMA_K, kNumThreads, kNumEpilogueWarps, A_STAGE_BYTES) are defined.
//
// A cluster of two CTAs (two SMs of one TPC) owns a 256 x BN output tile.  Each CTA stages only ITS 128 rows of A and
// ITS BN/2 rows of B per k block (32 KB instead of 48 KB at BN = 256); the leader CTA issues one 256 x BN x 16 MMA per
// UMMA_K that reads both halves, and each CTA's TMEM receives its own 128 rows of the accumulator.  This halves the
// B traffic through L2 and through each SM's shared memory (every staged byte is written once by TMA and read once by
// the tensor core, so staging bandwidth is what a single-CTA tile runs out of first).
//
// Synchronisation (all mbarriers live at the same smem offset in both CTAs):
//   full[s]        on the LEADER : leader's arrive.expect_tx(2 x stage bytes); both CTAs' TMA loads complete_tx on it
//   empty[s]       on BOTH       : leader's tcgen05.commit multicast -> each CTA's producer may refill stage s
//   tmem_full[a]   on BOTH       : leader's tcgen05.commit multicast -> each CTA's epilogue drains its 128 rows
//   tmem_empty[a]  on the LEADER : 8 epilogue warps x 2 CTAs arrive (remote arrive from the peer)
#pragma once

template <int BN> struct PairCfg {
    static constexpr int B_HALF = BN / 2;
    static constexpr int B_STAGE_BYTES = B_HALF * BK * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;  // per CTA
    // epilogue staging: every epilogue warp owns two 32-row x 128-byte (64 bf16) swizzled tiles for TMA stores
    static constexpr int EPI_TILE_BYTES = 32 * 128;
    static constexpr int EPI_BYTES = kNumEpilogueWarps * 2 * EPI_TILE_BYTES;  // 64 KB
    static constexpr int PIPE_BUDGET = 232448 - 1024 - 256 - EPI_BYTES;
    static constexpr int STAGES = (PIPE_BUDGET / STAGE_BYTES) > 8 ? 8 : (PIPE_BUDGET / STAGE_BYTES);
    static constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 + 256;
};

// Epilogue math for 64 consecutive accumulator columns of one row (thread = row), written as bf16 into the warp's
// swizzled staging tile(s); the caller then issues one TMA store per tile (TMA clips rows >= M / cols >= N).
template <int MODE>
__device__ __forceinline__ void epilogue_chunk64(const EpiParams &ep, int64_t row, bool row_ok, int col_base, int N,
                                                 const uint32_t r[64], uint8_t *stage0, uint8_t *stage1, int lane) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int col = col_base + 8 * j;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[8 * j + i]);
        const bool ok = row_ok && col < N;  // N % 8 == 0, so a group of 8 columns is entirely in or out
        if (MODE != ECGVIT_EPI_DGELU && ep.bias != nullptr && col < N) {
            const float4 b0 = __ldg(reinterpret_cast<const float4 *>(ep.bias + col));
            const float4 b1 = __ldg(reinterpret_cast<const float4 *>(ep.bias + col + 4));
            v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
            v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
        }
        const uint32_t soff = static_cast<uint32_t>(lane) * 128u + (static_cast<uint32_t>(j ^ (lane & 7)) << 4);
        if (MODE == ECGVIT_EPI_BIAS_RES) {
            if (ok) {
                float a[8];
                load8(reinterpret_cast<const bf16 *>(ep.aux) + row * ep.ldo + col, a);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] += a[i];
            }
        } else if (MODE == ECGVIT_EPI_DGELU) {
            if (ok) {
                float u[8];
                load8(reinterpret_cast<const bf16 *>(ep.aux) + row * ep.ldo + col, u);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] *= gelu_grad<false>(u[i]);
            }
        }
        uint4 packed;
        packed.x = pack_bf16x2(v[0], v[1]); packed.y = pack_bf16x2(v[2], v[3]);
        packed.z = pack_bf16x2(v[4], v[5]); packed.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4 *>(stage0 + soff) = packed;
        if (MODE == ECGVIT_EPI_BIAS_GELU) {
            // gelu of the ROUNDED pre-activation: that is the value backward differentiates
            float lo, hi, h[8];
            unpack_bf16x2(packed.x, lo, hi); h[0] = gelu_fwd<false>(lo); h[1] = gelu_fwd<false>(hi);
            unpack_bf16x2(packed.y, lo, hi); h[2] = gelu_fwd<false>(lo); h[3] = gelu_fwd<false>(hi);
            unpack_bf16x2(packed.z, lo, hi); h[4] = gelu_fwd<false>(lo); h[5] = gelu_fwd<false>(hi);
            unpack_bf16x2(packed.w, lo, hi); h[6] = gelu_fwd<false>(lo); h[7] = gelu_fwd<false>(hi);
            uint4 ph;
            ph.x = pack_bf16x2(h[0], h[1]); ph.y = pack_bf16x2(h[2], h[3]);
            ph.z = pack_bf16x2(h[4], h[5]); ph.w = pack_bf16x2(h[6], h[7]);
            *reinterpret_cast<uint4 *>(stage1 + soff) = ph;
        }
    }
}

template <int BN, bool A_MN, bool B_MN, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kNumThreads, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_out2, int M, int N,
                int K, int split_k, EpiParams ep) {
    using Cfg = PairCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int BM2 = 2 * BM;  // rows of the pair tile

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *epi_smem = smem + STAGES * Cfg::STAGE_BYTES;  // 1024-byte aligned (stage sizes are multiples of 8 KB)
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(epi_smem + Cfg::EPI_BYTES);
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *tmem_full_bar = empty_bar + STAGES;
    uint64_t *tmem_empty_bar = tmem_full_bar + 2;
    uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(tmem_empty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmap_a);
        ptx::prefetch_tensormap(&tmap_b);
        if (MODE != ECGVIT_EPI_ATOMIC_F32) ptx::prefetch_tensormap(&tmap_out);
        if (MODE == ECGVIT_EPI_BIAS_GELU) ptx::prefetch_tensormap(&tmap_out2);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(&tmem_full_bar[a], 1);
            ptx::mbar_init(&tmem_empty_bar[a], 2 * kNumEpilogueWarps);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 2) ptx::tmem_alloc_2sm(tmem_ptr_smem, Cfg::TMEM_COLS);
    ptx::tcgen05_fence_before();
    ptx::cluster_sync_all();
    ptx::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int tiles_m = (M + BM2 - 1) / BM2;
    const int tiles_n = (N + BN - 1) / BN;
    const int kb_total = (K + BK - 1) / BK;
    const int kb_per_split = (kb_total + split_k - 1) / split_k;
    const int num_units = tiles_m * tiles_n * split_k;

    if (warp == 0) {
        // ================================ TMA producer (both CTAs) ================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int u = cluster_id; u < num_units; u += num_clusters) {
                const int tile_n = u % tiles_n;
                const int tile_m = (u / tiles_n) % tiles_m;
                const int split = u / (tiles_n * tiles_m);
                const int kb0 = split * kb_per_split;
                const int kb1 = min(kb0 + kb_per_split, kb_total);
                const int row0 = tile_m * BM2 + rank * BM;          // this CTA's rows of A
                const int col0 = tile_n * BN + rank * Cfg::B_HALF;  // this CTA's rows of B
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t *sa = smem + stage * Cfg::STAGE_BYTES;
                    uint8_t *sb = sa + A_STAGE_BYTES;
                    if (leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
                    if (!A_MN) {
                        ptx::tma_load_2d_2sm(sa, &tmap_a, &full_bar[stage], kb * BK, row0);
                    } else {
#pragma unroll
                        for (int j = 0; j < BM / 64; ++j)
                            ptx::tma_load_2d_2sm(sa + j * 8192, &tmap_a, &full_bar[stage], row0 + j * 64, kb * BK);
                    }
                    if (!B_MN) {
                        ptx::tma_load_2d_2sm(sb, &tmap_b, &full_bar[stage], kb * BK, col0);
                    } else {
#pragma unroll
                        for (int j = 0; j < Cfg::B_HALF / 64; ++j)
                            ptx::tma_load_2d_2sm(sb + j * 8192, &tmap_b, &full_bar[stage], col0 + j * 64, kb * BK);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer (leader CTA only) ============================
        if (leader && lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(BM2, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int u = cluster_id; u < num_units; u += num_clusters, ++it) {
                const int split = u / (tiles_n * tiles_m);
                const int kb0 = split * kb_per_split;
                const int kb1 = min(kb0 + kb_per_split, kb_total);
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
                ptx::tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tcgen05_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t da = A_MN ? ptx::make_smem_desc(sa + k * 2048, 8192, 1024)
                                                 : ptx::make_smem_desc(sa + k * 32, 16, 1024);
                        const uint64_t db = B_MN ? ptx::make_smem_desc(sb + k * 2048, 8192, 1024)
                                                 : ptx::make_smem_desc(sb + k * 32, 16, 1024);
                        ptx::umma_bf16_2sm(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    ptx::umma_commit_2sm(&empty_bar[stage], 0b11);  // frees the stage in both CTAs
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit_2sm(&tmem_full_bar[acc], 0b11);  // both CTAs' epilogues may drain their half
            }
        }
    } else if (warp >= 4) {
        // ================================ epilogue (both CTAs, own 128 rows) ======================
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        constexpr int COLS_PER_WARP = BN / 2;
        uint8_t *stage0 = epi_smem + (warp - 4) * 2 * Cfg::EPI_TILE_BYTES;
        uint8_t *stage1 = stage0 + Cfg::EPI_TILE_BYTES;
        int it = 0;
        int buf = 0;
        for (int u = cluster_id; u < num_units; u += num_clusters, ++it) {
            const int tile_n = u % tiles_n;
            const int tile_m = (u / tiles_n) % tiles_m;
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
            ptx::tcgen05_fence_after();
            const int row_base = tile_m * BM2 + rank * BM + q * 32;
            const int64_t row = static_cast<int64_t>(row_base) + lane;
            if (MODE == ECGVIT_EPI_ATOMIC_F32) {
#pragma unroll 1
                for (int c = 0; c < COLS_PER_WARP / 32; ++c) {
                    const int col0 = half * COLS_PER_WARP + c * 32;
                    uint32_t r[32];
                    ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + col0, r);
                    ptx::tmem_ld_wait();
                    if (row < M) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int col = tile_n * BN + col0 + j * 8;
                            if (col < N) {
                                float v[8];
#pragma unroll
                                for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[j * 8 + i]);
                                epilogue_store<MODE, bf16, 8, false>(ep, row, col, v);
                            }
                        }
                    }
                }
            } else {
#pragma unroll 1
                for (int c = 0; c < COLS_PER_WARP / 64; ++c) {
                    const int col0 = half * COLS_PER_WARP + c * 64;
                    const int col_base = tile_n * BN + col0;
                    if (col_base >= N) break;  // warp-uniform: the whole 64-column chunk is outside the matrix
                    uint32_t r[64];
                    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + col0;
                    ptx::tmem_ld_32x32(taddr, r);
                    ptx::tmem_ld_32x32(taddr + 32, r + 32);
                    ptx::tmem_ld_wait();
                    uint8_t *s0, *s1;
                    if (MODE == ECGVIT_EPI_BIAS_GELU) {
                        s0 = stage0; s1 = stage1;
                        if (lane == 0) ptx::tma_store_wait_read<0>();  // previous chunk's two stores have read smem
                    } else {
                        s0 = buf ? stage1 : stage0; s1 = nullptr;
                        if (lane == 0) ptx::tma_store_wait_read<1>();  // the store that used this buffer two chunks ago
                        buf ^= 1;
                    }
                    __syncwarp();
                    epilogue_chunk64<MODE>(ep, row, row < M, col_base, N, r, s0, s1, lane);
                    ptx::fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the TMA (async proxy)
                    __syncwarp();
                    if (lane == 0) {
                        ptx::tma_store_2d(&tmap_out, s0, col_base, row_base);
                        if (MODE == ECGVIT_EPI_BIAS_GELU) ptx::tma_store_2d(&tmap_out2, s1, col_base, row_base);
                        ptx::tma_store_commit();
                    }
                }
            }
            ptx::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(&tmem_empty_bar[acc], 0);  // the leader's barrier
        }
        if (MODE != ECGVIT_EPI_ATOMIC_F32 && lane == 0) ptx::tma_store_wait_all<0>();
    }

    ptx::tcgen05_fence_before();
    ptx::cluster_sync_all();  // nobody exits (or frees TMEM) while the peer can still touch this CTA's smem / barriers
    if (warp == 2) {
        ptx::tcgen05_fence_after();
        ptx::tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
    }
}
