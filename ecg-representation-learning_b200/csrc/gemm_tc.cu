// tcgen05 / TMEM / TMA GEMM for sm_100a:  C[m,n] = sum_k A(m,k) * B(n,k), bf16 operands, fp32 accumulate.
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0      TMA producer   (cp.async.bulk.tensor 2-D, SWIZZLE_128B, mbarrier complete_tx)
//   warp 1      MMA issuer     (one elected lane issues tcgen05.mma.cta_group::1.kind::f16, 128 x BN x 16)
//   warp 2      TMEM allocator (2 accumulator buffers of BN fp32 columns -> epilogue overlaps next mainloop)
//   warps 4-11  epilogue       (tcgen05.ld 32x32b -> fused bias / residual / GELU / GELU' / fp32 red.add -> HBM)
// Both operands may be K-major (row-major [rows, K]) or MN-major (row-major [K, rows]); the latter is what the
// autograd dgrad (B = W[N,K] read as [K-out, n]) and wgrad (A = dY^T, B = X^T) contractions need, so no
// transposed copies of activations or weights are ever materialised.
//
// Replaces: nn.Linear forward/backward library GEMMs inside vit_pytorch (reference call site
// /root/reference/ecg_transformer/models/ecg_vit.py:141 and autograd at models/train.py:280).
#include "common.cuh"
#include "ptx.cuh"

#include <cudaTypedefs.h>
#include <stdlib.h>

namespace ecgvit {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kNumThreads = 384;
constexpr int kNumEpilogueWarps = 8;
constexpr int A_STAGE_BYTES = BM * BK * 2;

template <int BN> struct TileCfg {
    static constexpr int B_STAGE_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int STAGES = (BN == 256) ? 4 : 6;
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN, bool A_MN, bool B_MN, int MODE>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, int M, int N,
               int K, int split_k, EpiParams ep) {
    using Cfg = TileCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;

    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *tmem_full_bar = empty_bar + STAGES;
    uint64_t *tmem_empty_bar = tmem_full_bar + 2;
    uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(tmem_empty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmap_a);
        ptx::prefetch_tensormap(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(&tmem_full_bar[a], 1);
            ptx::mbar_init(&tmem_empty_bar[a], kNumEpilogueWarps);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 2) ptx::tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
    ptx::tcgen05_fence_before();
    __syncthreads();
    ptx::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int tiles_m = (M + BM - 1) / BM;
    const int tiles_n = (N + BN - 1) / BN;
    const int kb_total = (K + BK - 1) / BK;
    const int kb_per_split = (kb_total + split_k - 1) / split_k;
    const int num_units = tiles_m * tiles_n * split_k;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
                const int tile_n = u % tiles_n;
                const int tile_m = (u / tiles_n) % tiles_m;
                const int split = u / (tiles_n * tiles_m);
                const int kb0 = split * kb_per_split;
                const int kb1 = min(kb0 + kb_per_split, kb_total);
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t *sa = smem + stage * Cfg::STAGE_BYTES;
                    uint8_t *sb = sa + A_STAGE_BYTES;
                    ptx::mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                    if (!A_MN) {
                        ptx::tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, tile_m * BM);
                    } else {
#pragma unroll
                        for (int j = 0; j < BM / 64; ++j)
                            ptx::tma_load_2d(sa + j * 8192, &tmap_a, &full_bar[stage], tile_m * BM + j * 64, kb * BK);
                    }
                    if (!B_MN) {
                        ptx::tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BK, tile_n * BN);
                    } else {
#pragma unroll
                        for (int j = 0; j < BN / 64; ++j)
                            ptx::tma_load_2d(sb + j * 8192, &tmap_b, &full_bar[stage], tile_n * BN + j * 64, kb * BK);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int u = blockIdx.x; u < num_units; u += gridDim.x, ++it) {
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
                        // K-major: 8-row groups 1024 B apart, advance 32 B per UMMA_K inside the swizzle row.
                        // MN-major: 64-element MN chunks 8192 B apart (LBO), 8-k groups 1024 B apart (SBO),
                        //           advance two k groups (2048 B) per UMMA_K.
                        const uint64_t da = A_MN ? ptx::make_smem_desc(sa + k * 2048, 8192, 1024)
                                                 : ptx::make_smem_desc(sa + k * 32, 16, 1024);
                        const uint64_t db = B_MN ? ptx::make_smem_desc(sb + k * 2048, 8192, 1024)
                                                 : ptx::make_smem_desc(sb + k * 32, 16, 1024);
                        ptx::umma_bf16(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    ptx::umma_commit(&empty_bar[stage]);  // frees this smem stage when the MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit(&tmem_full_bar[acc]);  // accumulator ready for the epilogue
            }
        }
    } else if (warp >= 4) {
        // ================================ epilogue ====================================
        const int q = warp & 3;              // TMEM lane quadrant this warp may access
        const int half = (warp - 4) >> 2;    // which half of the BN columns
        constexpr int COLS_PER_WARP = BN / 2;
        int it = 0;
        for (int u = blockIdx.x; u < num_units; u += gridDim.x, ++it) {
            const int tile_n = u % tiles_n;
            const int tile_m = (u / tiles_n) % tiles_m;
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
            ptx::tcgen05_fence_after();
            const int64_t row = static_cast<int64_t>(tile_m) * BM + q * 32 + lane;
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
            ptx::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[acc]);
        }
    }

    ptx::tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tcgen05_fence_after();
        ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

#include "gemm_tc_pair.cuh"

// ---- host side ------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (fn == nullptr) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

// 2-D bf16 tensor map over a row-major [outer, inner] matrix with leading dimension ld (elements)
int make_tmap(CUtensorMap *tm, const void *base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
              uint32_t box_outer, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    PFN_cuTensorMapEncodeTiled_v12000 enc = get_encode_fn();
    if (enc == nullptr) return fail(-2, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(-3, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}

template <int BN, bool A_MN, bool B_MN, int MODE>
int launch(const ecgvit_gemm_args *g, int split_k, cudaStream_t stream) {
    using Cfg = TileCfg<BN>;
    CUtensorMap ta, tb;
    int rc;
    if (!A_MN) rc = make_tmap(&ta, g->A, g->K, g->M, g->lda, BK, BM);
    else rc = make_tmap(&ta, g->A, g->M, g->K, g->lda, 64, BK);
    if (rc) return rc;
    if (!B_MN) rc = make_tmap(&tb, g->B, g->K, g->N, g->ldb, BK, BN);
    else rc = make_tmap(&tb, g->B, g->N, g->K, g->ldb, 64, BK);
    if (rc) return rc;

    auto kern = gemm_tc_kernel<BN, A_MN, B_MN, MODE>;
    static bool attr_set = false;  // per template instantiation
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute(gemm_tc): %s", cudaGetErrorString(e));
        attr_set = true;
    }
    const int tiles_m = (g->M + BM - 1) / BM, tiles_n = (g->N + BN - 1) / BN;
    const int units = tiles_m * tiles_n * split_k;
    const int grid = units < sm_count() ? units : sm_count();
    EpiParams ep{g->out, g->out2, g->aux, g->bias, g->ldo, make_dropout(g->dropout_p, g->dropout_stream, g->dropout_seed)};
    kern<<<grid, kNumThreads, Cfg::SMEM_BYTES, stream>>>(ta, tb, g->M, g->N, g->K, split_k, ep);
    return check_launch("gemm_tc");
}

template <int BN> int dispatch(const ecgvit_gemm_args *g, int split_k, cudaStream_t s) {
    const bool a_mn = !g->a_kmajor, b_mn = !g->b_kmajor;
    const int mode = g->epilogue;
#define ECGVIT_CASE(AM, BMN, MODE) \
    if (a_mn == AM && b_mn == BMN && mode == MODE) return launch<BN, AM, BMN, MODE>(g, split_k, s);
    ECGVIT_CASE(false, false, ECGVIT_EPI_STORE)
    ECGVIT_CASE(false, false, ECGVIT_EPI_BIAS_RES)
    ECGVIT_CASE(false, false, ECGVIT_EPI_BIAS_GELU)
    ECGVIT_CASE(false, true, ECGVIT_EPI_STORE)
    ECGVIT_CASE(false, true, ECGVIT_EPI_DGELU)
    ECGVIT_CASE(true, true, ECGVIT_EPI_ATOMIC_F32)
    ECGVIT_CASE(true, true, ECGVIT_EPI_STORE)
    ECGVIT_CASE(true, false, ECGVIT_EPI_STORE)
#undef ECGVIT_CASE
    return fail(-1, "gemm(bf16): unsupported combination a_kmajor=%d b_kmajor=%d epilogue=%d", g->a_kmajor,
                g->b_kmajor, mode);
}


template <int BN, bool A_MN, bool B_MN, int MODE>
int launch_pair(const ecgvit_gemm_args *g, int split_k, cudaStream_t stream) {
    using Cfg = PairCfg<BN, B_MN>;
    CUtensorMap ta, tb;
    int rc;
    if (!A_MN) rc = make_tmap(&ta, g->A, g->K, g->M, g->lda, BK, BM);
    else rc = make_tmap(&ta, g->A, g->M, g->K, g->lda, 64, BK);
    if (rc) return rc;
    if (!B_MN) rc = make_tmap(&tb, g->B, g->K, g->N, g->ldb, BK, Cfg::B_HALF);
    else rc = make_tmap(&tb, g->B, g->N, g->K, g->ldb, 64, BK);
    if (rc) return rc;
    // outputs leave (and the residual / pre-activation operand arrives) through TMA on 32-row x 32-column tiles
    // with 64-byte swizzle; TMA clips at M x N
    CUtensorMap to = ta, to2 = ta, tx = ta;
    if (MODE != ECGVIT_EPI_ATOMIC_F32) {
        if ((rc = make_tmap(&to, g->out, g->N, g->M, g->ldo, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
        if (MODE == ECGVIT_EPI_BIAS_GELU &&
            (rc = make_tmap(&to2, g->out2, g->N, g->M, g->ldo, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
        if ((MODE == ECGVIT_EPI_BIAS_RES || MODE == ECGVIT_EPI_DGELU) &&
            (rc = make_tmap(&tx, g->aux, g->N, g->M, g->ldo, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    }
    auto kern = gemm_tc2_kernel<BN, A_MN, B_MN, MODE>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute(gemm_tc2): %s", cudaGetErrorString(e));
        attr_set = true;
    }
    const int tiles_m = (g->M + 2 * BM - 1) / (2 * BM), tiles_n = (g->N + BN - 1) / BN;
    const int units = tiles_m * tiles_n * split_k;
    const int clusters = sm_count() / 2;
    const int grid = 2 * (units < clusters ? units : clusters);
    EpiParams ep{g->out, g->out2, g->aux, g->bias, g->ldo, make_dropout(g->dropout_p, g->dropout_stream, g->dropout_seed)};
    cudaError_t le = launch_pdl(kern, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, stream, ta, tb, to, to2, tx, g->M,
                                g->N, g->K, split_k, ep);
    if (le != cudaSuccess) return fail((int)le, "gemm_tc2 launch: %s", cudaGetErrorString(le));
    return check_launch("gemm_tc2");
}

template <int BN> int dispatch_pair(const ecgvit_gemm_args *g, int split_k, cudaStream_t s) {
    const bool a_mn = !g->a_kmajor, b_mn = !g->b_kmajor;
    const int mode = g->epilogue;
#define ECGVIT_CASE(AM, BMN, MODE) \
    if (a_mn == AM && b_mn == BMN && mode == MODE) return launch_pair<BN, AM, BMN, MODE>(g, split_k, s);
    ECGVIT_CASE(false, false, ECGVIT_EPI_STORE)
    ECGVIT_CASE(false, false, ECGVIT_EPI_BIAS_RES)
    ECGVIT_CASE(false, false, ECGVIT_EPI_BIAS_GELU)
    ECGVIT_CASE(false, true, ECGVIT_EPI_STORE)
    ECGVIT_CASE(false, true, ECGVIT_EPI_DGELU)
    ECGVIT_CASE(true, true, ECGVIT_EPI_ATOMIC_F32)
    ECGVIT_CASE(true, true, ECGVIT_EPI_STORE)
    ECGVIT_CASE(true, false, ECGVIT_EPI_STORE)
#undef ECGVIT_CASE
    return fail(-1, "gemm(bf16): unsupported combination a_kmajor=%d b_kmajor=%d epilogue=%d", g->a_kmajor,
                g->b_kmajor, mode);
}

inline double b_kmajor_eff(int b_kmajor) { return b_kmajor ? 0.88 : 0.82; }  // MN-major B over-fetches at 192

// ECGVIT_GEMM_CTA_GROUP=1 selects the single-CTA kernel (kept for A/B measurements); default is the CTA pair
int cta_group_setting() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("ECGVIT_GEMM_CTA_GROUP");
        v = (e != nullptr && e[0] == '1') ? 1 : 2;
    }
    return v;
}

}  // namespace

int gemm_bf16_tc(const ecgvit_gemm_args *g, cudaStream_t stream) {
    ECGVIT_REQUIRE(g->M > 0 && g->N > 0 && g->K > 0, "gemm: empty problem M=%d N=%d K=%d", g->M, g->N, g->K);
    ECGVIT_REQUIRE(g->N % 8 == 0, "gemm(bf16): N=%d must be a multiple of 8", g->N);
    ECGVIT_REQUIRE(g->lda % 8 == 0 && g->ldb % 8 == 0, "gemm(bf16): lda=%lld ldb=%lld must be multiples of 8",
                   (long long)g->lda, (long long)g->ldb);
    ECGVIT_REQUIRE((reinterpret_cast<uintptr_t>(g->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(g->B) & 15) == 0,
                   "gemm(bf16): operand pointers must be 16-byte aligned");
    ECGVIT_REQUIRE(g->ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(g->out) & 15) == 0,
                   "gemm(bf16): output must be 16-byte aligned with ldo %% 8 == 0");
    const int sms = sm_count();
    const int kb_total = (g->K + BK - 1) / BK;
    int split_k = g->split_k > 1 ? g->split_k : 1;
    ECGVIT_REQUIRE(split_k == 1 || g->epilogue == ECGVIT_EPI_ATOMIC_F32, "gemm: split_k needs the atomic epilogue");
    if (g->epilogue == ECGVIT_EPI_ATOMIC_F32 && g->split_k <= 0) {
        // auto: smallest split (>= 8 k blocks each) whose work units fill >= 90 % of their last wave
        const int pair = cta_group_setting() == 2 ? 2 : 1;  // work units are 256-row tiles run by CTA pairs
        const int workers = sms / pair;
        const long tiles = (long)((g->M + pair * BM - 1) / (pair * BM)) * ((g->N + 255) / 256);
        int best = 1;
        double best_eff = 0.0;
        for (int s = 1; s <= 32 && s * 8 <= kb_total; ++s) {
            const long units = tiles * s;
            const double eff = (double)units / (double)(((units + workers - 1) / workers) * workers);
            if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
            if (eff >= 0.9) { best = s; break; }
        }
        split_k = best;
    }
    if (split_k > kb_total) split_k = kb_total;
    // every split must own at least one k block
    while (split_k > 1 && ((kb_total + split_k - 1) / split_k) * (split_k - 1) >= kb_total) --split_k;

    // Tile width.  Narrow tiles stage more bytes per flop (a 128-wide single-CTA tile needs the full 128 B/clk of
    // shared-memory bandwidth for operand reads alone), so they are only chosen when the wide tile would waste a
    // large part of its last wave.
    if (cta_group_setting() == 2) {
        // Shared memory moves 128 B/clk per SM and every staged byte is written once (TMA) and read once (MMA), so a
        // 256 x BN pair tile can sustain at most min(1, 128 / (2 * bytes per k block / MMA clocks per k block)) of the
        // tensor peak: 1.00 at BN = 256, 0.88 at 192, 0.67 at 128.  Pick the width with the best (tile efficiency x
        // last-wave occupancy).
        const int clusters = sms / 2;
        const long tiles_m2 = (g->M + 2 * BM - 1) / (2 * BM);
        auto score = [&](int bn, double eff) {
            const long units = tiles_m2 * ((g->N + bn - 1) / bn) * split_k;
            const long waves = (units + clusters - 1) / clusters;
            // time ~ waves * bn / eff  (per-tile MMA time is proportional to bn)
            return (double)waves * bn / eff;
        };
        const double s256 = score(256, 1.0), s192 = score(192, b_kmajor_eff(g->b_kmajor)), s128 = score(128, 0.67);
        if (g->N <= 128 || (s128 < s256 && s128 < s192)) return dispatch_pair<128>(g, split_k, stream);
        if (s192 < s256) return dispatch_pair<192>(g, split_k, stream);
        return dispatch_pair<256>(g, split_k, stream);
    }
    auto waves = [&](int bn) {
        const long units = (long)((g->M + BM - 1) / BM) * ((g->N + bn - 1) / bn) * split_k;
        return (units + sms - 1) / sms;
    };
    const bool use128 = (g->N <= 128) || (waves(128) * 128 * 1.6 < waves(256) * 256);
    return use128 ? dispatch<128>(g, split_k, stream) : dispatch<256>(g, split_k, stream);
}

}  // namespace ecgvit
