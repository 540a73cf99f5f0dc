// Softmax attention on the 5th-generation tensor cores (tcgen05 / TMEM / TMA) for head dim 64, bf16.
//
// Replaces chunk(3) + rearrange + q k^T * scale + Softmax + Dropout + attn v + rearrange back of vit_pytorch's
// Attention.forward (reached from /root/reference/ecg_transformer/models/ecg_vit.py:102-116,141) and its autograd
// backward (models/train.py:280).  Operates in place on the packed projection qkv [B*N, 3*H*64] (q | k | v, head-major)
// through 3-D TMA boxes of the [B][N][3*H*64] view, so sequences are zero-padded by the TMA unit, never in memory.
//
// Two geometries share every kernel:
//   packed (N <= 64; BASELINE.json configs[1]: N = 51): one 128-row tile holds TWO (batch, head) problems (64 padded
//     rows each); S = Q K^T is a 128 x 128 product whose diagonal 64 x 64 blocks are the two problems, P is kept
//     block-diagonal (zeros elsewhere) so one 128 x 64 x 128 product gives both outputs.  One key block per tile: these
//     kernels live on the TMA ring that keeps loads of the next tiles in flight (the shape is HBM-bound).
//   long (N > 64; configs[3]: N = 2401 per-lead tokens): a tile is 128 queries of one (batch, head), key blocks of 128 are
//     streamed with a running (max, sum); nothing of size N x N is ever written.
//
// Forward (attn_tc_fwd_kernel), one persistent CTA per SM, 12 warps:
//   warp 0   TMA producer: Q tiles (2-deep ring per tile slot) and K/V blocks (4-deep ring)
//   warp 1   MMA issuer : S_t = Q_t K^T (SS, 128x128x64) and O_t += P_t V (TS: P from TMEM, V MN-major from smem)
//   warp 2   TMEM allocator (512 columns: S0 S1 | O0 O1 | P0 P1)
//   warps 4-7 / 8-11  softmax warpgroup of tile 0 / tile 1 (thread = query row): S -> registers, running max / sum,
//            exp2, dropout, P (bf16, packed two per column) -> TMEM, lazy rescale of O, final O / l -> smem -> TMA store
// The two tiles of a CTA ping-pong: the tensor core runs tile 1's products while tile 0's warpgroup is in its softmax.
// In the long geometry both tiles are consecutive query tiles of the same head and share every K/V block.
#include "common.cuh"
#include "ptx.cuh"

#include <cudaTypedefs.h>
#include <stdlib.h>

namespace ecgvit {

namespace {

constexpr int DH = 64;               // head dim served by these kernels
constexpr int TM = 128;              // query rows per tile = TMEM lanes
constexpr int TN = 128;              // keys per block
constexpr int TILE_BYTES = TM * DH * 2;  // 16 KB: one [128 x 64] bf16 tile, 128-byte rows, SWIZZLE_128B
constexpr float LOG2E_F = 1.4426950408889634f;

// ---- forward shared-memory map (offsets from a 1024-byte aligned base) -----------------------------------------------
constexpr int F_Q_OFF = 0;                               // [tile 2][slot 2] Q tiles
constexpr int F_KV_OFF = F_Q_OFF + 4 * TILE_BYTES;       // [stage 4] K tile | V tile
constexpr int F_KV_STAGES = 4;
constexpr int F_OST_OFF = F_KV_OFF + F_KV_STAGES * 2 * TILE_BYTES;  // [tile 2] output staging
constexpr int F_BAR_OFF = F_OST_OFF + 2 * TILE_BYTES;
constexpr int F_SMEM_BYTES = F_BAR_OFF + 256 + 1024;     // + barrier block + alignment slack
// TMEM columns
constexpr uint32_t F_TM_S = 0, F_TM_O = 256, F_TM_P = 384;  // S_t at S + 128 t, O_t at O + 64 t, P_t at P + 64 t

struct FwdBars {
    uint64_t q_full[2][2], q_empty[2][2];
    uint64_t kv_full[F_KV_STAGES], kv_empty[F_KV_STAGES];
    uint64_t s_full[2], p_full[2], pv_done[2];
    uint32_t tmem_ptr;
};
static_assert(sizeof(FwdBars) <= 256, "barrier block");

// work decomposition shared by all roles (everything here is CTA-uniform)
template <bool kPacked> struct FwdItem {
    bool valid[2];
    int prob[2][2];  // packed: the two (b*H + h) problems of tile t; long: prob[t][0] = b*H + h
    int q0[2];       // long: first query row of tile t
};
template <bool kPacked>
__device__ __forceinline__ void fwd_decode(int item, int B, int N, int H, FwdItem<kPacked> &it) {
    if (kPacked) {
        const int n_prob = B * H, n_pt = (n_prob + 1) >> 1;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int pt = 2 * item + t;
            it.valid[t] = pt < n_pt;
            it.prob[t][0] = 2 * pt;
            it.prob[t][1] = 2 * pt + 1;  // may equal n_prob: its loads are fully out of bounds (zeros), its stores clipped
            it.q0[t] = 0;
        }
    } else {
        const int nqt = (N + TM - 1) / TM, pairs = (nqt + 1) >> 1;
        const int bh = item / pairs, qp = item - bh * pairs;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            it.valid[t] = 2 * qp + t < nqt;
            it.prob[t][0] = it.prob[t][1] = bh;
            it.q0[t] = (2 * qp + t) * TM;
        }
    }
}
template <bool kPacked> __device__ __forceinline__ int fwd_items(int B, int N, int H) {
    if (kPacked) return (((B * H + 1) >> 1) + 1) >> 1;
    const int nqt = (N + TM - 1) / TM;
    return B * H * ((nqt + 1) >> 1);
}

// byte offset of 16-byte chunk j (0..7) of row r inside a [rows x 128 B] SWIZZLE_128B tile (1024-byte aligned base)
__device__ __forceinline__ uint32_t sw128_offset(int r, int j) {
    return static_cast<uint32_t>(r) * 128u + (static_cast<uint32_t>(j ^ (r & 7)) << 4);
}

template <bool kPacked>
__global__ void __launch_bounds__(384, 1)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_o,
                   float *__restrict__ lse, int B, int N, int H, float scale, DropoutParams drop) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    FwdBars *bars = reinterpret_cast<FwdBars *>(smem + F_BAR_OFF);

    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int inner = H * DH;
    const int n_items = fwd_items<kPacked>(B, N, H);
    const int nkb = kPacked ? 1 : (N + TN - 1) / TN;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tm_qkv);
        ptx::prefetch_tensormap(&tm_o);
    }
    if (warp == 1 && lane == 0) {
        for (int t = 0; t < 2; ++t) {
            for (int s = 0; s < 2; ++s) {
                ptx::mbar_init(&bars->q_full[t][s], 1);
                ptx::mbar_init(&bars->q_empty[t][s], 1);
            }
            ptx::mbar_init(&bars->s_full[t], 1);
            ptx::mbar_init(&bars->p_full[t], 4);   // one arrival per warp of the tile's softmax warpgroup
            ptx::mbar_init(&bars->pv_done[t], 1);
        }
        for (int s = 0; s < F_KV_STAGES; ++s) {
            ptx::mbar_init(&bars->kv_full[s], 1);
            ptx::mbar_init(&bars->kv_empty[s], 1);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 2) ptx::tmem_alloc(&bars->tmem_ptr, 512);
    ptx::tcgen05_fence_before();
    __syncthreads();
    ptx::tcgen05_fence_after();
    const uint32_t tmem_base = bars->tmem_ptr;
    pdl_wait();

    if (warp == 0) {
        // ================================ TMA producer ==================================================
        int qcnt[2] = {0, 0}, kvcnt = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            FwdItem<kPacked> it;
            fwd_decode<kPacked>(item, B, N, H, it);
            auto load_q = [&](int t) {
                const int slot = qcnt[t] & 1;
                ptx::mbar_wait(&bars->q_empty[t][slot], ((qcnt[t] >> 1) & 1) ^ 1);
                if (ptx::elect_one()) {
                    uint8_t *dst = smem + F_Q_OFF + (t * 2 + slot) * TILE_BYTES;
                    ptx::mbar_arrive_expect_tx(&bars->q_full[t][slot], TILE_BYTES);
                    if (kPacked) {
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const int p = it.prob[t][i];
                            ptx::tma_load_3d(dst + i * (TILE_BYTES / 2), &tm_qkv, &bars->q_full[t][slot], (p % H) * DH, 0,
                                             p / H);
                        }
                    } else {
                        const int p = it.prob[t][0];
                        ptx::tma_load_3d(dst, &tm_qkv, &bars->q_full[t][slot], (p % H) * DH, it.q0[t], p / H);
                    }
                }
                __syncwarp();
                ++qcnt[t];
            };
            // K block (rows k0.. of problem(s) `prob`) and the matching V block into the next ring stage
            auto load_kv = [&](const int prob[2], int k0) {
                const int st = kvcnt % F_KV_STAGES;
                ptx::mbar_wait(&bars->kv_empty[st], ((kvcnt / F_KV_STAGES) & 1) ^ 1);
                if (ptx::elect_one()) {
                    uint8_t *dst = smem + F_KV_OFF + st * 2 * TILE_BYTES;
                    ptx::mbar_arrive_expect_tx(&bars->kv_full[st], 2 * TILE_BYTES);
                    if (kPacked) {
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const int p = prob[i], c = (p % H) * DH, b = p / H;
                            ptx::tma_load_3d(dst + i * (TILE_BYTES / 2), &tm_qkv, &bars->kv_full[st], inner + c, 0, b);
                            ptx::tma_load_3d(dst + TILE_BYTES + i * (TILE_BYTES / 2), &tm_qkv, &bars->kv_full[st],
                                             2 * inner + c, 0, b);
                        }
                    } else {
                        const int p = prob[0], c = (p % H) * DH, b = p / H;
                        ptx::tma_load_3d(dst, &tm_qkv, &bars->kv_full[st], inner + c, k0, b);
                        ptx::tma_load_3d(dst + TILE_BYTES, &tm_qkv, &bars->kv_full[st], 2 * inner + c, k0, b);
                    }
                }
                __syncwarp();
                ++kvcnt;
            };
            if (kPacked) {
#pragma unroll
                for (int t = 0; t < 2; ++t)
                    if (it.valid[t]) {
                        load_q(t);
                        load_kv(it.prob[t], 0);
                    }
            } else {
#pragma unroll
                for (int t = 0; t < 2; ++t)
                    if (it.valid[t]) load_q(t);
                for (int j = 0; j < nkb; ++j) load_kv(it.prob[0], j * TN);
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ====================================================
        constexpr uint32_t idesc_qk = ptx::make_idesc_bf16(TM, TN, 0, 0);
        constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(TM, DH, 0, 1);
        const uint32_t smem_base = ptx::smem_u32(smem);
        int qcnt[2] = {0, 0}, kvcnt = 0;
        uint32_t pph[2] = {0, 0};
        // S_t = Q_t K^T with K from ring stage `st`
        auto issue_qk = [&](int t, int slot, int st) {
            if (ptx::elect_one()) {
                const uint64_t da = ptx::make_smem_desc(smem_base + F_Q_OFF + (t * 2 + slot) * TILE_BYTES, 16, 1024);
                const uint64_t db = ptx::make_smem_desc(smem_base + F_KV_OFF + st * 2 * TILE_BYTES, 16, 1024);
#pragma unroll
                for (int k = 0; k < DH / 16; ++k)
                    ptx::umma_bf16(tmem_base + F_TM_S + t * TN, da + 2 * k, db + 2 * k, idesc_qk, k > 0 ? 1u : 0u);
                ptx::umma_commit(&bars->s_full[t]);
            }
            __syncwarp();
        };
        // O_t (+)= P_t V with V from ring stage `st` (MN-major: one key per 128-byte row, 16 keys = 2048 B per UMMA_K)
        auto issue_pv = [&](int t, int st, bool acc, bool release_kv) {
            if (ptx::elect_one()) {
                const uint64_t db = ptx::make_smem_desc(smem_base + F_KV_OFF + st * 2 * TILE_BYTES + TILE_BYTES, 8192, 1024);
#pragma unroll
                for (int k = 0; k < TN / 16; ++k)
                    ptx::umma_bf16_ts(tmem_base + F_TM_O + t * DH, tmem_base + F_TM_P + t * (TN / 2) + 8 * k,
                                      db + 128 * k, idesc_pv, (acc || k > 0) ? 1u : 0u);
                ptx::umma_commit(&bars->pv_done[t]);
                if (release_kv) ptx::umma_commit(&bars->kv_empty[st]);
            }
            __syncwarp();
        };
        auto release_q = [&](int t, int slot) {
            if (ptx::elect_one()) ptx::umma_commit(&bars->q_empty[t][slot]);
            __syncwarp();
        };
        auto wait_kv = [&](int cnt) {
            ptx::mbar_wait(&bars->kv_full[cnt % F_KV_STAGES], (cnt / F_KV_STAGES) & 1);
        };
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            FwdItem<kPacked> it;
            fwd_decode<kPacked>(item, B, N, H, it);
            const int last_t = it.valid[1] ? 1 : 0;
            if (kPacked) {
                int st[2] = {0, 0};
#pragma unroll
                for (int t = 0; t < 2; ++t)
                    if (it.valid[t]) {
                        const int slot = qcnt[t] & 1;
                        st[t] = kvcnt % F_KV_STAGES;
                        ptx::mbar_wait(&bars->q_full[t][slot], (qcnt[t] >> 1) & 1);
                        wait_kv(kvcnt);
                        ptx::tcgen05_fence_after();
                        issue_qk(t, slot, st[t]);
                        release_q(t, slot);
                        ++qcnt[t];
                        ++kvcnt;
                    }
#pragma unroll
                for (int t = 0; t < 2; ++t)
                    if (it.valid[t]) {
                        ptx::mbar_wait(&bars->p_full[t], pph[t]);
                        pph[t] ^= 1;
                        ptx::tcgen05_fence_after();
                        issue_pv(t, st[t], false, true);
                    }
            } else {
                int slot[2] = {qcnt[0] & 1, qcnt[1] & 1};
                wait_kv(kvcnt);
#pragma unroll
                for (int t = 0; t < 2; ++t)
                    if (it.valid[t]) {
                        ptx::mbar_wait(&bars->q_full[t][slot[t]], (qcnt[t] >> 1) & 1);
                        ptx::tcgen05_fence_after();
                        issue_qk(t, slot[t], kvcnt % F_KV_STAGES);
                        if (nkb == 1) release_q(t, slot[t]);
                    }
                for (int j = 0; j < nkb; ++j) {
                    const int st = (kvcnt + j) % F_KV_STAGES;
#pragma unroll
                    for (int t = 0; t < 2; ++t)
                        if (it.valid[t]) {
                            ptx::mbar_wait(&bars->p_full[t], pph[t]);
                            pph[t] ^= 1;
                            ptx::tcgen05_fence_after();
                            issue_pv(t, st, j > 0, t == last_t);
                            if (j + 1 < nkb) {
                                if (t == 0) wait_kv(kvcnt + j + 1);  // tile 0 is always valid
                                ptx::tcgen05_fence_after();
                                issue_qk(t, slot[t], (kvcnt + j + 1) % F_KV_STAGES);
                                if (j + 2 == nkb) release_q(t, slot[t]);
                            }
                        }
                }
                kvcnt += nkb;
#pragma unroll
                for (int t = 0; t < 2; ++t)
                    if (it.valid[t]) ++qcnt[t];
            }
        }
    } else if (warp >= 4) {
        // ================================ softmax warpgroups (tile = (warp - 4) / 4) ==================
        const int t = (warp - 4) >> 2;
        const int q = warp & 3;                 // TMEM lane quadrant of this warp
        const int r = q * 32 + lane;            // row of the tile owned by this thread
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const uint32_t tS = lane_addr + F_TM_S + t * TN, tO = lane_addr + F_TM_O + t * DH,
                       tP = lane_addr + F_TM_P + t * (TN / 2);
        const float sl2 = scale * LOG2E_F;
        const bool dropping = drop.threshold != 0;
        const uint32_t seed = dropping ? __ldg(drop.seed) : 0u;
        const int Np = (N + 63) / 64 * 64;      // row pitch of the dropout counter (shared with the other kernels)
        // packed: this thread's problem is half `hf` of the tile; it reads S columns [64 hf, 64 hf + 64)
        const int hf = kPacked ? (r >> 6) : 0;
        const int col0 = kPacked ? 64 * hf : 0;
        constexpr int NCH = kPacked ? 2 : 4;    // 32-column chunks of S read per key block
        uint8_t *ost = smem + F_OST_OFF + t * TILE_BYTES;
        if (kPacked) {
            // the off-diagonal half of P never changes: zero it once (TMEM is not cleared by allocation)
            uint32_t z[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) z[i] = 0u;
            ptx::tmem_st_32x16(tP + (hf ? 0 : 32), z);
            ptx::tmem_st_32x16(tP + (hf ? 0 : 32) + 16, z);
            ptx::tmem_st_wait();
        }
        uint32_t sph = 0, dph = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            FwdItem<kPacked> it;
            fwd_decode<kPacked>(item, B, N, H, it);
            if (!it.valid[t]) continue;
            const int prob = it.prob[t][hf];
            const int qrow = kPacked ? (r & 63) : it.q0[t] + r;    // query index inside the problem
            float m = -INFINITY, l = 0.f;
            for (int j = 0; j < nkb; ++j) {
                const int nvalid = kPacked ? N : min(TN, N - j * TN);  // valid keys of this block (>= 1)
                ptx::mbar_wait(&bars->s_full[t], sph);
                sph ^= 1;
                ptx::tcgen05_fence_after();
                // ---- pass 1: running max
                float mx = m;
#pragma unroll 1
                for (int c = 0; c < NCH; ++c) {
                    const int lim = nvalid - 32 * c;   // valid columns of this chunk
                    if (lim <= 0) break;
                    uint32_t s[32];
                    ptx::tmem_ld_32x32(tS + col0 + 32 * c, s);
                    ptx::tmem_ld_wait();
                    if (lim >= 32) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(s[i]));
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, i < lim ? __uint_as_float(s[i]) : -INFINITY);
                    }
                }
                const float alpha = ex2_approx((m - mx) * sl2);   // 0 on the first block (m = -inf)
                m = mx;
                const float nm = -m * sl2;
                if (j > 0) {
                    // P_t and O_t are free once the previous block's P V product has retired
                    ptx::mbar_wait(&bars->pv_done[t], dph);
                    dph ^= 1;
                    ptx::tcgen05_fence_after();
                    if (__any_sync(0xffffffffu, alpha != 1.0f)) {   // lazy: the running max rarely moves after a few blocks
#pragma unroll 1
                        for (int c = 0; c < DH / 32; ++c) {
                            uint32_t o[32];
                            ptx::tmem_ld_32x32(tO + 32 * c, o);
                            ptx::tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                            ptx::tmem_st_32x32(tO + 32 * c, o);
                        }
                    }
                }
                // ---- pass 2: p = exp2(s * sl2 - m * sl2), row sum, dropout, bf16 pack -> TMEM
                float sum = 0.f;
                const uint32_t e_row = (static_cast<uint32_t>(prob) * Np + static_cast<uint32_t>(qrow)) * Np +
                                       static_cast<uint32_t>(j * TN);
#pragma unroll 1
                for (int c = 0; c < NCH; ++c) {
                    const int lim = nvalid - 32 * c;
                    uint32_t pk[16];
                    if (lim <= 0) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) pk[i] = 0u;
                    } else {
                        uint32_t s[32];
                        ptx::tmem_ld_32x32(tS + col0 + 32 * c, s);
                        ptx::tmem_ld_wait();
                        float p[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) p[i] = ex2_approx(fmaf(__uint_as_float(s[i]), sl2, nm));
                        if (lim < 32) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) p[i] = i < lim ? p[i] : 0.f;
                        }
#pragma unroll
                        for (int i = 0; i < 32; ++i) sum += p[i];
                        if (dropping) dropout_apply<32>(drop, seed, e_row + 32 * c, p);
#pragma unroll
                        for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(p[2 * i], p[2 * i + 1]);
                    }
                    ptx::tmem_st_32x16(tP + (col0 >> 1) + 16 * c, pk);
                }
                l = l * alpha + sum;
                ptx::tmem_st_wait();
                ptx::tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&bars->p_full[t]);
            }
            // ---- epilogue: O / l -> bf16 -> swizzled staging tile -> TMA store; log-sum-exp of the scaled scores
            ptx::mbar_wait(&bars->pv_done[t], dph);
            dph ^= 1;
            ptx::tcgen05_fence_after();
            if (q == 0 && lane == 0) ptx::tma_store_wait_read<0>();  // the previous store of this tile has read `ost`
            ptx::named_bar_sync(1 + t, 128);
            const float inv_l = 1.0f / l;
#pragma unroll 1
            for (int c = 0; c < DH / 32; ++c) {
                uint32_t o[32];
                ptx::tmem_ld_32x32(tO + 32 * c, o);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    uint4 v;
                    v.x = pack_bf16x2(__uint_as_float(o[8 * jj + 0]) * inv_l, __uint_as_float(o[8 * jj + 1]) * inv_l);
                    v.y = pack_bf16x2(__uint_as_float(o[8 * jj + 2]) * inv_l, __uint_as_float(o[8 * jj + 3]) * inv_l);
                    v.z = pack_bf16x2(__uint_as_float(o[8 * jj + 4]) * inv_l, __uint_as_float(o[8 * jj + 5]) * inv_l);
                    v.w = pack_bf16x2(__uint_as_float(o[8 * jj + 6]) * inv_l, __uint_as_float(o[8 * jj + 7]) * inv_l);
                    *reinterpret_cast<uint4 *>(ost + sw128_offset(r, 4 * c + jj)) = v;
                }
            }
            ptx::tcgen05_fence_before();   // O_t has been read: the next item's first P V may overwrite it
            if (qrow < N && prob < B * H) lse[static_cast<int64_t>(prob) * N + qrow] = m * scale + logf(l);
            ptx::fence_proxy_async_smem();
            ptx::named_bar_sync(1 + t, 128);
            if (q == 0 && lane == 0) {
                if (kPacked) {
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int p = it.prob[t][i];
                        ptx::tma_store_3d(&tm_o, ost + i * (TILE_BYTES / 2), (p % H) * DH, 0, p / H);
                    }
                } else {
                    const int p = it.prob[t][0];
                    ptx::tma_store_3d(&tm_o, ost, (p % H) * DH, it.q0[t], p / H);
                }
                ptx::tma_store_commit();
            }
        }
        if (q == 0 && lane == 0) ptx::tma_store_wait_all<0>();
    }

    ptx::tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tcgen05_fence_after();
        ptx::tmem_dealloc(tmem_base, 512);
    }
}

// ---- host side ----------------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
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

// bf16 [B][N][cols] view (row pitch `ld` elements, batch pitch N * ld) with a {64 features, box_rows tokens, 1} box
int make_tmap3(CUtensorMap *tm, const void *base, int cols, int N, int B, int64_t ld, int box_rows) {
    PFN_cuTensorMapEncodeTiled_v12000 enc = encode_fn();
    if (enc == nullptr) return fail(-2, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)N, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)N * (cuuint64_t)ld * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(-3, "cuTensorMapEncodeTiled(3d) failed with CUresult %d", (int)r);
    return 0;
}

template <bool kPacked>
int launch_fwd(const void *qkv, void *o, float *lse, int B, int N, int H, float scale, DropoutParams drop,
               cudaStream_t stream) {
    const int inner = H * DH;
    CUtensorMap tq, to;
    int rc;
    if ((rc = make_tmap3(&tq, qkv, 3 * inner, N, B, 3 * (int64_t)inner, kPacked ? 64 : TM))) return rc;
    if ((rc = make_tmap3(&to, o, inner, N, B, inner, kPacked ? 64 : TM))) return rc;
    auto kern = attn_tc_fwd_kernel<kPacked>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM_BYTES);
        if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute(attn_tc_fwd): %s", cudaGetErrorString(e));
        attr_set = true;
    }
    int items;
    if (kPacked) items = (((B * H + 1) >> 1) + 1) >> 1;
    else items = B * H * ((((N + TM - 1) / TM) + 1) >> 1);
    const int grid = items < sm_count() ? items : sm_count();
    cudaError_t le = launch_pdl(kern, dim3(grid), dim3(384), F_SMEM_BYTES, stream, tq, to, lse, B, N, H, scale, drop);
    if (le != cudaSuccess) return fail((int)le, "attn_tc_fwd launch: %s", cudaGetErrorString(le));
    return check_launch("attn_tc_fwd");
}

}  // namespace

// ECGVIT_ATTN=mma keeps the warp-level mma.sync kernels (A/B measurements)
bool attention_tc_enabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("ECGVIT_ATTN");
        v = (e != nullptr && e[0] == 'm') ? 0 : 1;
    }
    return v == 1;
}

bool attention_tc_supported(int N, int dh) { return attention_tc_enabled() && dh == DH && N >= 1; }

int attention_fwd_tc(const void *qkv, void *o, float *lse, int B, int N, int H, int dh, float scale, DropoutParams drop,
                     cudaStream_t stream) {
    ECGVIT_REQUIRE(dh == DH, "attention_fwd_tc: head dim %d (only 64)", dh);
    ECGVIT_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(o) & 15) == 0,
                   "attention_fwd_tc: qkv / o must be 16-byte aligned");
    if (N <= 64) return launch_fwd<true>(qkv, o, lse, B, N, H, scale, drop, stream);
    return launch_fwd<false>(qkv, o, lse, B, N, H, scale, drop, stream);
}

}  // namespace ecgvit
