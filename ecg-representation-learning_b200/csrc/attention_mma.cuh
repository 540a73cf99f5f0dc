// Warp-level tensor-core building blocks shared by the short-sequence (attention_mma.cu) and the tiled
// (attention_flash.cu) attention kernels: ldmatrix / mma.sync m16n8k16 bf16 wrappers, 64-row smem tiles with a
// +8 element pad (conflict-free ldmatrix), and the three products every kernel is made of.
// Include inside namespace ecgvit { namespace { ... } }.
#pragma once

constexpr int NMAX = 64;  // padded sequence length handled by one CTA
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ void ldsm_x4(uint32_t r[4], const void *p) {
    const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t r[4], const void *p) {
    const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
// D(16x8, fp32) += A(16x16, bf16, row) * B(16x8, bf16, col)
__device__ __forceinline__ void mma16816(float c[4], const uint32_t a[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
                 "{%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}
__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
    return v;
}

__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(dst))),
                 "l"(src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// copy an [N x DH] head tile (row stride `ld` elements) into smem [NMAX][DH + 8] with cp.async (no register staging),
// zero-filling rows >= N; the caller waits with cp_async_wait_all() + __syncthreads()
template <int DH>
__device__ __forceinline__ void load_tile(bf16 (*dst)[DH + 8], const bf16 *src, int ld, int N) {
    // 128 threads: thread t owns 16-byte chunk (t % VPR) of rows t / VPR + k * (128 / VPR); pointers advance by a
    // constant per step (the generic i / VPR, i % VPR form cost ~30 integer instructions per cp.async)
    constexpr int VPR = DH / 8;
    constexpr int ROWS_PER_STEP = 128 / VPR;
    const int c = (threadIdx.x % VPR) * 8;
    int r = threadIdx.x / VPR;
    const bf16 *p = src + static_cast<int64_t>(r) * ld + c;
    bf16 *q = &dst[r][c];
    const int64_t pstep = static_cast<int64_t>(ROWS_PER_STEP) * ld;
#pragma unroll
    for (int k = 0; k < NMAX / ROWS_PER_STEP; ++k) {
        if (r < N) cp_async16(q, p);
        else *reinterpret_cast<uint4 *>(q) = make_uint4(0u, 0u, 0u, 0u);
        r += ROWS_PER_STEP;
        p += pstep;
        q += ROWS_PER_STEP * (DH + 8);
    }
}

// s[j][:] (16 rows x 64 keys, C-fragment layout) = X[m0:m0+16, :DH] * Y[:, :DH]^T, both row-major in smem
template <int DH, int NT>
__device__ __forceinline__ void rows_times_transposed(float s[2 * NT][4], bf16 (*X)[DH + 8], bf16 (*Y)[DH + 8], int m0,
                                                      int lane) {
#pragma unroll
    for (int j = 0; j < 2 * NT; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) s[j][i] = 0.f;
#pragma unroll
    for (int kk = 0; kk < DH / 16; ++kk) {
        uint32_t a[4];
        ldsm_x4(a, &X[m0 + (lane & 7) + ((lane >> 3) & 1) * 8][kk * 16 + (lane >> 4) * 8]);
#pragma unroll
        for (int np = 0; np < NT; ++np) {
            uint32_t b[4];
            ldsm_x4(b, &Y[np * 16 + (lane & 7) + (lane >> 4) * 8][kk * 16 + ((lane >> 3) & 1) * 8]);
            mma16816(s[2 * np], a, b[0], b[1]);
            mma16816(s[2 * np + 1], a, b[2], b[3]);
        }
    }
}

// acc[DH/8][4] (16 rows x DH) += P(16 x 64, given as A fragments per 16-key tile) * Z[:, :DH], Z row-major [key][d]
template <int DH, int NT>
__device__ __forceinline__ void frag_times_rows(float acc[DH / 8][4], const uint32_t pa[NT][4], bf16 (*Z)[DH + 8],
                                                int lane) {
#pragma unroll
    for (int kk = 0; kk < NT; ++kk) {
#pragma unroll
        for (int dp = 0; dp < DH / 16; ++dp) {
            uint32_t b[4];
            ldsm_x4_t(b, &Z[kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][dp * 16 + (lane >> 4) * 8]);
            mma16816(acc[2 * dp], pa[kk], b[0], b[1]);
            mma16816(acc[2 * dp + 1], pa[kk], b[2], b[3]);
        }
    }
}

// acc (16 rows j0.. x DH) += W^T[j0:j0+16, :] * Z, with W stored [q][key] (pitch 72) and Z stored [q][d]
template <int DH, int NT>
__device__ __forceinline__ void transposed_times_rows(float acc[DH / 8][4], bf16 (*W)[NMAX + 8], bf16 (*Z)[DH + 8],
                                                      int j0, int lane) {
#pragma unroll
    for (int kk = 0; kk < NT; ++kk) {
        uint32_t a[4];
        const int mi = lane >> 3;
        ldsm_x4_t(a, &W[kk * 16 + (lane & 7) + (mi >> 1) * 8][j0 + (mi & 1) * 8]);
#pragma unroll
        for (int dp = 0; dp < DH / 16; ++dp) {
            uint32_t b[4];
            ldsm_x4_t(b, &Z[kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][dp * 16 + (lane >> 4) * 8]);
            mma16816(acc[2 * dp], a, b[0], b[1]);
            mma16816(acc[2 * dp + 1], a, b[2], b[3]);
        }
    }
}

template <int DH>
__device__ __forceinline__ void store_rows(bf16 *dst, int ld, float acc[DH / 8][4], int r0, int r1, int N, int t,
                                           float mul0, float mul1) {
    bf16 *p0 = dst + r0 * ld + 2 * t, *p1 = dst + r1 * ld + 2 * t;
#pragma unroll
    for (int j = 0; j < DH / 8; ++j) {
        if (r0 < N) *reinterpret_cast<uint32_t *>(p0 + 8 * j) = pack_bf16x2(acc[j][0] * mul0, acc[j][1] * mul0);
        if (r1 < N) *reinterpret_cast<uint32_t *>(p1 + 8 * j) = pack_bf16x2(acc[j][2] * mul1, acc[j][3] * mul1);
    }
}

