// Whole-sequence softmax attention per (batch, head) CTA, fp32 math on the FFMA pipe.
// Used for the fp32 parity mode (any head dim) and for bf16 shapes the tensor-core kernel does not cover.
// Reads the packed projection qkv[B*N, 3*H*dh] (q | k | v, head-major) in place and writes o[B*N, H*dh]:
// no [B,H,N,N] tensor and no permuted copies are materialised (vit_pytorch Attention.forward does both).
#include "common.cuh"

namespace ecgvit {

namespace {

template <typename T>
__device__ __forceinline__ void load_head_tile(float *dst, const T *src, int64_t row_stride, int N, int dh, int pitch,
                                               float mul) {
    // src points at (row 0, first column of this head); vectorised by 8 along dh
    const int vec = dh / 8;
    for (int i = threadIdx.x; i < N * vec; i += blockDim.x) {
        const int r = i / vec, c = (i - r * vec) * 8;
        float v[8];
        load8(src + r * row_stride + c, v);
#pragma unroll
        for (int k = 0; k < 8; ++k) dst[r * pitch + c + k] = v[k] * mul;
    }
}

template <typename T>
__global__ void __launch_bounds__(128) attention_fwd_simt_kernel(const T *__restrict__ qkv, T *__restrict__ o,
                                                                  float *__restrict__ lse, int N, int H, int dh,
                                                                  float scale, DropoutParams drop) {
    extern __shared__ float sm[];
    const int pitch = dh + 1, spitch = N + 1;
    float *Qs = sm, *Ks = Qs + N * pitch, *Vs = Ks + N * pitch, *S = Vs + N * pitch;
    const int h = blockIdx.x % H, b = blockIdx.x / H;
    const int inner = H * dh;
    const int64_t ld = 3 * (int64_t)inner;
    const T *base = qkv + (int64_t)b * N * ld + (int64_t)h * dh;
    load_head_tile(Qs, base, ld, N, dh, pitch, scale);
    load_head_tile(Ks, base + inner, ld, N, dh, pitch, 1.f);
    load_head_tile(Vs, base + 2 * inner, ld, N, dh, pitch, 1.f);
    __syncthreads();
    for (int idx = threadIdx.x; idx < N * N; idx += blockDim.x) {
        const int i = idx / N, j = idx - i * N;
        float s = 0.f;
        for (int d = 0; d < dh; ++d) s = fmaf(Qs[i * pitch + d], Ks[j * pitch + d], s);
        S[i * spitch + j] = s;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int i = warp; i < N; i += nwarp) {
        float m = -INFINITY;
        for (int j = lane; j < N; j += 32) m = fmaxf(m, S[i * spitch + j]);
        m = warp_max(m);
        float sum = 0.f;
        for (int j = lane; j < N; j += 32) {
            const float e = expf(S[i * spitch + j] - m);
            S[i * spitch + j] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        if (drop.threshold != 0) {
            // element index = ((b*H + h) * pitch + query) * pitch + key, pitch = N rounded up to 64
            const uint32_t seed = __ldg(drop.seed), np = (N + 63) / 64 * 64;
            for (int j = lane; j < N; j += 32)
                S[i * spitch + j] *= inv * dropout_one(drop, seed, (blockIdx.x * np + i) * np + j);
        } else {
            for (int j = lane; j < N; j += 32) S[i * spitch + j] *= inv;
        }
        if (lane == 0) lse[((int64_t)b * H + h) * N + i] = m + logf(sum);
    }
    __syncthreads();
    T *ob = o + (int64_t)b * N * inner + (int64_t)h * dh;
    for (int idx = threadIdx.x; idx < N * dh; idx += blockDim.x) {
        const int i = idx / dh, d = idx - i * dh;
        float acc = 0.f;
        for (int j = 0; j < N; ++j) acc = fmaf(S[i * spitch + j], Vs[j * pitch + d], acc);
        ob[(int64_t)i * inner + d] = from_f32<T>(acc);
    }
}

template <typename T>
__global__ void __launch_bounds__(128) attention_bwd_simt_kernel(const T *__restrict__ qkv, const T *__restrict__ o,
                                                                  const T *__restrict__ d_o,
                                                                  const float *__restrict__ lse, T *__restrict__ dqkv,
                                                                  int N, int H, int dh, float scale,
                                                                  DropoutParams drop) {
    extern __shared__ float sm[];
    const int pitch = dh + 1, spitch = N + 1;
    float *Qs = sm, *Ks = Qs + N * pitch, *Vs = Ks + N * pitch, *dOs = Vs + N * pitch;
    float *P = dOs + N * pitch, *Dv = P + N * spitch;  // Dv[N]
    const int h = blockIdx.x % H, b = blockIdx.x / H;
    const int inner = H * dh;
    const int64_t ld = 3 * (int64_t)inner;
    const T *base = qkv + (int64_t)b * N * ld + (int64_t)h * dh;
    const T *ob = o + (int64_t)b * N * inner + (int64_t)h * dh;
    const T *dob = d_o + (int64_t)b * N * inner + (int64_t)h * dh;
    load_head_tile(Qs, base, ld, N, dh, pitch, 1.f);
    load_head_tile(Ks, base + inner, ld, N, dh, pitch, 1.f);
    load_head_tile(Vs, base + 2 * inner, ld, N, dh, pitch, 1.f);
    load_head_tile(dOs, dob, (int64_t)inner, N, dh, pitch, 1.f);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    // D[i] = sum_d dO[i,d] * O[i,d]
    for (int i = warp; i < N; i += nwarp) {
        float s = 0.f;
        for (int d = lane; d < dh; d += 32) s = fmaf(dOs[i * pitch + d], to_f32(ob[(int64_t)i * inner + d]), s);
        s = warp_sum(s);
        if (lane == 0) Dv[i] = s;
    }
    // P = exp(scale * q k^T - lse)
    const float *lse_bh = lse + ((int64_t)b * H + h) * N;
    for (int idx = threadIdx.x; idx < N * N; idx += blockDim.x) {
        const int i = idx / N, j = idx - i * N;
        float s = 0.f;
        for (int d = 0; d < dh; ++d) s = fmaf(Qs[i * pitch + d], Ks[j * pitch + d], s);
        P[i * spitch + j] = expf(s * scale - lse_bh[i]);
    }
    __syncthreads();
    T *dq = dqkv + (int64_t)b * N * ld + (int64_t)h * dh;
    T *dk = dq + inner, *dv = dq + 2 * inner;
    const bool dropping = drop.threshold != 0;
    const uint32_t seed = dropping ? __ldg(drop.seed) : 0u, np = (N + 63) / 64 * 64;
    // dV[j,d] = sum_i Pdrop[i,j] dO[i,d]
    for (int idx = threadIdx.x; idx < N * dh; idx += blockDim.x) {
        const int j = idx / dh, d = idx - j * dh;
        float acc = 0.f;
        for (int i = 0; i < N; ++i) {
            const float mk = dropping ? dropout_one(drop, seed, (blockIdx.x * np + i) * np + j) : 1.f;
            acc = fmaf(P[i * spitch + j] * mk, dOs[i * pitch + d], acc);
        }
        dv[(int64_t)j * ld + d] = from_f32<T>(acc);
    }
    __syncthreads();
    // dS = P * (mask * dO V^T - D), in place over P
    for (int idx = threadIdx.x; idx < N * N; idx += blockDim.x) {
        const int i = idx / N, j = idx - i * N;
        float dp = 0.f;
        for (int d = 0; d < dh; ++d) dp = fmaf(dOs[i * pitch + d], Vs[j * pitch + d], dp);
        const float mk = dropping ? dropout_one(drop, seed, (blockIdx.x * np + i) * np + j) : 1.f;
        P[i * spitch + j] *= (dp * mk - Dv[i]);
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < N * dh; idx += blockDim.x) {
        const int i = idx / dh, d = idx - i * dh;
        float aq = 0.f, ak = 0.f;
        for (int j = 0; j < N; ++j) {
            aq = fmaf(P[i * spitch + j], Ks[j * pitch + d], aq);   // dQ[i,d] = sum_j dS[i,j] K[j,d]
            ak = fmaf(P[j * spitch + i], Qs[j * pitch + d], ak);   // dK[i,d] = sum_j dS[j,i] Q[j,d]
        }
        dq[(int64_t)i * ld + d] = from_f32<T>(aq * scale);
        dk[(int64_t)i * ld + d] = from_f32<T>(ak * scale);
    }
}

size_t fwd_smem(int N, int dh) { return sizeof(float) * (3 * (size_t)N * (dh + 1) + (size_t)N * (N + 1)); }
size_t bwd_smem(int N, int dh) { return sizeof(float) * (4 * (size_t)N * (dh + 1) + (size_t)N * (N + 1) + N); }

template <typename K> int set_smem(K kern, size_t bytes, const char *what) {
    if (bytes > 227 * 1024) return fail(-1, "%s: sequence too long for the whole-sequence kernel (%zu bytes of smem)", what, bytes);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return fail((int)e, "%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
    return 0;
}

}  // namespace

int attention_fwd_simt(const void *qkv, void *o, float *lse, int B, int N, int H, int dh, float scale, int dtype,
                       DropoutParams drop, cudaStream_t stream) {
    const size_t smem = fwd_smem(N, dh);
    int rc;
    if (dtype == ECGVIT_BF16) {
        if ((rc = set_smem(attention_fwd_simt_kernel<bf16>, smem, "attention_fwd"))) return rc;
        attention_fwd_simt_kernel<bf16><<<B * H, 128, smem, stream>>>((const bf16 *)qkv, (bf16 *)o, lse, N, H, dh, scale, drop);
    } else {
        if ((rc = set_smem(attention_fwd_simt_kernel<float>, smem, "attention_fwd"))) return rc;
        attention_fwd_simt_kernel<float><<<B * H, 128, smem, stream>>>((const float *)qkv, (float *)o, lse, N, H, dh, scale, drop);
    }
    return check_launch("attention_fwd_simt");
}

int attention_bwd_simt(const void *qkv, const void *o, const void *d_o, const float *lse, void *dqkv, int B, int N,
                       int H, int dh, float scale, int dtype, DropoutParams drop, cudaStream_t stream) {
    const size_t smem = bwd_smem(N, dh);
    int rc;
    if (dtype == ECGVIT_BF16) {
        if ((rc = set_smem(attention_bwd_simt_kernel<bf16>, smem, "attention_bwd"))) return rc;
        attention_bwd_simt_kernel<bf16><<<B * H, 128, smem, stream>>>((const bf16 *)qkv, (const bf16 *)o, (const bf16 *)d_o, lse, (bf16 *)dqkv, N, H, dh, scale, drop);
    } else {
        if ((rc = set_smem(attention_bwd_simt_kernel<float>, smem, "attention_bwd"))) return rc;
        attention_bwd_simt_kernel<float><<<B * H, 128, smem, stream>>>((const float *)qkv, (const float *)o, (const float *)d_o, lse, (float *)dqkv, N, H, dh, scale, drop);
    }
    return check_launch("attention_bwd_simt");
}

}  // namespace ecgvit
