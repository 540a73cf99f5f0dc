// CLS pool + mlp_head (LayerNorm + Linear(d -> n_class)) + BCE-with-logits loss, forward and backward.
// Replaces x[:, 0] -> mlp_head (vit_pytorch ViT.forward) and nn.BCEWithLogitsLoss
// (/root/reference/ecg_transformer/models/ecg_vit.py:118,148).  B x n_class is tiny (256 x 71): FFMA + shuffles.
#include "common.cuh"

namespace ecgvit {

namespace {

constexpr int HEAD_MAXV = 4;  // d <= 1024

// one CTA (4 warps) per sample: every warp LayerNorms the CLS row redundantly (768 elements, cheaper than a barrier +
// smem round trip), then the warps split the n_class dot products
template <typename T>
__global__ void __launch_bounds__(128) head_fwd_kernel(const T *__restrict__ tok, const float *__restrict__ gamma,
                                                        const float *__restrict__ beta, const float *__restrict__ w,
                                                        const float *__restrict__ bias, float *__restrict__ xn,
                                                        float *__restrict__ mean_out, float *__restrict__ rstd_out,
                                                        float *__restrict__ logits, int B, int N, int d, int n_class,
                                                        float eps) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int b = blockIdx.x;
    if (b >= B) return;
    const T *xr = tok + (int64_t)b * N * d;
    float v[HEAD_MAXV][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < HEAD_MAXV; ++i) {
        const int c = (i * 32 + lane) * 8;
        if (c < d) {
            load8(xr + c, v[i]);
#pragma unroll
            for (int k = 0; k < 8; ++k) s += v[i][k];
        }
    }
    const float mean = warp_sum(s) / (float)d;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < HEAD_MAXV; ++i) {
        const int c = (i * 32 + lane) * 8;
        if (c < d) {
#pragma unroll
            for (int k = 0; k < 8; ++k) { const float t = v[i][k] - mean; sq += t * t; }
        }
    }
    const float rstd = rsqrtf(warp_sum(sq) / (float)d + eps);
#pragma unroll
    for (int i = 0; i < HEAD_MAXV; ++i) {
        const int c = (i * 32 + lane) * 8;
        if (c < d) {
            float g[8], bt[8];
            load8(gamma + c, g);
            load8(beta + c, bt);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[i][k] = fmaf((v[i][k] - mean) * rstd, g[k], bt[k]);
            if (warp == 0) store8(xn + (int64_t)b * d + c, v[i]);
        }
    }
    if (warp == 0 && lane == 0) { mean_out[b] = mean; rstd_out[b] = rstd; }
    for (int cls = warp; cls < n_class; cls += nwarp) {
        const float *wr = w + (int64_t)cls * d;
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < HEAD_MAXV; ++i) {
            const int c = (i * 32 + lane) * 8;
            if (c < d) {
                float wv[8];
                load8(wr + c, wv);
#pragma unroll
                for (int k = 0; k < 8; ++k) acc = fmaf(v[i][k], wv[k], acc);
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) logits[(int64_t)b * n_class + cls] = acc + bias[cls];
    }
}

__device__ __forceinline__ float bce_with_logits(float z, float y) {
    // max(z,0) - z*y + log1p(exp(-|z|))   (ATen's stable form)
    return fmaxf(z, 0.f) - z * y + log1pf(expf(-fabsf(z)));
}

// `weight[labels.long()]` of ecg_vit.py:144-147: the element weight is looked up by the integer value of the label
__device__ __forceinline__ float label_weight(const float *__restrict__ table, int n_weight, float y) {
    if (table == nullptr) return 1.f;
    const int i = min(max(static_cast<int>(y), 0), n_weight - 1);
    return __ldg(table + i);
}

// single CTA, deterministic tree reduction
__global__ void __launch_bounds__(1024) bce_loss_kernel(const float *__restrict__ logits,
                                                         const float *__restrict__ labels, float *__restrict__ loss,
                                                         int n, int reduction, const float *__restrict__ loss_weight,
                                                         int n_weight) {
    __shared__ float red[32];
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float l = label_weight(loss_weight, n_weight, labels[i]) * bce_with_logits(logits[i], labels[i]);
        if (reduction == ECGVIT_REDUCTION_NONE) loss[i] = l;
        s += l;
    }
    if (reduction == ECGVIT_REDUCTION_NONE) return;
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
        t = warp_sum(t);
        if (threadIdx.x == 0) loss[0] = (reduction == ECGVIT_REDUCTION_MEAN) ? t / (float)n : t;
    }
}

// backward A: per sample, dlogits -> dxn = dlogits W -> LayerNorm' -> CLS row of dtok
template <typename T, typename TX = T>
__global__ void __launch_bounds__(128) head_bwd_rows_kernel(const TX *__restrict__ tok, const float *__restrict__ gamma,
                                                             const float *__restrict__ w,
                                                             const float *__restrict__ labels,
                                                             const float *__restrict__ mean_in,
                                                             const float *__restrict__ rstd_in,
                                                             const float *__restrict__ logits, T *__restrict__ dtok,
                                                             float *__restrict__ dxn_out, float *__restrict__ dlog_out,
                                                             int B, int N, int d, int n_class, float coef_host,
                                                             const float *__restrict__ coef_dev,
                                                             const float *__restrict__ loss_weight, int n_weight) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    // upstream gradient of the loss: a host scalar and / or a DEVICE scalar (autograd's grad_output: no host sync)
    const float coef = coef_dev != nullptr ? coef_host * __ldg(coef_dev) : coef_host;
    float dxn[HEAD_MAXV][8];
#pragma unroll
    for (int i = 0; i < HEAD_MAXV; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) dxn[i][k] = 0.f;
    for (int cls = 0; cls < n_class; ++cls) {
        const float z = logits[(int64_t)b * n_class + cls], y = labels[(int64_t)b * n_class + cls];
        const float dl = coef * label_weight(loss_weight, n_weight, y) * (1.0f / (1.0f + expf(-z)) - y);
        if (lane == 0) dlog_out[(int64_t)b * n_class + cls] = dl;
        const float *wr = w + (int64_t)cls * d;
#pragma unroll
        for (int i = 0; i < HEAD_MAXV; ++i) {
            const int c = (i * 32 + lane) * 8;
            if (c < d) {
                float wv[8];
                load8(wr + c, wv);
#pragma unroll
                for (int k = 0; k < 8; ++k) dxn[i][k] = fmaf(dl, wv[k], dxn[i][k]);
            }
        }
    }
    const float mean = mean_in[b], rstd = rstd_in[b];
    const TX *xr = tok + (int64_t)b * N * d;
    float g[HEAD_MAXV][8], xh[HEAD_MAXV][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < HEAD_MAXV; ++i) {
        const int c = (i * 32 + lane) * 8;
        if (c < d) {
            float xv[8], gm[8];
            load8(xr + c, xv);
            load8(gamma + c, gm);
            store8(dxn_out + (int64_t)b * d + c, dxn[i]);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                xh[i][k] = (xv[k] - mean) * rstd;
                g[i][k] = dxn[i][k] * gm[k];
                s1 += g[i][k];
                s2 = fmaf(g[i][k], xh[i][k], s2);
            }
        }
    }
    s1 = warp_sum(s1) / (float)d;
    s2 = warp_sum(s2) / (float)d;
    T *dr = dtok + (int64_t)b * N * d;
#pragma unroll
    for (int i = 0; i < HEAD_MAXV; ++i) {
        const int c = (i * 32 + lane) * 8;
        if (c < d) {
            float o[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = rstd * (g[i][k] - s1 - xh[i][k] * s2);
            store8(dr + c, o);
        }
    }
}

// backward B: per column k: dgamma, dbeta of the head LayerNorm and the column sum of the CLS-row gradients.
// block = 32 columns x 8 sample groups, smem tree at the end
template <typename T, typename TX = T>
__global__ void __launch_bounds__(256) head_bwd_cols_kernel(const TX *__restrict__ tok, const T *__restrict__ dtok,
                                                             const float *__restrict__ dxn,
                                                             const float *__restrict__ mean_in,
                                                             const float *__restrict__ rstd_in,
                                                             float *__restrict__ dgamma, float *__restrict__ dbeta,
                                                             float *__restrict__ dcolsum, int B, int N, int d) {
    __shared__ float red[3][8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int k = blockIdx.x * 32 + tx;
    float sg = 0.f, sb = 0.f, sc = 0.f;
    if (k < d) {
        for (int b = ty; b < B; b += 8) {
            const float xh = (to_f32(tok[(int64_t)b * N * d + k]) - mean_in[b]) * rstd_in[b];
            const float dv = dxn[(int64_t)b * d + k];
            sg = fmaf(dv, xh, sg);
            sb += dv;
            sc += to_f32(dtok[(int64_t)b * N * d + k]);
        }
    }
    red[0][ty][tx] = sg; red[1][ty][tx] = sb; red[2][ty][tx] = sc;
    __syncthreads();
    if (ty == 0 && k < d) {
        float a = 0.f, bsum = 0.f, c = 0.f;
#pragma unroll
        for (int y = 0; y < 8; ++y) { a += red[0][y][tx]; bsum += red[1][y][tx]; c += red[2][y][tx]; }
        dgamma[k] += a;
        dbeta[k] += bsum;
        if (dcolsum != nullptr) dcolsum[k] += c;
    }
}

// backward C: dW[c,k] += sum_b dlogits[b,c] xn[b,k];  db[c] += sum_b dlogits[b,c]
// grid (d/32, n_class), block = 32 columns x 8 sample groups
__global__ void __launch_bounds__(256) head_bwd_weight_kernel(const float *__restrict__ dlog,
                                                               const float *__restrict__ xn, float *__restrict__ dw,
                                                               float *__restrict__ db, int B, int d, int n_class) {
    __shared__ float red[2][8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int cls = blockIdx.y;
    const int k = blockIdx.x * 32 + tx;
    float s = 0.f, sb = 0.f;
    if (k < d) {
        for (int b = ty; b < B; b += 8) {
            const float dl = dlog[(int64_t)b * n_class + cls];
            s = fmaf(dl, xn[(int64_t)b * d + k], s);
            sb += dl;
        }
    }
    red[0][ty][tx] = s; red[1][ty][tx] = sb;
    __syncthreads();
    if (ty == 0 && k < d) {
        float a = 0.f, bsum = 0.f;
#pragma unroll
        for (int y = 0; y < 8; ++y) { a += red[0][y][tx]; bsum += red[1][y][tx]; }
        dw[(int64_t)cls * d + k] += a;
        if (k == 0) db[cls] += bsum;
    }
}

}  // namespace
}  // namespace ecgvit

using namespace ecgvit;

extern "C" {

int ecgvit_head_fwd(const void *tok, const float *gamma, const float *beta, const float *w, const float *b,
                    const float *labels, const float *loss_weight, int n_weight, float *xn, float *mean, float *rstd,
                    float *logits, float *loss, int B, int N, int d, int n_class, int reduction, float eps, int dtype,
                    void *stream) {
    ECGVIT_REQUIRE(tok && gamma && beta && w && b && xn && mean && rstd && logits, "head_fwd: null argument");
    ECGVIT_REQUIRE(B > 0 && N > 0 && n_class > 0, "head_fwd: bad sizes");
    ECGVIT_REQUIRE(d % 8 == 0 && d <= 8 * 32 * HEAD_MAXV, "head_fwd: d=%d must be a multiple of 8 and <= %d", d,
                   8 * 32 * HEAD_MAXV);
    ECGVIT_REQUIRE(labels == nullptr || loss != nullptr, "head_fwd: labels given but loss is null");
    ECGVIT_REQUIRE(loss_weight == nullptr || n_weight >= 1, "head_fwd: loss_weight table needs n_weight >= 1");
    const int grid = B;
    if (dtype == ECGVIT_BF16)
        head_fwd_kernel<bf16><<<grid, 128, 0, as_stream(stream)>>>((const bf16 *)tok, gamma, beta, w, b, xn, mean, rstd, logits, B, N, d, n_class, eps);
    else if (dtype == ECGVIT_F32 || dtype == ECGVIT_BF16_RES32)   // fp32 tokens (parity mode / fp32 residual stream)
        head_fwd_kernel<float><<<grid, 128, 0, as_stream(stream)>>>((const float *)tok, gamma, beta, w, b, xn, mean, rstd, logits, B, N, d, n_class, eps);
    else return fail(-1, "head_fwd: unknown dtype %d", dtype);
    int rc = check_launch("head_fwd");
    if (rc) return rc;
    if (labels != nullptr) {
        bce_loss_kernel<<<1, 1024, 0, as_stream(stream)>>>(logits, labels, loss, B * n_class, reduction, loss_weight,
                                                              n_weight);
        rc = check_launch("bce_loss");
    }
    return rc;
}

int ecgvit_head_bwd(const void *tok, const float *gamma, const float *w, const float *labels, const float *loss_weight,
                    int n_weight, const float *xn, const float *mean, const float *rstd, const float *logits, void *dtok,
                    float *dw, float *db,
                    float *dgamma, float *dbeta, float *dcolsum, float *scratch, int B, int N, int d, int n_class,
                    int reduction, float grad_scale, const float *grad_scale_dev, int dtype, void *stream) {
    ECGVIT_REQUIRE(tok && gamma && w && labels && xn && mean && rstd && logits && dtok && dw && db && dgamma &&
                       dbeta && scratch,
                   "head_bwd: null argument");
    ECGVIT_REQUIRE(reduction == ECGVIT_REDUCTION_MEAN || reduction == ECGVIT_REDUCTION_SUM,
                   "head_bwd: reduction must be mean or sum");
    ECGVIT_REQUIRE(loss_weight == nullptr || n_weight >= 1, "head_bwd: loss_weight table needs n_weight >= 1");
    ECGVIT_REQUIRE(d % 8 == 0 && d <= 8 * 32 * HEAD_MAXV, "head_bwd: d=%d must be a multiple of 8 and <= %d", d,
                   8 * 32 * HEAD_MAXV);
    cudaStream_t s = as_stream(stream);
    const size_t esz = dtype == ECGVIT_F32 ? 4 : 2;   // dtok is bf16 in both bf16 modes
    cudaError_t e = cudaMemsetAsync(dtok, 0, (size_t)B * N * d * esz, s);
    if (e != cudaSuccess) return fail((int)e, "head_bwd: memset: %s", cudaGetErrorString(e));
    const float coef = grad_scale * (reduction == ECGVIT_REDUCTION_MEAN ? 1.0f / ((float)B * (float)n_class) : 1.0f);
    float *dxn = scratch, *dlog = scratch + (size_t)B * d;
    const int grid = (B + 3) / 4;
    if (dtype == ECGVIT_BF16) {
        head_bwd_rows_kernel<bf16><<<grid, 128, 0, s>>>((const bf16 *)tok, gamma, w, labels, mean, rstd, logits, (bf16 *)dtok, dxn, dlog, B, N, d, n_class, coef, grad_scale_dev, loss_weight, n_weight);
        head_bwd_cols_kernel<bf16><<<(d + 31) / 32, 256, 0, s>>>((const bf16 *)tok, (const bf16 *)dtok, dxn, mean, rstd, dgamma, dbeta, dcolsum, B, N, d);
    } else if (dtype == ECGVIT_BF16_RES32) {
        head_bwd_rows_kernel<bf16, float><<<grid, 128, 0, s>>>((const float *)tok, gamma, w, labels, mean, rstd, logits, (bf16 *)dtok, dxn, dlog, B, N, d, n_class, coef, grad_scale_dev, loss_weight, n_weight);
        head_bwd_cols_kernel<bf16, float><<<(d + 31) / 32, 256, 0, s>>>((const float *)tok, (const bf16 *)dtok, dxn, mean, rstd, dgamma, dbeta, dcolsum, B, N, d);
    } else if (dtype == ECGVIT_F32) {
        head_bwd_rows_kernel<float><<<grid, 128, 0, s>>>((const float *)tok, gamma, w, labels, mean, rstd, logits, (float *)dtok, dxn, dlog, B, N, d, n_class, coef, grad_scale_dev, loss_weight, n_weight);
        head_bwd_cols_kernel<float><<<(d + 31) / 32, 256, 0, s>>>((const float *)tok, (const float *)dtok, dxn, mean, rstd, dgamma, dbeta, dcolsum, B, N, d);
    } else return fail(-1, "head_bwd: unknown dtype %d", dtype);
    dim3 gw((d + 31) / 32, n_class);
    head_bwd_weight_kernel<<<gw, 256, 0, s>>>(dlog, xn, dw, db, B, d, n_class);
    return check_launch("head_bwd");
}

}  // extern "C"
