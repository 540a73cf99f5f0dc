// Attention front door: picks the kernel for (dtype, N, dh).
#include "common.cuh"

namespace ecgvit {
int attention_fwd_simt(const void *qkv, void *o, float *lse, int B, int N, int H, int dh, float scale, int dtype,
                       DropoutParams drop, cudaStream_t stream);
int attention_bwd_simt(const void *qkv, const void *o, const void *d_o, const float *lse, void *dqkv, int B, int N,
                       int H, int dh, float scale, int dtype, DropoutParams drop, cudaStream_t stream);
int attention_fwd_mma(const void *qkv, void *o, float *lse, int B, int N, int H, int dh, float scale,
                      DropoutParams drop, cudaStream_t stream);
int attention_bwd_mma(const void *qkv, const void *o, const void *d_o, const float *lse, void *dqkv, int B, int N,
                      int H, int dh, float scale, DropoutParams drop, cudaStream_t stream);
bool attention_mma_supported(int N, int dh);
}  // namespace ecgvit

using namespace ecgvit;

extern "C" {

int ecgvit_attention_fwd(const void *qkv, void *o, float *lse, int B, int N, int H, int dh, float scale,
                         float dropout_p, int dropout_stream, const uint32_t *dropout_seed, int dtype, void *stream) {
    const DropoutParams drop = make_dropout(dropout_p, dropout_stream, dropout_seed);
    ECGVIT_REQUIRE(qkv && o && lse && B > 0 && N > 0 && H > 0, "attention_fwd: bad arguments");
    ECGVIT_REQUIRE(dh % 8 == 0, "attention_fwd: head dim %d must be a multiple of 8", dh);
    ECGVIT_REQUIRE(dtype == ECGVIT_F32 || dtype == ECGVIT_BF16, "attention_fwd: unknown dtype %d", dtype);
    if (dtype == ECGVIT_BF16 && attention_mma_supported(N, dh))
        return attention_fwd_mma(qkv, o, lse, B, N, H, dh, scale, drop, as_stream(stream));
    return attention_fwd_simt(qkv, o, lse, B, N, H, dh, scale, dtype, drop, as_stream(stream));
}

int ecgvit_attention_bwd(const void *qkv, const void *o, const void *d_o, const float *lse, void *dqkv, int B,
                         int N, int H, int dh, float scale, float dropout_p, int dropout_stream,
                         const uint32_t *dropout_seed, int dtype, void *stream) {
    const DropoutParams drop = make_dropout(dropout_p, dropout_stream, dropout_seed);
    ECGVIT_REQUIRE(qkv && o && d_o && lse && dqkv && B > 0 && N > 0 && H > 0, "attention_bwd: bad arguments");
    ECGVIT_REQUIRE(dh % 8 == 0, "attention_bwd: head dim %d must be a multiple of 8", dh);
    ECGVIT_REQUIRE(dtype == ECGVIT_F32 || dtype == ECGVIT_BF16, "attention_bwd: unknown dtype %d", dtype);
    if (dtype == ECGVIT_BF16 && attention_mma_supported(N, dh))
        return attention_bwd_mma(qkv, o, d_o, lse, dqkv, B, N, H, dh, scale, drop, as_stream(stream));
    return attention_bwd_simt(qkv, o, d_o, lse, dqkv, B, N, H, dh, scale, dtype, drop, as_stream(stream));
}

}  // extern "C"
