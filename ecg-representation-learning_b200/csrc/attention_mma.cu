// placeholder until the tensor-core attention kernel lands
#include "common.cuh"
namespace ecgvit {
bool attention_mma_supported(int, int) { return false; }
int attention_fwd_mma(const void *, void *, float *, int, int, int, int, float, cudaStream_t) { return fail(-1, "attention_mma: not built"); }
int attention_bwd_mma(const void *, const void *, const void *, const float *, void *, int, int, int, int, float, cudaStream_t) { return fail(-1, "attention_mma: not built"); }
}
