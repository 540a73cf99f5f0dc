// Evaluation metrics of the reference's eval loop on the device: replaces the sklearn calls of
// ecg_transformer/util/train.py:12-56 `get_accuracy` (roc_auc_score per class, accuracy / balanced accuracy /
// classification_report over all B x n_class decisions) that force logits and labels back to the host.
//
// AUROC of class c is the Mann-Whitney statistic, counted exactly in integers:
//     U_c = #{(i, j): y_i = 1, y_j = 0, p_i > p_j} + 1/2 #{(i, j): y_i = 1, y_j = 0, p_i = p_j},  auc_c = U_c / (n+ n-)
// which is what sklearn's trapezoid over the ROC curve evaluates to (ties share a diagonal segment).  n+ is small
// (multi-hot labels, ~3 of 71 set), so the n+ x n- comparisons are cheaper than 71 sorts and need no scratch.
#include "common.cuh"

namespace ecgvit {
namespace {

constexpr int MT = 256;  // rows per tile / threads per block

// counts[0..3] = TP, FP, TN, FN over every (row, class) with decision p >= 0.5; per class: n_pos at counts[4 + c]
__global__ void __launch_bounds__(MT) confusion_kernel(const float *__restrict__ preds, const float *__restrict__ labels,
                                                        int64_t n_rows, int n_class, unsigned long long *counts) {
    const int c = blockIdx.y;
    unsigned int tp = 0, fp = 0, tn = 0, fn = 0;
    for (int64_t r = (int64_t)blockIdx.x * MT + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * MT) {
        const bool hit = preds[r * n_class + c] >= 0.5f;
        const bool pos = labels[r * n_class + c] == 1.0f;  // sklearn compares the flattened float labels with {0, 1}
        tp += hit && pos; fp += hit && !pos; tn += !hit && !pos; fn += !hit && pos;
    }
    tp = __reduce_add_sync(0xffffffffu, tp); fp = __reduce_add_sync(0xffffffffu, fp);
    tn = __reduce_add_sync(0xffffffffu, tn); fn = __reduce_add_sync(0xffffffffu, fn);
    if ((threadIdx.x & 31) == 0) {  // integer atomics: the totals do not depend on arrival order
        atomicAdd(&counts[0], (unsigned long long)tp);
        atomicAdd(&counts[1], (unsigned long long)fp);
        atomicAdd(&counts[2], (unsigned long long)tn);
        atomicAdd(&counts[3], (unsigned long long)fn);
        atomicAdd(&counts[4 + c], (unsigned long long)(tp + fn));
    }
}

// grid (row tile j, class c): thread = one row j; walks all rows i in smem tiles and, for the positive ones, counts
// p_i > p_j and p_i == p_j when row j is a negative.  pairs[2c] += greater, pairs[2c + 1] += equal.
__global__ void __launch_bounds__(MT) auroc_pairs_kernel(const float *__restrict__ preds,
                                                          const float *__restrict__ labels, int64_t n_rows, int n_class,
                                                          unsigned long long *pairs) {
    __shared__ float s_p[MT];
    __shared__ unsigned char s_y[MT];
    const int c = blockIdx.y;
    const int64_t j = (int64_t)blockIdx.x * MT + threadIdx.x;
    const float pj = j < n_rows ? preds[j * n_class + c] : 0.f;
    // "positive" = label 1 (the reference feeds float multi-hot labels; any other value counts as negative, as in
    // sklearn's binary roc_auc_score with labels {0, 1})
    const bool is_neg = j < n_rows && labels[j * n_class + c] != 1.0f;
    unsigned int gt = 0, eq = 0;
    for (int64_t i0 = 0; i0 < n_rows; i0 += MT) {
        const int64_t i = i0 + threadIdx.x;
        __syncthreads();
        s_p[threadIdx.x] = i < n_rows ? preds[i * n_class + c] : 0.f;
        s_y[threadIdx.x] = (i < n_rows && labels[i * n_class + c] == 1.0f) ? 1 : 0;
        __syncthreads();
        const int lim = (int)min((int64_t)MT, n_rows - i0);
        if (is_neg) {
            for (int k = 0; k < lim; ++k) {
                if (s_y[k]) {  // block-uniform branch
                    const float pi = s_p[k];
                    gt += pi > pj;
                    eq += pi == pj;
                }
            }
        }
    }
    gt = __reduce_add_sync(0xffffffffu, gt);
    eq = __reduce_add_sync(0xffffffffu, eq);
    if ((threadIdx.x & 31) == 0 && (gt | eq)) {
        atomicAdd(&pairs[2 * c], (unsigned long long)gt);
        atomicAdd(&pairs[2 * c + 1], (unsigned long long)eq);
    }
}

// out[0..3] = binary_accuracy, weighted_binary_accuracy, binary_negative_recall, binary_positive_recall (the names and
// the swapped arguments of util/train.py:44-55 are kept), out[4] = macro_auc (NaN when no class has both labels),
// out[5] = number of classes with both labels, out[6 + c] = auc of class c (NaN when undefined)
__global__ void metrics_finalize_kernel(const unsigned long long *__restrict__ counts,
                                        const unsigned long long *__restrict__ pairs, int64_t n_rows, int n_class,
                                        double *__restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double tp = (double)counts[0], fp = (double)counts[1], tn = (double)counts[2], fn = (double)counts[3];
    const double total = tp + fp + tn + fn;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    out[0] = (tp + tn) / total;  // accuracy_score(labels, preds_bin)
    // balanced_accuracy_score(labels, preds_bin): mean recall over the classes present in the labels
    double bal = 0.0;
    int present = 0;
    if (tp + fn > 0) { bal += tp / (tp + fn); ++present; }
    if (tn + fp > 0) { bal += tn / (tn + fp); ++present; }
    out[1] = present ? bal / present : nan;
    // classification_report(preds_bin, labels): y_true = preds_bin, y_pred = labels (util/train.py:47-49), and
    // `rec_pos, rec_neg = [report[k]['recall'] for k in ['neg', 'pos']]` (:50); zero_division = 0
    const double recall_of_pred_neg = (tn + fn) > 0 ? tn / (tn + fn) : 0.0;  // report['neg']['recall'] -> rec_pos
    const double recall_of_pred_pos = (tp + fp) > 0 ? tp / (tp + fp) : 0.0;  // report['pos']['recall'] -> rec_neg
    out[2] = recall_of_pred_pos;  // binary_negative_recall = rec_neg
    out[3] = recall_of_pred_neg;  // binary_positive_recall = rec_pos
    double sum = 0.0;
    int valid = 0;
    for (int c = 0; c < n_class; ++c) {
        const double n_pos = (double)counts[4 + c], n_neg = (double)n_rows - n_pos;
        if (n_pos > 0 && n_neg > 0) {  // torch.any(labels != labels[0], dim=0) for {0, 1} labels (util/train.py:29)
            const double auc = ((double)pairs[2 * c] + 0.5 * (double)pairs[2 * c + 1]) / (n_pos * n_neg);
            out[6 + c] = auc;
            sum += auc;
            ++valid;
        } else {
            out[6 + c] = nan;
        }
    }
    out[4] = valid ? sum / valid : nan;
    out[5] = (double)valid;
}

}  // namespace
}  // namespace ecgvit

using namespace ecgvit;

extern "C" {

int64_t ecgvit_eval_metrics_scratch_bytes(int n_class) { return (int64_t)(4 + 3 * n_class) * 8; }

int ecgvit_eval_metrics(const float *preds, const float *labels, int64_t n_rows, int n_class, int with_auc,
                        void *scratch, double *out, void *stream) {
    ECGVIT_REQUIRE(preds && labels && scratch && out && n_rows > 0 && n_class > 0, "eval_metrics: bad arguments");
    ECGVIT_REQUIRE(n_rows < (int64_t)1 << 31, "eval_metrics: %lld rows overflow the per-thread pair counters",
                   (long long)n_rows);
    cudaStream_t st = as_stream(stream);
    unsigned long long *counts = reinterpret_cast<unsigned long long *>(scratch);
    unsigned long long *pairs = counts + 4 + n_class;
    cudaError_t e = cudaMemsetAsync(scratch, 0, (size_t)ecgvit_eval_metrics_scratch_bytes(n_class), st);
    if (e != cudaSuccess) return fail((int)e, "eval_metrics: memset: %s", cudaGetErrorString(e));
    const int tiles = (int)((n_rows + MT - 1) / MT);
    confusion_kernel<<<dim3(tiles < 64 ? tiles : 64, n_class), MT, 0, st>>>(preds, labels, n_rows, n_class, counts);
    int rc = check_launch("eval_metrics_confusion");
    if (rc) return rc;
    if (with_auc) {
        auroc_pairs_kernel<<<dim3(tiles, n_class), MT, 0, st>>>(preds, labels, n_rows, n_class, pairs);
        rc = check_launch("eval_metrics_auroc");
        if (rc) return rc;
    }
    metrics_finalize_kernel<<<1, 32, 0, st>>>(counts, pairs, n_rows, n_class, out);
    return check_launch("eval_metrics_finalize");
}

}  // extern "C"
