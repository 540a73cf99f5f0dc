"""
ORACLE (test infrastructure) -- CPU restatement of the reference's evaluation metrics.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this file.

Follows `/root/reference/ecg_transformer/util/train.py:12-56 get_accuracy` with the sklearn calls it makes written out
(scikit-learn is third party; `roc_auc_score(average=None)` per column = area under the ROC curve = Mann-Whitney U with
half credit for ties; `accuracy_score`, `balanced_accuracy_score`, `classification_report(...)[..]['recall']`).

PINNED: tests/golden/eval_metrics.npz holds the outputs of the reference's own `get_accuracy` (sklearn underneath,
tests/golden/make_golden_metrics.py); tests/test_oracle.py compares.
"""
import numpy as np


def auroc(scores, y):
    """exact Mann-Whitney: O(n+ n-) comparisons"""
    pos, neg = scores[y == 1], scores[y != 1]
    gt = (pos[:, None] > neg[None, :]).sum(dtype=np.int64)
    eq = (pos[:, None] == neg[None, :]).sum(dtype=np.int64)
    return (gt + 0.5 * eq) / (len(pos) * len(neg))


def get_accuracy(preds, labels):
    """preds, labels: [n, n_class] numpy fp32.  Returns (scalars[5], per_class_auc[n_class] with NaN where undefined)"""
    preds_bin = (preds >= 0.5).astype(np.float32)
    n_class = preds.shape[1]
    per_class = np.full(n_class, np.nan)
    two = np.any(labels != labels[0], axis=0)                      # util/train.py:29
    for c in np.nonzero(two)[0]:
        per_class[c] = auroc(preds[:, c], labels[:, c])
    macro = np.nanmean(per_class) if two.any() else np.nan
    p, y = preds_bin.flatten(), labels.flatten()
    tp, fp = np.sum((p == 1) & (y == 1)), np.sum((p == 1) & (y != 1))
    tn, fn = np.sum((p == 0) & (y != 1)), np.sum((p == 0) & (y == 1))
    acc = (tp + tn) / p.size                                       # accuracy_score(labels, preds_bin)
    recalls = [r for r, d in ((tp / max(tp + fn, 1), tp + fn), (tn / max(tn + fp, 1), tn + fp)) if d > 0]
    bal = float(np.mean(recalls))                                  # balanced_accuracy_score(labels, preds_bin)
    # classification_report(preds_bin, labels): y_true = preds_bin, y_pred = labels; zero_division = 0 (:47-49)
    rec_of_neg = tn / (tn + fn) if (tn + fn) > 0 else 0.0          # report['neg']['recall'] -> `rec_pos` (:50)
    rec_of_pos = tp / (tp + fp) if (tp + fp) > 0 else 0.0          # report['pos']['recall'] -> `rec_neg` (:50)
    return np.array([acc, bal, rec_of_pos, rec_of_neg, macro], dtype=np.float64), per_class
