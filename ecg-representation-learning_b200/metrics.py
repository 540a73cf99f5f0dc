"""
Evaluation on the device: the forward-only pass of `MyTrainer.evaluate`
(`ecg_transformer/models/train.py:321-378`) and the metrics of `ecg_transformer/util/train.py:12-56 get_accuracy`,
without moving logits / labels to the host and without sklearn: one kernel family (`csrc/metrics.cu`) counts the
confusion matrix and the exact Mann-Whitney pairs of every class, and a single 8 * (6 + n_class)-byte read returns all
numbers.
"""
import math

import torch

from . import _lib


def get_accuracy(preds, labels, return_auc=True, id2code=None):
    """Same contract as the reference's `get_accuracy(preds, labels, return_auc)`: `preds` are per-class probabilities
    (the caller applies `torch.sigmoid`, train.py:366), `labels` the multi-hot ground truth, both [n, n_class] CUDA
    tensors.  Returns the same dict keys.  `id2code` maps class index -> name for `per_class_auc` (the reference reads
    `config('datasets.PTB-XL.code.id2code')`, which is data and not shipped here); indices are used when omitted.

    Reference quirks kept on purpose: the `classification_report` call swaps y_true / y_pred and the two recalls are
    then bound to the opposite names (util/train.py:47-55), so `binary_negative_recall` is TP / (TP + FP) and
    `binary_positive_recall` is TN / (TN + FN)."""
    if not preds.is_cuda:
        raise RuntimeError('ecg_b200.metrics.get_accuracy runs on the GPU: pass CUDA tensors (there is no CPU path)')
    lib = _lib.load()
    preds = preds.float().contiguous()
    labels = labels.to(device=preds.device, dtype=torch.float32).contiguous()
    assert preds.dim() == 2 and preds.shape == labels.shape
    n, n_class = preds.shape
    scratch = torch.empty(int(lib.ecgvit_eval_metrics_scratch_bytes(n_class)), device=preds.device, dtype=torch.uint8)
    out = torch.empty(6 + n_class, device=preds.device, dtype=torch.float64)
    _lib.check(lib.ecgvit_eval_metrics(preds.data_ptr(), labels.data_ptr(), n, n_class, 1 if return_auc else 0,
                                       scratch.data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream),
               'eval_metrics')
    vals = out.tolist()  # the one device -> host read
    macro_auc, code2auroc = None, None
    if return_auc and vals[5] > 0:
        name = (lambda i: id2code[i]) if id2code is not None else (lambda i: i)
        code2auroc = {name(c): vals[6 + c] for c in range(n_class) if not math.isnan(vals[6 + c])}
        macro_auc = vals[4]
    return dict(binary_accuracy=vals[0], weighted_binary_accuracy=vals[1], binary_negative_recall=vals[2],
                binary_positive_recall=vals[3], macro_auc=macro_auc, per_class_auc=code2auroc)


@torch.no_grad()
def evaluate(model, batches, loss_reduction='mean', return_predictions=False, id2code=None):
    """`MyTrainer.evaluate` (train.py:321-378) for an iterable of dict(sample_values=, labels=) batches: eval mode,
    forward only, `loss_reduction` 'mean' (mean of per-batch means) or 'none' (per-sample loss = mean over classes),
    metrics over the concatenated logits.  Everything stays on the device until the final read."""
    assert loss_reduction in ('mean', 'none')
    red_ori, training = model.loss_reduction, model.training
    model.loss_reduction = loss_reduction
    model.eval()
    losses, logits, labels = [], [], []
    for inputs in batches:
        x, y = inputs['sample_values'].cuda(non_blocking=True), inputs['labels'].cuda(non_blocking=True)
        output = model(sample_values=x, labels=y)
        losses.append(output.loss.reshape(1) if loss_reduction == 'mean' else output.loss.mean(dim=-1))
        logits.append(output.logits)
        labels.append(y)
    logits, labels = torch.cat(logits, dim=0), torch.cat(labels, dim=0)
    loss_all = torch.cat(losses)
    d_log = {'eval/loss': loss_all.mean().item() if loss_reduction == 'mean' else loss_all.cpu().numpy()}
    d_log.update({f'eval/{k}': v for k, v in get_accuracy(torch.sigmoid(logits), labels, id2code=id2code).items()})
    model.loss_reduction = red_ori
    model.train(training)
    return dict(metrics=d_log, predictions=dict(labels=labels, logits=logits)) if return_predictions else d_log
