"""Generates tests/golden/eval_metrics.npz by running the REFERENCE's `get_accuracy`
(/root/reference/ecg_transformer/util/train.py:12-56, sklearn underneath) on seeded logits / labels.
Run in the build container only:  python tests/golden/make_golden_metrics.py"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

ref_shim.install()
U = importlib.import_module('ecg_transformer.util.train')
id2code = importlib.import_module('ecg_transformer.util').config('datasets.PTB-XL.code.id2code')
codes = [id2code[i] for i in range(71)]

out = {}
g = torch.Generator().manual_seed(77)
for name, (n, p_pos, quant) in dict(sparse=(300, 3 / 71, None), ties=(257, 0.2, 8), tiny=(5, 0.3, None)).items():
    logits = torch.randn(n, 71, generator=g) * 2
    if quant:  # many exactly equal scores
        logits = torch.round(logits * quant) / quant
    labels = (torch.rand(n, 71, generator=g) < p_pos).float()
    labels[:, 5] = 0   # a class with no positives
    labels[:, 9] = 1   # a class with no negatives
    preds = torch.sigmoid(logits)
    d = U.get_accuracy(preds, labels)
    out[f'{name}/logits'], out[f'{name}/labels'], out[f'{name}/preds'] = logits.numpy(), labels.numpy(), preds.numpy()
    out[f'{name}/scalars'] = np.array([d['binary_accuracy'], d['weighted_binary_accuracy'], d['binary_negative_recall'],
                                       d['binary_positive_recall'], d['macro_auc']], dtype=np.float64)
    auc = np.full(71, np.nan)
    for i, c in enumerate(codes):
        if c in d['per_class_auc']:
            auc[i] = d['per_class_auc'][c]
    out[f'{name}/per_class_auc'] = auc
    print(name, out[f'{name}/scalars'], int(np.isfinite(auc).sum()), 'valid classes')
np.savez_compressed(os.path.join(ROOT, 'tests/golden/eval_metrics.npz'), **out)
