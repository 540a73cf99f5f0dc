"""GPU tier: device evaluation metrics against what the reference's own `get_accuracy` (sklearn)
returned for the same probabilities / labels (tests/golden/eval_metrics.npz), and the eval loop against the oracle."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import ecg_b200
from ecg_b200 import EcgVit, EcgVitConfig
from oracle.ecg_vit_oracle import OracleConfig, OracleEcgVit, synthetic_batch
from oracle import metrics as oracle_metrics

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'eval_metrics.npz'))
KEYS = ['binary_accuracy', 'weighted_binary_accuracy', 'binary_negative_recall', 'binary_positive_recall', 'macro_auc']


@pytest.mark.parametrize('case', ['sparse', 'ties', 'tiny'])
def test_get_accuracy_matches_reference_vectors(case):
    preds, labels = torch.from_numpy(G[f'{case}/preds']).cuda(), torch.from_numpy(G[f'{case}/labels']).cuda()
    d = ecg_b200.get_accuracy(preds, labels)
    want = G[f'{case}/scalars']
    for k, w in zip(KEYS, want):
        assert abs(d[k] - w) <= 1e-12 * abs(w), (k, d[k], w)
    auc = G[f'{case}/per_class_auc']
    assert sorted(d['per_class_auc']) == [int(i) for i in np.nonzero(~np.isnan(auc))[0]]
    for c, v in d['per_class_auc'].items():
        assert abs(v - auc[c]) <= 1e-12, (c, v, auc[c])
    d2 = ecg_b200.get_accuracy(preds, labels, return_auc=False)
    assert d2['macro_auc'] is None and d2['per_class_auc'] is None and d2['binary_accuracy'] == d['binary_accuracy']
    names = {i: f'code{i}' for i in range(71)}
    assert set(ecg_b200.get_accuracy(preds, labels, id2code=names)['per_class_auc']) == {f'code{c}' for c in d['per_class_auc']}


def test_get_accuracy_large_random_vs_oracle():
    g = torch.Generator().manual_seed(5)
    n = 3000
    preds = torch.sigmoid(torch.round(torch.randn(n, 71, generator=g) * 16) / 16)
    labels = (torch.rand(n, 71, generator=g) < 0.05).float()
    d = ecg_b200.get_accuracy(preds.cuda(), labels.cuda())
    scalars, per_class = oracle_metrics.get_accuracy(preds.numpy(), labels.numpy())
    for k, w in zip(KEYS, scalars):
        assert abs(d[k] - w) <= 1e-12, k
    for c, v in d['per_class_auc'].items():
        assert abs(v - per_class[c]) <= 1e-12


def test_get_accuracy_degenerate_and_cpu():
    preds = torch.full((4, 3), 0.7).cuda()
    labels = torch.ones(4, 3).cuda()
    d = ecg_b200.get_accuracy(preds, labels)
    assert d['macro_auc'] is None and d['per_class_auc'] is None and d['binary_accuracy'] == 1.0
    assert d['binary_positive_recall'] == 0.0          # zero_division=0: nothing was predicted negative
    with pytest.raises(RuntimeError):
        ecg_b200.get_accuracy(preds.cpu(), labels.cpu())


def test_evaluate_loop_matches_oracle():
    cfg = dict(max_signal_length=500, patch_size=50, num_channels=12, hidden_size=64, num_hidden_layers=2,
               num_attention_heads=4, intermediate_size=128, hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1)
    torch.manual_seed(11)
    oracle = OracleEcgVit(config=OracleConfig(**cfg), loss_reduction='none').eval()
    model = EcgVit(config=EcgVitConfig(compute_dtype='fp32', **cfg))
    model.load_state_dict(oracle.state_dict(), strict=True)
    model.cuda().train()
    batches = []
    for s in range(3):
        x, y = synthetic_batch(8 if s < 2 else 5, length=500, seed=100 + s)   # ragged last batch
        batches.append(dict(sample_values=x, labels=y))
    for red in ('mean', 'none'):
        out = ecg_b200.evaluate(model, batches, loss_reduction=red, return_predictions=True)
        assert model.training and model.loss_reduction == 'mean'       # both restored (train.py:340,377)
        oracle.loss_reduction = red
        with torch.no_grad():
            outs = [oracle(**b) for b in batches]
        logits = torch.cat([o.logits for o in outs])
        labels = torch.cat([b['labels'] for b in batches])
        assert (out['predictions']['logits'].cpu() - logits).norm() / logits.norm() < 1e-5
        if red == 'mean':
            want = float(np.mean([o.loss.item() for o in outs]))
            assert abs(out['metrics']['eval/loss'] - want) < 1e-5 * want
        else:
            want = torch.cat([o.loss.mean(dim=-1) for o in outs]).numpy()
            assert out['metrics']['eval/loss'].shape == (21,)
            assert np.allclose(out['metrics']['eval/loss'], want, rtol=1e-5)
        scalars, per_class = oracle_metrics.get_accuracy(torch.sigmoid(out['predictions']['logits']).cpu().numpy(),
                                                         labels.numpy())
        assert abs(out['metrics']['eval/binary_accuracy'] - scalars[0]) < 1e-12
        if not math.isnan(scalars[4]):
            assert abs(out['metrics']['eval/macro_auc'] - scalars[4]) < 1e-12
