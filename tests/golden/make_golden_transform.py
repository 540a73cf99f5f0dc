"""Generates tests/golden/input_transform.npz by running the REFERENCE's own transform classes
(/root/reference/ecg_transformer/preprocess/transform.py, imported through oracle/ref_shim.py) on seeded records.
Run in the build container only (the reference does not travel):  python tests/golden/make_golden_transform.py"""
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

ref_shim.install()
T = importlib.import_module('ecg_transformer.preprocess.transform')
cfg = json.load(open('/root/reference/ecg_transformer/util/config.json'))
stats = cfg['datasets']['PTB-XL']['train-stats']['denoised']  # what get_ptbxl_dataset(std_norm=True) uses

out = dict(mean=np.asarray(stats['mean'], dtype=np.float64), std=np.asarray(stats['std'], dtype=np.float64))
rng = np.random.default_rng(77)
for name, (B, L, k) in dict(ragged=(3, 245, 50), full_block=(2, 250, 50), long=(2, 500, 25)).items():
    rec = (rng.standard_normal((B, 12, L)) * 0.2).astype(np.float64)  # h5 stores float64 (dataset.py:88)
    tsf = [T.Normalize(mean=stats['mean'], std=stats['std']),
           T.TimeEndPad(k, pad_kwargs=dict(mode='constant', constant_values=0))]
    eval_out = []
    for b in range(B):
        a = rec[b].astype(np.float32)
        for t in tsf:
            a = t(a)
        eval_out.append(a)
    # training split: + TimeOut; record the spans it drew by replaying its RNG calls under the same seed
    torch.manual_seed(1234)
    to = T.TimeOut()
    train_out = []
    for b in range(B):
        a = rec[b].astype(np.float32)
        for t in tsf + [to]:
            a = t(a)
        train_out.append(a)
    torch.manual_seed(1234)
    sampler = torch.distributions.Uniform(low=0, high=0.5)
    Lp = eval_out[0].shape[-1]
    spans = []
    for b in range(B):
        r = sampler.sample().item()
        l_crop = round(r * Lp)
        spans.append((torch.randint(high=Lp - l_crop, size=(1,)).item(), l_crop))
    out[f"{name}/records"] = rec.astype(np.float32)  # the cast `dataset.py:88` applies first
    out[f'{name}/k'] = np.int64(k)
    out[f'{name}/eval'] = np.stack(eval_out)
    out[f'{name}/train'] = np.stack(train_out)
    out[f'{name}/spans'] = np.asarray(spans, dtype=np.int64)
    # sanity: the replayed spans are exactly the zeroed samples
    for b, (s, l) in enumerate(spans):
        ref = eval_out[b].copy()
        ref[..., s:s + l] = 0
        assert np.array_equal(ref, train_out[b]), name
np.savez_compressed(os.path.join(ROOT, 'tests/golden/input_transform.npz'), **out)
print({k: getattr(v, 'shape', v) for k, v in out.items()})
