"""GPU tier: the reference's per-record input transforms fused into the patch gather.
Bit-exact against vectors the reference's own `transform.py` classes produced (tests/golden/input_transform.npz)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import ecg_b200
from ecg_b200 import EcgVit, EcgVitConfig, FusedTrainer, InputPipeline, _lib as L
from oracle.ecg_vit_oracle import patch_matrix

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'input_transform.npz'))
CASES = ['ragged', 'full_block', 'long']


def stream():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize('case', CASES)
@pytest.mark.parametrize('mode', ['eval', 'train'])
def test_patchify_transform_bit_exact(case, mode):
    lib = L.load()
    rec, k = torch.from_numpy(G[f'{case}/records']).cuda(), int(G[f'{case}/k'])
    want_sig = torch.from_numpy(G[f'{case}/{mode}'])            # what the reference's dataset returns per record
    B, C, L_raw = rec.shape
    Lp = want_sig.shape[-1]
    P = k if case != 'long' else 25
    n = Lp // P
    mean = torch.from_numpy(G['mean']).float().cuda()           # Normalize casts its stats to fp32 (transform.py:27)
    std = torch.from_numpy(G['std']).float().cuda()
    spans = torch.from_numpy(G[f'{case}/spans']).to(torch.int32).cuda() if mode == 'train' else None
    a = torch.full((B * n, P * C), 9.0, device='cuda')
    L.check(lib.ecgvit_patchify_transform(rec.data_ptr(), mean.data_ptr(), std.data_ptr(), L.ptr(spans), a.data_ptr(),
                                          B, C, rec.stride(1), L_raw, n, P, L.F32, stream()), 'patchify_transform')
    want = patch_matrix(want_sig, P)
    assert torch.equal(a.cpu(), want)                           # normalisation, padding and masking bit-exact
    ab = torch.empty(B * n, P * C, device='cuda', dtype=torch.bfloat16)
    L.check(lib.ecgvit_patchify_transform(rec.data_ptr(), mean.data_ptr(), std.data_ptr(), L.ptr(spans), ab.data_ptr(),
                                          B, C, rec.stride(1), L_raw, n, P, L.BF16, stream()), 'patchify_transform')
    assert torch.equal(ab.cpu(), want.bfloat16())
    # no Normalize, no TimeOut: pure gather + zero padding
    L.check(lib.ecgvit_patchify_transform(rec.data_ptr(), None, None, None, a.data_ptr(), B, C, rec.stride(1), L_raw,
                                          n, P, L.F32, stream()), 'patchify_transform')
    padded = torch.nn.functional.pad(rec.cpu(), (0, Lp - L_raw))
    assert torch.equal(a.cpu(), patch_matrix(padded, P))


def test_patchify_transform_rejects_bad_arguments():
    lib = L.load()
    x = torch.zeros(1, 12, 100, device='cuda')
    a = torch.zeros(2, 600, device='cuda')
    m = torch.zeros(12, device='cuda')
    assert lib.ecgvit_patchify_transform(x.data_ptr(), m.data_ptr(), None, None, a.data_ptr(), 1, 12, 100, 100, 2, 50,
                                         L.F32, stream()) != 0          # mean without std
    assert lib.ecgvit_patchify_transform(x.data_ptr(), None, None, None, a.data_ptr(), 1, 12, 100, 100, 1, 50,
                                         L.F32, stream()) != 0          # n_patch * P would drop samples
    assert b'patchify_transform' in lib.ecgvit_last_error()


CFG = dict(max_signal_length=300, patch_size=50, num_channels=12, hidden_size=64, num_hidden_layers=2,
           num_attention_heads=4, intermediate_size=128, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)


@pytest.mark.parametrize('dtype', ['fp32', 'bf16'])
@pytest.mark.parametrize('case', ['ragged', 'full_block'])
def test_model_on_raw_records_equals_model_on_reference_transformed(case, dtype):
    """EcgVit with an InputPipeline on RAW records == the same EcgVit on what the reference's dataset would have fed it"""
    torch.manual_seed(3)
    model = EcgVit(config=EcgVitConfig(compute_dtype=dtype, **CFG)).cuda()
    rec, k = torch.from_numpy(G[f'{case}/records']).cuda(), int(G[f'{case}/k'])
    y = (torch.rand(rec.shape[0], 71, device='cuda') < 0.1).float()
    pipe = InputPipeline(normalize=dict(mean=G['mean'], std=G['std']), pad=k, timeout=True)
    spans = torch.from_numpy(G[f'{case}/spans'])
    for mode in ('eval', 'train'):
        model.train(mode == 'train')
        model.input_pipeline = None
        want = model(sample_values=torch.from_numpy(G[f'{case}/{mode}']).cuda(), labels=y)
        want_logits, want_loss = want.logits.clone(), want.loss.clone()
        model.input_pipeline = pipe
        got = model(sample_values=rec, labels=y, time_out_spans=spans)
        assert torch.equal(got.logits, want_logits) and torch.equal(got.loss, want_loss)
    # spans drawn by the model itself follow the reference's RNG stream
    model.train()
    torch.manual_seed(1234)
    got = model(sample_values=rec, labels=y)
    assert torch.equal(got.logits, want_logits)


def test_fused_trainer_with_pipeline_graph_equals_eager():
    case = 'ragged'
    rec, k = torch.from_numpy(G[f'{case}/records']).cuda(), int(G[f'{case}/k'])
    y = (torch.rand(rec.shape[0], 71, device='cuda') < 0.1).float()
    results = []
    for graph in (False, True):
        torch.manual_seed(3)
        model = EcgVit(config=EcgVitConfig(compute_dtype='fp32', **CFG)).cuda().train()
        model.input_pipeline = InputPipeline(normalize=dict(mean=G['mean'], std=G['std']), pad=k, timeout=True)
        tr = FusedTrainer(model, learning_rate=1e-3, weight_decay=1e-2, schedule='constant', n_warmup=0, n_step=10,
                          max_grad_norm=1.0, use_cuda_graph=graph, data_parallel=False)
        torch.manual_seed(1234)
        losses = []
        for _ in range(3):
            loss, _ = tr.step(rec, y)   # a new TimeOut span per record every step, read by the captured graph
            losses.append(float(loss))
        results.append((losses, model._flat_p.clone()))
    for a, b in zip(results[0][0], results[1][0]):
        assert abs(a - b) < 1e-5 * abs(a)            # fp32 split-K atomics reorder between runs
    assert float((results[0][1] - results[1][1]).norm() / results[0][1].norm()) < 1e-5
    assert len(set(results[0][0])) == 3
