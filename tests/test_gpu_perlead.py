"""GPU tier: BASELINE.json configs[3] geometry -- per-lead tokens (every lead tokenised on its own, patch_dim = P) and
the tiled attention kernels it needs (N > 64).  Oracle: vit_pytorch's ViT(image_size=(C, L), patch_size=(1, P),
channels=1) restated in oracle/vit_restated.py (not constructible through the reference wrapper)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import ecg_b200
from ecg_b200 import EcgVit, EcgVitConfig, FusedTrainer
from oracle.ecg_vit_oracle import OracleConfig, OracleEcgVit, OracleTrainer, synthetic_batch

from test_gpu_parity import rel, cosine, FP32_TOL, BF16_TOL, BF16_GRAD_COS


def make(cfg, dtype, batch, seed=77):
    torch.manual_seed(seed)
    oracle = OracleEcgVit(config=OracleConfig(**cfg)).train()
    model = EcgVit(config=EcgVitConfig(compute_dtype=dtype, **cfg))
    assert list(model.state_dict().keys()) == list(oracle.state_dict().keys())
    assert all(a.shape == b.shape for a, b in zip(model.state_dict().values(), oracle.state_dict().values()))
    model.load_state_dict(oracle.state_dict(), strict=True)
    model.cuda().train()
    x, y = synthetic_batch(batch, num_channels=cfg['num_channels'], length=cfg['max_signal_length'], seed=seed)
    return oracle, model, x, y


SHORT = dict(max_signal_length=300, patch_size=25, num_channels=3, hidden_size=64, num_hidden_layers=2,
             num_attention_heads=4, intermediate_size=128, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
             per_lead_tokens=True)   # N = 3 * 12 + 1 = 37, patch_dim 25 -> padded to 32
LONG = dict(SHORT, max_signal_length=500, num_channels=12, hidden_size=128, num_attention_heads=2)  # N = 241, dh = 64


@pytest.mark.parametrize('dtype', ['fp32', 'bf16'])
def test_per_lead_forward_backward_and_steps(dtype):
    oracle, model, x, y = make(SHORT, dtype, 4)
    assert model.vit.pos_embedding.shape == (1, 37, 64) and model.vit.to_patch_embedding[1].weight.shape == (64, 25)
    want = oracle(sample_values=x, labels=y)
    want.loss.backward()
    got = model(sample_values=x.cuda(), labels=y.cuda())
    got.loss.backward()
    tol = FP32_TOL if dtype == 'fp32' else BF16_TOL
    assert rel(got.logits, want.logits) < tol and rel(got.loss, want.loss) < tol
    for (k, p), (_, q) in zip(model.named_parameters(), oracle.named_parameters()):
        if dtype == 'fp32':
            assert rel(p.grad, q.grad) < 2e-4, (k, rel(p.grad, q.grad))
        assert cosine(p.grad, q.grad) > (0.999999 if dtype == 'fp32' else BF16_GRAD_COS), k
    if dtype == 'fp32':
        oracle2, model2, _, _ = make(SHORT, dtype, 4)
        ot = OracleTrainer(oracle2, learning_rate=1e-3, weight_decay=1e-2)
        tr = FusedTrainer(model2, learning_rate=1e-3, weight_decay=1e-2, use_cuda_graph=True, data_parallel=False)
        for _ in range(3):
            ot.step(x, y)
            tr.step(x.cuda(), y.cuda())
        for (k, p), (_, q) in zip(model2.named_parameters(), oracle2.named_parameters()):
            assert rel(p, q) < 1e-5, k


def test_per_lead_long_sequence_uses_tiled_attention():
    """N = 241 > 64: flash kernels forward and backward, against the oracle at the bf16 bar"""
    oracle, model, x, y = make(LONG, 'bf16', 3)
    want = oracle(sample_values=x, labels=y)
    want.loss.backward()
    got = model(sample_values=x.cuda(), labels=y.cuda())
    got.loss.backward()
    assert rel(got.logits, want.logits) < BF16_TOL and rel(got.loss, want.loss) < BF16_TOL
    for (k, p), (_, q) in zip(model.named_parameters(), oracle.named_parameters()):
        assert cosine(p.grad, q.grad) > BF16_GRAD_COS, (k, cosine(p.grad, q.grad))
    # fp32 parity mode keeps whole sequences in shared memory and says so instead of falling back
    _, model32, _, _ = make(LONG, 'fp32', 1)
    with pytest.raises(RuntimeError, match='sequence too long'):
        model32(sample_values=x[:1].cuda(), labels=y[:1].cuda())


def test_per_lead_long_sequence_dropout_masks_match_host_replica():
    from test_gpu_dropout import inject
    cfg = dict(LONG, hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.2)
    oracle, model, x, y = make(cfg, 'bf16', 2)
    got = model(sample_values=x.cuda(), labels=y.cuda())
    got.loss.backward()
    seed = int(model._engine.rng[0])
    inject(oracle, seed, cfg['attention_probs_dropout_prob'], cfg['hidden_dropout_prob'])
    want = oracle(sample_values=x, labels=y)
    want.loss.backward()
    assert rel(got.logits, want.logits) < BF16_TOL
    for (k, p), (_, q) in zip(model.named_parameters(), oracle.named_parameters()):
        assert cosine(p.grad, q.grad) > BF16_GRAD_COS, (k, cosine(p.grad, q.grad))
