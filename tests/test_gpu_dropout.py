"""GPU tier: training-mode dropout (the reference default is p = 0.1 at 5 sites per block + the embedding,
ecg_vit.py:38-39,113-114).  torch's Philox stream cannot be reproduced by another implementation, so parity is shown the
other way round: the kernels' counter-based masks are recomputed on the host (`_lib.dropout_keep_mask`) and INJECTED into
the CPU oracle's nn.Dropout modules; forward, loss and every gradient must then agree to the usual tolerances."""
import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu

from conftest import GOLDEN_CFG
import ecg_b200
from ecg_b200 import EcgVit, EcgVitConfig, FusedTrainer
from oracle.ecg_vit_oracle import OracleConfig, OracleEcgVit, synthetic_batch


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def cosine(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-300))


class InjectedDropout(nn.Module):
    """nn.Dropout with the keep-mask of the CUDA kernels: y = x * mask (mask already holds 1 / (1 - p))"""

    def __init__(self, seed, stream, p, attention=False):
        super().__init__()
        self.seed, self.stream, self.p, self.attention = seed, stream, p, attention

    def forward(self, x):
        if self.attention:  # [B, H, N, N] -> ((b*H + h) * Np + i) * Np + j, Np = N rounded up to 64
            B, H, N, _ = x.shape
            Np = (N + 63) // 64 * 64
            bh = torch.arange(B * H).reshape(B, H, 1, 1)
            i = torch.arange(N).reshape(1, 1, N, 1)
            j = torch.arange(N).reshape(1, 1, 1, N)
            idx = (bh * Np + i) * Np + j
        else:  # [B, N, F] stored as [B*N, F]: row * F + col
            idx = torch.arange(x.numel()).reshape(x.shape)
        return x * ecg_b200._lib.dropout_keep_mask(self.seed, self.stream, self.p, idx)


def inject(oracle, seed, p_emb, p_blk):
    oracle.vit.dropout = InjectedDropout(seed, 0, p_emb)
    for l, (attn, ff) in enumerate(oracle.vit.transformer.layers):
        attn.fn.dropout = InjectedDropout(seed, 1 + 4 * l, p_blk, attention=True)
        attn.fn.to_out[1] = InjectedDropout(seed, 2 + 4 * l, p_blk)
        ff.fn.net[2] = InjectedDropout(seed, 3 + 4 * l, p_blk)
        ff.fn.net[4] = InjectedDropout(seed, 4 + 4 * l, p_blk)


CFG = dict(GOLDEN_CFG, hidden_size=128, num_attention_heads=4, intermediate_size=256, hidden_dropout_prob=0.1,
           attention_probs_dropout_prob=0.2)


@pytest.mark.parametrize('dtype,tol', [('fp32', 1e-5), ('bf16', 1e-2)])
def test_dropout_forward_backward_match_oracle_with_injected_masks(dtype, tol):
    torch.manual_seed(3)
    oracle = OracleEcgVit(config=OracleConfig(**CFG)).train()
    model = EcgVit(config=EcgVitConfig(compute_dtype=dtype, **CFG))
    model.load_state_dict(oracle.state_dict(), strict=True)
    model.cuda().train()
    x, y = synthetic_batch(6, length=500, seed=9)
    out = model(sample_values=x.cuda(), labels=y.cuda())
    out.loss.backward()
    seed = int(model._engine.rng[0])
    inject(oracle, seed, CFG['attention_probs_dropout_prob'], CFG['hidden_dropout_prob'])
    ref = oracle(sample_values=x, labels=y)
    ref.loss.backward()
    assert rel(out.logits, ref.logits) < tol and rel(out.loss, ref.loss) < tol
    # the masks really are active: the eval-mode forward differs
    with torch.no_grad():
        assert rel(model.eval()(x.cuda()).logits, ref.logits) > 10 * tol
    for (k, p), (_, q) in zip(model.named_parameters(), oracle.named_parameters()):
        if dtype == 'fp32':
            assert rel(p.grad, q.grad) < 5e-4, (k, rel(p.grad, q.grad))
        else:
            assert cosine(p.grad, q.grad) > 0.999, (k, cosine(p.grad, q.grad))


def test_dropout_fused_steps_track_oracle_and_draw_fresh_masks():
    torch.manual_seed(4)
    oracle = OracleEcgVit(config=OracleConfig(**CFG)).train()
    model = EcgVit(config=EcgVitConfig(compute_dtype='fp32', **CFG))
    model.load_state_dict(oracle.state_dict(), strict=True)
    model.cuda().train()
    x, y = synthetic_batch(6, length=500, seed=10)
    opt = torch.optim.AdamW(oracle.parameters(), lr=1e-3, weight_decay=1e-2)
    tr = FusedTrainer(model, learning_rate=1e-3, weight_decay=1e-2, max_grad_norm=1.0, use_cuda_graph=True)
    seeds = []
    for _ in range(3):
        loss, _ = tr.step(x.cuda(), y.cuda())
        seeds.append(int(model._engine.rng[0]))
        inject(oracle, seeds[-1], CFG['attention_probs_dropout_prob'], CFG['hidden_dropout_prob'])
        opt.zero_grad()
        o = oracle(sample_values=x, labels=y)
        o.loss.backward()
        nn.utils.clip_grad_norm_(oracle.parameters(), 1.0, error_if_nonfinite=True)
        opt.step()
        assert rel(loss, o.loss) < 1e-5
    assert len(set(seeds)) == 3, 'every step must draw a new mask'
    for k, v in model.state_dict().items():
        assert rel(v, oracle.state_dict()[k]) < 2e-5, k


def test_dropout_keep_rate_on_device():
    """FF1's epilogue output h = drop(gelu(u)): zeros appear at rate p (gelu itself is never exactly 0 for these inputs)"""
    cfg = dict(CFG, hidden_dropout_prob=0.25, attention_probs_dropout_prob=0.0)
    model = EcgVit(config=EcgVitConfig(compute_dtype='bf16', **cfg)).cuda().train()
    x, y = synthetic_batch(16, length=500, seed=2)
    model(sample_values=x.cuda(), labels=y.cuda())
    h = model._engine._cur.h[0].float()
    u = model._engine._cur.u[0].float()
    zero_rate = float(((h == 0) & (u.abs() > 1e-3)).float().mean())
    assert abs(zero_rate - 0.25) < 0.01, zero_rate
