"""GPU tier: the tcgen05 / TMEM attention kernels (csrc/attention_tc.cu; bf16, head dim 64) through the C ABI against a
plain fp32 torch statement of vit_pytorch's Attention core (softmax(q k^T * scale) [dropout] v) on the same bf16 inputs.

Covers both geometries -- packed (N <= 64: two (batch, head) problems per 128-row tile, incl. an odd number of problems
and single-token sequences) and long (N > 64: streamed key blocks, incl. N = 2401 of BASELINE.json configs[3], partial
last blocks and an odd number of query tiles) -- the saved log-sum-exp, and attention-probability dropout with the
kernels' counter-based masks replicated on the host."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from ecg_b200 import _lib as L

BF16 = L.BF16


def stream():
    return torch.cuda.current_stream().cuda_stream


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def reference(qkv, B, N, H, dh, scale, mask=None):
    """fp32 attention on the bf16-rounded projection; returns (o [B*N, inner], lse [B, H, N], leaf for autograd)"""
    leaf = qkv.float().requires_grad_(True)
    q, k, v = (t.reshape(B, N, H, dh).permute(0, 2, 1, 3) for t in leaf.chunk(3, dim=-1))
    s = (q @ k.transpose(-1, -2)) * scale
    p = s.softmax(-1)
    if mask is not None:
        p = p * mask
    o = (p @ v).permute(0, 2, 1, 3).reshape(B * N, H * dh)
    return o, torch.logsumexp(s, -1), leaf


def keep_mask(seed, site, p, B, N, H):
    Np = (N + 63) // 64 * 64
    bh = torch.arange(B * H).reshape(B, H, 1, 1)
    i = torch.arange(N).reshape(1, 1, N, 1)
    j = torch.arange(N).reshape(1, 1, 1, N)
    return L.dropout_keep_mask(seed, site, p, (bh * Np + i) * Np + j).cuda()


def run_fwd(lib, qkv, B, N, H, dh, p=0.0, site=0, seed_t=None):
    inner = H * dh
    o = torch.full((B * N, inner), float('nan'), device='cuda', dtype=torch.bfloat16)
    lse = torch.full((B, H, N), float('nan'), device='cuda')
    L.check(lib.ecgvit_attention_fwd(qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), B, N, H, dh, dh ** -0.5, p, site,
                                     None if seed_t is None else seed_t.data_ptr(), BF16, stream()), 'attn')
    torch.cuda.synchronize()
    return o, lse


CASES = [(3, 51, 12), (256, 51, 12), (1, 51, 1), (3, 51, 3), (5, 64, 1), (2, 1, 2), (7, 17, 5), (2, 33, 4),   # packed
         (2, 65, 3), (1, 128, 2), (1, 129, 1), (2, 200, 2), (1, 321, 1), (1, 640, 3), (1, 2401, 2)]            # long


@pytest.mark.parametrize('B,N,H', CASES)
def test_attention_tc_forward(B, N, H):
    lib = L.load()
    dh = 64
    g = torch.Generator(device='cuda').manual_seed(B * 1000 + N)
    qkv = (torch.randn(B * N, 3 * H * dh, device='cuda', generator=g) * 1.5).bfloat16()
    o, lse = run_fwd(lib, qkv, B, N, H, dh)
    want, want_lse, _ = reference(qkv, B, N, H, dh, dh ** -0.5)
    assert torch.isfinite(o.float()).all() and torch.isfinite(lse).all()
    assert rel(o, want) < 8e-3, rel(o, want)
    assert rel(lse, want_lse) < 1e-5, rel(lse, want_lse)
    # element-wise: nothing further off than bf16 rounding of P and O allows
    assert float((o.float() - want).abs().max()) < 0.03 * float(want.abs().max())


@pytest.mark.parametrize('B,N,H', [(3, 51, 4), (1, 200, 2), (1, 300, 1)])
def test_attention_tc_forward_dropout_masks_match_host_replica(B, N, H):
    lib = L.load()
    dh, p, site, seed = 64, 0.2, 9, 424242
    qkv = torch.randn(B * N, 3 * H * dh, device='cuda').bfloat16()
    seed_t = torch.tensor([seed], dtype=torch.int32, device='cuda')
    o, lse = run_fwd(lib, qkv, B, N, H, dh, p, site, seed_t)
    mask = keep_mask(seed, site, p, B, N, H)
    want, want_lse, _ = reference(qkv, B, N, H, dh, dh ** -0.5, mask)
    assert rel(o, want) < 8e-3, rel(o, want)
    assert rel(lse, want_lse) < 1e-5           # the saved statistics are those of the undropped softmax
    plain, _, _ = reference(qkv, B, N, H, dh, dh ** -0.5)
    assert rel(o, plain) > 0.1                 # and the masks really were applied


def test_attention_tc_forward_large_scores_are_stable():
    """running max / rescale: scores far above fp32 exp range, maxima that move from block to block"""
    lib = L.load()
    B, N, H, dh = 1, 500, 2, 64
    qkv = torch.randn(B * N, 3 * H * dh, device='cuda')
    qkv[:, :2 * H * dh] *= 6.0
    ramp = torch.linspace(0.2, 3.0, N, device='cuda').unsqueeze(1)   # later keys score higher: the max keeps moving
    qkv[:, H * dh:2 * H * dh] *= ramp
    qkv = qkv.bfloat16()
    o, lse = run_fwd(lib, qkv, B, N, H, dh)
    want, want_lse, _ = reference(qkv, B, N, H, dh, dh ** -0.5)
    assert torch.isfinite(o.float()).all()
    assert rel(o, want) < 8e-3 and rel(lse, want_lse) < 1e-5


def run_bwd(lib, qkv, o, lse, d_o, B, N, H, dh, p=0.0, site=0, seed_t=None):
    dqkv = torch.full_like(qkv, float('nan'))
    n_scr = int(lib.ecgvit_attention_bwd_scratch_floats(B, N, H, dh, BF16))
    scr = torch.empty(max(n_scr, 1), device='cuda')
    L.check(lib.ecgvit_attention_bwd(qkv.data_ptr(), o.data_ptr(), d_o.data_ptr(), lse.data_ptr(), dqkv.data_ptr(),
                                     scr.data_ptr() if n_scr else None, B, N, H, dh, dh ** -0.5, p, site,
                                     None if seed_t is None else seed_t.data_ptr(), BF16, stream()), 'attn_bwd')
    torch.cuda.synchronize()
    return dqkv


def parts(t, H, dh):
    return t.float().chunk(3, dim=-1)


@pytest.mark.parametrize('B,N,H', CASES)
def test_attention_tc_backward(B, N, H):
    lib = L.load()
    dh = 64
    g = torch.Generator(device='cuda').manual_seed(B * 1000 + N + 1)
    qkv = torch.randn(B * N, 3 * H * dh, device='cuda', generator=g).bfloat16()
    d_o = torch.randn(B * N, H * dh, device='cuda', generator=g).bfloat16()
    o, lse = run_fwd(lib, qkv, B, N, H, dh)
    dqkv = run_bwd(lib, qkv, o, lse, d_o, B, N, H, dh)
    want, _, leaf = reference(qkv, B, N, H, dh, dh ** -0.5)
    want.backward(d_o.float())
    assert torch.isfinite(dqkv.float()).all()
    for name, got, ref in zip('qkv', parts(dqkv, H, dh), parts(leaf.grad, H, dh)):
        assert rel(got, ref) < 1.5e-2, (name, rel(got, ref))


@pytest.mark.parametrize('B,N,H', [(3, 51, 4), (1, 200, 2), (1, 300, 1)])
def test_attention_tc_backward_dropout_masks_match_host_replica(B, N, H):
    lib = L.load()
    dh, p, site, seed = 64, 0.2, 5, 77
    qkv = torch.randn(B * N, 3 * H * dh, device='cuda').bfloat16()
    d_o = torch.randn(B * N, H * dh, device='cuda').bfloat16()
    seed_t = torch.tensor([seed], dtype=torch.int32, device='cuda')
    o, lse = run_fwd(lib, qkv, B, N, H, dh, p, site, seed_t)
    dqkv = run_bwd(lib, qkv, o, lse, d_o, B, N, H, dh, p, site, seed_t)
    want, _, leaf = reference(qkv, B, N, H, dh, dh ** -0.5, keep_mask(seed, site, p, B, N, H))
    want.backward(d_o.float())
    for name, got, ref in zip('qkv', parts(dqkv, H, dh), parts(leaf.grad, H, dh)):
        assert rel(got, ref) < 1.5e-2, (name, rel(got, ref))
