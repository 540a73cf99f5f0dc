"""GPU tier: every C-ABI kernel against plain PyTorch on the same seeded inputs (called through ctypes)."""
import ctypes
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

import ecg_b200
from ecg_b200 import _lib as L
from oracle.ecg_vit_oracle import patch_matrix

DT = {L.F32: torch.float32, L.BF16: torch.bfloat16}


@pytest.fixture(scope='module')
def lib():
    assert torch.cuda.is_available()
    lib = L.load()
    assert lib.ecgvit_device_ok() == 1, 'these kernels are sm_100a only'
    return lib


def stream():
    return torch.cuda.current_stream().cuda_stream


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def gemm(lib, A, B, M, N, K, a_k, b_k, epi, out, dtype, out2=None, aux=None, bias=None, split_k=1):
    g = L.GemmArgs(M, N, K, A.data_ptr(), A.stride(0), a_k, B.data_ptr(), B.stride(0), b_k, epi, out.data_ptr(),
                   out.stride(0), L.ptr(out2), L.ptr(aux), L.ptr(bias), dtype, split_k, 0, 0.0, None)
    L.check(lib.ecgvit_gemm(ctypes.byref(g), stream()), 'gemm')


# ---------------------------------------------------------------------------------------------------------------
def test_patchify_is_bit_exact(lib):
    for (B, C, Lg, P) in [(2, 12, 500, 50), (3, 12, 2500, 50), (2, 12, 2560, 64), (1, 1, 5000, 25)]:
        x = torch.randn(B, C, Lg, device='cuda')
        n = Lg // P
        want = patch_matrix(x.cpu(), P)
        a = torch.empty(B * n, P * C, device='cuda')
        L.check(lib.ecgvit_patchify(x.data_ptr(), a.data_ptr(), B, C, x.stride(1), n, P, L.F32, stream()), 'patchify')
        assert torch.equal(a.cpu(), want)
        ab = torch.empty(B * n, P * C, device='cuda', dtype=torch.bfloat16)
        L.check(lib.ecgvit_patchify(x.data_ptr(), ab.data_ptr(), B, C, x.stride(1), n, P, L.BF16, stream()), 'patchify')
        assert torch.equal(ab.cpu(), want.bfloat16())
    # integer index check on a shorter-than-stride view (L < max_signal_length)
    idx = torch.arange(2 * 12 * 600, dtype=torch.float32, device='cuda').reshape(2, 12, 600)[:, :, :500]
    a = torch.empty(20, 600, device='cuda')
    L.check(lib.ecgvit_patchify(idx.data_ptr(), a.data_ptr(), 2, 12, idx.stride(1), 10, 50, L.F32, stream()), 'patchify')
    assert torch.equal(a.cpu(), patch_matrix(idx.cpu().contiguous(), 50))


def _operands(M, N, K, a_k, b_k, dtype, seed=0):
    g = torch.Generator(device='cuda').manual_seed(seed)
    A = torch.randn((M, K) if a_k else (K, M), device='cuda', generator=g).to(dtype)
    B = torch.randn((N, K) if b_k else (K, N), device='cuda', generator=g).to(dtype)
    Am = A.float() if a_k else A.float().t()
    Bm = B.float() if b_k else B.float().t()
    return A, B, Am @ Bm.t()


@pytest.mark.parametrize('a_k,b_k', [(1, 1), (1, 0), (0, 0), (0, 1)])
def test_gemm_f32_layouts(lib, a_k, b_k):
    M, N, K = 200, 132, 77
    A, B, want = _operands(M, N, K, a_k, b_k, torch.float32)
    out = torch.zeros(M, N, device='cuda')
    gemm(lib, A, B, M, N, K, a_k, b_k, L.EPI_STORE, out, L.F32)
    assert rel(out, want) < 1e-6


@pytest.mark.parametrize('dtype', [L.F32, L.BF16])
def test_gemm_epilogues(lib, dtype):
    td = DT[dtype]
    tol = 1e-6 if dtype == L.F32 else 6e-3
    M, N, K = 304, 256, 192
    A, B, acc = _operands(M, N, K, 1, 1, td, seed=1)
    bias = torch.randn(N, device='cuda')
    aux = torch.randn(M, N, device='cuda').to(td)
    # STORE + bias
    out = torch.zeros(M, N, device='cuda', dtype=td)
    gemm(lib, A, B, M, N, K, 1, 1, L.EPI_STORE, out, dtype, bias=bias)
    assert rel(out, acc + bias) < tol
    # BIAS_RES
    out = torch.zeros(M, N, device='cuda', dtype=td)
    gemm(lib, A, B, M, N, K, 1, 1, L.EPI_BIAS_RES, out, dtype, aux=aux, bias=bias)
    assert rel(out, acc + bias + aux.float()) < tol
    # BIAS_GELU
    out2 = torch.zeros(M, N, device='cuda', dtype=td)
    gemm(lib, A, B, M, N, K, 1, 1, L.EPI_BIAS_GELU, out, dtype, out2=out2, bias=bias)
    u = acc + bias
    assert rel(out, u) < tol
    assert rel(out2, torch.nn.functional.gelu(out.float())) < tol  # gelu of the stored pre-activation
    # DGELU (B MN-major like the real dgrad)
    A2, B2, acc2 = _operands(M, N, K, 1, 0, td, seed=2)
    gemm(lib, A2, B2, M, N, K, 1, 0, L.EPI_DGELU, out, dtype, aux=aux)
    uu = aux.float().requires_grad_(True)
    torch.nn.functional.gelu(uu).sum().backward()
    assert rel(out, acc2 * uu.grad) < tol
    # ATOMIC_F32 accumulates
    A3, B3, acc3 = _operands(M, N, K, 0, 0, td, seed=3)
    o32 = torch.ones(M, N, device='cuda')
    gemm(lib, A3, B3, M, N, K, 0, 0, L.EPI_ATOMIC_F32, o32, dtype, split_k=0)
    assert rel(o32, acc3 + 1.0) < (1e-6 if dtype == L.F32 else 1e-5)


@pytest.mark.parametrize('a_k,b_k', [(1, 1), (1, 0), (0, 0), (0, 1)])
@pytest.mark.parametrize('shape', [(128, 256, 64), (128, 128, 128), (256, 512, 256), (333, 264, 600), (1000, 768, 768),
                                   (44, 64, 64)])
def test_gemm_bf16_tcgen05_layouts(lib, a_k, b_k, shape):
    M, N, K = shape
    if (not a_k and M % 8) or (not b_k and N % 8):
        pytest.skip('MN-major operands need a leading dimension that is a multiple of 8')
    A, B, want = _operands(M, N, K, a_k, b_k, torch.bfloat16, seed=M + N + K)
    out = torch.zeros(M, N, device='cuda', dtype=torch.bfloat16)
    gemm(lib, A, B, M, N, K, a_k, b_k, L.EPI_STORE, out, L.BF16)
    torch.cuda.synchronize()
    assert rel(out, want) < 4e-3, f'rel err {rel(out, want)}'
    # element-wise: every entry within bf16 rounding of the fp32 product
    assert float((out.float() - want).abs().max()) <= 0.02 * float(want.abs().max())


@pytest.mark.parametrize('split_k', [0, 1, 3, 7])
def test_gemm_bf16_wgrad_split_k(lib, split_k):
    # dW[N_out, K_in] = dY^T X with the contraction over M = 2000 rows (both operands MN-major)
    Mrows, n_out, k_in = 2000, 384, 600
    dY = torch.randn(Mrows, n_out, device='cuda').bfloat16()
    X = torch.randn(Mrows, k_in, device='cuda').bfloat16()
    want = dY.float().t() @ X.float()
    out = torch.zeros(n_out, k_in, device='cuda')
    gemm(lib, dY, X, n_out, k_in, Mrows, 0, 0, L.EPI_ATOMIC_F32, out, L.BF16, split_k=split_k)
    assert rel(out, want) < 1e-5


def test_gemm_bf16_persistent_many_tiles(lib):
    # more tiles than SMs: exercises the TMEM double buffering and the smem ring wrap-around
    M, N, K = 128 * 40, 256 * 6, 64 * 13
    A, B, want = _operands(M, N, K, 1, 1, torch.bfloat16, seed=9)
    out = torch.zeros(M, N, device='cuda', dtype=torch.bfloat16)
    gemm(lib, A, B, M, N, K, 1, 1, L.EPI_STORE, out, L.BF16)
    assert rel(out, want) < 4e-3


@pytest.mark.parametrize('dtype', [L.F32, L.BF16])
@pytest.mark.parametrize('d', [64, 256, 768, 1024])
def test_layernorm_fwd_bwd(lib, dtype, d):
    td = DT[dtype]
    M = 777
    x = (torch.randn(M, d, device='cuda') * 2 + 0.5).to(td)
    gamma, beta = torch.randn(d, device='cuda'), torch.randn(d, device='cuda')
    y = torch.empty(M, d, device='cuda', dtype=td)
    mean, rstd = torch.empty(M, device='cuda'), torch.empty(M, device='cuda')
    L.check(lib.ecgvit_layernorm_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), mean.data_ptr(),
                                     rstd.data_ptr(), M, d, 1e-5, dtype, stream()), 'ln')
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (d,), gr, br, 1e-5)
    tol = 2e-6 if dtype == L.F32 else 4e-3
    assert rel(y, yr) < tol
    assert rel(mean, xr.mean(-1)) < 1e-5
    dy = torch.randn(M, d, device='cuda').to(td)
    dres = torch.randn(M, d, device='cuda').to(td)
    yr.backward(dy.float())
    dx = torch.empty(M, d, device='cuda', dtype=td)
    dg, db, dc = (torch.zeros(d, device='cuda') for _ in range(3))
    scratch = torch.empty(int(lib.ecgvit_layernorm_bwd_scratch_floats(d)), device='cuda')
    L.check(lib.ecgvit_layernorm_bwd(dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                     dres.data_ptr(), dx.data_ptr(), dg.data_ptr(), db.data_ptr(), dc.data_ptr(),
                                     scratch.data_ptr(), None, 0.0, 0, None, M, d, 0, dtype, stream()), 'ln_bwd')
    want_dx = xr.grad + dres.float()
    assert rel(dx, want_dx) < (1e-5 if dtype == L.F32 else 5e-3)
    assert rel(dg, gr.grad) < 1e-4 and rel(db, br.grad) < 1e-4
    assert rel(dc, want_dx.sum(0)) < (1e-4 if dtype == L.F32 else 1e-3)  # column sums of the unrounded result
    # the dropout variant: same dx plus dxm = mask * dx / (1 - p), column sums of the masked values
    p, site = 0.25, 7
    seed = torch.tensor([1234567], dtype=torch.int64, device='cuda').to(torch.int32)
    dx2, dxm = torch.empty_like(dx), torch.empty_like(dx)
    dg2, db2, dc2 = (torch.zeros(d, device='cuda') for _ in range(3))
    L.check(lib.ecgvit_layernorm_bwd(dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                     dres.data_ptr(), dx2.data_ptr(), dg2.data_ptr(), db2.data_ptr(), dc2.data_ptr(),
                                     scratch.data_ptr(), dxm.data_ptr(), p, site, seed.data_ptr(), M, d, 1, dtype,
                                     stream()), 'ln_bwd_drop')
    assert float(dg2.abs().max()) == 0.0                      # deferred: nothing folded yet
    L.check(lib.ecgvit_layernorm_bwd_finalize(scratch.data_ptr(), dg2.data_ptr(), db2.data_ptr(), dc2.data_ptr(), M, d,
                                              stream()), 'ln_bwd_finalize')
    assert torch.equal(dx2, dx) and rel(dg2, dg) < 1e-6 and rel(db2, db) < 1e-6
    mult = L.dropout_keep_mask(1234567, site, p, torch.arange(M * d)).reshape(M, d).cuda()
    want_m = want_dx * mult
    assert rel(dxm, want_m) < (1e-5 if dtype == L.F32 else 5e-3)
    assert bool(((dxm == 0) | (mult > 0)).all())
    assert rel(dc2, want_m.sum(0)) < (1e-4 if dtype == L.F32 else 1e-3)


@pytest.mark.parametrize('dtype', [L.F32, L.BF16])
@pytest.mark.parametrize('cfg', [(3, 51, 12, 64), (2, 11, 4, 16), (2, 41, 8, 32), (1, 64, 2, 64), (2, 7, 3, 8),
                                 (2, 65, 3, 64), (1, 200, 2, 64), (2, 129, 2, 32), (1, 321, 1, 64)])  # > 64: tiled kernels
def test_attention_fwd_bwd(lib, dtype, cfg):
    td = DT[dtype]
    B, N, H, dh = cfg
    if dtype == L.F32 and N > 140:
        pytest.skip('fp32 parity mode keeps the whole sequence in shared memory (N <= ~140)')
    inner = H * dh
    qkv = torch.randn(B * N, 3 * inner, device='cuda').to(td)
    d_o = torch.randn(B * N, inner, device='cuda').to(td)
    scale = dh ** -0.5
    o = torch.empty(B * N, inner, device='cuda', dtype=td)
    lse = torch.empty(B, H, N, device='cuda')
    L.check(lib.ecgvit_attention_fwd(qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), B, N, H, dh, scale, 0.0, 0, None, dtype,
                                     stream()), 'attn')
    ref_in = qkv.float().requires_grad_(True)
    q, k, v = (t.reshape(B, N, H, dh).permute(0, 2, 1, 3) for t in ref_in.chunk(3, dim=-1))
    s = (q @ k.transpose(-1, -2)) * scale
    want = (s.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(B * N, inner)
    tol = 2e-6 if dtype == L.F32 else 8e-3
    assert rel(o, want) < tol
    assert rel(lse, torch.logsumexp(s, -1)) < 1e-5 if dtype == L.F32 else True
    want.backward(d_o.float())
    dqkv = torch.empty_like(qkv)
    n_scr = int(lib.ecgvit_attention_bwd_scratch_floats(B, N, H, dh, dtype))
    scr = torch.empty(max(n_scr, 1), device='cuda')
    L.check(lib.ecgvit_attention_bwd(qkv.data_ptr(), o.data_ptr(), d_o.data_ptr(), lse.data_ptr(), dqkv.data_ptr(),
                                     scr.data_ptr() if n_scr else None, B, N,
                                     H, dh, scale, 0.0, 0, None, dtype, stream()), 'attn_bwd')
    assert rel(dqkv, ref_in.grad) < (1e-5 if dtype == L.F32 else 1.5e-2)


@pytest.mark.parametrize('dtype', [L.F32, L.BF16])
@pytest.mark.parametrize('reduction', ['mean', 'sum'])
@pytest.mark.parametrize('weighted', [False, True])
def test_head_fwd_bwd(lib, dtype, reduction, weighted):
    td = DT[dtype]
    table = torch.tensor([0.7, 3.5], device='cuda') if weighted else None  # EcgVit.loss_weight (ecg_vit.py:144-147)
    tp, nw = (table.data_ptr(), 2) if weighted else (None, 0)
    B, N, d, C = 9, 11, 256, 71
    tok = torch.randn(B * N, d, device='cuda').to(td)
    gamma, beta = torch.randn(d, device='cuda'), torch.randn(d, device='cuda')
    w, b = torch.randn(C, d, device='cuda') * 0.1, torch.randn(C, device='cuda')
    labels = (torch.rand(B, C, device='cuda') < 0.1).float()
    xn, mean, rstd = torch.empty(B, d, device='cuda'), torch.empty(B, device='cuda'), torch.empty(B, device='cuda')
    logits, loss = torch.empty(B, C, device='cuda'), torch.empty(1, device='cuda')
    red = L.REDUCTION[reduction]
    L.check(lib.ecgvit_head_fwd(tok.data_ptr(), gamma.data_ptr(), beta.data_ptr(), w.data_ptr(), b.data_ptr(),
                                labels.data_ptr(), tp, nw, xn.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                logits.data_ptr(), loss.data_ptr(), B, N, d, C, red, 1e-5, dtype, stream()), 'head')
    ew = table[labels.long()] if weighted else None
    tr = tok.float().requires_grad_(True)
    pr = [t.clone().requires_grad_(True) for t in (gamma, beta, w, b)]
    cls = tr.reshape(B, N, d)[:, 0]
    z = torch.nn.functional.linear(torch.nn.functional.layer_norm(cls, (d,), pr[0], pr[1], 1e-5), pr[2], pr[3])
    want_loss = torch.nn.functional.binary_cross_entropy_with_logits(z, labels, weight=ew, reduction=reduction)
    assert rel(logits, z) < 1e-5 and rel(loss[0], want_loss) < 1e-5
    want_loss.backward()
    dtok = torch.full((B * N, d), 7.0, device='cuda').to(td)
    grads = [torch.zeros_like(t) for t in (w, b, gamma, beta)]
    dcol = torch.zeros(d, device='cuda')
    scratch = torch.empty(B * d + B * C, device='cuda')
    L.check(lib.ecgvit_head_bwd(tok.data_ptr(), gamma.data_ptr(), w.data_ptr(), labels.data_ptr(), tp, nw,
                                xn.data_ptr(), mean.data_ptr(), rstd.data_ptr(), logits.data_ptr(), dtok.data_ptr(),
                                grads[0].data_ptr(), grads[1].data_ptr(), grads[2].data_ptr(), grads[3].data_ptr(),
                                dcol.data_ptr(), scratch.data_ptr(), B, N, d, C, red, 1.0, None, dtype, stream()), 'head_bwd')
    tol = 1e-5 if dtype == L.F32 else 5e-3
    assert rel(dtok, tr.grad) < tol
    assert rel(grads[0], pr[2].grad) < 1e-5 and rel(grads[1], pr[3].grad) < 1e-5
    assert rel(grads[2], pr[0].grad) < 1e-5 and rel(grads[3], pr[1].grad) < 1e-5
    assert rel(dcol, dtok.float().sum(0)) < 1e-5
    # 'none' reduction, forward only
    ln = torch.empty(B, C, device='cuda')
    L.check(lib.ecgvit_head_fwd(tok.data_ptr(), gamma.data_ptr(), beta.data_ptr(), w.data_ptr(), b.data_ptr(),
                                labels.data_ptr(), tp, nw, xn.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                logits.data_ptr(), ln.data_ptr(), B, N, d, C, L.REDUCTION['none'], 1e-5, dtype,
                                stream()), 'head')
    assert rel(ln, torch.nn.functional.binary_cross_entropy_with_logits(z, labels, weight=ew, reduction='none')) < 1e-5


@pytest.mark.parametrize('dtype', [L.F32, L.BF16])
def test_embed_assemble_and_colsum(lib, dtype):
    td = DT[dtype]
    B, n, d = 5, 10, 128
    e = torch.randn(B * n, d, device='cuda').to(td)
    cls, pos = torch.randn(d, device='cuda'), torch.randn(n + 1, d, device='cuda')
    tok = torch.empty(B * (n + 1), d, device='cuda', dtype=td)
    L.check(lib.ecgvit_embed_assemble(e.data_ptr(), cls.data_ptr(), pos.data_ptr(), tok.data_ptr(), B, n, d, 0.0, 0, None,
                                      dtype, stream()), 'assemble')
    want = torch.cat([cls.expand(B, 1, d), e.float().reshape(B, n, d)], 1) + pos
    assert rel(tok, want.reshape(-1, d)) < (1e-7 if dtype == L.F32 else 4e-3)
    dtok = torch.randn(B * (n + 1), d, device='cuda').to(td)
    de = torch.empty(B * n, d, device='cuda', dtype=td)
    dcls, dpos, dbias = torch.zeros(d, device='cuda'), torch.zeros(n + 1, d, device='cuda'), torch.zeros(d, device='cuda')
    L.check(lib.ecgvit_embed_assemble_bwd(dtok.data_ptr(), de.data_ptr(), dcls.data_ptr(), dpos.data_ptr(),
                                          dbias.data_ptr(), B, n, d, 0.0, 0, None, dtype, stream()), 'assemble_bwd')
    g = dtok.float().reshape(B, n + 1, d)
    assert torch.equal(de.reshape(B, n, d), dtok.reshape(B, n + 1, d)[:, 1:])
    assert rel(dpos, g.sum(0)) < 1e-6 and rel(dcls, g[:, 0].sum(0)) < 1e-6 and rel(dbias, g[:, 1:].sum((0, 1))) < 1e-5
    x = torch.randn(1234, 264, device='cuda').to(td)
    out = torch.ones(264, device='cuda')
    L.check(lib.ecgvit_colsum(x.data_ptr(), out.data_ptr(), 1234, 264, 264, dtype, stream()), 'colsum')
    assert rel(out, x.float().sum(0) + 1) < 1e-5


@pytest.mark.parametrize('max_norm', [1.0, 0.05, 0.0])
def test_clip_and_adamw_match_torch(lib, max_norm):
    n = 64 * 1000 + 192
    torch.manual_seed(3)
    p0, g0 = torch.randn(n, device='cuda'), torch.randn(n, device='cuda') * 0.01
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref], lr=3e-4, weight_decay=1e-2)
    p, m, v = p0.clone(), torch.zeros(n, device='cuda'), torch.zeros(n, device='cuda')
    shadow = torch.empty(n, device='cuda', dtype=torch.bfloat16)
    hyper, stats = torch.zeros(16, device='cuda'), torch.zeros(L.STATS_FLOATS, device='cuda')
    for t in range(1, 4):
        g = g0 * t
        ref.grad = g.clone()
        if max_norm > 0:
            norm = torch.nn.utils.clip_grad_norm_([ref], max_norm, error_if_nonfinite=True)
        else:
            norm = g.norm()
        opt.step()
        hyper.copy_(torch.tensor(L.adamw_hyper(3e-4, 0.9, 0.999, 1e-8, 1e-2, t, max_norm, 1.0)))
        L.check(lib.ecgvit_grad_sumsq(g.data_ptr(), L.F32, n, hyper.data_ptr(), stats.data_ptr(), stream()), 'sumsq')
        L.check(lib.ecgvit_adamw_step(p.data_ptr(), m.data_ptr(), v.data_ptr(), g.data_ptr(), L.F32, shadow.data_ptr(), n,
                                      hyper.data_ptr(), stats.data_ptr(), 0, stream()), 'adamw')
        assert abs(float(stats[2]) - float(norm)) < 1e-5 * float(norm)
        assert rel(p, ref.data) < 1e-6
        assert torch.equal(shadow, p.bfloat16())
    st = opt.state[ref]
    assert rel(m, st['exp_avg']) < 1e-5 and rel(v, st['exp_avg_sq']) < 1e-5


def test_adamw_skips_update_on_nonfinite_gradients(lib):
    n = 4096
    p0 = torch.randn(n, device='cuda')
    p, m, v = p0.clone(), torch.zeros(n, device='cuda'), torch.zeros(n, device='cuda')
    g = torch.randn(n, device='cuda')
    g[17] = float('inf')
    hyper, stats = torch.zeros(16, device='cuda'), torch.zeros(L.STATS_FLOATS, device='cuda')
    hyper.copy_(torch.tensor(L.adamw_hyper(3e-4, 0.9, 0.999, 1e-8, 1e-2, 1, 1.0, 1.0)))
    L.check(lib.ecgvit_grad_sumsq(g.data_ptr(), L.F32, n, hyper.data_ptr(), stats.data_ptr(), stream()), 'sumsq')
    L.check(lib.ecgvit_adamw_step(p.data_ptr(), m.data_ptr(), v.data_ptr(), g.data_ptr(), L.F32, None, n, hyper.data_ptr(),
                                  stats.data_ptr(), 0, stream()), 'adamw')
    assert torch.equal(p, p0) and float(m.abs().sum()) == 0.0
    assert float(stats[1]) == 1.0 and not math.isfinite(float(stats[2]))


def test_cast_shadow(lib):
    src = torch.randn(100003, device='cuda')
    dst = torch.empty(100003, device='cuda', dtype=torch.bfloat16)
    L.check(lib.ecgvit_cast_f32_to_bf16(src.data_ptr(), dst.data_ptr(), src.numel(), stream()), 'cast')
    assert torch.equal(dst, src.bfloat16())


# ---- fp32 residual stream (ECGVIT_BF16_RES32 / ECGVIT_EPI_BIAS_RES_F32) -----------------------------------------------
@pytest.mark.parametrize('shape', [(304, 256, 192), (1000, 768, 768), (2048, 1024, 4096), (333, 264, 600)])
def test_gemm_fp32_residual_epilogue(lib, shape):
    M, N, K = shape
    A, B, acc = _operands(M, N, K, 1, 1, torch.bfloat16, seed=M + 1)
    bias = torch.randn(N, device='cuda')
    aux = torch.randn(M, N, device='cuda') * 30.0          # a stream much larger than the update: bf16 would lose it
    out = torch.zeros(M, N, device='cuda')
    gemm(lib, A, B, M, N, K, 1, 1, L.EPI_BIAS_RES_F32, out, L.BF16, aux=aux, bias=bias)
    want = acc + bias + aux
    assert rel(out, want) < 2e-5, rel(out, want)   # fp32 accumulation order over K; a bf16 stream would sit at 2e-3
    assert float((out - want).abs().max()) < 1e-3 * float(acc.abs().max())
    # in place (out == aux), as the engine never does but the ABI allows with distinct tiles per warp
    with pytest.raises(RuntimeError):
        gemm(lib, A.float(), B.float(), M, N, K, 1, 1, L.EPI_BIAS_RES_F32, out, L.F32, aux=aux, bias=bias)


@pytest.mark.parametrize('d', [64, 256, 768, 1024])
def test_layernorm_fp32_stream_in_bf16_out(lib, d):
    M = 777
    x = torch.randn(M, d, device='cuda') * 2 + 0.5
    gamma, beta = torch.randn(d, device='cuda'), torch.randn(d, device='cuda')
    y = torch.empty(M, d, device='cuda', dtype=torch.bfloat16)
    mean, rstd = torch.empty(M, device='cuda'), torch.empty(M, device='cuda')
    L.check(lib.ecgvit_layernorm_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), mean.data_ptr(),
                                     rstd.data_ptr(), M, d, 1e-5, L.BF16_RES32, stream()), 'ln')
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (d,), gr, br, 1e-5)
    assert rel(y, yr) < 4e-3 and rel(mean, xr.mean(-1)) < 1e-5
    assert rel(rstd, 1.0 / torch.sqrt(xr.var(-1, unbiased=False) + 1e-5)) < 1e-5
    dy = torch.randn(M, d, device='cuda').bfloat16()
    dres = torch.randn(M, d, device='cuda').bfloat16()
    yr.backward(dy.float())
    dx = torch.empty(M, d, device='cuda', dtype=torch.bfloat16)
    dg, db, dc = (torch.zeros(d, device='cuda') for _ in range(3))
    scratch = torch.empty(int(lib.ecgvit_layernorm_bwd_scratch_floats(d)), device='cuda')
    L.check(lib.ecgvit_layernorm_bwd(dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                     dres.data_ptr(), dx.data_ptr(), dg.data_ptr(), db.data_ptr(), dc.data_ptr(),
                                     scratch.data_ptr(), None, 0.0, 0, None, M, d, 0, L.BF16_RES32, stream()), 'ln_bwd')
    want_dx = xr.grad + dres.float()
    assert rel(dx, want_dx) < 5e-3
    assert rel(dg, gr.grad) < 1e-4 and rel(db, br.grad) < 1e-4 and rel(dc, want_dx.sum(0)) < 1e-3


def test_embed_assemble_and_head_take_fp32_stream(lib):
    B, n, d, C = 5, 10, 128, 71
    e = torch.randn(B * n, d, device='cuda').bfloat16()
    cls, pos = torch.randn(d, device='cuda'), torch.randn(n + 1, d, device='cuda')
    tok = torch.empty(B * (n + 1), d, device='cuda')
    L.check(lib.ecgvit_embed_assemble(e.data_ptr(), cls.data_ptr(), pos.data_ptr(), tok.data_ptr(), B, n, d, 0.0, 0, None,
                                      L.BF16_RES32, stream()), 'assemble')
    want = torch.cat([cls.expand(B, 1, d), e.float().reshape(B, n, d)], 1) + pos
    assert rel(tok, want.reshape(-1, d)) < 1e-7
    # head on fp32 tokens == the fp32-mode head; its token gradient comes back in bf16
    gamma, beta = torch.randn(d, device='cuda'), torch.randn(d, device='cuda')
    w, b = torch.randn(C, d, device='cuda') * 0.1, torch.randn(C, device='cuda')
    labels = (torch.rand(B, C, device='cuda') < 0.1).float()
    N = n + 1
    outs = {}
    for code, td in ((L.F32, torch.float32), (L.BF16_RES32, torch.bfloat16)):
        xn, mean, rstd = torch.empty(B, d, device='cuda'), torch.empty(B, device='cuda'), torch.empty(B, device='cuda')
        logits, loss = torch.empty(B, C, device='cuda'), torch.empty(1, device='cuda')
        L.check(lib.ecgvit_head_fwd(tok.data_ptr(), gamma.data_ptr(), beta.data_ptr(), w.data_ptr(), b.data_ptr(),
                                    labels.data_ptr(), None, 0, xn.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                    logits.data_ptr(), loss.data_ptr(), B, N, d, C, 0, 1e-5, code, stream()), 'head')
        dtok = torch.full((B * N, d), 7.0, device='cuda').to(td)
        grads = [torch.zeros_like(t) for t in (w, b, gamma, beta)]
        dcol = torch.zeros(d, device='cuda')
        scratch = torch.empty(B * d + B * C, device='cuda')
        L.check(lib.ecgvit_head_bwd(tok.data_ptr(), gamma.data_ptr(), w.data_ptr(), labels.data_ptr(), None, 0,
                                    xn.data_ptr(), mean.data_ptr(), rstd.data_ptr(), logits.data_ptr(), dtok.data_ptr(),
                                    grads[0].data_ptr(), grads[1].data_ptr(), grads[2].data_ptr(), grads[3].data_ptr(),
                                    dcol.data_ptr(), scratch.data_ptr(), B, N, d, C, 0, 1.0, None, code, stream()), 'head_bwd')
        outs[code] = (logits, loss, dtok.float(), grads)
    a, bb = outs[L.F32], outs[L.BF16_RES32]
    assert torch.equal(a[0], bb[0]) and torch.equal(a[1], bb[1])
    assert rel(bb[2], a[2]) < 4e-3
    for ga, gb in zip(a[3], bb[3]):
        assert rel(gb, ga) < 1e-3


@pytest.mark.parametrize('epi', ['res', 'res_f32', 'dgelu'])
def test_gemm_aux_epilogues_many_row_tiles_partial_last_column_tile(lib, epi):
    """persistent loop: every cluster walks several tiles, and the last column tile is partial (N = 1024 on 192-wide tiles,
    the shape of the 'large' model's out-proj): the residual / pre-activation tiles of units outside the matrix are never
    fetched, so their barriers must not be waited on (nor their phase advanced)"""
    M, N, K = 256 * 90, 1024, 256
    A, B, acc = _operands(M, N, K, 1, 1 if epi != 'dgelu' else 0, torch.bfloat16, seed=17)
    bias = torch.randn(N, device='cuda')
    if epi == 'res_f32':
        aux = torch.randn(M, N, device='cuda')
        out = torch.zeros(M, N, device='cuda')
        gemm(lib, A, B, M, N, K, 1, 1, L.EPI_BIAS_RES_F32, out, L.BF16, aux=aux, bias=bias)
        assert rel(out, acc + bias + aux) < 2e-5
    elif epi == 'res':
        aux = torch.randn(M, N, device='cuda').bfloat16()
        out = torch.zeros(M, N, device='cuda', dtype=torch.bfloat16)
        gemm(lib, A, B, M, N, K, 1, 1, L.EPI_BIAS_RES, out, L.BF16, aux=aux, bias=bias)
        assert rel(out, acc + bias + aux.float()) < 4e-3
    else:
        aux = torch.randn(M, N, device='cuda').bfloat16()
        out = torch.zeros(M, N, device='cuda', dtype=torch.bfloat16)
        gemm(lib, A, B, M, N, K, 1, 0, L.EPI_DGELU, out, L.BF16, aux=aux)
        uu = aux.float().requires_grad_(True)
        torch.nn.functional.gelu(uu).sum().backward()
        assert rel(out, acc * uu.grad) < 5e-3


def test_clip_adamw_read_bf16_gradients(lib):
    """data-parallel runs all-reduce the gradients in bf16: norm and AdamW then read the bf16 image (grad_dtype BF16)"""
    n = 4096 * 5 + 8
    g32 = torch.randn(n, device='cuda') * 0.3
    g16 = g32.bfloat16()
    outs = []
    for g, code in ((g16.float(), L.F32), (g16, L.BF16)):
        p = torch.linspace(-1, 1, n, device='cuda')
        m, v = torch.zeros(n, device='cuda'), torch.zeros(n, device='cuda')
        shadow = torch.empty(n, device='cuda', dtype=torch.bfloat16)
        hyper = torch.tensor(L.adamw_hyper(1e-2, 0.9, 0.999, 1e-8, 0.1, 1, 0.5, 0.125), device='cuda', dtype=torch.float32)
        stats = torch.zeros(L.STATS_FLOATS, device='cuda')
        L.check(lib.ecgvit_grad_sumsq(g.data_ptr(), code, n, hyper.data_ptr(), stats.data_ptr(), stream()), 'sumsq')
        L.check(lib.ecgvit_adamw_step(p.data_ptr(), m.data_ptr(), v.data_ptr(), g.data_ptr(), code, shadow.data_ptr(), n,
                                      hyper.data_ptr(), stats.data_ptr(), 0, stream()), 'adamw')
        outs.append((p, m, v, shadow, float(stats[2])))
    for a, b in zip(outs[0][:4], outs[1][:4]):
        assert torch.equal(a, b)                      # bf16 -> fp32 is exact: same arithmetic either way
    want = float((g16.float() * 0.125).norm())
    assert abs(outs[1][4] - want) < 1e-5 * want


def test_grad_norm_slice_by_slice_matches_whole_buffer(lib):
    """ecgvit_grad_sumsq_partial over uneven slices issued in a scrambled order + ecgvit_grad_sumsq_finalize == the
    one-pass norm == torch's; twice the same bits (no atomics)"""
    torch.manual_seed(5)
    n = 3_000_004
    g = torch.randn(n, device='cuda') * 0.3
    hyper = torch.zeros(16, device='cuda')
    hyper.copy_(torch.tensor(L.adamw_hyper(3e-4, 0.9, 0.999, 1e-8, 1e-2, 1, 1.0, 0.5)))   # grad_scale 0.5
    bounds = [0, 1024, 700_000, 700_064, 2_200_000, n]          # 4-element aligned starts (16-byte aligned pointers)
    per = 37
    order = [3, 0, 4, 2, 1]
    got = []
    for _ in range(2):
        stats = torch.full((L.STATS_FLOATS,), float('nan'), device='cuda')
        for i in order:
            lo, hi = bounds[i], bounds[i + 1]
            L.check(lib.ecgvit_grad_sumsq_partial(g.data_ptr() + 4 * lo, L.F32, hi - lo, hyper.data_ptr(), stats.data_ptr(),
                                                  i * per, per, stream()), 'partial')
        L.check(lib.ecgvit_grad_sumsq_finalize(stats.data_ptr(), per * (len(bounds) - 1), stream()), 'finalize')
        torch.cuda.synchronize()
        got.append(stats[:3].clone())
    assert torch.equal(got[0], got[1])
    want = float((g.double() * 0.5).norm())
    assert abs(float(got[0][2]) - want) < 1e-5 * want
    assert float(got[0][1]) == 0.0
    whole = torch.zeros(L.STATS_FLOATS, device='cuda')
    L.check(lib.ecgvit_grad_sumsq(g.data_ptr(), L.F32, n, hyper.data_ptr(), whole.data_ptr(), stream()), 'sumsq')
    assert abs(float(whole[2]) - float(got[0][2])) < 1e-5 * want
    # a non-finite gradient anywhere raises the flag
    g[2_500_000] = float('inf')
    stats = torch.zeros(L.STATS_FLOATS, device='cuda')
    for i in order:
        lo, hi = bounds[i], bounds[i + 1]
        L.check(lib.ecgvit_grad_sumsq_partial(g.data_ptr() + 4 * lo, L.F32, hi - lo, hyper.data_ptr(), stats.data_ptr(),
                                              i * per, per, stream()), 'partial')
    L.check(lib.ecgvit_grad_sumsq_finalize(stats.data_ptr(), per * (len(bounds) - 1), stream()), 'finalize')
    assert float(stats[1]) == 1.0
