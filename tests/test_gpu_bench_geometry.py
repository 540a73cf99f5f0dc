"""GPU tier: the configurations the BENCH / BASELINE.json lines are quoted on, against the CPU oracle.

 * configs[1] geometry -- base d768 / 12 layers / 12 heads, 12 x 2500, patch 50 (N = 51 tokens): the exact shape
   `bench.py` times, bf16 (north_star bar: logits / loss <= 1e-2 relative, every gradient cosine >= 0.999) and fp32
   parity mode (<= 1e-5 relative);
 * configs[3] geometry -- 12 x 5000, patch 25, per-lead tokens (N = 2401), d768 / 12 heads, 2 layers, batch 1: the
   tiled attention kernels walking all their key tiles, against the oracle's materialised N x N softmax;
 * the contraction shapes of one cfg2 step (M = 13 056 token rows; K = 13 056 for the weight gradients) one by one
   against a plain fp32 torch matmul of the same bf16 operands.
"""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu

from ecg_b200 import EcgVit, EcgVitConfig, FusedTrainer
from ecg_b200 import _lib as L
from oracle.ecg_vit_oracle import OracleConfig, OracleEcgVit, OracleTrainer, synthetic_batch

from test_gpu_parity import rel, cosine, FP32_TOL, BF16_TOL, BF16_GRAD_COS

BASE_2500 = dict(max_signal_length=2500, patch_size=50, num_channels=12, hidden_size=768, num_hidden_layers=12,
                 num_attention_heads=12, intermediate_size=3072, hidden_dropout_prob=0.0,
                 attention_probs_dropout_prob=0.0)
CFG4 = dict(max_signal_length=5000, patch_size=25, num_channels=12, hidden_size=768, num_hidden_layers=2,
            num_attention_heads=12, intermediate_size=3072, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
            per_lead_tokens=True)


def make(cfg, dtype, batch, seed=77):
    torch.manual_seed(seed)
    oracle = OracleEcgVit(config=OracleConfig(**cfg)).train()
    model = EcgVit(config=EcgVitConfig(compute_dtype=dtype, **cfg))
    model.load_state_dict(oracle.state_dict(), strict=True)
    model.cuda().train()
    x, y = synthetic_batch(batch, num_channels=cfg['num_channels'], length=cfg['max_signal_length'], seed=seed)
    return oracle, model, x, y


@pytest.fixture(scope='module')
def base_oracle_run():
    """one oracle forward + backward of the bench model on 8 seeded records, shared by the bf16 and fp32 tests"""
    torch.manual_seed(77)
    oracle = OracleEcgVit(config=OracleConfig(**BASE_2500)).train()
    x, y = synthetic_batch(8, length=2500, seed=77)
    out = oracle(sample_values=x, labels=y)
    out.loss.backward()
    return oracle, x, y, out


@pytest.mark.parametrize('dtype', ['bf16', 'fp32'])
def test_bench_geometry_forward_backward_vs_oracle(base_oracle_run, dtype):
    oracle, x, y, want = base_oracle_run
    model = EcgVit(config=EcgVitConfig(compute_dtype=dtype, **BASE_2500))
    model.load_state_dict(oracle.state_dict(), strict=True)
    model.cuda().train()
    assert model.vit.pos_embedding.shape == (1, 51, 768)
    got = model(sample_values=x.cuda(), labels=y.cuda())
    got.loss.backward()
    tol = BF16_TOL if dtype == 'bf16' else FP32_TOL
    assert rel(got.logits, want.logits) < tol, rel(got.logits, want.logits)
    assert rel(got.loss, want.loss) < tol, rel(got.loss, want.loss)
    worst = min((cosine(p.grad, q.grad), k) for (k, p), (_, q) in zip(model.named_parameters(), oracle.named_parameters()))
    assert worst[0] >= (BF16_GRAD_COS if dtype == 'bf16' else 0.99999), worst
    if dtype == 'fp32':
        for (k, p), (_, q) in zip(model.named_parameters(), oracle.named_parameters()):
            assert rel(p.grad, q.grad) < 1e-3, (k, rel(p.grad, q.grad))


def test_bench_geometry_three_fused_steps_vs_oracle_trainer():
    """the captured (CUDA graph) step of bench.py, bf16, three optimisation steps against the oracle's trainer"""
    oracle, model, x, y = make(BASE_2500, 'bf16', 8)
    ot = OracleTrainer(oracle, learning_rate=3e-4, weight_decay=1e-2, schedule='constant', max_grad_norm=1.0)
    tr = FusedTrainer(model, learning_rate=3e-4, weight_decay=1e-2, schedule='constant', max_grad_norm=1.0,
                      use_cuda_graph=True, data_parallel=False)
    xd, yd = x.cuda(), y.cuda()
    for _ in range(3):
        o_loss, _, o_norm = ot.step(x, y)
        loss, _ = tr.step(xd, yd)
        assert rel(loss, o_loss) < 2 * BF16_TOL, (float(loss), float(o_loss))
        assert abs(tr.grad_norm() - float(o_norm)) < 5e-2 * float(o_norm)
    tr.check_finite()


def test_cfg4_long_signal_per_lead_tokens_vs_oracle():
    """N = 2401 tokens (12 leads x 200 windows + CLS), d768, 12 heads of 64, 2 layers, one record"""
    oracle, model, x, y = make(CFG4, 'bf16', 1)
    assert model.vit.pos_embedding.shape == (1, 2401, 768)
    want = oracle(sample_values=x, labels=y)
    want.loss.backward()
    got = model(sample_values=x.cuda(), labels=y.cuda())
    got.loss.backward()
    assert rel(got.logits, want.logits) < BF16_TOL, rel(got.logits, want.logits)
    assert rel(got.loss, want.loss) < BF16_TOL
    worst = min((cosine(p.grad, q.grad), k) for (k, p), (_, q) in zip(model.named_parameters(), oracle.named_parameters()))
    assert worst[0] >= BF16_GRAD_COS, worst


# ---- the contractions of one cfg2 step at their real sizes -------------------------------------------------------
M_TOK = 256 * 51  # 13 056 token rows


def _gemm(lib, A, B, M, N, K, a_k, b_k, epi, out, aux=None, out2=None, bias=None, split_k=1):
    g = L.GemmArgs(M, N, K, A.data_ptr(), A.stride(0), a_k, B.data_ptr(), B.stride(0), b_k, epi, out.data_ptr(),
                   out.stride(0), L.ptr(out2), L.ptr(aux), L.ptr(bias), L.BF16, split_k, 0, 0.0, None)
    L.check(lib.ecgvit_gemm(ctypes.byref(g), torch.cuda.current_stream().cuda_stream), 'gemm')


def _rand(shape, seed):
    g = torch.Generator(device='cuda').manual_seed(seed)
    return torch.randn(shape, device='cuda', generator=g).bfloat16()


@pytest.mark.parametrize('N,K', [(2304, 768), (768, 768), (3072, 768), (768, 3072), (768, 600)])
def test_step_gemm_forward_shapes(N, K):
    """y = x W^T (+ bias): QKV, out-proj, FF1, FF2 and the patch embedding at M = 13 056 (12 800 for the embedding)"""
    lib = L.load()
    M = M_TOK if K != 600 else 256 * 50
    X, W = _rand((M, K), 1), _rand((N, K), 2)
    bias = torch.randn(N, device='cuda')
    out = torch.empty(M, N, device='cuda', dtype=torch.bfloat16)
    _gemm(lib, X, W, M, N, K, 1, 1, L.EPI_STORE, out, bias=bias)
    want = X.float() @ W.float().t() + bias
    assert rel(out, want) < 4e-3, rel(out, want)
    assert float((out.float() - want).abs().max()) <= 0.02 * float(want.abs().max())


@pytest.mark.parametrize('N,K', [(768, 2304), (768, 768), (768, 3072), (3072, 768)])
def test_step_gemm_dgrad_shapes(N, K):
    """dx = dy W (W stored [K_out, N] row-major, read MN-major)"""
    lib = L.load()
    dY, W = _rand((M_TOK, K), 3), _rand((K, N), 4)
    out = torch.empty(M_TOK, N, device='cuda', dtype=torch.bfloat16)
    _gemm(lib, dY, W, M_TOK, N, K, 1, 0, L.EPI_STORE, out)
    want = dY.float() @ W.float()
    assert rel(out, want) < 4e-3, rel(out, want)


@pytest.mark.parametrize('n_out,k_in', [(3072, 768), (768, 3072), (2304, 768), (768, 768), (768, 600)])
def test_step_gemm_wgrad_shapes_auto_split(n_out, k_in):
    """dW[n_out, k_in] += dY^T X with the contraction over all 13 056 token rows, split-K chosen by the library
    (fp32 TMA reduce-add into the gradient buffer): the largest single share of the step"""
    lib = L.load()
    rows = M_TOK if k_in != 600 else 256 * 50
    dY, X = _rand((rows, n_out), 5), _rand((rows, k_in), 6)
    out = torch.full((n_out, k_in), 0.5, device='cuda')
    _gemm(lib, dY, X, n_out, k_in, rows, 0, 0, L.EPI_ATOMIC_F32, out, split_k=0)
    want = dY.float().t() @ X.float() + 0.5
    assert rel(out, want) < 2e-5, rel(out, want)


def test_step_gemm_gelu_pair_shapes():
    """FF1 forward (bias + GELU, two outputs) and FF2 dgrad (x GELU') at M = 13 056"""
    lib = L.load()
    X, W1 = _rand((M_TOK, 768), 7), _rand((3072, 768), 8) * 0.05
    bias = torch.randn(3072, device='cuda') * 0.1
    u = torch.empty(M_TOK, 3072, device='cuda', dtype=torch.bfloat16)
    h = torch.empty_like(u)
    _gemm(lib, X, W1, M_TOK, 3072, 768, 1, 1, L.EPI_BIAS_GELU, u, out2=h, bias=bias)
    want_u = X.float() @ W1.float().t() + bias
    assert rel(u, want_u) < 4e-3
    assert rel(h, torch.nn.functional.gelu(u.float())) < 4e-3
    dZ, W2 = _rand((M_TOK, 768), 9), _rand((768, 3072), 10) * 0.05
    du = torch.empty_like(u)
    _gemm(lib, dZ, W2, M_TOK, 3072, 768, 1, 0, L.EPI_DGELU, du, aux=u)
    uu = u.float().requires_grad_(True)
    torch.nn.functional.gelu(uu).sum().backward()
    want = (dZ.float() @ W2.float()) * uu.grad
    assert rel(du, want) < 5e-3, rel(du, want)
