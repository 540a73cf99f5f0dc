"""GPU tier: the trainer shell around the fused step -- the re-hosted `MyTrainer.train` loop (fit / evaluate / patience /
save / resume) against the oracle's trainer, `torch.optim`-style checkpointing of FusedAdamW, the autograd bridge's
guards, the sticky non-finite counter and what the captured CUDA graph is keyed on."""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from ecg_b200 import EcgVit, EcgVitConfig, FusedTrainer, FusedAdamW, clip_grad_norm_, fused_train_step
from oracle.ecg_vit_oracle import OracleConfig, OracleEcgVit, OracleTrainer, synthetic_batch

from test_gpu_parity import rel, CFG_REFDEFAULT, FP32_TOL

CFG = dict(CFG_REFDEFAULT)   # 12 x 2560, patch 64, d128, 2 layers, dropout 0


def pair(dtype='fp32', seed=7, **over):
    cfg = dict(CFG, **over)
    torch.manual_seed(seed)
    oracle = OracleEcgVit(config=OracleConfig(**cfg)).train()
    model = EcgVit(config=EcgVitConfig(compute_dtype=dtype, **cfg))
    model.load_state_dict(oracle.state_dict(), strict=True)
    return oracle, model.cuda().train()


def epoch_batches(n, bsz, seed, first_step):
    """the order FusedTrainer.fit draws for an epoch that starts at optimizer step `first_step`"""
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(seed + first_step))
    return [perm[i:i + bsz] for i in range(0, n, bsz)]


def test_fit_two_epochs_and_resume_track_the_oracle_trainer(tmp_path):
    """2 epochs of fit() (cosine schedule with warm-up, clipping active, a ragged last batch), interrupted after epoch 1
    and resumed in a fresh process-like state (new model + new trainer from the saved files), against the ORACLE's
    trainer fed the same batches: parameters equal to fp32 tolerance, eval loss equal, optimizer state really restored"""
    n, bsz, seed = 22, 8, 5
    x, y = synthetic_batch(n, length=2560, seed=31)
    xe, ye = synthetic_batch(10, length=2560, seed=32)
    steps = 2 * math.ceil(n / bsz)
    kw = dict(learning_rate=1e-3, weight_decay=1e-2, schedule='cosine', n_warmup=2, n_step=steps, max_grad_norm=0.05)
    oracle, model = pair()
    ot = OracleTrainer(oracle, **kw)
    tr = FusedTrainer(model, use_cuda_graph=True, data_parallel=False, **kw)
    logs = tr.fit((x, y), (xe, ye), num_train_epoch=1, train_batch_size=bsz, patience=8, output_dir=str(tmp_path),
                  save_every_n_epoch=1, log_every=2, shuffle_seed=seed)
    assert os.path.exists(tmp_path / 'model - ep1.pt') and os.path.exists(tmp_path / 'trainer - ep1.pt')
    o_losses = []
    for idx in epoch_batches(n, bsz, seed, 0):
        o_losses.append(float(ot.step(x[idx], y[idx])[0]))
    train_logs = [r for r in logs if 'train/loss' in r]
    assert [r['step'] for r in train_logs] == [2, 3]          # every 2nd step and the epoch's last
    for r in train_logs:
        assert abs(r['train/loss'] - o_losses[r['step'] - 1]) < 1e-5 * o_losses[r['step'] - 1]
    ev = [r for r in logs if 'eval/loss' in r][0]
    oracle.eval()
    with torch.no_grad():
        want = float(oracle(sample_values=xe[:8], labels=ye[:8]).loss + oracle(sample_values=xe[8:], labels=ye[8:]).loss) / 2
    oracle.train()
    assert abs(ev['eval/loss'] - want) < 1e-5 * want and 0.0 <= ev['eval/binary_accuracy'] <= 1.0
    # ---- resume from the files in a fresh model / trainer
    _, model2 = pair(seed=99)                                   # different init: everything must come from the files
    tr2 = FusedTrainer(model2, use_cuda_graph=True, data_parallel=False, learning_rate=123.0)
    tr2.resume(str(tmp_path), 'ep1')
    assert tr2.step_count == 3 and tr2.schedule == 'cosine' and tr2.lr == 1e-3 and tr2.n_step == steps
    assert tr2.max_grad_norm == 0.05
    tr2.fit((x, y), (xe, ye), num_train_epoch=1, train_batch_size=bsz, shuffle_seed=seed)
    for idx in epoch_batches(n, bsz, seed, 3):
        ot.step(x[idx], y[idx])
    for k, v in model2.state_dict().items():
        assert rel(v, oracle.state_dict()[k]) < 2e-5, k
    # a resume that forgot the optimizer state would not match: moments matter after 3 steps
    _, model3 = pair(seed=99)
    model3.load_state_dict(torch.load(tmp_path / 'model - ep1.pt'), strict=True)
    tr3 = FusedTrainer(model3, data_parallel=False, **kw)
    tr3.step_count = 3
    for idx in epoch_batches(n, bsz, seed, 3):
        tr3.step(x[idx].cuda(), y[idx].cuda())
    assert max(rel(v, oracle.state_dict()[k]) for k, v in model3.state_dict().items()) > 1e-4


def test_fit_stops_early_on_patience():
    _, model = pair()
    x, y = synthetic_batch(16, length=2560, seed=1)
    tr = FusedTrainer(model, data_parallel=False)
    logs = tr.fit((x, y), (x[:8], y[:8]), num_train_epoch=5, train_batch_size=8, patience=0)
    assert logs[-1].get('early_stop') and logs[-1]['epoch'] == 1 and tr.step_count == 2


def test_fused_adamw_state_dict_round_trip_resumes_exactly():
    """torch.save(optimizer.state_dict()) style checkpoint of the drop-in optimizer: moments and step come back"""
    _, m1 = pair()
    x, y = synthetic_batch(6, length=2560, seed=2)
    xd, yd = x.cuda(), y.cuda()

    def loop_step(model, opt):
        opt.zero_grad()
        model(sample_values=xd, labels=yd).loss.backward()
        clip_grad_norm_(model, 1.0)
        opt.step()

    o1 = FusedAdamW(m1, lr=1e-3, weight_decay=1e-2)
    for _ in range(3):
        loop_step(m1, o1)
    sd_m, sd_o = {k: v.clone() for k, v in m1.state_dict().items()}, o1.state_dict()
    assert sd_o['flat']['step'] == 3 and float(sd_o['flat']['exp_avg_sq'].sum()) > 0
    loop_step(m1, o1)
    _, m2 = pair(seed=5)
    m2.load_state_dict(sd_m, strict=True)
    o2 = FusedAdamW(m2, lr=1e-3, weight_decay=1e-2)
    o2.load_state_dict(sd_o)
    loop_step(m2, o2)
    assert rel(m2._flat_p, m1._flat_p) < 1e-6      # (not bitwise: the split-K weight gradients add in any order)
    # without the optimizer state the step differs (bias correction restarts, moments are zero)
    _, m3 = pair(seed=5)
    m3.load_state_dict(sd_m, strict=True)
    o3 = FusedAdamW(m3, lr=1e-3, weight_decay=1e-2)
    loop_step(m3, o3)
    assert rel(m3._flat_p, m1._flat_p) > 1e-4


def test_fused_adamw_leaves_parameters_without_gradient_untouched():
    oracle, model = pair()
    x, y = synthetic_batch(4, length=2560, seed=3)
    frozen = model.vit.mlp_head[1].weight
    frozen.requires_grad_(False)
    oracle.vit.mlp_head[1].weight.requires_grad_(False)
    before = frozen.detach().clone()
    opt = FusedAdamW(model, lr=1e-2, weight_decay=0.1)
    ref = torch.optim.AdamW([p for p in oracle.parameters()], lr=1e-2, weight_decay=0.1)
    for _ in range(2):
        opt.zero_grad()
        model(sample_values=x.cuda(), labels=y.cuda()).loss.backward()
        assert frozen.grad is None
        opt.step()
        ref.zero_grad()
        oracle(sample_values=x, labels=y).loss.backward()
        ref.step()
    assert torch.equal(frozen.detach(), before)                 # no weight decay, no update: torch's behaviour
    for k, v in model.state_dict().items():
        assert rel(v, oracle.state_dict()[k]) < 1e-5, k


def test_backward_after_a_second_forward_of_the_same_shape_raises():
    _, model = pair()
    x, y = synthetic_batch(4, length=2560, seed=4)
    xd, yd = x.cuda(), y.cuda()
    l1 = model(sample_values=xd, labels=yd).loss
    l2 = model(sample_values=xd, labels=yd).loss      # same shape: overwrites the cached workspace
    with pytest.raises(RuntimeError, match='another forward of the same input shape'):
        l1.backward()
    l2.backward()                                     # its own backward is fine


def test_backward_survives_a_forward_of_another_shape_and_redraws_its_own_dropout_masks():
    cfg = dict(hidden_dropout_prob=0.2, attention_probs_dropout_prob=0.2)
    _, model = pair(dtype='fp32', **cfg)
    x, y = synthetic_batch(6, length=2560, seed=6)
    xd, yd = x.cuda(), y.cuda()
    model._engine_seed = None
    out = model(sample_values=xd, labels=yd)
    seed_a = model._engine.last_seed
    model(sample_values=xd[:3], labels=yd[:3])        # monitoring pass on another batch size: new seed, other workspace
    assert model._engine.last_seed != seed_a
    out.loss.backward()
    got = model._flat_g.clone()
    assert model._engine.last_seed == seed_a          # the backward put its own seed back
    # reference: the same forward / backward with nothing in between
    model.zero_grad()
    model._engine.upload_seed(*seed_a)
    loss, _ = model._engine.forward(xd, yd, 'mean')
    model._engine.backward()
    assert rel(got, model._flat_g) < 1e-6


def test_nonfinite_gradient_is_counted_and_reported_at_a_later_poll():
    _, model = pair()
    x, y = synthetic_batch(4, length=2560, seed=8)
    tr = FusedTrainer(model, data_parallel=False)
    tr.step(x.cuda(), y.cuda())
    tr.check_finite()
    p_before = model._flat_p.clone()
    bad = x.clone()
    bad[0, 0, 0] = float('inf')
    tr.step(bad.cuda(), y.cuda())                      # skipped on the device
    assert torch.equal(model._flat_p, p_before)
    tr.step(x.cuda(), y.cuda())                        # a good step afterwards must not hide it
    with pytest.raises(RuntimeError, match='non-finite'):
        tr.check_finite()
    tr.check_finite()                                  # cleared by the poll


def test_captured_graph_is_keyed_on_what_it_bakes_in():
    _, model = pair(dtype='bf16', hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1)
    x, y = synthetic_batch(4, length=2560, seed=9)
    tr = FusedTrainer(model, use_cuda_graph=True, data_parallel=False)
    tr.step(x.cuda(), y.cuda())
    g1 = tr._graph
    tr.step(x.cuda(), y.cuda())
    assert tr._graph is g1
    model.config.hidden_dropout_prob = 0.0              # a different dropout wiring is a different kernel sequence
    tr.step(x.cuda(), y.cuda())
    assert tr._graph is not g1
    g2 = tr._graph
    model.loss_reduction = 'sum'
    tr.step(x.cuda(), y.cuda())
    assert tr._graph is not g2


def test_fused_train_step_keeps_its_trainer_on_the_model():
    _, model = pair()
    x, y = synthetic_batch(4, length=2560, seed=10)
    out = fused_train_step(model, {'sample_values': x, 'labels': y}, lr=1e-3)
    assert out.loss.ndim == 0 and model._fused_trainer.step_count == 1
    fused_train_step(model, {'sample_values': x, 'labels': y}, lr=1e-3)
    assert model._fused_trainer.step_count == 2


@pytest.mark.parametrize('graph', [False, True])
def test_deferred_optimizer_is_the_same_training_run(graph):
    """defer_optimizer=True moves clip + AdamW of step k next to the forward pass of step k + 1 (slice by slice, each
    layer waiting for its own weights): same losses, same gradient norms, and after flush() the same parameters and
    moments as the immediate trainer -- cosine schedule with warm-up and an active clip, so a shifted hyper-parameter
    block or a double-applied update would show"""
    kw = dict(learning_rate=1e-3, weight_decay=1e-2, schedule='cosine', n_warmup=2, n_step=6, max_grad_norm=0.05)
    _, m_now = pair()
    _, m_def = pair()
    t_now = FusedTrainer(m_now, use_cuda_graph=graph, data_parallel=False, **kw)
    t_def = FusedTrainer(m_def, use_cuda_graph=graph, data_parallel=False, defer_optimizer=True, **kw)
    x, y = synthetic_batch(6, length=2560, seed=12)
    xd, yd = x.cuda(), y.cuda()
    for step in range(5):
        l_now, _ = t_now.step(xd, yd)
        l_def, _ = t_def.step(xd, yd)
        assert rel(l_def, l_now) < 1e-6, (step, float(l_def), float(l_now))
        assert abs(t_def.grad_norm() - t_now.grad_norm()) < 1e-5 * t_now.grad_norm()
        if step == 2:
            # reading the weights mid-run applies the pending update (and must not apply it twice afterwards)
            sd = m_def.state_dict()
            assert rel(sd['vit.pos_embedding'], m_now.state_dict()['vit.pos_embedding']) < 1e-6
    assert t_def._pending_step == 5
    assert rel(m_def._flat_p, m_now._flat_p) > 1e-6          # the last update is still pending ...
    t_def.flush()
    assert rel(m_def._flat_p, m_now._flat_p) < 1e-6          # ... and flush() applies exactly it
    assert rel(t_def.exp_avg, t_now.exp_avg) < 1e-5 and rel(t_def.exp_avg_sq, t_now.exp_avg_sq) < 1e-5
    if m_def._shadow is not None:
        assert torch.equal(m_def._shadow, m_def._flat_p.bfloat16())
    t_def.check_finite()
    t_def.flush()                                             # idempotent
    assert rel(m_def._flat_p, m_now._flat_p) < 1e-6
