"""CPU tier: host-side mirror of the reference interface (config, module tree / state_dict keys, schedules,
bucket layout) and the world_size-2 gradient exchange over gloo."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import GOLDEN_CFG, ROOT
import ecg_b200
from ecg_b200 import EcgVit, EcgVitConfig
from oracle.ecg_vit_oracle import OracleConfig, OracleEcgVit, lr_lambda


def test_config_fields_defaults_and_named_sizes():
    c = EcgVitConfig()
    assert (c.max_signal_length, c.patch_size, c.num_channels, c.hidden_size, c.num_hidden_layers,
            c.num_attention_heads, c.intermediate_size, c.hidden_dropout_prob, c.attention_probs_dropout_prob,
            c.num_class, c.size) == (2560, 64, 12, 512, 8, 8, 2048, 0.1, 0.1, 71, None)
    b = EcgVitConfig.from_defined('ecg-vit-base')
    assert (b.hidden_size, b.num_hidden_layers, b.num_attention_heads, b.intermediate_size, b.size) == \
        (768, 12, 12, 3072, 'base')
    assert (b.max_signal_length, b.patch_size) == (2560, 64)  # from_defined never touches these (ecg_vit.py:56-92)
    with pytest.raises(ValueError):
        EcgVitConfig.from_defined('ecg-vit-huge')


def test_state_dict_keys_shapes_and_round_trip_with_oracle():
    kw = dict(GOLDEN_CFG)
    ours = EcgVit(config=EcgVitConfig(**kw))
    ref = OracleEcgVit(config=OracleConfig(**kw))
    sd_o, sd_r = ours.state_dict(), ref.state_dict()
    assert list(sd_o.keys()) == list(sd_r.keys())
    assert all(sd_o[k].shape == sd_r[k].shape for k in sd_r)
    ours.load_state_dict(sd_r, strict=True)
    ref.load_state_dict(ours.state_dict(), strict=True)
    assert [n for n, _ in ours.named_parameters()] == [n for n, _ in ref.named_parameters()]


def test_base_model_key_count_and_param_count():
    c = EcgVitConfig.from_defined('ecg-vit-base')
    c.max_signal_length, c.patch_size = 2500, 50
    with torch.device('meta'):
        m = EcgVit(config=c)
    assert len(m.state_dict()) == 140
    assert sum(p.numel() for p in m.parameters()) == 85_584_455
    assert m.meta == {'name': 'EcgVit', 'input shape': '12 x 2500', '#patch': 50, '#layer': 12, '#head': 12}
    assert m.meta_str == '{nm=EcgVit, in-sp=12x2500, #p=50, #l=12, #h=12}'
    assert m.to_str() == 'EcgVit, base'


def test_short_name_map_is_a_bijection():
    m = EcgVit(config=EcgVitConfig(**GOLDEN_CFG))
    names = m._short_names()
    assert len(set(names.values())) == len(names)
    assert names['vit.transformer.layers.1.0.fn.to_qkv.weight'] == 'l1.qkv.w'
    assert names['vit.transformer.layers.0.0.fn.to_out.0.bias'] == 'l0.out.b'
    assert names['vit.transformer.layers.0.1.fn.net.3.weight'] == 'l0.ff2.w'
    assert names['vit.transformer.layers.0.1.norm.bias'] == 'l0.ln2.b'
    assert names['vit.mlp_head.0.weight'] == 'head.ln.w' and names['vit.mlp_head.1.bias'] == 'head.b'


def test_cpu_forward_fails_loudly():
    m = EcgVit(config=EcgVitConfig(**GOLDEN_CFG))
    with pytest.raises(RuntimeError, match='no CPU path'):
        m.eval()(torch.zeros(1, 12, 500))
    with pytest.raises(RuntimeError):
        m.vit(torch.zeros(1, 12, 1, 500))  # the parameter containers have no eager forward


def test_loss_reduction_property():
    m = EcgVit(config=EcgVitConfig(**dict(GOLDEN_CFG, hidden_dropout_prob=0.1)))
    m.loss_reduction = 'none'
    assert m.loss_reduction == 'none'


def test_dropout_mask_replica_statistics():
    """host replica of the kernels' counter-based dropout: keep rate, scale, determinism, stream independence"""
    idx = torch.arange(0, 200000)
    a = ecg_b200._lib.dropout_keep_mask(1234, 7, 0.1, idx)
    b = ecg_b200._lib.dropout_keep_mask(1234, 7, 0.1, idx)
    c = ecg_b200._lib.dropout_keep_mask(1234, 8, 0.1, idx)
    assert torch.equal(a, b) and not torch.equal(a, c)
    keep = float((a > 0).float().mean())
    assert abs(keep - 0.9) < 5e-3
    assert abs(float(a.max()) - 1.0 / (1.0 - 6554 / 65536)) < 1e-6
    assert abs(float(a.mean()) - 1.0) < 1e-2  # unbiased
    assert torch.equal(ecg_b200._lib.dropout_keep_mask(1, 0, 0.0, idx[:10]), torch.ones(10))


def test_lr_multiplier_matches_oracle_schedule():
    for sched in ('constant', 'cosine'):
        for s in range(0, 60, 3):
            assert ecg_b200.lr_multiplier(sched, s, 5, 50) == lr_lambda(sched, s, 5, 50)


def test_train_args_defaults():
    a = ecg_b200.get_train_args(n_train=17441)
    assert (a['learning_rate'], a['weight_decay'], a['warmup_ratio'], a['schedule'], a['train_batch_size']) == \
        (3e-4, 1e-2, 0.05, 'cosine', 64)
    assert a['steps_per_epoch'] == 17441 // 64 and a['n_step'] == 3 * (17441 // 64)
    with pytest.raises(ValueError):
        ecg_b200.get_train_args(dict(schedule='linear'))


def test_bucket_slices_cover_the_flat_buffer_in_backward_order():
    from ecg_b200.parallel import bucket_slices
    offs = [100, 200, 300, 400, 500]
    sl = bucket_slices(offs, 650, 5, 2)
    assert sl == [(3, 400, 650), (1, 200, 400), (0, 100, 200), (-1, 0, 100)]
    covered = sorted((lo, hi) for _, lo, hi in sl)
    assert covered[0][0] == 0 and covered[-1][1] == 650
    assert all(a[1] == b[0] for a, b in zip(covered, covered[1:]))


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["ECG_ROOT"])
from ecg_b200.parallel import BucketedGradReducer
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["ECG_PORT"],
                        rank=int(os.environ["ECG_RANK"]), world_size=2)
rank = dist.get_rank()
torch.manual_seed(100 + rank)
flat = torch.randn(650)
mine = flat.clone()
red = BucketedGradReducer(flat_g=flat, layer_offsets=[100, 200, 300, 400, 500], depth=5, bucket_layers=2)
red.begin()
for layer in (4, 3, 2, 1, 0, -1):
    red.on_layer_done(layer)
red.finish()
torch.manual_seed(100 + (1 - rank))
other = torch.randn(650)
assert torch.allclose(flat, mine + other), "bucketed all-reduce != sum of both ranks"
gathered = [torch.zeros(650) for _ in range(2)]
dist.all_gather(gathered, flat)
assert torch.equal(gathered[0], gathered[1]), "ranks disagree after the exchange"
dist.destroy_process_group()
print("ok", rank)
'''


def test_bucketed_reducer_world_size_2_gloo(tmp_path):
    import socket
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / 'worker.py'
    script.write_text(_WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, ECG_ROOT=ROOT, ECG_PORT=str(port), ECG_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), '\n'.join(outs)


def test_input_pipeline_host_side_matches_reference_draws():
    """InputPipeline.padded_length / draw_spans replay TimeEndPad's length rule and TimeOut's RNG calls
    (transform.py:147-151,180-183): under the seed the golden vectors were made with, the spans are identical"""
    import numpy as np
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'input_transform.npz'))
    for case in ('ragged', 'full_block', 'long'):
        rec, k = g[f'{case}/records'], int(g[f'{case}/k'])
        pipe = ecg_b200.InputPipeline(normalize=dict(mean=g['mean'], std=g['std']), pad=k, timeout=True)
        Lp = pipe.padded_length(rec.shape[-1])
        assert Lp == g[f'{case}/eval'].shape[-1]
        torch.manual_seed(1234)
        spans = pipe.draw_spans(rec.shape[0], Lp)
        assert spans.dtype == torch.int32 and np.array_equal(spans.numpy(), g[f'{case}/spans'])
    assert ecg_b200.InputPipeline(pad=50).padded_length(2500) == 2550
    assert ecg_b200.InputPipeline().padded_length(2500) == 2500
    with pytest.raises(AssertionError):
        ecg_b200.InputPipeline(pad=True)
