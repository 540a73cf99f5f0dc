"""Side measurements for the BASELINE.md results table (not the bench.py contract): the fused step of other named
configurations on ONE GPU, device-resident inputs, CUDA events around `steps` graph replays.
Usage: python scripts/bench_configs.py [name ...]   names: base_p0 base_p01 large_b512 large_b256 cfg1_fp32 cfg1_bf16"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ecg_b200
from ecg_b200 import synthetic_batch

CASES = {
    'base_p0': ('ecg-vit-base', 'bf16', 256, 0.0), 'base_p01': ('ecg-vit-base', 'bf16', 256, 0.1),
    'large_b512': ('ecg-vit-large', 'bf16', 512, 0.1), 'large_b256': ('ecg-vit-large', 'bf16', 256, 0.1),
    'cfg4_b32': ('cfg4', 'bf16', 32, 0.1), 'cfg4_b16_p0': ('cfg4', 'bf16', 16, 0.0),
    'cfg1_bf16': ('cfg1', 'bf16', 32, 0.1), 'cfg1_fp32': ('cfg1', 'fp32', 32, 0.1),
}
CFG1 = dict(hidden_size=256, num_hidden_layers=4, num_attention_heads=8, intermediate_size=1024)  # BASELINE configs[0]
FLOPS = {'ecg-vit-base': 26.370e9, 'ecg-vit-large': 93.299e9, 'cfg1': 1.025e9, 'cfg4': 1861.3e9}
names = sys.argv[1:] or list(CASES)
rows = []
for name in names:
    key, dtype, batch, p = CASES[name]
    conf = ecg_b200.EcgVitConfig(**CFG1) if key == 'cfg1' else ecg_b200.EcgVitConfig.from_defined(
        'ecg-vit-base' if key == 'cfg4' else key)
    conf.max_signal_length, conf.patch_size, conf.compute_dtype = 2500, 50, dtype
    if key == 'cfg4':  # BASELINE configs[3]: 12x5000, patch 25, per-lead tokens -> N = 2401
        conf.max_signal_length, conf.patch_size, conf.per_lead_tokens = 5000, 25, True
    conf.hidden_dropout_prob = conf.attention_probs_dropout_prob = p
    torch.manual_seed(77)
    model = ecg_b200.EcgVit(config=conf).cuda().train()
    tr = ecg_b200.FusedTrainer(model, use_cuda_graph=True, data_parallel=False)
    x, y = synthetic_batch(batch, length=conf.max_signal_length, seed=77)
    x, y = x.cuda(), y.cuda()
    for _ in range(5 if key != 'cfg4' else 2):
        tr.step(x, y)
    torch.cuda.synchronize()
    steps = 30 if key != 'cfg4' else 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss, _ = tr.step(x, y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    sps = batch / ms * 1e3
    row = dict(name=name, model=key, dtype=dtype, batch=batch, dropout=p, ms_per_step=round(ms, 3),
               samples_per_s=round(sps, 1), loss=float(loss),
               tflops=None if FLOPS[key] is None else round(sps * FLOPS[key] / 1e12, 1),
               mem_gb=round(torch.cuda.max_memory_allocated() / 2 ** 30, 2))
    print(json.dumps(row), flush=True)
    rows.append(row)
    del model, tr, x, y
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
os.makedirs('gpurun_out', exist_ok=True)
json.dump(rows, open('gpurun_out/bench_configs.json', 'w'), indent=1)
