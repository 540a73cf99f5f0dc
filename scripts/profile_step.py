"""In-graph per-kernel device times of the training step via torch.profiler (CUPTI): what each kernel really costs
inside the CUDA-graph replay (warm L2, no launch gaps), unlike ncu's serialised cold-cache replays.
Usage: python scripts/profile_step.py [dropout] [out.json]"""
import collections
import json
import os
import re
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ecg_b200
from bench import BASE_CFG
from ecg_b200 import synthetic_batch

p = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
out = sys.argv[2] if len(sys.argv) > 2 else 'gpurun_out/profiler_step.json'
cfg = dict(BASE_CFG, hidden_dropout_prob=p, attention_probs_dropout_prob=p)
torch.manual_seed(77)
model = ecg_b200.EcgVit(config=ecg_b200.EcgVitConfig(compute_dtype='bf16', **cfg)).cuda().train()
tr = ecg_b200.FusedTrainer(model, use_cuda_graph=True)
x, y = synthetic_batch(256)
x, y = x.cuda(), y.cuda()
for _ in range(5):
    tr.step(x, y)
torch.cuda.synchronize()
STEPS = 5
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(STEPS):
        tr.step(x, y)
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for ev in prof.events():
    if ev.device_type != torch.autograd.DeviceType.CUDA:
        continue
    name = ev.name.replace('(anonymous namespace)::', '').replace('void ', '').replace('ecgvit::', '')
    name = re.sub(r'\(.*', '', name)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ev.device_time if hasattr(ev, 'device_time') else ev.cuda_time
tot = sum(v[1] for v in agg.values())
rows = [{'kernel': k, 'launches_per_step': v[0] / STEPS, 'us_per_step': v[1] / STEPS, 'avg_us': v[1] / v[0],
         'share': v[1] / tot} for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])]
os.makedirs(os.path.dirname(out) or '.', exist_ok=True)
json.dump({'dropout': p, 'kernel_time_us_per_step': tot / STEPS, 'kernels': rows}, open(out, 'w'), indent=1)
print(f'dropout {p}: sum of kernel times {tot / STEPS:.1f} us/step')
for r in rows[:30]:
    print(f"{r['us_per_step']:9.1f}us {r['launches_per_step']:6.1f}x {r['avg_us']:8.1f}us {r['share'] * 100:5.1f}%  {r['kernel'][:90]}")
