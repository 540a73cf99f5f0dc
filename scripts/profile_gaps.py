"""Timeline of one replayed step (torch.profiler chrome trace): per stream, the busy time, and the idle gaps between
consecutive kernels, grouped by the kernel that FOLLOWS the gap.  Run with the default engine settings (PDL and the
weight-gradient side stream on) to see where the step's wall time goes beyond the kernels themselves.
Usage: python scripts/profile_gaps.py [dropout]"""
import collections
import json
import os
import re
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ecg_b200
from bench import BASE_CFG
from ecg_b200 import synthetic_batch

p = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
cfg = dict(BASE_CFG, hidden_dropout_prob=p, attention_probs_dropout_prob=p)
torch.manual_seed(77)
model = ecg_b200.EcgVit(config=ecg_b200.EcgVitConfig(compute_dtype='bf16', **cfg)).cuda().train()
tr = ecg_b200.FusedTrainer(model, use_cuda_graph=True)
x, y = synthetic_batch(256)
x, y = x.cuda(), y.cuda()
for _ in range(5):
    tr.step(x, y)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        tr.step(x, y)
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), 'trace.json')
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))['traceEvents'] if e.get('cat') in ('kernel', 'gpu_memcpy', 'gpu_memset')]
ev.sort(key=lambda e: e['ts'])


def short(n):
    n = n.replace('(anonymous namespace)::', '').replace('void ', '').replace('ecgvit::', '')
    return re.sub(r'\(.*', '', n)[:60]


# the middle step: between the 1st and 2nd adamw kernels
adam = [i for i, e in enumerate(ev) if 'adamw' in e['name']]
lo, hi = adam[0] + 1, adam[1] + 1
step = ev[lo:hi]
t0, t1 = step[0]['ts'], step[-1]['ts'] + step[-1]['dur']
print(f'step span {t1 - t0:.1f} us, {len(step)} device activities')
streams = collections.defaultdict(list)
for e in step:
    streams[e['args'].get('stream', e.get('tid'))].append(e)
for s, es in streams.items():
    busy = sum(e['dur'] for e in es)
    print(f'stream {s}: {len(es)} activities, busy {busy:.1f} us')
# union busy time over all streams
iv = sorted((e['ts'], e['ts'] + e['dur']) for e in step)
union, cur0, cur1 = 0.0, iv[0][0], iv[0][1]
for a, b in iv[1:]:
    if a > cur1:
        union += cur1 - cur0
        cur0, cur1 = a, b
    else:
        cur1 = max(cur1, b)
union += cur1 - cur0
print(f'device busy (union over streams) {union:.1f} us -> idle {t1 - t0 - union:.1f} us')
main = max(streams.values(), key=len)
gaps = collections.defaultdict(lambda: [0, 0.0])
for a, b in zip(main, main[1:]):
    g = b['ts'] - (a['ts'] + a['dur'])
    k = f'{short(a["name"])} -> {short(b["name"])}'
    gaps[k][0] += 1
    gaps[k][1] += g
print('main-stream gaps by (previous -> next) kernel:')
for k, (n, g) in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f'{g:8.1f} us total {n:4d}x {g / n:6.2f} us  {k}')
print(f'sum of main-stream gaps {sum(g for _, g in gaps.values()):.1f} us')
