"""Turns the raw ncu outputs brought back in gpurun_out/ into the small, tracked summaries under profiles/.
  python scripts/summarize_profiles.py step gpurun_out/step_metrics.csv profiles/<name>.json "<command that produced it>"
  python scripts/summarize_profiles.py full gpurun_out/prof.ncu-rep profiles/<name>.json "<command>"
"""
import collections
import csv
import json
import re
import subprocess
import sys


def clean(name):
    name = name.replace('void ', '').replace('ecgvit::<unnamed>::', '').replace('unnamed>::', '')
    return re.sub(r'\(.*', '', name)


def step(path, out, cmd):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr = rows[start]
    ki, mi, vi, ii = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
    per = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= vi:
            continue
        d = per.setdefault(r[ii], {'name': clean(r[ki])})
        d[r[mi]] = float(r[vi].replace(',', ''))
    agg = collections.OrderedDict()
    for d in per.values():
        a = agg.setdefault(d['name'], collections.Counter())
        a['launches'] += 1
        a['time_us'] += d.get('gpu__time_duration.sum', 0) / 1000
        a['dram'] += d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)
        a['tensor'] += d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0)
        a['dram_pct'] += d.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 0)
    tot = sum(a['time_us'] for a in agg.values())
    res = {'command': cmd, 'note': 'ncu per-launch times are cold-cache and serialised: compare SHARES, not absolutes',
           'launches': len(per), 'total_time_us': round(tot, 1), 'kernels': []}
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['time_us']):
        n = a['launches']
        res['kernels'].append({'kernel': k, 'launches': n, 'time_us': round(a['time_us'], 1), 'share': round(a['time_us'] / tot, 4),
                               'avg_us': round(a['time_us'] / n, 2), 'dram_bytes_per_launch': round(a['dram'] / n),
                               'tensor_pipe_active_pct': round(a['tensor'] / n, 1), 'dram_throughput_pct': round(a['dram_pct'] / n, 1)})
    g = [k for k in res['kernels'] if k['kernel'].startswith('gemm_tc2')]
    if g:
        n = sum(k['launches'] for k in g)
        res['gemm_family'] = {'launches': n, 'share': round(sum(k['share'] for k in g), 4),
                              'dram_bytes_per_launch': round(sum(k['dram_bytes_per_launch'] * k['launches'] for k in g) / n)}
    json.dump(res, open(out, 'w'), indent=1)
    print(json.dumps(res.get('gemm_family')), 'total', res['total_time_us'])


def full(path, out, cmd):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ['Kernel Name', 'gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
            'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
            'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
            'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
            'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size', 'launch__cluster_dim_x',
            'smsp__inst_executed.sum', 'sm__inst_executed_pipe_uniform.sum']
    idx = {w: hdr.index(w) for w in want if w in hdr}
    res = {'command': cmd, 'units': {w: units[i] for w, i in idx.items()}, 'launches': []}
    for r in rows[2:]:
        res['launches'].append({w: r[i] for w, i in idx.items()})
    json.dump(res, open(out, 'w'), indent=1)
    for l in res['launches']:
        print(l['Kernel Name'][:60], l.get('gpu__time_duration.sum'), l.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'))


if __name__ == '__main__':
    {'step': step, 'full': full}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else '')
