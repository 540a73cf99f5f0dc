#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples from `ncu -i X.ncu-rep --page source --csv` (stdin or file)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
# several launches in one report print one table each: pick the one asked for (third argument, default the first)
hdr_all = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hdr_i = hdr_all[which]
rows = rows[:hdr_all[which + 1] - 1] if which + 1 < len(hdr_all) else rows
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
total = sum(int(r[col['# Samples']]) for r in data)
print('total samples', total, 'instructions', len(data))
agg = {}
for s in stalls:
    agg[s] = sum(int(r[col[s]] or 0) for r in data)
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
order = sorted(range(len(data)), key=lambda i: -int(data[i][col['# Samples']]))[:top]
for i in sorted(order):
    r = data[i]
    n = int(r[col['# Samples']])
    dom = max(stalls, key=lambda s: int(r[col[s]] or 0))
    print(f'{i:5d} {n:6d} {100.0 * n / total:5.1f}%  {dom:22s} exec={r[col["Instructions Executed"]]:>8s}  {r[col["Source"]].strip()[:110]}')
