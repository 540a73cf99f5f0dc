#!/usr/bin/env python
"""profiles/sass_summary.json: per-kernel counts of the SASS mnemonics that tell a Blackwell-native kernel from a legacy
one (B200_PROFILING.md): UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UTMAREDG / UBLKCP (TMA),
HMMA (mma.sync), LDGSTS (cp.async), MUFU, and whether the kernel allocates TMEM.  Runs `cuobjdump -sass` on the built
library; no GPU needed.   python scripts/sass_summary.py [out.json]"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'ecg-representation-learning_b200', 'libecgvit_b200.so')
PATTERNS = collections.OrderedDict([
    ('UTCHMMA', r'\bUTCHMMA'), ('UTCHMMA.2CTA', r'\bUTCHMMA\.2CTA'), ('LDTM', r'\bLDTM'), ('STTM', r'\bSTTM'),
    ('UTMALDG', r'\bUTMALDG'), ('UTMASTG', r'\bUTMASTG'), ('UTMAREDG', r'\bUTMAREDG'), ('UBLKCP', r'\bUBLKCP'),
    ('UTCBAR', r'\bUTCBAR'), ('UTCATOMSWS (tmem alloc)', r'\bUTCATOMSWS'), ('HMMA (mma.sync)', r'\bHMMA'),
    ('LDGSTS (cp.async)', r'\bLDGSTS'), ('MUFU', r'\bMUFU'), ('REDG / RED (global red.add)', r'(?<![.\w])REDG?\b'), ('ATOMG (global atomic)', r'(?<![.\w])ATOMG?\b'),
    ('STAS (st.async to the peer CTA)', r'\bSTAS\b'), ('UCGABAR (cluster barrier)', r'\bUCGABAR'),
    ('FFMA2 / FMUL2 / FADD2', r'\bF(FMA|MUL|ADD)2\b'),
])


def demangle(names):
    out = subprocess.run(['c++filt'] + names, capture_output=True, text=True).stdout.splitlines()
    clean = []
    for n in out:
        n = re.sub(r'\(anonymous namespace\)::|ecgvit::|void ', '', n)
        clean.append(re.sub(r'\(.*', '', n))
    return clean


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'profiles', 'sass_summary.json')
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None or '/*' not in line:
            continue
        kernels[cur]['instructions'] += 1
        for key, pat in PATTERNS.items():
            if re.search(pat, line):
                kernels[cur][key] += 1
    names = demangle(list(kernels))
    res = {'library': os.path.relpath(LIB, ROOT), 'how': 'cuobjdump -sass, mnemonic counts per kernel (static instruction counts)',
           'kernels': []}
    for (mangled, c), name in zip(kernels.items(), names):
        row = {'kernel': name, 'instructions': c['instructions']}
        row.update({k: c[k] for k in PATTERNS if c[k]})
        native = c['UTCHMMA'] > 0
        row['tensor_path'] = 'tcgen05 (UTC*MMA)' if native else ('mma.sync (HMMA)' if c['HMMA (mma.sync)'] else 'none')
        res['kernels'].append(row)
    res['kernels'].sort(key=lambda r: (r['tensor_path'] != 'tcgen05 (UTC*MMA)', r['kernel']))
    json.dump(res, open(out_path, 'w'), indent=1)
    t = [r for r in res['kernels'] if r['tensor_path'].startswith('tcgen05')]
    print(f'{len(res["kernels"])} kernels, {len(t)} on tcgen05; attention:',
          [(r['kernel'][:28], r.get('UTCHMMA'), r.get('LDTM'), r.get('UTMALDG')) for r in t if 'attn' in r['kernel']])


if __name__ == '__main__':
    main()
