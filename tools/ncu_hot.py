#!/usr/bin/env python
"""Hot-spot view of an `ncu --page source --csv --print-source sass` export: per kernel, the SASS
instructions with the most executed instructions / stall samples, plus stall-reason totals.
usage: python tools/ncu_hot.py export.csv [top_n]"""
import csv
import sys
from collections import defaultdict

fn = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kernels = []
cur = None
rd = csv.reader(open(fn, errors='replace'))
hdr = None
for row in rd:
    if not row:
        continue
    if row[0] == 'Kernel Name':
        cur = {'name': row[1], 'rows': []}
        kernels.append(cur)
        hdr = None
        continue
    if row[0] == 'Address':
        hdr = row
        continue
    if hdr and cur is not None:
        cur['rows'].append(dict(zip(hdr, row)))

def f(x):
    try:
        return float(x)
    except Exception:
        return 0.0

for k in kernels:
    rows = k['rows']
    tot_i = sum(f(r['Instructions Executed']) for r in rows)
    tot_s = sum(f(r['# Samples']) for r in rows)
    print('=' * 100)
    print(k['name'], ' SASS lines', len(rows), ' warp-instr %.4g' % tot_i, ' samples %d' % tot_s)
    stall = defaultdict(float)
    for r in rows:
        for key, v in r.items():
            if key.startswith('stall_') and 'Not Issued' not in key:
                stall[key] += f(v)
    print('stalls:', ', '.join('%s %.1f%%' % (a[6:], 100 * b / max(tot_s, 1)) for a, b in sorted(stall.items(), key=lambda x: -x[1])[:8]))
    op = defaultdict(float)
    for r in rows:
        m = r['Source'].split()
        name = m[1] if m and m[0].startswith('@') else (m[0] if m else '')
        op[name.split('.')[0]] += f(r['Instructions Executed'])
    print('opcode mix:', ', '.join('%s %.1f%%' % (a, 100 * b / max(tot_i, 1)) for a, b in sorted(op.items(), key=lambda x: -x[1])[:14]))
    print('--- top by samples')
    for i, r in sorted(enumerate(rows), key=lambda x: -f(x[1]['# Samples']))[:top]:
        print('%5d %6.2f%% smp %6.2f%% ins  thr %4.1f  %s' % (i, 100 * f(r['# Samples']) / max(tot_s, 1), 100 * f(r['Instructions Executed']) / max(tot_i, 1), f(r['Avg. Predicated-On Threads Executed']), r['Source'].strip()[:90]))
