#!/usr/bin/env python
"""Per-CUDA-line totals from `ncu --page source --csv --print-source cuda,sass`.
usage: python tools/ncu_lines.py export.csv [top_n]"""
import csv
import sys

fn = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rd = csv.reader(open(fn, errors='replace'))
path = func = None
hdr = None
data = {}
for row in rd:
    if not row:
        continue
    if row[0] == 'File Path':
        path = row[1].split('/')[-1]; continue
    if row[0] == 'Function Name':
        func = row[1]; continue
    if row[0] == 'Line No':
        hdr = row; continue
    if hdr and row[0] not in ('', '...'):
        d = dict(zip(hdr[4:], row[4:]))
        try:
            ins = float(d['Instructions Executed']); smp = float(d['# Samples'])
        except Exception:
            continue
        data.setdefault(func, []).append((path, int(row[0]), row[1].strip(), ins, smp))
for func, rows in data.items():
    ti = sum(r[3] for r in rows); ts = sum(r[4] for r in rows)
    print('=' * 100); print(func, 'instr %.4g samples %d' % (ti, ts))
    for r in sorted(rows, key=lambda r: -(r[3] / max(ti, 1) + r[4] / max(ts, 1)))[:top]:
        print('%-18s %5d %6.2f%% ins %6.2f%% smp  %s' % (r[0][:18], r[1], 100 * r[3] / max(ti, 1), 100 * r[4] / max(ts, 1), r[2][:100]))
