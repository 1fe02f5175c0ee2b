#!/usr/bin/env python
"""Turn ncu CSV logs (launch list / per-kernel metric passes, `--csv --log-file`) into the compact
tables committed under profiles/.  Usage:
    python profiles/summarize_ncu.py launches  gpurun_out/r01_launches.csv        > profiles/r01_launches.md
    python profiles/summarize_ncu.py metrics   gpurun_out/r01_kernel_metrics.csv  > profiles/r01_kernel_metrics.md
"""
import csv
import io
import re
import sys
from collections import OrderedDict, defaultdict


def load(fn):
    lines = open(fn, errors='replace').read().splitlines()
    start = [k for k, l in enumerate(lines) if l.startswith('"ID"')][0]
    return list(csv.DictReader(io.StringIO('\n'.join(lines[start:]))))


def short(name):
    name = re.sub(r'\s+', ' ', name)
    name = re.sub(r'^void ', '', name)
    m = re.match(r'(at::[\w:<>]*?(?:kernel|Copy)\w*)', name)
    if m:
        return m.group(1)[:60]
    return name.split('(')[0][:60]


def launches(fn):
    rows = load(fn)
    agg = OrderedDict()
    for r in rows:
        k = (short(r['Kernel Name']), r['Block Size'], r['Grid Size'])
        agg.setdefault(k, []).append(float(r['Metric Value']))
    total = sum(sum(v) for v in agg.values())
    print('| kernel | block | grid | launches | mean us | total us | share |')
    print('|---|---|---|---:|---:|---:|---:|')
    for (k, b, g), v in agg.items():
        print('| `%s` | %s | %s | %d | %.1f | %.1f | %.1f %% |' % (
            k, b, g, len(v), sum(v) / len(v) / 1e3, sum(v) / 1e3, 100 * sum(v) / total))
    print('\nTotal device time of the listed launches: %.3f ms (ncu: serialised, cold cache -- compare shares).' % (total / 1e6))


def metrics(fn):
    rows = load(fn)
    per = defaultdict(lambda: defaultdict(list))
    order = []
    for r in rows:
        k = short(r['Kernel Name'])
        if k not in order:
            order.append(k)
        try:
            per[k][r['Metric Name']].append(float(r['Metric Value'].replace(',', '')))
        except ValueError:
            pass
    names = sorted({m for k in per for m in per[k]})
    for k in order:
        print('### `%s`  (%d launches, means)\n' % (k, len(next(iter(per[k].values())))))
        print('| metric | value |')
        print('|---|---:|')
        d = {}
        for m in names:
            if m in per[k]:
                d[m] = sum(per[k][m]) / len(per[k][m])
                print('| %s | %s |' % (m, ('%.4g' % d[m]) if abs(d[m]) < 1e6 else '%.6e' % d[m]))
        if 'gpu__time_duration.sum' in d and 'dram__bytes_read.sum' in d:
            t = d['gpu__time_duration.sum'] * 1e-9
            tr = d['dram__bytes_read.sum'] + d.get('dram__bytes_write.sum', 0.0)
            print('| **DRAM traffic / launch** | %.4e B |' % tr)
            print('| **DRAM throughput** | %.0f GB/s |' % (tr / t / 1e9))
        print()


if __name__ == '__main__':
    {'launches': launches, 'metrics': metrics}[sys.argv[1]](sys.argv[2])
