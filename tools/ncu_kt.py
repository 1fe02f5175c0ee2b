"""Per-launch kernel durations from an `ncu --metrics gpu__time_duration.sum --csv` log of
tools/bucket_breakdown.py (6 fwd/bwd pairs per pooled size) or tools/c3_breakdown.py (23 fwd, then 23
bwd launches per call; then the module test).  Prints the median forward / backward duration per group."""
import csv
import statistics
import sys

fn, kind = sys.argv[1], sys.argv[2]
rows = []
for r in csv.reader(open(fn, errors='replace')):
    if len(r) > 5 and r[0].isdigit():
        rows.append(r)
hdr = None
for r in csv.reader(open(fn, errors='replace')):
    if r and r[0] == 'ID':
        hdr = r
        break
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
def us(r):
    v = float(r[vi].replace(',', ''))
    u = r[ui]
    return v / 1000.0 if u in ('ns', 'nsecond') else (v * 1000.0 if u in ('ms', 'msecond') else v)
L = [('bwd' if '<1' in r[ki] or '(bool)1, ' in r[ki] and r[ki].index('(bool)1') < r[ki].index(',') else 'fwd', us(r)) for r in rows]
# kernel names look like ra_kernel<0, 1> / ra_kernel<(bool)0, (bool)1>: first template argument = BWD
def is_bwd(name):
    a = name[name.index('<') + 1:].split(',')[0]
    return a.strip().endswith('1')
L = [('bwd' if is_bwd(r[ki]) else 'fwd', us(r)) for r in rows]
out = []
if kind == 'buckets':
    for b, P in enumerate((14, 28, 56, 112)):
        seg = L[12 * b:12 * b + 12]
        f = [t for k, t in seg if k == 'fwd'][2:]
        bw = [t for k, t in seg if k == 'bwd'][2:]
        if f and bw:
            out.append('P%d f %.1f b %.1f' % (P, statistics.median(f), statistics.median(bw)))
else:
    for c, name in enumerate(('bbox7', 'mask14', 'sem56')):
        seg = L[46 * c:46 * c + 46]
        f = [t for k, t in seg if k == 'fwd'][3:]
        bw = [t for k, t in seg if k == 'bwd'][3:]
        if f and bw:
            out.append('%s f %.1f b %.1f' % (name, statistics.median(f), statistics.median(bw)))
print(' us: ' + '  '.join(out), ' (%d launches)' % len(L))
