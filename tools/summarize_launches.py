"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list:
per kernel launches, total/avg device time and share.  Usage:
    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/<name>.txt"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    u = r[ui]
    v *= {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1., 'msecond': 1., 's': 1e3,
          'second': 1e3}.get(u, 1e-6)
    a = agg.setdefault(r[ki], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print('# %s' % ' '.join(sys.argv[2:]))
print('%-44s %8s %12s %10s %7s' % ('kernel', 'launches', 'total_ms', 'avg_ms', 'share'))
for k, a in agg.items():
    print('%-44s %8d %12.3f %10.3f %6.1f%%' % (k[:44], a[0], a[1], a[1] / a[0], 100 * a[1] / tot))
