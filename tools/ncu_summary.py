"""Summarises an `ncu --set full` report: per kernel launch the duration, DRAM
traffic, pipe utilisation, occupancy, registers and the top stall reasons; with
--json also writes {kernel: {"dram_bytes": read+write per launch (mean), ...}}.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--json profiles/x.json] > profiles/x.txt
"""
import collections
import csv
import io
import json
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
unit_scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.}
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio']
stalls = [h for h in hdr if h.startswith('smsp__pcsamp_warps_issue_stalled')
          and not h.endswith('not_issued')]
agg = collections.OrderedDict()
print('# %s' % rep)
for r in rows[2:]:
    name = r[col['Kernel Name']]
    print('== %s' % name)
    for w in WANT:
        if w in col:
            print('   %-62s %s %s' % (w, r[col[w]], units[col[w]]))
    tot = sum(float(r[col[h]]) for h in stalls) or 1.
    top = sorted(((float(r[col[h]]), h.replace('smsp__pcsamp_warps_issue_stalled_', ''))
                  for h in stalls), reverse=True)[:5]
    print('   top stall reasons: ' + ', '.join('%s %.0f%%' % (h, 100 * v / tot) for v, h in top))
    rd = float(r[col['dram__bytes_read.sum']]) * unit_scale[units[col['dram__bytes_read.sum']]]
    wr = float(r[col['dram__bytes_write.sum']]) * unit_scale[units[col['dram__bytes_write.sum']]]
    a = agg.setdefault(name, {'launches': 0, 'dram_bytes': 0., 'ms': 0.})
    a['launches'] += 1
    a['dram_bytes'] += rd + wr
    tu = units[col['gpu__time_duration.sum']]
    a['ms'] += float(r[col['gpu__time_duration.sum']]) * {'ms': 1., 'us': 1e-3, 'ns': 1e-6,
                                                          's': 1e3}.get(tu, 1.)
for a in agg.values():
    a['dram_bytes'] /= a['launches']
    a['ms'] /= a['launches']
if '--json' in sys.argv:
    with open(sys.argv[sys.argv.index('--json') + 1], 'w') as f:
        json.dump({'source': rep, 'config': '2-D Euler 2048^2 N=3 Rusanov', 'kernels': agg}, f,
                  indent=1)
