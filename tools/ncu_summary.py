"""Summarises an `ncu --set full` report: per kernel launch the duration, DRAM
traffic, pipe utilisation, occupancy, registers and the top stall reasons; with
--into FILE --config NAME also merges, under FILE[NAME], {"source": ..., "kernels":
{kernel: {"dram_bytes": read+write per launch (mean), "fp64_flops": 2 DFMA + DMUL + DADD
thread instructions executed per launch, "pipe_fp64_pct": ..., "ms": ...}}} — the
per-launch numbers bench.py's roofline object quotes for that configuration.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--into profiles/r2_ncu_kernels.json \
        --config c2] > profiles/x.txt
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
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed',
        'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed',
        'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed',
        'smsp__cycles_elapsed.max', 'lts__t_bytes.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum',
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
    if rd != rd or wr != wr:
        # (ncu occasionally loses a multi-pass launch: all of its counters read nan)
        print('   (incomplete capture: not averaged)')
        continue
    a = agg.setdefault(name, {'launches': 0, 'dram_bytes': 0., 'ms': 0., 'fp64_flops': 0.,
                              'pipe_fp64_pct': 0., 'registers': 0})
    a['launches'] += 1
    a['dram_bytes'] += rd + wr
    tu = units[col['gpu__time_duration.sum']]
    a['ms'] += float(r[col['gpu__time_duration.sum']]) * {'ms': 1., 'us': 1e-3, 'ns': 1e-6,
                                                          's': 1e3}.get(tu, 1.)

    def rate(op):
        k = 'smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed' % op
        return float(r[col[k]]) if k in col and r[col[k]] else 0.

    if 'smsp__cycles_elapsed.max' in col:
        cyc = float(r[col['smsp__cycles_elapsed.max']])
        flops = (2. * rate('dfma') + rate('dmul') + rate('dadd')) * cyc
        a['fp64_flops'] += flops
        print('   FP64 flops executed (2 DFMA + DMUL + DADD thread inst.): %.4e  -> %.2f TFLOP/s '
              'under ncu' % (flops, flops / (float(r[col['gpu__time_duration.sum']]) *
                                             {'ms': 1e-3, 'us': 1e-6, 'ns': 1e-9, 's': 1.}.get(tu, 1e-3))
                             / 1e12))
    k = 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'
    if k in col and r[col[k]]:
        a['pipe_fp64_pct'] += float(r[col[k]])
    if 'launch__registers_per_thread' in col:
        a['registers'] = int(float(r[col['launch__registers_per_thread']]))
for a in agg.values():
    for k in ('dram_bytes', 'ms', 'fp64_flops', 'pipe_fp64_pct'):
        a[k] /= a['launches']
if '--into' in sys.argv:
    path = sys.argv[sys.argv.index('--into') + 1]
    name = sys.argv[sys.argv.index('--config') + 1]
    try:
        with open(path) as f:
            allc = json.load(f)
    except (OSError, ValueError):
        allc = {}
    cur = allc.setdefault(name, {'source': '', 'kernels': {}})
    cur['source'] = (cur['source'] + '; ' if cur['source'] else '') + rep + \
        ' (ncu --set full --clock-control none)'
    cur['kernels'].update(agg)
    with open(path, 'w') as f:
        json.dump(allc, f, indent=1)
