"""Per-kernel markdown table from an `ncu --metrics ... --csv --log-file X.csv` run (long format: one row per metric):
launches, device time, DRAM bytes and achieved bandwidth against the measured HBM peak, L1/TEX and issue utilisation,
tensor-core activity, occupancy, registers.

    python tools/ncu_long_table.py gpurun_out/step_metrics.csv [name-regex] > profiles/rNN_step_ncu.md
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 6538.6
pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
if os.path.exists(pk):
    peak = json.load(open(pk))['hbm_gbs']
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ix = {h: i for i, h in enumerate(hdr)}
launch = {}
order = []
for r in rd:
    if len(r) < len(hdr):
        continue
    kid = r[ix['ID']]
    d = launch.get(kid)
    if d is None:
        d = launch[kid] = {'name': r[ix['Kernel Name']], 'grid': r[ix['Grid Size']] if 'Grid Size' in ix else '',
                           'block': r[ix['Block Size']] if 'Block Size' in ix else ''}
        order.append(kid)
    try:
        v = float(r[ix['Metric Value']].replace(',', ''))
    except ValueError:
        continue
    u = r[ix['Metric Unit']].lower()
    name = r[ix['Metric Name']]
    if 'bytes' in name:
        v *= {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(u, 1)
    if name == 'gpu__time_duration.sum':
        v *= {'ns': 1e-3, 'nsecond': 1e-3, 'us': 1, 'usecond': 1, 'ms': 1e3, 'msecond': 1e3}.get(u, 1)
    d[name] = v

M = {'dur': 'gpu__time_duration.sum', 'rd': 'dram__bytes_read.sum', 'wr': 'dram__bytes_write.sum',
     'l1': 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'issue': 'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
     'occ': 'sm__warps_active.avg.pct_of_peak_sustained_active', 'regs': 'launch__registers_per_thread'}
tensor_keys = sorted({k for d in launch.values() for k in d if 'tensor' in k})
agg = {}
for kid in order:
    d = launch[kid]
    name = re.sub(r'\(anonymous namespace\)::|<?unnamed>::|sed::|void ', '', d['name'])
    name = re.sub(r'\(.*$', '', name)
    if pat and not pat.search(name):
        continue
    a = agg.setdefault(name, {'n': 0, 'dur': 0.0, 'bytes': 0.0, 'l1': 0.0, 'issue': 0.0, 'occ': 0.0, 'tensor': 0.0, 'regs': 0,
                              'cfg': ''})
    dur = d.get(M['dur'], 0.0)
    a['n'] += 1
    a['dur'] += dur
    a['bytes'] += d.get(M['rd'], 0.0) + d.get(M['wr'], 0.0)
    for k in ('l1', 'issue', 'occ'):
        a[k] += d.get(M[k], 0.0) * dur
    a['tensor'] += max([d.get(k, 0.0) for k in tensor_keys] + [0.0]) * dur
    a['regs'] = int(d.get(M['regs'], 0))
    a['cfg'] = '%sx%s' % (d['grid'], d['block'])
total = sum(a['dur'] for a in agg.values()) or 1.0
print('| kernel | launches | total us | share | DRAM GB/launch | GB/s | of HBM peak (%.0f GB/s) | L1/TEX %% | issue %% | tensor pipe %% | occupancy %% | regs | last grid x block |' % peak)
print('|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|')
for name in sorted(agg, key=lambda n: -agg[n]['dur']):
    a = agg[name]
    w = a['dur'] or 1.0
    gbs = a['bytes'] / w / 1e3
    print('| `%s` | %d | %.1f | %.1f%% | %.3f | %.0f | %.2f | %.0f | %.0f | %.1f | %.0f | %d | %s |' % (
        name[:64], a['n'], a['dur'], 100 * a['dur'] / total, a['bytes'] / a['n'] / 1e9, gbs, gbs / peak, a['l1'] / w,
        a['issue'] / w, a['tensor'] / w, a['occ'] / w, a['regs'], a['cfg']))
print()
print('tensor pipe %% = max over: %s' % ', '.join('`%s`' % k for k in tensor_keys))
print('total: %d launches, %.3f ms of device time (ncu replays every kernel alone with cold caches: compare shares, '
      'not absolutes).' % (sum(a['n'] for a in agg.values()), total / 1e3))
