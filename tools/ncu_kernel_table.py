"""Markdown table of what bounds every kernel of a captured step, from `ncu -i X.ncu-rep --page raw --csv`:
launches, device time, DRAM bytes and achieved bandwidth against the measured HBM peak, L1/TEX and issue utilisation,
executed tensor-core operations against their peak (the op-count metric, not the realtime cycle sampler), occupancy.

    ncu -i gpurun_out/step_prof.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_kernel_table.py raw.csv [name-regex] > profiles/rNN_step_ncu.md
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
peak = 6538.6
pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
if os.path.exists(pk):
    peak = json.load(open(pk))['hbm_gbs']


def col(suffix):
    for i, h in enumerate(hdr):
        if h == suffix or h.endswith('.' + suffix) or h.endswith(suffix):
            return i
    return None


C = {k: col(v) for k, v in {
    'name': 'Kernel Name', 'dur': 'gpu__time_duration.sum', 'rd': 'dram__bytes_read.sum', 'wr': 'dram__bytes_write.sum',
    'l1': 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'issue': 'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
    'tensor': 'sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
    'occ': 'sm__warps_active.avg.pct_of_peak_sustained_active', 'regs': 'launch__registers_per_thread',
    'grid': 'launch__grid_size', 'block': 'launch__block_size'}.items()}


def num(r, key, scale_bytes=False):
    i = C.get(key)
    if i is None:
        return float('nan')
    try:
        v = float(r[i].replace(',', ''))
    except ValueError:
        return float('nan')
    u = units[i].lower()
    if scale_bytes:
        v *= {'byte': 1, 'bytes': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(u, 1)
    if key == 'dur':
        v *= {'ns': 1e-3, 'nsecond': 1e-3, 'us': 1, 'usecond': 1, 'ms': 1e3, 'msecond': 1e3, 's': 1e6, 'second': 1e6}.get(u, 1)
    return v


agg = {}
order = []
for r in data:
    name = r[C['name']]
    name = re.sub(r'\(anonymous namespace\)::|<?unnamed>::|sed::|void ', '', name)
    name = re.sub(r'\(.*$', '', name)
    if pat and not pat.search(name):
        continue
    a = agg.get(name)
    if a is None:
        a = agg[name] = {'n': 0, 'dur': 0.0, 'bytes': 0.0, 'l1': 0.0, 'issue': 0.0, 'tensor': 0.0, 'occ': 0.0, 'regs': 0, 'cfg': ''}
        order.append(name)
    d = num(r, 'dur')
    a['n'] += 1
    a['dur'] += d
    a['bytes'] += num(r, 'rd', True) + num(r, 'wr', True)
    for k in ('l1', 'issue', 'tensor', 'occ'):
        v = num(r, k)
        a[k] += (0.0 if v != v else v) * d                        # duration-weighted
    a['regs'] = int(num(r, 'regs')) if num(r, 'regs') == num(r, 'regs') else 0
    a['cfg'] = '%sx%s' % (r[C['grid']].replace(',', ''), r[C['block']].replace(',', '')) if C.get('grid') is not None else ''
total = sum(a['dur'] for a in agg.values())
print('| kernel | launches | total us | share | DRAM GB/launch | GB/s | of HBM peak (%.0f) | L1/TEX %% | issue %% | tensor ops %% | occupancy %% | regs | last grid x block |' % peak)
print('|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|')
for name in sorted(order, key=lambda n: -agg[n]['dur']):
    a = agg[name]
    gbs = a['bytes'] / a['dur'] / 1e3 if a['dur'] else 0.0
    w = a['dur'] or 1.0
    print('| `%s` | %d | %.1f | %.1f%% | %.3f | %.0f | %.2f | %.0f | %.0f | %.1f | %.0f | %d | %s |' % (
        name[:60], a['n'], a['dur'], 100 * a['dur'] / total, a['bytes'] / a['n'] / 1e9, gbs, gbs / peak,
        a['l1'] / w, a['issue'] / w, a['tensor'] / w, a['occ'] / w, a['regs'], a['cfg']))
print()
print('total: %d launches, %.3f ms of device time (ncu replays every kernel alone with cold caches: compare shares, '
      'not absolutes).' % (sum(a['n'] for a in agg.values()), total / 1e3))
