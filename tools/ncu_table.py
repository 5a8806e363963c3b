"""Tabulate the metrics that decide what bounds a tcgen05 kernel from `ncu -i X.ncu-rep --page raw --csv`.
usage: ncu_table.py raw.csv [kernel-name-regex]"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
COLS = [('dur_us', 'gpu__time_duration.sum'),
        ('tensor%', 'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed'),
        ('tc_smem%', 'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'),
        ('lsu_smem%', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'),
        ('tma_ld_GB', 'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum'),
        ('xbar2l1%', 'l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed'),
        ('lts%', 'lts__throughput.avg.pct_of_peak_sustained_elapsed'),
        ('l2hit%', 'lts__t_sector_hit_rate.pct'),
        ('dramR_GB', 'dram__bytes_read.sum'), ('dramW_GB', 'dram__bytes_write.sum'),
        ('dram%', 'dram__throughput.avg.pct_of_peak_sustained_elapsed'),
        ('sm_cycles', 'sm__cycles_elapsed.max')]
for _, c in COLS:
    if c not in idx:
        alt = [h for h in hdr if h.endswith(c)]
        if alt:
            idx[c] = idx[alt[0]]


def val(r, c):
    if c not in idx:
        return float('nan')
    try:
        v = float(r[idx[c]].replace(',', ''))
    except ValueError:
        return float('nan')
    u = units[idx[c]]
    if u in ('byte', 'bytes'):
        return v / 1e9
    if u == 'Kbyte':
        return v / 1e6
    if u == 'Mbyte':
        return v / 1e3
    if u == 'Gbyte':
        return v
    if u in ('ns', 'nsecond'):
        return v / 1e3
    if u in ('ms', 'msecond'):
        return v * 1e3
    return v


print('| kernel | ' + ' | '.join(n for n, _ in COLS) + ' |')
print('|---|' + '---:|' * len(COLS))
for r in data:
    name = r[idx['Kernel Name']]
    if pat and not pat.search(name):
        continue
    m = re.search(r'(\w+)<([^>]*)>', name)
    nm = (m.group(1) + '<' + m.group(2) + '>') if m else name[:40]
    print('| `%s` | ' % nm + ' | '.join('%.2f' % val(r, c) for _, c in COLS) + ' |')
