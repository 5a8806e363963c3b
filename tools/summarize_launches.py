"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launch count,
total device time and share.  usage: summarize_launches.py launches.csv [skip_first_n] > summary.md"""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r'\(anonymous namespace\)::', '', name)
    name = re.sub(r'^void ', '', name)
    name = name.replace('sed::<unnamed>::', 'sed::')
    depth, out = 0, []
    for ch in name:                     # cut at the '(' that opens the parameter list
        if ch == '<':
            depth += 1
        elif ch == '>':
            depth -= 1
        elif ch == '(' and depth == 0:
            break
        out.append(ch)
    return ''.join(out)[:90]


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        rows.append((short(r['Kernel Name']), float(r['Metric Value'].replace(',', '')), r['Grid Size'], r['Block Size']))
    rows = rows[skip:]
    agg = OrderedDict()
    for n, t, g, b in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += t
    total = sum(v[1] for v in agg.values())
    print('| kernel | launches | total ms | share |')
    print('|---|---:|---:|---:|')
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('| `%s` | %d | %.3f | %.1f%% |' % (n, c, t / 1e6, 100 * t / total))
    print('\ntotal: %d launches, %.3f ms of device time (ncu: cold-cache, serialised -- compare SHARES)' % (len(rows), total / 1e6))


if __name__ == '__main__':
    main()
