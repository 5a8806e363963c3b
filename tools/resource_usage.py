"""Registers / static shared memory of every kernel in libsedb200.so (cuobjdump --dump-resource-usage).

    python tools/resource_usage.py [substring ...]
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'sound_event_detection_dcase2017_task4_b200', 'libsedb200.so')


def main():
    out = subprocess.run(['cuobjdump', '--dump-resource-usage', LIB], capture_output=True, text=True).stdout
    names, rows = [], []
    cur = None
    for line in out.splitlines():
        m = re.search(r'Function (\S+):', line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r'REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)', line)
        if m and cur:
            names.append(cur)
            rows.append(tuple(int(x) for x in m.groups()))
            cur = None
    dem = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True).stdout.splitlines()
    filt = sys.argv[1:]
    for n, (reg, stack, sh, loc) in sorted(zip(dem, rows)):
        n = re.sub(r'\(anonymous namespace\)::', '', n)
        n = re.sub(r'\(.*', '', n)
        n = re.sub(r'^void ', '', n)
        if filt and not any(f in n for f in filt):
            continue
        print('%-72s regs %3d  stack %4d  smem %6d  local %d' % (n[:72], reg, stack, sh, loc))


if __name__ == '__main__':
    main()
