#!/usr/bin/env python
"""Per-SASS-instruction executed counts of one kernel from an .ncu-rep, grouped into runs.  Usage: ncu_hot.py rep regex [--full]"""
import csv
import subprocess
import sys

out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--kernel-name', 'regex:' + sys.argv[2]], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[1]
ia, ie = h.index('Source'), h.index('Instructions Executed')
data = []
for r in rows[2:]:
    if len(r) > ie and r[ie].isdigit():
        if data and r[ia].strip().startswith('LDC R1, c[0x0][0x37c]'):
            break  # second launch of the same kernel
        data.append((r[ia].strip(), int(r[ie])))
tot = sum(e for _, e in data)
print(len(data), 'instructions,', tot, 'executed (warp level)')
if '--full' in sys.argv:
    for i, (s, e) in enumerate(data):
        print(f"{i:5d} {e:9d} {s[:100]}")
else:
    start = 0
    for i in range(1, len(data) + 1):
        if i == len(data) or abs(data[i][1] - data[start][1]) > 0.15 * max(data[start][1], 1):
            c = sum(e for _, e in data[start:i])
            if c > 0.004 * tot:
                print(f"[{start:5d}-{i - 1:5d}] n={i - start:4d} exec/instr~{data[start][1]:9d} share={100 * c / tot:5.1f}%  {data[start][0][:70]}")
            start = i
