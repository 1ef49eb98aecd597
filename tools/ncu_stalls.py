#!/usr/bin/env python
"""Top stall sites of one kernel from an .ncu-rep (source page).  Usage: ncu_stalls.py rep regex [N]"""
import csv
import subprocess
import sys

out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--kernel-name', 'regex:' + sys.argv[2]], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[1]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
ia, isamp, ie = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
data = []
for k, r in enumerate(rows[2:]):
    if len(r) > ie and r[isamp].isdigit():
        st = {c: int(r[i]) for i, c in stall_cols if r[i].isdigit() and int(r[i])}
        data.append((k, r[ia].strip(), int(r[isamp]), int(r[ie]) if r[ie].isdigit() else 0, st))
tot = sum(d[2] for d in data)
print('total samples', tot)
agg = {}
for d in data:
    for c, v in d[4].items():
        agg[c] = agg.get(c, 0) + v
print('by reason:', ', '.join(f"{c[6:]}={100 * v / tot:.1f}%" for c, v in sorted(agg.items(), key=lambda x: -x[1])))
for d in sorted(data, key=lambda d: -d[2])[:n]:
    top = ', '.join(f"{c[6:]}={v}" for c, v in sorted(d[4].items(), key=lambda x: -x[1])[:3])
    print(f"{d[0]:5d} {100 * d[2] / tot:5.1f}% exec={d[3]:8d} {d[1][:60]:60s} {top}")
