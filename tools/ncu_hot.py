#!/usr/bin/env python
"""SASS hot-region summary from an ncu report's source page (no GPU needed).
usage: ncu_hot.py report.ncu-rep kernel_regex [min_share]"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.004
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ia, ie, isamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
data = [(r[ia], int(r[ie]), int(r[isamp])) for r in rows[2:] if len(r) > isamp and r[ie].isdigit()]
# the csv holds the listing twice (two views); keep the first copy
half = len(data) // 2
if half and [d[0] for d in data[:half]] == [d[0] for d in data[half:2 * half]]:
    data = data[:half]
tot = sum(d[1] for d in data)
ts = sum(d[2] for d in data)
print('total warp-instr', tot, 'lines', len(data), 'samples', ts)
i = 0
while i < len(data):
    j = i
    while j < len(data) and data[j][1] == data[i][1]:
        j += 1
    n, e = j - i, data[i][1]
    if e * n > tot * min_share:
        ops = {}
        for s, _, _ in data[i:j]:
            t = s.strip().split()
            op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
            ops[op] = ops.get(op, 0) + 1
        print(f'[{i}:{j}] n={n} exec={e} share={e*n/tot:.3f} samples={sum(d[2] for d in data[i:j])/ts:.3f}',
              dict(sorted(ops.items(), key=lambda kv: -kv[1])))
    i = j
