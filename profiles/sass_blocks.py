#!/usr/bin/env python
"""Basic-block level instruction / stall-sample shares from an `ncu --page source --csv` SASS dump.
    python profiles/sass_blocks.py gpurun_out/r02b_stream_src_sass.csv [min_share_pct]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
h = rows[1]
isrc, ie, ism = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
data = [(r[isrc].strip(), int(r[ie]), int(r[ism])) for r in rows[2:] if len(r) > ie and r[ie].isdigit()]
tot = sum(d[1] for d in data)
stot = sum(d[2] for d in data)
print("total warp instr", tot, "n sass", len(data), "samples", stot)
runs, start = [], 0
for i in range(1, len(data) + 1):
    if i == len(data) or data[i][1] != data[start][1]:
        runs.append((start, i, data[start][1]))
        start = i
for (a, b, c) in runs:
    w = (b - a) * c
    samples = sum(d[2] for d in data[a:b])
    if 100 * w / tot > thr or 100 * samples / stot > thr:
        ops = [d[0].split()[1] if d[0].startswith('@') else d[0].split()[0] for d in data[a:b]]
        cnt = Counter(o.split('.')[0] for o in ops)
        print(f"[{a:5d},{b:5d}) n={b - a:4d} exec={c:12d} instr={100 * w / tot:5.1f}% samples={100 * samples / stot:5.1f}% "
              f"top={cnt.most_common(5)}")
