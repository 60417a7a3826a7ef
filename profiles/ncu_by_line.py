#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export by CUDA source line.

    ncu -i prof.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:<k> > src.csv
    python profiles/ncu_by_line.py src.csv [top_n]
"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rows = list(csv.reader(open(path)))
cur_file = None
hdr = None
agg = defaultdict(lambda: [0.0, 0.0, ""])
cur_line = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) > 6 and r[0] == "Line No":
        hdr = r
        si = hdr.index("# Samples")
        ii = hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0].strip():
        cur_line = (cur_file, int(r[0]))
        agg[cur_line][2] = r[1].strip()
    if cur_line is None:
        continue
    try:
        agg[cur_line][0] += float(r[si] or 0)
        agg[cur_line][1] += float(r[ii] or 0)
    except ValueError:
        pass
tot_s = sum(v[0] for v in agg.values()) or 1.0
tot_i = sum(v[1] for v in agg.values()) or 1.0
print(f"total samples {tot_s:.0f}, warp instructions {tot_i:.0f}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top_n]:
    print(f"{v[0] / tot_s * 100:5.1f}% smp {v[1] / tot_i * 100:5.1f}% inst | {k[0]}:{k[1]:<4} {v[2][:100]}")
