#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X` launch list.

    python profiles/launch_summary.py gpurun_out/launches.csv > profiles/rN_launches.txt
"""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
name_i, val_i, met_i = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
tot, cnt, order = OrderedDict(), {}, []
for r in rows[1:]:
    if r[met_i] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"^void ", "", r[name_i])
    name = re.sub(r"<unnamed>::|\(anonymous namespace\)::", "", name)
    name = re.sub(r"\(.*$", "", name)[:70]
    ns = float(r[val_i].replace(",", ""))
    tot[name] = tot.get(name, 0.0) + ns
    cnt[name] = cnt.get(name, 0) + 1
    order.append((name, ns))
total = sum(tot.values())
print(f"# total GPU time {total / 1e6:.2f} ms over {len(order)} launches")
for name, ns in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{name:72s} launches {cnt[name]:4d}  total_ms {ns / 1e6:10.3f}  share {100 * ns / total:5.1f}%  mean_us {ns / cnt[name] / 1e3:10.1f}")
