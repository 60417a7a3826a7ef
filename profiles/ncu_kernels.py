#!/usr/bin/env python
"""Key metrics of every kernel in an .ncu-rep (reads it through `ncu -i ... --page raw --csv`).

    python profiles/ncu_kernels.py gpurun_out/x.ncu-rep > profiles/rN_x_summary.txt
"""
import csv
import io
import re
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
PAT = (r"^gpu__time_duration.sum$|^dram__bytes_(read|write).sum$|dram_throughput.avg.pct|^lts__t_sector_hit_rate.pct|^l1tex__t_sector_hit_rate.pct|"
       r"smsp__issue_active.avg.pct|sm__warps_active.avg.pct|registers_per_thread$|^smsp__inst_executed.sum$|pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active|"
       r"thread_inst_executed_per_inst_executed.ratio|issue_stalled.*per_issue_active.ratio$|local_(load|store)s$|launch__grid_size|launch__block_size|"
       r"launch__occupancy_limit|sm__throughput.avg.pct|l1tex__throughput.avg.pct|lts__throughput.avg.pct|shared_mem_per_block")
keys = [k for k in hdr if re.search(PAT, k)]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")])
    for k in keys:
        i = hdr.index(k)
        try:
            v = float(r[i].replace(",", ""))
        except ValueError:
            v = None
        if "stalled" in k and v is not None and v < 0.2:
            continue
        print(f"   {k:86s} {r[i]} {units[i]}")
