#!/usr/bin/env python
"""Table of an ncu --csv launch list with several metrics per launch: python tools/launch_table.py launches.csv"""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); mi = hdr.index("Metric Name"); vi = hdr.index("Metric Value"); ii = hdr.index("ID")
d = {}
for r in rows[1:]:
    d.setdefault((int(r[ii]), r[ki][:70]), {})[r[mi]] = float(r[vi])
tot = sum(m['gpu__time_duration.sum'] for m in d.values()) / 1e6
print(f"total {tot:.3f} ms over {len(d)} launches")
for (i, k), m in sorted(d.items()):
    t = m['gpu__time_duration.sum'] / 1e6
    if t > 0.03:
        print(f"{i:3d} {k:70s} {t:7.3f} ms {100*t/tot:5.1f}%  inst {m.get('smsp__inst_executed.sum',0)/1e6:8.1f}M  warps {m.get('sm__warps_active.avg.pct_of_peak_sustained_active',0):5.1f}%  issue {m.get('smsp__issue_active.avg.pct_of_peak_sustained_active',0):5.1f}%  dram R {m.get('dram__bytes_read.sum',0)/1e9:6.2f} W {m.get('dram__bytes_write.sum',0)/1e9:6.2f} GB")
