#!/usr/bin/env python
"""Prints (and optionally saves as csv) the key metrics of every kernel in an .ncu-rep, plus the top
source lines by executed instructions:  python tools/ncu_summary.py rep.ncu-rep [--out profiles/x.csv] [--src regex]"""
import argparse
import csv
import io
import subprocess

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers"]

ap = argparse.ArgumentParser()
ap.add_argument("rep")
ap.add_argument("--out")
ap.add_argument("--src", help="kernel-name regex for the source page")
ap.add_argument("--top", type=int, default=30)
a = ap.parse_args()

raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
for r in rows[2:]:
    print("-----")
    for w, i in idx:
        print(f"  {w:84s} {r[i]} {units[i]}")
if a.out:
    with open(a.out, "w") as f:
        w = csv.writer(f)
        w.writerow([n for n, _ in idx])
        w.writerow([units[i] for _, i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for _, i in idx])
if a.src:
    src = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          f"regex:{a.src}", "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    fname, hdr, agg = "?", None, {}
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif len(r) > 8 and r[0] == "Line No":
            hdr = r
            ie, ti, ss = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        elif hdr and len(r) > 8 and r[0].isdigit():
            try:
                agg[(fname, int(r[0]), r[1].strip())] = (int(r[ie]), int(r[ti]), int(r[ss]))
            except ValueError:
                pass
    tot = sum(v[0] for v in agg.values())
    tots = sum(v[2] for v in agg.values())
    print(f"== CUDA lines by instructions executed (total {tot}, samples {tots}) ==")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[: a.top]:
        print(f"{100 * v[0] / tot:5.1f}% inst {100 * v[2] / max(1, tots):5.1f}% stall  thr/inst {v[1] / max(1, v[0]):5.1f}  {k[0]}:{k[1]}  {k[2][:100]}")
