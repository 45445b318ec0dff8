#!/usr/bin/env python
"""Per-source-line instruction and stall-sample shares of one launch in an .ncu-rep:
   python tools/ncu_lines.py rep.ncu-rep --skip 2 [--by samples|inst] [--top 40]"""
import argparse, csv, io, subprocess
ap = argparse.ArgumentParser()
ap.add_argument("rep"); ap.add_argument("--skip", type=int, default=0); ap.add_argument("--by", default="samples")
ap.add_argument("--top", type=int, default=40)
a = ap.parse_args()
src = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(a.skip),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
fname, hdr, agg, fn = "?", None, {}, "?"
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": fname = r[1].split("/")[-1]
    elif len(r) >= 2 and r[0] == "Function Name": fn = r[1]
    elif len(r) > 8 and r[0] == "Line No":
        hdr = r; ie, ti, ss = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    elif hdr and len(r) > 8 and r[0].isdigit():
        try:
            k = (fname, int(r[0]), r[1].strip())
            if k not in agg: agg[k] = (int(r[ie]), int(r[ti]), int(r[ss]))
        except ValueError: pass
tot = sum(v[0] for v in agg.values()); tots = sum(v[2] for v in agg.values())
print(fn); print(f"total warp-inst {tot}  samples {tots}")
idx = 2 if a.by == "samples" else 0
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][idx])[: a.top]:
    print(f"{100*v[2]/max(1,tots):5.1f}% smp {100*v[0]/tot:5.1f}% inst thr/inst {v[1]/max(1,v[0]):5.1f}  {k[0]}:{k[1]}  {k[2][:105]}")
