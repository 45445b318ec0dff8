#!/usr/bin/env python
"""Phase timing of the end-to-end path (pinned host columns -> ingest -> build -> rows -> host), both input forms:
copy all four columns (round 1) and copy session + ts, read the tails of aid / type over PCIe (zero_copy)."""
import argparse, pathlib, sys, time
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
import __graft_entry__ as g
ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0)
args = ap.parse_args()
g.build()
from otto_multi_objective_recommender_system_b200 import covisit, synth
dev = torch.device("cuda:0")
frame = synth.generate(synth.SynthSpec.scaled("train", args.scale), device=dev)
host = synth.EventFrame(*(t.cpu().pin_memory() for t in (frame.session, frame.aid, frame.ts, frame.type)), n_aids=frame.n_aids)
del frame
def T():
    torch.cuda.synchronize(); return time.perf_counter()
b = None
for mode in ("copy", "zero_copy", "copy", "zero_copy"):
    t0 = T()
    if mode == "copy":
        f = synth.EventFrame(host.session.to(dev, non_blocking=True), host.aid.to(dev, non_blocking=True),
                             host.ts.to(dev, non_blocking=True), host.type.to(dev, non_blocking=True), host.n_aids)
        t1 = T()
        c = covisit.ingest(f, "desc", device=dev)
    else:
        t1 = t0
        c = covisit.ingest(host, "asc", device=dev, zero_copy=True)
    t2 = T()
    nb = covisit.CovisitBuilder(c, covisit.CLICKS)
    if b is not None:
        nb.workspace, nb.records, nb.scratch, nb.table = b.workspace, b.records, b.scratch, b.table
    b = nb
    t3 = T()
    b.count_begin()
    t4 = T()
    b.count_finish(); b.scatter(); t = b.reduce()
    t5 = T()
    rows = t.to_rows()
    t6 = T()
    out = [x.cpu() for x in rows]
    t7 = T()
    print(f"{mode:9s} h2d {1e3*(t1-t0):.1f}  ingest {1e3*(t2-t1):.1f}  builder-init {1e3*(t3-t2):.1f}  count_begin {1e3*(t4-t3):.1f}  "
          f"rest of build {1e3*(t5-t4):.1f}  to_rows {1e3*(t6-t5):.1f}  d2h(pageable) {1e3*(t7-t6):.1f}  total {1e3*(t7-t0):.1f} ms")
    del c, rows, out
