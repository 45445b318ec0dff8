#!/usr/bin/env python
"""Phase timing of the end-to-end path (host columns -> device -> ingest -> build -> rows -> host)."""
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
for it in range(3):
    t0 = T()
    f = synth.EventFrame(host.session.to(dev, non_blocking=True), host.aid.to(dev, non_blocking=True),
                         host.ts.to(dev, non_blocking=True), host.type.to(dev, non_blocking=True), host.n_aids)
    t1 = T()
    c = covisit.ingest(f, "desc", device=dev)
    t2 = T()
    b = covisit.CovisitBuilder(c, covisit.CLICKS)
    t3 = T()
    t = b.build()
    t4 = T()
    rows = t.to_rows()
    t5 = T()
    out = [x.cpu() for x in rows]
    t6 = T()
    print(f"h2d {1e3*(t1-t0):.1f}  ingest {1e3*(t2-t1):.1f}  builder-init {1e3*(t3-t2):.1f}  build {1e3*(t4-t3):.1f}  to_rows {1e3*(t5-t4):.1f}  d2h {1e3*(t6-t5):.1f}  total {1e3*(t6-t0):.1f} ms")
    del f, c, b, t, rows, out
