#!/usr/bin/env python
"""Staged scatter on ONE GPU: pass A (pairs into coarse buckets of a staging buffer) and pass B (place at the owner)
timed against the direct scatter, full-scale clicks frame, G simulated ranks (every buffer local, so pass B shows the
kernel's own cost without NVLink).  Prints one JSON line.

  python tools/time_staged.py --scale 1.0 --world 1
"""
import argparse
import json
import pathlib
import sys
from dataclasses import replace

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch

import __graft_entry__ as g

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--world", type=int, default=1)
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()
g.build()
from otto_multi_objective_recommender_system_b200 import covisit, distributed, synth

dev = torch.device("cuda:0")
frame = synth.generate(synth.SynthSpec.scaled("train", args.scale), device=dev)
csr = covisit.ingest(frame, "desc", device=dev)
del frame
G, S = args.world, csr.n_sessions
spec = replace(covisit.CLICKS, global_events=csr.n_events)


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


# the direct scatter of the whole frame, for reference
b = covisit.CovisitBuilder(csr, spec)
b.build()
direct = []
for _ in range(args.steps):
    b.count_begin()
    b.count_finish()
    t0 = ev()
    b.scatter()
    t1 = ev()
    torch.cuda.synchronize()
    direct.append(t0.elapsed_time(t1))
single = b.reduce()
digest = int(single.len.sum().item()), int(single.aid_y[single.len > 0][:, 0].to(torch.int64).sum().item())
del b

ranks = [distributed.GpuRankBackend(csr.slice_sessions(r * S // G, (r + 1) * S // G), spec) for r in range(G)]
out = {"plan": [], "pass_a": [], "pass_b": [], "partition": []}
staged = None
for step in range(args.steps + 1):
    counts = torch.stack([r.count_begin().clone() for r in ranks])
    cuts = None
    for i, r in enumerate(ranks):
        cuts, before = r.plan_owners(counts, G, i)
        r.count_finish_owned(cuts, i, before)
    t = [ev()]
    totals = [r.b.stage_plan(counts, G, i) for i, r in enumerate(ranks)][0]
    t.append(ev())
    if staged is None:
        staged = [torch.empty(max(n, 1), dtype=torch.int64, device=dev) for n in totals]
    for i, r in enumerate(ranks):
        r.b.scatter_staged(G, staged[i].data_ptr())
    t.append(ev())
    for i, r in enumerate(ranks):
        r.b.place_staged([s.data_ptr() for s in staged], cuts[i], cuts[i + 1])
    t.append(ev())
    parts = [r.partition() for r in ranks]
    t.append(ev())
    torch.cuda.synchronize()
    if step:
        for k, a, c in zip(out, t, t[1:]):
            out[k].append(a.elapsed_time(c))
# parity of the staged path against the direct build, owned rows
ok = True
for i, r in enumerate(ranks):
    records, bin_off = parts[i]
    bc = r.owner_bin_cuts(cuts, None)
    table = r.reduce([(records, bin_off[bc[i]:bc[i + 1] + 1])], bc[i], bc[i + 1], cuts[i], cuts[i + 1])
    lo, hi = cuts[i], cuts[i + 1]
    ok = ok and torch.equal(table.aid_y[lo:hi], single.aid_y[lo:hi]) and torch.equal(table.wgt[lo:hi], single.wgt[lo:hi]) \
        and torch.equal(table.len[lo:hi], single.len[lo:hi])
print(json.dumps({"scale": args.scale, "simulated_ranks": G, "direct_scatter_plus_partition_ms": min(direct),
                  "staged_ms": {k: min(v) for k, v in out.items()}, "staged_records": totals, "rows_equal": bool(ok),
                  "note": "all ranks on one device: times are sums over the simulated ranks, no NVLink"}))
