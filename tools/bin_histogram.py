#!/usr/bin/env python
"""Distribution of bin sizes (records per bin) of one build: how many bins and how many records fall into each
power-of-two size class, and how many rows end with fewer than K entries.  Input of the reduce tier design."""
import argparse
import json
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch

import __graft_entry__ as g

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--variant", default="clicks")
args = ap.parse_args()
g.build()
from otto_multi_objective_recommender_system_b200 import covisit, synth

dev = torch.device("cuda:0")
frame = synth.generate(synth.SynthSpec.scaled("train", args.scale), device=dev)
csr = covisit.ingest(frame, "desc", device=dev)
del frame
spec = {"clicks": covisit.CLICKS, "carts_orders": covisit.CARTS_ORDERS, "buy2buy": covisit.BUY2BUY}[args.variant]
b = covisit.CovisitBuilder(csr, spec)
t = b.build()
torch.cuda.synchronize()
st = b.stats.as_dict()
B = st["bins"]
off = b.views()["bin_offsets"][:B + 1]
n = (off[1:] - off[:-1]).to(torch.int64)
cls = torch.where(n > 0, torch.floor(torch.log2(n.clamp(min=1).double())).long() + 1, torch.zeros_like(n))
out = {"stats": st, "classes": []}
for c in range(int(cls.max().item()) + 1):
    m = cls == c
    out["classes"].append({"records_lt": 0 if c == 0 else 2 ** c, "bins": int(m.sum().item()), "records": int(n[m].sum().item())})
ln = t.len
out["rows_len_lt_k"] = int((ln < spec.k).sum().item())
out["rows_len_0"] = int((ln == 0).sum().item())
out["rows"] = int(ln.numel())
print(json.dumps(out, indent=1))
