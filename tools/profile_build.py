#!/usr/bin/env python
"""One warm build followed by one build inside an NVTX range "step" - the target of the ncu recipes:

  ncu --nvtx --nvtx-include "step/" --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_build.py --scale 1.0
  ncu --nvtx --nvtx-include "step/" --set full --clock-control none --import-source on -k regex:reduce \
      -o gpurun_out/reduce python tools/profile_build.py --scale 0.2
"""
import argparse
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch

import __graft_entry__ as g

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--variant", default="clicks")
ap.add_argument("--split-ub", type=int, default=0)
ap.add_argument("--what", default="build", choices=["build", "candidates"])
args = ap.parse_args()

g.build()
from dataclasses import replace

from otto_multi_objective_recommender_system_b200 import covisit, synth

dev = torch.device("cuda:0")
frame = synth.generate(synth.SynthSpec.scaled("train", args.scale), device=dev)
csr = covisit.ingest(frame, "desc", device=dev)
del frame
spec = {"clicks": covisit.CLICKS, "carts_orders": covisit.CARTS_ORDERS, "buy2buy": covisit.BUY2BUY}[args.variant]
if args.split_ub:
    spec = replace(spec, split_ub=args.split_ub)
b = covisit.CovisitBuilder(csr, spec)
b.build()
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("step")
b.build()
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print(b.stats.as_dict())
