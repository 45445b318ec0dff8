#!/usr/bin/env python
"""Candidate generation once warm, once inside an NVTX range "step" (for ncu --nvtx-include "step/")."""
import argparse, pathlib, sys, time
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
import __graft_entry__ as g
ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=0.2)
ap.add_argument("--top-n", type=int, default=20)
args = ap.parse_args()
g.build()
from otto_multi_objective_recommender_system_b200 import candidates, covisit, synth
dev = torch.device("cuda:0")
frame = synth.generate(synth.SynthSpec.scaled("train", args.scale), device=dev)
csr = covisit.ingest(frame, "desc", device=dev)
A = csr.n_aids
tables = {stem: covisit.build_topk(csr, spec)[0] for stem, spec in covisit.VARIANTS.items()}
del frame, csr
test = synth.generate(synth.SynthSpec.scaled("test", args.scale), device=dev)
sess = covisit.ingest(test, "asc", device=dev)
gen = candidates.CandidateGenerator(tables, candidates.reference_spec(tables.keys(), args.top_n), A)
mlen = candidates.max_session_len(sess)
gen(sess, mlen)
torch.cuda.synchronize()
t0 = time.perf_counter()
torch.cuda.nvtx.range_push("step")
gen(sess, mlen)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print(f"sessions {sess.n_sessions} events {sess.n_events} max_len {mlen}: {1e3 * (time.perf_counter() - t0):.3f} ms")
