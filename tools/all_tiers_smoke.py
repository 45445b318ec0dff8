#!/usr/bin/env python
"""A small build that reaches every reduce tier (rows of 31 .. 6145 records, hand-overs, split rows), both CSR orders
and the candidate kernels in a few seconds - the case to run under a checker or after touching a kernel
(compute-sanitizer is closed on this pool, so in round 2 it only ran plain)."""
import pathlib, sys
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import pandas as pd
import torch
import __graft_entry__ as g
g.build()
from otto_multi_objective_recommender_system_b200 import candidates, covisit, synth

sizes = [31, 33, 384, 385, 1536, 1537, 3072, 3073, 6144, 6145]
rng = np.random.default_rng(17)
rows, s, ts = [], 0, 1660000000
for i, n in enumerate(sizes):
    for x, pool in ((i, n), (50 + i, 50)):
        for j in range(n):
            y = 100 + (j if pool == n else int(rng.integers(0, pool)))
            rows += [(s, x, ts, int(rng.integers(0, 3))), (s, y, ts + 1, int(rng.integers(0, 3)))]
            s += 1
            ts += int(rng.integers(1, 60))
for k in range(6):                       # tied entries: hand-over to the hash-table kernel
    rows += [(s, 90, ts, 0)] + [(s, 7000 + 29 * k + j, ts, 0) for j in range(29)]
    s += 1
df = pd.DataFrame(rows, columns=["session", "aid", "ts", "type"])
frame = synth.EventFrame.from_pandas(df, 7300)
for order in ("desc", "asc"):
    csr = covisit.ingest(frame, order, device="cuda:0")
    for name, spec in covisit.VARIANTS.items():
        t, st = covisit.build_topk(csr, spec, exact=True)
        torch.cuda.synchronize()
        print(order, name, st["pairs"], st["distinct"], st["tier_records"], st["split_rows"])
train = synth.generate(synth.SynthSpec("train", 1500, 200, seed=3))
test = synth.generate(synth.SynthSpec("test", 400, 200, seed=4, first_session=1500))
csr = covisit.ingest(train, "desc", device="cuda:0")
tables = {stem: covisit.build_topk(csr, spec)[0] for stem, spec in covisit.VARIANTS.items()}
sess = covisit.ingest(test, "asc", device="cuda:0")
cand = candidates.generate_candidates(sess, tables, candidates.reference_spec(tables.keys(), 20))
pred, long_s = candidates.assemble_predictions(sess, cand, {t: list(range(20)) for t in ("click", "cart", "order")}, 20)
candidates.recency_long_predictions(sess, tables, pred, long_s, 20)
torch.cuda.synchronize()
print("candidates ok", int(long_s.sum()))
