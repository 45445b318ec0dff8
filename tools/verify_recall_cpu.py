#!/usr/bin/env python
"""The recall@20 check of `bench.py`'s pipeline, recomputed WITHOUT a GPU from tables the plain-C oracle builds.

bench.py (recall_check) takes the first --recall-sample test sessions, cuts each at a seeded random event, runs the CUDA
candidate path on the histories with the three CUDA-built tables and compares the lists and the three recalls with the
restated reference loops (oracle/candidates_oracle.py) fed the SAME table rows; it prints both recalls in the bench
line (`pipeline.recall_at_20`, `pipeline.recall_at_20_oracle`).  Here the three tables (clicks / time_weighted,
carts-orders / cart_weighted, buy2buy / cart_order) come from oracle/covisit_oracle.c instead, at the same scale, and
the same reference loops run on them: if the recalls equal the ones a GPU run printed, the CUDA tables of all three
variants agree with the oracle's on every row the sampled sessions touch, and the candidate lists built from them score
the same.  (Whole-table equality of the clicks matrix: tools/verify_digest_cpu.py.)

  python tools/verify_recall_cpu.py --scale 1.0 --bench-line profiles/r02_bench_n1.json --out profiles/r02_cpu_recall_full_scale.json
"""
import argparse
import json
import multiprocessing as mp
import pathlib
import sys
import time

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import numpy as np
import pandas as pd

from oracle import candidates_oracle as oc
from oracle import covisit_oracle as co
from oracle import covisit_oracle_c as cc

STEMS = {"time_weighted": co.CLICKS, "cart_weighted": co.CARTS_ORDERS, "cart_order": co.BUY2BUY}   # covisit.VARIANTS
_DF = None
_NEED = None


def _rows_of_range(args):
    stem, lo, hi = args
    t = cc.build_c(_DF, STEMS[stem], x_range=(lo, hi))
    t = t.loc[np.isin(t["aid_x"].to_numpy(), _NEED), ["aid_x", "aid_y"]]
    return t


def main():
    global _DF, _NEED
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--recall-sample", type=int, default=10000)
    ap.add_argument("--ranges", type=int, default=12)
    ap.add_argument("--workers", type=int, default=3)
    ap.add_argument("--bench-line", default=None, help="a bench.py JSON line whose pipeline.recall_at_20(_oracle) to compare with")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from otto_multi_objective_recommender_system_b200 import synth
    cc.lib()
    t0 = time.perf_counter()
    # ---- the sampled histories and their labels, exactly as bench.py's recall_check forms them
    test = synth.generate(synth.SynthSpec.scaled("test", args.scale))
    ses, aid_all, typ_all = test.session.numpy(), test.aid.numpy(), test.type.numpy()
    starts = np.flatnonzero(np.r_[True, ses[1:] != ses[:-1]])
    off = np.r_[starts, len(ses)].astype(np.int64)
    n = min(args.recall_sample, len(starts))
    rng = np.random.default_rng(7)
    rows, labels = [], {"click": [], "cart": [], "order": []}
    for i in range(n):
        a, t = aid_all[off[i]:off[i + 1]].tolist(), typ_all[off[i]:off[i + 1]].tolist()
        if len(a) < 2:
            continue
        cut = int(rng.integers(0, len(a) - 1))
        (ha, ht), (c, k, o) = oc.split_for_recall(a, t, cut)
        rows.append(pd.DataFrame({"session": int(ses[off[i]]), "aid": ha, "ts": np.arange(len(ha)), "type": ht}))
        labels["click"].append(c)
        labels["cart"].append(k)
        labels["order"].append(o)
    hdf = pd.concat(rows, ignore_index=True)
    _NEED = np.unique(hdf["aid"].to_numpy())
    # ---- the three tables from the C oracle, only the rows the histories can touch
    train = synth.generate(synth.SynthSpec.scaled("train", args.scale))
    n_aids = train.n_aids
    _DF = train.to_pandas()
    del train
    edges = np.linspace(0, n_aids, args.ranges + 1).astype(np.int64)
    otables = {}
    for stem in STEMS:
        jobs = [(stem, int(edges[r]), int(edges[r + 1])) for r in range(args.ranges)]
        if args.workers > 1:
            with mp.get_context("fork").Pool(args.workers) as pool:
                parts = pool.map(_rows_of_range, jobs, chunksize=1)
        else:
            parts = [_rows_of_range(j) for j in jobs]
        t = pd.concat(parts, ignore_index=True)
        otables[stem] = {int(x): [int(v) for v in g] for x, g in t.groupby("aid_x", sort=True)["aid_y"]}
    # ---- the restated reference loops on them
    hl = oc.session_lists(hdf)
    pops = [list(range(20))] * 3
    want = {"click": [], "cart": [], "order": []}
    n_long = 0
    for t in hl.itertuples():
        if len(set(t.aid)) >= 20:
            w = oc.recency_predictions(t.aid, t.type, otables, 20)
            n_long += 1
        else:
            w = oc.standalone_predictions(t.aid, t.type, otables, pops, 20)
        for ti, name in enumerate(("click", "cart", "order")):
            want[name].append(list(w[ti]))
    recall = {name: oc.recall_at_20(want[name], [[x] if not isinstance(x, (list, set, tuple)) else list(x) for x in labels[name]])
              for name in ("click", "cart", "order")}
    out = {"scale": args.scale, "recall_sample_sessions": len(hl), "long_sessions_in_sample": n_long,
           "table_rows_touched": {s: len(v) for s, v in otables.items()}, "recall_at_20_cpu_tables": recall,
           "seconds": round(time.perf_counter() - t0, 1),
           "how": "tables from oracle/covisit_oracle.c, lists from oracle/candidates_oracle.py; no GPU involved"}
    if args.bench_line:
        line = json.loads(pathlib.Path(args.bench_line).read_text().strip().splitlines()[-1])["pipeline"]
        out["bench_line"] = args.bench_line
        out["recall_at_20_gpu_lists"] = {k: v for k, v in line["recall_at_20"].items() if k != "weighted"}
        out["recall_at_20_oracle_on_gpu_tables"] = line["recall_at_20_oracle"]
        out["sample_sessions_in_bench_line"] = line["recall_sample_sessions"]
        out["equal"] = (out["recall_at_20_gpu_lists"] == recall and line["recall_at_20_oracle"] == recall
                        and line["recall_sample_sessions"] == len(hl))
    text = json.dumps(out)
    print(text)
    if args.out:
        pathlib.Path(args.out).write_text(text + "\n")
    return 0 if out.get("equal", True) else 1


if __name__ == "__main__":
    sys.exit(main())
