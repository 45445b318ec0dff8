#!/usr/bin/env python
"""Whole-table parity of the CUDA build against the plain-C oracle WITHOUT a GPU: recomputes, on the host, the digest
that `bench.py` prints as `parity.digest` (and commits under tests/golden/bench_digest.json when run with
--write-digest on a GPU) from the C restatement of the build recipe (oracle/covisit_oracle.c), and compares.

The digest is a wrapping 64-bit sum over every kept entry of mix(aid_x * k + rank, aid_y, float bits of wgt), plus the
row count, pairs P and distinct pairs D: equal digests mean every row of the GPU's top-K table - every aid_y, its
rank and every weight bit - equals the oracle's.  The matrix is accumulated one aid_x range at a time (a row depends
on nothing else), a few ranges in parallel processes, so that full scale (1.06 G pairs) fits a 64 GB host.

  python tools/verify_digest_cpu.py --scale 1.0 --ranges 8 --workers 3 --out profiles/r02_cpu_digest_full_scale.json
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

from oracle import covisit_oracle as co
from oracle import covisit_oracle_c as cc

_DF = None
SPECS = {"clicks": co.CLICKS, "carts_orders": co.CARTS_ORDERS, "buy2buy": co.BUY2BUY}


def _one_range(args):
    variant, lo, hi = args
    t0 = time.perf_counter()
    t = cc.build_c(_DF, SPECS[variant], x_range=(lo, hi))
    return t, t.attrs["pairs"], t.attrs["distinct"], time.perf_counter() - t0


def main():
    global _DF
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--variant", default="clicks", choices=sorted(SPECS))
    ap.add_argument("--ranges", type=int, default=8)
    ap.add_argument("--workers", type=int, default=3)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from otto_multi_objective_recommender_system_b200 import synth
    cc.lib()
    t0 = time.perf_counter()
    frame = synth.generate(synth.SynthSpec.scaled("train", args.scale))
    n_aids = frame.n_aids
    _DF = frame.to_pandas()
    del frame
    t_gen = time.perf_counter() - t0
    edges = np.linspace(0, n_aids, args.ranges + 1).astype(np.int64)
    jobs = [(args.variant, int(edges[r]), int(edges[r + 1])) for r in range(args.ranges)]
    t0 = time.perf_counter()
    if args.workers > 1:
        with mp.get_context("fork").Pool(args.workers) as pool:
            parts = pool.map(_one_range, jobs, chunksize=1)
    else:
        parts = [_one_range(j) for j in jobs]
    t_build = time.perf_counter() - t0
    table = pd.concat([p[0] for p in parts], ignore_index=True)
    spec = SPECS[args.variant]
    digest = cc.table_digest(table, spec.k)
    digest.update(pairs=int(sum(p[1] for p in parts)), distinct=int(sum(p[2] for p in parts)))
    digest["pair_checksum"] = digest["pairs"]       # the GPU's checksum is the sum of cnt over the distinct pairs = P
    key = f"{args.variant}@{args.scale:g}"
    golden = json.load(open(ROOT / "tests" / "golden" / "bench_digest.json")).get(key)
    out = {"key": key, "events": int(len(_DF)), "cpu_digest": digest, "gpu_digest_committed": golden,
           "equal": golden is not None and all(golden[k] == digest[k] for k in golden),
           "seconds": {"generate": round(t_gen, 1), "build": round(t_build, 1), "per_range": [round(p[3], 1) for p in parts]},
           "how": f"oracle/covisit_oracle.c, {args.ranges} aid_x ranges, {args.workers} processes; golden = tests/golden/bench_digest.json "
                  "(written by `bench.py --write-digest` on one B200; every N-GPU bench line is compared with it)"}
    line = json.dumps(out)
    print(line)
    if args.out:
        pathlib.Path(args.out).write_text(line + "\n")
    return 0 if out["equal"] else 1


if __name__ == "__main__":
    sys.exit(main())
