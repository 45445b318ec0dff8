"""CPU oracle for the covisitation-matrix BUILD half of the hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package never does.

PARITY UNPINNED: /root/reference contains no covisitation builder - every script under
src/covisitation and src/ranker only *reads* pre-built top_15_<stem>_<part>.pqt files
(src/covisitation/inference.py:87-111, src/ranker/covisitation_candidate_generation.py:49-73).  The
matrices were made out of tree with cuDF 22.10 (requirements.txt:23).  This file therefore restates the
north_star recipe in order-preserving pandas, using the session self-merge idiom
that is visible in the reference at src/matrix_factorization/torch_trainer.py:198-223
(`df.merge(df, on='session')`, `aid_x != aid_y`, `ts_x` / `ts_y`, `groupby(['aid_x', 'aid_y'])`).
There are no golden vectors in the reference for this half; tests/golden holds vectors produced by THIS
oracle plus the hand-derived EDA session-747 fixture (notebook cell 37).

Steps (Appendix A numbering):
  1 type pre-filter   2 sort (session asc, ts desc, stable)   3 keep the 30 most recent per session
  4 self-merge on session; keep |ts_x - ts_y| < W (strict), aid_x != aid_y (and optional x/y type masks)
  5 drop_duplicates(session, aid_x, aid_y) keeps the first row (smallest i, then smallest j)
  6 weight from the winner row   7 groupby(aid_x, aid_y).sum()   8 stable top-K (wgt desc, aid_y asc)
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import pandas as pd

TS_MIN = 1659304800          # dataset min ts (EDA notebook cell 6)
TS_MAX = 1662328791          # dataset max ts incl. test week

WEIGHT_UNIT, WEIGHT_TYPE, WEIGHT_TIME = 0, 1, 2


@dataclass
class OracleSpec:
    weight_mode: int = WEIGHT_TIME
    type_weight: tuple = (1.0, 6.0, 3.0)      # {0:1, 1:6, 2:3}: baseline/aid_weight.py:34
    event_types: tuple = (0, 1, 2)            # pre-filter before the tail cut (buy2buy: (1, 2))
    x_types: tuple = (0, 1, 2)                # pair-level filters on type_x / type_y (generic spec)
    y_types: tuple = (0, 1, 2)
    window_s: int = 86400
    tail_n: int = 30
    k: int = 20
    ts_min: int = TS_MIN
    ts_max: int = TS_MAX
    chunk_sessions: int = 100_000             # utilities/split_dataset_writer_parquet.py:23


CLICKS = OracleSpec(WEIGHT_TIME, k=20)
CARTS_ORDERS = OracleSpec(WEIGHT_TYPE, k=15)
BUY2BUY = OracleSpec(WEIGHT_UNIT, event_types=(1, 2), window_s=14 * 86400, k=15)


def dedup_pairs(df: pd.DataFrame, spec: OracleSpec) -> pd.DataFrame:
    """Steps 1-6 for one frame: the per-session deduplicated pair rows with their weight.

    Returns columns session, aid_x, aid_y, ts_x, type_y, wgt (float32) in merge order."""
    df = df[["session", "aid", "ts", "type"]]
    if set(spec.event_types) != {0, 1, 2}:
        df = df.loc[df["type"].isin(spec.event_types)]
    df = df.sort_values(["session", "ts"], ascending=[True, False], kind="stable").reset_index(drop=True)
    df = df.loc[df.groupby("session").cumcount() < spec.tail_n]
    m = df.merge(df, on="session")
    keep = ((m["ts_x"].astype(np.int64) - m["ts_y"].astype(np.int64)).abs() < spec.window_s) & (m["aid_x"] != m["aid_y"])
    if set(spec.x_types) != {0, 1, 2}:
        keep &= m["type_x"].isin(spec.x_types)
    if set(spec.y_types) != {0, 1, 2}:
        keep &= m["type_y"].isin(spec.y_types)
    m = m.loc[keep]
    m = m.drop_duplicates(["session", "aid_x", "aid_y"])
    if spec.weight_mode == WEIGHT_TIME:
        w = 1.0 + 3.0 * (m["ts_x"].astype(np.float64) - spec.ts_min) / float(spec.ts_max - spec.ts_min)
    elif spec.weight_mode == WEIGHT_TYPE:
        w = m["type_y"].map({0: spec.type_weight[0], 1: spec.type_weight[1], 2: spec.type_weight[2]})
    else:
        w = pd.Series(1.0, index=m.index)
    m = m.assign(wgt=w.astype(np.float32))
    return m[["session", "aid_x", "aid_y", "ts_x", "type_y", "wgt"]]


def accumulate(df: pd.DataFrame, spec: OracleSpec, exact: bool = False) -> pd.DataFrame:
    """Steps 1-7 with the reference's 100k-session chunking; one row per distinct (aid_x, aid_y).

    With exact=True the integer accumulators (cnt, tsum = sum(ts_x - ts_min)) ride along so that the GPU's
    integer form of the time weight can be compared bit-exact before any float is formed."""
    sessions = np.sort(df["session"].unique())
    acc = None
    acc_int = None
    for lo in range(0, len(sessions), spec.chunk_sessions):
        chunk_ids = sessions[lo: lo + spec.chunk_sessions]
        part = df.loc[(df["session"] >= chunk_ids[0]) & (df["session"] <= chunk_ids[-1])]
        pairs = dedup_pairs(part, spec)
        s = pairs.groupby(["aid_x", "aid_y"])["wgt"].sum()
        acc = s if acc is None else acc.add(s, fill_value=0)
        if exact:
            pairs = pairs.assign(cnt=np.int64(1), tsum=pairs["ts_x"].astype(np.int64) - spec.ts_min)
            si = pairs.groupby(["aid_x", "aid_y"])[["cnt", "tsum"]].sum()
            acc_int = si if acc_int is None else acc_int.add(si, fill_value=0)
    if acc is None:
        cols = {"aid_x": np.int32, "aid_y": np.int32, "wgt": np.float32}
        out = pd.DataFrame({c: np.zeros(0, t) for c, t in cols.items()})
        if exact:
            out["cnt"] = np.zeros(0, np.int64)
            out["tsum"] = np.zeros(0, np.int64)
        return out
    out = acc.astype(np.float32).reset_index()
    if exact:
        ai = acc_int.astype(np.int64).reset_index()
        out["cnt"] = ai["cnt"].to_numpy()
        out["tsum"] = ai["tsum"].to_numpy()
    return out


def topk(acc: pd.DataFrame, k: int) -> pd.DataFrame:
    """Step 8: rows are (aid_x, aid_y) ascending on entry; the stable sort breaks wgt ties by aid_y asc."""
    t = acc.sort_values(["aid_x", "wgt"], ascending=[True, False], kind="stable").reset_index(drop=True)
    t = t.loc[t.groupby("aid_x").cumcount() < k].reset_index(drop=True)
    return t.astype({"aid_x": np.int32, "aid_y": np.int32, "wgt": np.float32})


def build(df: pd.DataFrame, spec: OracleSpec, exact: bool = False) -> pd.DataFrame:
    """Whole build: frame -> top-K table with columns aid_x:int32, aid_y:int32, wgt:float32 (step 9)."""
    return topk(accumulate(df, spec, exact=exact), spec.k)


def split_parts(table: pd.DataFrame, n_parts: int, n_aids: int) -> list:
    """Disjoint contiguous aid_x ranges, like the reference's top_15_<stem>_<part>.pqt pieces."""
    edges = np.linspace(0, n_aids, n_parts + 1).astype(np.int64)
    return [table.loc[(table["aid_x"] >= edges[p]) & (table["aid_x"] < edges[p + 1])].reset_index(drop=True)
            for p in range(n_parts)]


def pair_stats(df: pd.DataFrame, spec: OracleSpec) -> dict:
    """P (pairs after in-session dedupe) and D (distinct pairs) for publishing beside a run."""
    acc = accumulate(df, spec, exact=True)
    return {"pairs": int(acc["cnt"].sum()), "distinct": int(len(acc))}
