"""ctypes face of the C restatement of the build-half oracle (oracle/covisit_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module; the
product package never does.  PARITY UNPINNED like the pandas oracle beside it (the reference ships no builder,
SURVEY.md section 0.1): this is a third, independent statement of SURVEY.md Appendix A, steps 1-7, returning the exact
integer accumulators (cnt, tsum, wsum) of every distinct pair; steps 8-9 (stable top-K) reuse covisit_oracle.topk.

The library is compiled on first use with gcc into oracle/_ref/ (git-ignored; __graft_entry__.build() does the same).
"""
from __future__ import annotations

import ctypes as C
import pathlib
import subprocess

import numpy as np
import pandas as pd

from . import covisit_oracle as co

_HERE = pathlib.Path(__file__).resolve().parent
SOURCE = _HERE / "covisit_oracle.c"
LIB_PATH = _HERE / "_ref" / "libcovisit_oracle.so"
_lib = None


class _Frame(C.Structure):
    _fields_ = [("n_events", C.c_int64), ("session", C.c_void_p), ("aid", C.c_void_p), ("ts", C.c_void_p), ("type", C.c_void_p)]


class _Recipe(C.Structure):
    _fields_ = [("event_type_mask", C.c_uint32), ("x_type_mask", C.c_uint32), ("y_type_mask", C.c_uint32),
                ("window_s", C.c_int32), ("tail_n", C.c_int32), ("ts_min", C.c_int32), ("type_weight", C.c_int32 * 3),
                ("x_lo", C.c_int32), ("x_hi", C.c_int32)]


def compile_library(force: bool = False) -> pathlib.Path:
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < SOURCE.stat().st_mtime:
        LIB_PATH.parent.mkdir(parents=True, exist_ok=True)
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", str(LIB_PATH), str(SOURCE)], check=True)
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        l = C.CDLL(str(compile_library()))
        l.covisit_oracle_accumulate.restype = C.c_void_p
        l.covisit_oracle_accumulate.argtypes = [C.POINTER(_Frame), C.POINTER(_Recipe)]
        l.covisit_oracle_count.restype = C.c_int64
        l.covisit_oracle_count.argtypes = [C.c_void_p]
        l.covisit_oracle_pairs.restype = C.c_int64
        l.covisit_oracle_pairs.argtypes = [C.c_void_p]
        l.covisit_oracle_fetch.restype = None
        l.covisit_oracle_fetch.argtypes = [C.c_void_p] * 6
        l.covisit_oracle_topk.restype = C.c_int64
        l.covisit_oracle_topk.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_double] + [C.c_void_p] * 5
        l.covisit_oracle_free.restype = None
        l.covisit_oracle_free.argtypes = [C.c_void_p]
        _lib = l
    return _lib


def _mask(types) -> int:
    return sum(1 << int(t) for t in set(types))


def _run(df: pd.DataFrame, spec: co.OracleSpec, x_range):
    tw = [int(w) for w in spec.type_weight]
    if any(float(a) != float(b) for a, b in zip(tw, spec.type_weight)):
        raise ValueError("the C oracle carries integer type weights")
    cols = [np.ascontiguousarray(df[c].to_numpy(), dtype=d) for c, d in
            (("session", np.int32), ("aid", np.int32), ("ts", np.int32), ("type", np.uint8))]
    frame = _Frame(len(df), *[c.ctypes.data for c in cols])
    x_lo, x_hi = (0, 2 ** 31 - 1) if x_range is None else (int(x_range[0]), int(x_range[1]))
    recipe = _Recipe(_mask(spec.event_types), _mask(spec.x_types), _mask(spec.y_types), int(spec.window_s), int(spec.tail_n),
                     int(spec.ts_min), (C.c_int32 * 3)(*tw), x_lo, x_hi)
    h = lib().covisit_oracle_accumulate(C.byref(frame), C.byref(recipe))
    if not h:
        raise MemoryError("covisit_oracle_accumulate failed (memory, 2^31 or more events, or tail_n outside 1..512)")
    return h


def accumulate(df: pd.DataFrame, spec: co.OracleSpec, x_range=None) -> pd.DataFrame:
    """Appendix A steps 1-7 over the whole frame (no chunking: integer sums do not depend on it).  -> one row per
    distinct pair, (aid_x, aid_y) ascending: aid_x, aid_y (int32), cnt, tsum = sum(ts_x - ts_min), wsum = sum of the
    type weights of the winner rows' type_y (int64).  `pairs` (rows after the in-session dedupe) is in .attrs.
    x_range = (lo, hi): only the rows lo <= aid_x < hi of the matrix (a row depends on nothing else)."""
    l = lib()
    h = _run(df, spec, x_range)
    try:
        n = int(l.covisit_oracle_count(h))
        out = {"aid_x": np.zeros(n, np.int32), "aid_y": np.zeros(n, np.int32), "cnt": np.zeros(n, np.int64),
               "tsum": np.zeros(n, np.int64), "wsum": np.zeros(n, np.int64)}
        l.covisit_oracle_fetch(h, *[a.ctypes.data for a in out.values()])
        pairs = int(l.covisit_oracle_pairs(h))
    finally:
        l.covisit_oracle_free(h)
    res = pd.DataFrame(out)
    res.attrs["pairs"] = pairs
    return res


def build_c(df: pd.DataFrame, spec: co.OracleSpec, x_range=None) -> pd.DataFrame:
    """Steps 1-8 entirely in C (for frames whose distinct pairs should not pass through pandas): the top-K table with
    columns aid_x, aid_y (int32), wgt (float32, formed once from the exact integers like weights()), cnt, tsum;
    .attrs carries pairs and distinct."""
    l = lib()
    h = _run(df, spec, x_range)
    try:
        n = int(l.covisit_oracle_count(h))
        out = {"aid_x": np.zeros(n, np.int32), "aid_y": np.zeros(n, np.int32), "wgt": np.zeros(n, np.float32),
               "cnt": np.zeros(n, np.int64), "tsum": np.zeros(n, np.int64)}
        w_scale = 3.0 / float(spec.ts_max - spec.ts_min)
        rows = int(l.covisit_oracle_topk(h, int(spec.weight_mode), int(spec.k), w_scale, *[a.ctypes.data for a in out.values()]))
        if rows < 0:
            raise ValueError("covisit_oracle_topk: k outside 1..64")
        pairs = int(l.covisit_oracle_pairs(h))
    finally:
        l.covisit_oracle_free(h)
    res = pd.DataFrame({c: a[:rows] for c, a in out.items()})
    res.attrs["pairs"], res.attrs["distinct"] = pairs, n
    return res


def table_digest(table: pd.DataFrame, k: int) -> dict:
    """bench.py's parity digest (table_digest there, on torch tensors) of a top-K table given as rows: wrapping 64-bit sum
    over the entries of mix(aid_x * k + rank inside the row, aid_y, float bits of wgt), and the number of rows."""
    ax = table["aid_x"].to_numpy().astype(np.int64)
    ay = table["aid_y"].to_numpy().astype(np.int64)
    wb = table["wgt"].to_numpy().astype(np.float32).view(np.int32).astype(np.int64)
    first = np.r_[True, ax[1:] != ax[:-1]] if len(ax) else np.zeros(0, bool)
    start = np.maximum.accumulate(np.where(first, np.arange(len(ax)), 0)) if len(ax) else np.zeros(0, np.int64)
    slot = ax * k + (np.arange(len(ax)) - start)
    with np.errstate(over="ignore"):
        h = (slot * np.int64(-7046029254386353131) + ay) * np.int64(-4658895280553007687)
        h = (h ^ (h >> np.int64(29))) * np.int64(-7723592293110705685) + wb * np.int64(2654435761)
        h = h ^ (h >> np.int64(32))
        total = int(h.sum(dtype=np.int64)) if len(h) else 0
    return {"rows": f"{total & 0xFFFFFFFFFFFFFFFF:016x}", "row_len_sum": int(len(ax))}


def weights(acc: pd.DataFrame, spec: co.OracleSpec) -> np.ndarray:
    """Step 6-7 weight of every distinct pair as ONE float32 formed from the exact integers: type / unit sums are exact
    integers; the time weight sum is cnt + 3 * tsum / (ts_max - ts_min) evaluated in fp64 and rounded once (the
    product's definition; the pandas oracle's float32 running sums agree within 1e-5 relative)."""
    if spec.weight_mode == co.WEIGHT_TIME:
        w = acc["cnt"].to_numpy().astype(np.float64) + (3.0 / float(spec.ts_max - spec.ts_min)) * acc["tsum"].to_numpy().astype(np.float64)
    elif spec.weight_mode == co.WEIGHT_TYPE:
        w = acc["wsum"].to_numpy().astype(np.float64)
    else:
        w = acc["cnt"].to_numpy().astype(np.float64)
    return w.astype(np.float32)


def build(df: pd.DataFrame, spec: co.OracleSpec) -> pd.DataFrame:
    """Whole build through the C accumulators: top-K table aid_x:int32, aid_y:int32, wgt:float32 (+ cnt, tsum)."""
    acc = accumulate(df, spec)
    acc = acc.assign(wgt=weights(acc, spec))
    return co.topk(acc[["aid_x", "aid_y", "wgt", "cnt", "tsum"]], spec.k)
