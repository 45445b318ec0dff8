"""Shared comparison helpers for the parity tests (oracle = checker, never the thing under test)."""
import numpy as np
import pandas as pd

from oracle import covisit_oracle as co

RTOL_TIME = 1e-5   # north_star: time-weighted sums agree within 1e-5 relative


def oracle_spec(spec) -> co.OracleSpec:
    """Product CovisitSpec -> OracleSpec (same fields, float type weights)."""
    return co.OracleSpec(weight_mode=spec.weight_mode, type_weight=tuple(float(w) for w in spec.type_weight),
                         event_types=tuple(spec.event_types), x_types=tuple(spec.x_types), y_types=tuple(spec.y_types),
                         window_s=spec.window_s, tail_n=spec.tail_n, k=spec.k, ts_min=spec.ts_min, ts_max=spec.ts_max)


def gpu_formula_topk(acc: pd.DataFrame, spec) -> pd.DataFrame:
    """Top-K of the oracle's exact integer accumulators under the GPU's weight definition
    wgt = float32(cnt + 3 * tsum / (ts_max - ts_min)); ties by aid_y ascending."""
    w = (acc["cnt"].to_numpy().astype(np.float64)
         + (3.0 / float(spec.ts_max - spec.ts_min)) * acc["tsum"].to_numpy().astype(np.float64)).astype(np.float32)
    t = acc.assign(wgt=w)
    return co.topk(t[["aid_x", "aid_y", "wgt", "cnt", "tsum"]], spec.k)


def assert_int_table_equal(got: pd.DataFrame, want: pd.DataFrame, what: str) -> None:
    assert len(got) == len(want), f"{what}: {len(got)} rows vs oracle {len(want)}"
    for c in ("aid_x", "aid_y"):
        assert np.array_equal(got[c].to_numpy(), want[c].to_numpy()), f"{what}: column {c} differs"
    assert np.array_equal(got["wgt"].to_numpy().view(np.uint32), want["wgt"].to_numpy().view(np.uint32)), \
        f"{what}: weights differ"


def assert_time_table_close(got: pd.DataFrame, acc: pd.DataFrame, k: int, what: str) -> None:
    """got = GPU rows; acc = oracle accumulate() rows (float32 pandas sums).  Every GPU pair must carry the
    oracle's weight within RTOL_TIME, and the GPU row set may differ from the oracle's top-K only where the
    oracle's weights tie with the K-th weight within that tolerance."""
    key = lambda d: d["aid_x"].to_numpy().astype(np.int64) << 32 | d["aid_y"].to_numpy().astype(np.int64)
    ow = pd.Series(acc["wgt"].to_numpy(), index=key(acc))
    gk = key(got)
    assert ow.index.is_unique
    missing = ~np.isin(gk, ow.index.to_numpy())
    assert not missing.any(), f"{what}: GPU emitted pairs the oracle never saw"
    w_or = ow.loc[gk].to_numpy()
    np.testing.assert_allclose(got["wgt"].to_numpy(), w_or, rtol=RTOL_TIME, atol=0, err_msg=what)
    want = co.topk(acc, k)
    assert len(got) == len(want), f"{what}: {len(got)} rows vs oracle {len(want)}"
    assert np.array_equal(got["aid_x"].to_numpy(), want["aid_x"].to_numpy()), what
    diff = np.nonzero(got["aid_y"].to_numpy() != want["aid_y"].to_numpy())[0]
    if diff.size:
        # any disagreement must be a swap among near-tied weights
        w_want = want["wgt"].to_numpy()[diff]
        w_got = w_or[diff]
        np.testing.assert_allclose(w_got, w_want, rtol=2 * RTOL_TIME, atol=0,
                                   err_msg=f"{what}: top-K differs beyond weight ties")


# ---- vectors produced by executing the reference's own loop bodies (tests/golden/make_reference_vectors.py) ----

REFERENCE_VECTOR_FILES = ["reference_candidates.json", "reference_candidates_wide.json"]


def reference_vectors(name="reference_candidates.json"):
    """reference_candidates.json: 60 short sessions, 80 aids, rows of <= 6; reference_candidates_wide.json: K = 15 rows,
    sessions of up to 458 events, more than 100 distinct candidates (most_common(100) truncates)."""
    import json
    import pathlib
    return json.load(open(pathlib.Path(__file__).resolve().parent / "golden" / name))


def reference_vector_inputs(g):
    """-> (event frame as DataFrame, {stem: {aid_x: [aid_y...]}}, labels {event: {session: set}}, popular lists)."""
    rows = []
    for s in g["sessions"]:
        rows += [(s["session"], a, 1661724000 + i, t) for i, (a, t) in enumerate(zip(s["aid"], s["type"]))]
    df = pd.DataFrame(rows, columns=["session", "aid", "ts", "type"])
    tables = {stem: {int(x): ys for x, ys in r.items()} for stem, r in g["tables"].items()}
    labels = {"click": {}, "cart": {}, "order": {}}
    for s in g["sessions"]:
        if s["click_labels"] != []:
            labels["click"][s["session"]] = {s["click_labels"]}
        if s["cart_labels"]:
            labels["cart"][s["session"]] = set(s["cart_labels"])
        if s["order_labels"]:
            labels["order"][s["session"]] = set(s["order_labels"])
    return df, tables, labels, g["popular"]


def table_rows(rows: dict) -> pd.DataFrame:
    """{aid_x: [aid_y best first]} -> the (aid_x, aid_y, wgt) rows of a part file (weights only encode the order)."""
    out = [(x, y, float(len(ys) - i)) for x, ys in sorted(rows.items()) for i, y in enumerate(ys)]
    return pd.DataFrame(out, columns=["aid_x", "aid_y", "wgt"])


def check_candidate_frames(g, family: str, frames: dict, score_column: str = "candidate_scores", with_labels: bool = True):
    """frames = {event: exploded frame (session, candidates, candidate_scores[, candidate_labels])} of ALL sessions of the
    vectors, in session order; family = 'ranker' | 'recency' | 'regular'."""
    for event in ("click", "cart", "order"):
        f = frames[event]
        by_session = {s: grp for s, grp in f.groupby("session", sort=False)}
        for s, want in zip(g["sessions"], g[family]):
            aids, scores, labels = want[event]
            grp = by_session.get(s["session"])
            got_aids = [] if grp is None else [int(a) for a in grp["candidates"]]
            assert got_aids == aids, (family, event, s["session"])
            got_scores = [] if grp is None else grp[score_column].tolist()
            if score_column == "candidate_scores":       # the pickled column is float32
                assert got_scores == [float(np.float32(w)) for w in scores], (family, event, s["session"])
            else:                                        # fp64 Counter values, bit for bit
                assert got_scores == scores, (family, event, s["session"])
            if with_labels:
                got_labels = [] if grp is None else [int(l) for l in grp["candidate_labels"]]
                assert got_labels == labels, (family, event, s["session"])
        assert list(dict.fromkeys(f["session"])) == [s["session"] for s, w in zip(g["sessions"], g[family]) if w[event][0]]


def check_standalone_predictions(g, preds: dict):
    """preds = {event: [list of aids per session]} (padding removed)."""
    for event in ("click", "cart", "order"):
        for s, want, got in zip(g["sessions"], g["standalone"], preds[event]):
            assert [int(a) for a in got] == want[event], (event, s["session"])
