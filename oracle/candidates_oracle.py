"""CPU oracle for the covisitation CANDIDATE-GENERATION half of the hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package never does.

Parity is pinned by re-execution: the loops below restate, statement for statement, the per-session
bodies of
  * src/ranker/covisitation_candidate_generation.py:108-141 / :248-281  (ranker form, most_common(100))
  * src/covisitation/inference.py:204-247 / :396-441                    (standalone form, most_common(20)
    + history + popular fill), with the fastText/Annoy neighbour term dropped because neither library
    nor the model file exists offline.
The reference modules themselves cannot be imported (top-level `import polars`, `import settings` opens a
log file under /home/gunes, and all logic sits under `if __name__ == '__main__'`).  The one importable
unit, covisitation_df_to_dict, is restated verbatim-in-behaviour below and checked in tests against the
reference's own source text executed from /root/reference when that tree is present.

Tables that a caller does not supply are treated as empty dicts, which is what the reference's
`if aid in table` guards do for absent keys; with only the three north_star tables (time_weighted,
cart_weighted, cart_order) carts/orders are exactly the reference's lists and clicks reduces to the same
three lists.
"""
from __future__ import annotations

import itertools
from collections import Counter

import numpy as np
import pandas as pd

STEMS = ("time_weighted", "click_weighted", "cart_weighted", "order_weighted", "click_cart", "click_order",
         "cart_order")


def covisitation_df_to_dict(df: pd.DataFrame) -> dict:
    """src/ranker/covisitation_candidate_generation.py:16-32: only aid_x / aid_y and the row order matter."""
    return df.groupby("aid_x")["aid_y"].apply(list).to_dict()


def session_lists(df: pd.DataFrame) -> pd.DataFrame:
    """src/ranker/covisitation_candidate_generation.py:77: event order inside a session is file order."""
    return df.groupby("session")[["aid", "type"]].agg(list).reset_index()


def _gather(session_aids, session_event_types, tables):
    """:110-124 of the ranker script (identical to covisitation/inference.py:206-218)."""
    t = {s: tables.get(s, {}) for s in STEMS}
    session_unique_aids = list(dict.fromkeys(session_aids[::-1]))
    a = np.array(session_aids)
    e = np.array(session_event_types)
    session_unique_click_and_cart_aids = np.unique(a[e <= 1]).tolist()
    session_unique_cart_and_order_aids = np.unique(a[e >= 1]).tolist()
    time_weighted = list(itertools.chain(*[t["time_weighted"][aid] for aid in session_unique_aids if aid in t["time_weighted"]]))
    click_weighted = list(itertools.chain(*[t["click_weighted"][aid] for aid in session_unique_click_and_cart_aids if aid in t["click_weighted"]]))
    cart_weighted = list(itertools.chain(*[t["cart_weighted"][aid] for aid in session_unique_click_and_cart_aids if aid in t["cart_weighted"]]))
    order_weighted = list(itertools.chain(*[t["order_weighted"][aid] for aid in session_unique_cart_and_order_aids if aid in t["order_weighted"]]))  # computed, unused (:122)
    click_cart = list(itertools.chain(*[t["click_cart"][aid] for aid in session_unique_click_and_cart_aids if aid in t["click_cart"]]))
    cart_order = list(itertools.chain(*[t["cart_order"][aid] for aid in session_unique_click_and_cart_aids if aid in t["cart_order"]]))  # over C01, not C12 (:124)
    del order_weighted
    clicks = time_weighted + click_weighted + cart_weighted + click_cart + cart_order      # :127
    carts = time_weighted + cart_weighted + cart_order                                     # :133
    orders = time_weighted + cart_weighted + cart_order                                    # :138
    return session_unique_aids, clicks, carts, orders


def ranker_candidates(session_aids, session_event_types, tables, n: int = 100):
    """One session of the ranker form: [(aids, counts)] for clicks, carts, orders (:127-141)."""
    unique, *lists = _gather(session_aids, session_event_types, tables)
    out = []
    for concat in lists:
        kept = [(aid, count) for aid, count in Counter(concat).most_common(n) if aid not in unique]
        out.append(([a for a, _ in kept], [c for _, c in kept]))
    return out


def regular_candidates(session_aids, session_event_types, tables, n: int = 100):
    """One session of ranker/regular_candidate_generation.py:139-180 (fastText term dropped): the session's unique
    aids (most recent first, scores |H| .. 1) followed by the ranker-form votes -> [(aids, scores)] per target."""
    unique, *lists = _gather(session_aids, session_event_types, tables)
    out = []
    for concat in lists:
        kept = [(aid, weight) for aid, weight in Counter(concat).most_common(n) if aid not in unique]
        weights = np.arange(1, len(unique) + 1).tolist()[::-1] + [weight for _, weight in kept]     # :163
        out.append((unique + [aid for aid, _ in kept], weights))
    return out


def regular_frame(df_events: pd.DataFrame, tables: dict, n: int = 100) -> dict:
    """All sessions -> the exploded frames of ranker/regular_candidate_generation.py:225-257."""
    sess = session_lists(df_events)
    rows = {"click": [], "cart": [], "order": []}
    for t in sess.itertuples():
        res = regular_candidates(t.aid, t.type, tables, n)
        for name, (aids, weights) in zip(("click", "cart", "order"), res):
            rows[name].extend((t.session, a, w) for a, w in zip(aids, weights))
    out = {}
    for name, r in rows.items():
        f = pd.DataFrame(r, columns=["session", "candidates", "candidate_scores"])
        f["candidates"] = f["candidates"].astype(np.uint64)
        f["candidate_scores"] = f["candidate_scores"].astype(np.float32)
        out[name] = f
    return out


def standalone_predictions(session_aids, session_event_types, tables, popular, n: int = 20):
    """One session of covisitation/inference.py:227-243 (covisitation branch; fastText term dropped).

    popular = (click, cart, order) most-frequent-aid lists (data/aid_frequencies/*.json, :76-83)."""
    unique, *lists = _gather(session_aids, session_event_types, tables)
    out = []
    for concat, pop in zip(lists, popular):
        sorted_aids = [aid for aid, count in Counter(concat).most_common(n) if aid not in unique]
        pred = unique + sorted_aids[:n - len(unique)]
        pred = pred + pop[:n - len(pred)]
        out.append(pred)
    return out


def recency_predictions(session_aids, session_event_types, tables, n: int = 20):
    """One long session (>= 20 unique aids) of covisitation/inference.py:142-199 (= :336-392), statement for
    statement, without the fastText / Annoy bonus (:165-170)."""
    event_type_coefficient = {0: 1, 1: 9, 2: 6}                                             # :72
    t = {s: tables.get(s, {}) for s in STEMS}
    a = np.array(session_aids)
    e = np.array(session_event_types)
    session_unique_click_aids = np.unique(a[e == 0]).tolist()
    session_unique_click_and_cart_aids = np.unique(a[e <= 1]).tolist()
    session_unique_cart_and_order_aids = np.unique(a[e >= 1]).tolist()
    click_recency_weights = np.logspace(0.1, 1, len(session_aids), base=2, endpoint=True) - 1
    cart_recency_weights = np.logspace(0.5, 1, len(session_aids), base=2, endpoint=True) - 1
    order_recency_weights = np.logspace(0.5, 1, len(session_aids), base=2, endpoint=True) - 1
    session_aid_click_weights = Counter()
    session_aid_cart_weights = Counter()
    session_aid_order_weights = Counter()
    for aid, event_type, cw, kw, ow in zip(session_aids, session_event_types, click_recency_weights, cart_recency_weights, order_recency_weights):
        session_aid_click_weights[aid] += (cw * event_type_coefficient[event_type])
        session_aid_cart_weights[aid] += (kw * event_type_coefficient[event_type])
        session_aid_order_weights[aid] += (ow * event_type_coefficient[event_type])
    for aid in itertools.chain(*[t["time_weighted"][x] for x in session_unique_click_aids if x in t["time_weighted"]]):
        session_aid_click_weights[aid] += 0.05
    sorted_click_aids = [aid for aid, weight in session_aid_click_weights.most_common(n)]
    for aid in itertools.chain(*[t["cart_weighted"][x] for x in session_unique_click_and_cart_aids if x in t["cart_weighted"]]):
        session_aid_cart_weights[aid] += 0.05
    sorted_cart_aids = [aid for aid, weight in session_aid_cart_weights.most_common(n)]
    for aid in itertools.chain(*[t["cart_order"][x] for x in session_unique_cart_and_order_aids if x in t["cart_order"]]):
        session_aid_order_weights[aid] += 0.15
    sorted_order_aids = [aid for aid, weight in session_aid_order_weights.most_common(n)]
    return [sorted_click_aids, sorted_cart_aids, sorted_order_aids]


def recency_weighted_candidates(session_aids, session_event_types):
    """One session of ranker/recency_weighted_candidate_generator.py:61-93 (= :169-201 for the test set), statement
    for statement: -> [(aids, weights)] for click, cart, order; every unique aid of the session, weight desc."""
    event_type_coefficient = {0: 1, 1: 6, 2: 1}                                             # :25
    session_unique_aids = list(dict.fromkeys(session_aids[::-1]))
    click_recency_weights = np.logspace(0.1, 1, len(session_aids), base=2, endpoint=True) - 1
    cart_recency_weights = np.logspace(0.5, 1, len(session_aids), base=2, endpoint=True) - 1
    order_recency_weights = np.logspace(0.5, 1, len(session_aids), base=2, endpoint=True) - 1
    session_aid_click_weights = Counter()
    session_aid_cart_weights = Counter()
    session_aid_order_weights = Counter()
    for aid, event_type, cw, kw, ow in zip(session_aids, session_event_types, click_recency_weights, cart_recency_weights, order_recency_weights):
        session_aid_click_weights[aid] += (cw * event_type_coefficient[event_type])
        session_aid_cart_weights[aid] += (kw * event_type_coefficient[event_type])
        session_aid_order_weights[aid] += (ow * event_type_coefficient[event_type])
    out = []
    for counter in (session_aid_click_weights, session_aid_cart_weights, session_aid_order_weights):
        ranked = counter.most_common(len(session_unique_aids))
        out.append(([aid for aid, weight in ranked], [weight for aid, weight in ranked]))
    return out


def recency_weighted_frame(df_events: pd.DataFrame) -> dict:
    """All sessions -> the exploded frames the script pickles (:117-144): session, candidates uint64,
    candidate_scores float32."""
    sess = session_lists(df_events)
    rows = {"click": [], "cart": [], "order": []}
    for t in sess.itertuples():
        res = recency_weighted_candidates(list(t.aid), list(t.type))
        for name, (aids, weights) in zip(("click", "cart", "order"), res):
            rows[name].extend((t.session, a, w) for a, w in zip(aids, weights))
    out = {}
    for name, r in rows.items():
        f = pd.DataFrame(r, columns=["session", "candidates", "candidate_scores"])
        f["candidates"] = f["candidates"].astype(np.uint64)
        f["candidate_scores_f64"] = f["candidate_scores"].astype(np.float64)
        f["candidate_scores"] = f["candidate_scores"].astype(np.float32)
        out[name] = f
    return out


def ranker_frame(df_events: pd.DataFrame, tables: dict, n: int = 100) -> dict:
    """All sessions -> the three exploded candidate frames the ranker script pickles (:177-197 / :290-307):
    columns session, candidates uint64, candidate_scores float32."""
    sess = session_lists(df_events)
    rows = {"click": [], "cart": [], "order": []}
    for t in sess.itertuples():
        res = ranker_candidates(t.aid, t.type, tables, n)
        for name, (aids, counts) in zip(("click", "cart", "order"), res):
            rows[name].extend((t.session, a, c) for a, c in zip(aids, counts))
    out = {}
    for name, r in rows.items():
        f = pd.DataFrame(r, columns=["session", "candidates", "candidate_scores"])
        f["candidates"] = f["candidates"].astype(np.uint64)
        f["candidate_scores"] = f["candidate_scores"].astype(np.float32)
        out[name] = f
    return out


def recall_at_20(pred: list, labels: list) -> float:
    """covisitation/inference.py:251-257: sum |pred ∩ label| / sum min(|label|, 20)."""
    hits = sum(len(set(p).intersection(set(l))) for p, l in zip(pred, labels))
    denom = sum(min(len(l), 20) for l in labels)
    return hits / denom if denom else 0.0


def split_for_recall(session_aids, session_event_types, cutoff: int):
    """validation.py:9-52 label semantics at a given cutoff index: the events up to and including `cutoff`
    are the history; labels are the next click and all later carts / orders."""
    hist_a, hist_t = session_aids[:cutoff + 1], session_event_types[:cutoff + 1]
    fut_a, fut_t = session_aids[cutoff + 1:], session_event_types[cutoff + 1:]
    click = [a for a, t in zip(fut_a, fut_t) if t == 0][:1]
    carts = list(dict.fromkeys(a for a, t in zip(fut_a, fut_t) if t == 1))
    orders = list(dict.fromkeys(a for a, t in zip(fut_a, fut_t) if t == 2))
    return (hist_a, hist_t), (click, carts, orders)
