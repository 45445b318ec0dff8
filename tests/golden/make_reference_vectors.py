"""Golden vectors for candidate generation produced by EXECUTING THE REFERENCE'S OWN per-session loop bodies.

The reference scripts cannot be imported (logic under `if __name__ == '__main__'`, `import polars`, `import settings`
opens a log file under /home/gunes), but their per-session loops are plain Python over numpy / Counter.  This script
cuts the loop bodies out of the files under /root/reference/src as TEXT (unmodified, only dedented), executes them per
session on a small seeded input and stores inputs and outputs in tests/golden/reference_candidates.json:

  ranker        ranker/covisitation_candidate_generation.py   first `for t in tqdm(df_val.itertuples()` loop
  standalone    covisitation/inference.py                     recency loop (>= 20 unique aids) + covisitation loop
  recency       ranker/recency_weighted_candidate_generator.py first `for idx, row in tqdm(df_val.iterrows()` loop
  regular       ranker/regular_candidate_generation.py        first `for t in tqdm(df_val.itertuples()` loop

The submission-mode twins of these loops (`df_test`) are executed too and must give identical lists.

The fastText / Annoy neighbour lookup (no model offline) is neutralised by a stub index that returns only the query
item, so `fasttext_similar_aids` is empty - exactly the term DESIGN.md states as dropped.  Everything else is the
reference's code.  tests/test_oracle.py checks the oracle restatements against these vectors (CPU), and
tests/test_zz_reference_vectors_gpu.py checks the CUDA path against them (GPU).

Run from the repo root in the build container (needs /root/reference):  python tests/golden/make_reference_vectors.py
"""
import collections
import itertools
import json
import pathlib
import re
import textwrap
from collections import Counter

import numpy as np

REF = pathlib.Path("/root/reference/src")
OUT = pathlib.Path(__file__).resolve().parent


def loop_body(path: pathlib.Path, for_prefix: str, nth: int = 0) -> str:
    """Source text of the body of the nth `for` loop whose header starts with for_prefix, dedented."""
    lines = path.read_text().split("\n")
    hits = [i for i, l in enumerate(lines) if l.strip().startswith(for_prefix)]
    start = hits[nth]
    indent = len(lines[start]) - len(lines[start].lstrip())
    body = []
    for l in lines[start + 1:]:
        if l.strip() and len(l) - len(l.lstrip()) <= indent:
            break
        body.append(l)
    return textwrap.dedent("\n".join(body))


class Recorder:
    """Stands in for df_val: `df_val.at[index, column] = value` is all the loop bodies do with it."""

    def __init__(self):
        self.cells = collections.defaultdict(dict)
        self.at = self

    def __setitem__(self, key, value):
        index, column = key
        self.cells[index][column] = value


class StubAnnoy:
    def get_nns_by_item(self, i, n, search_k=-1, include_distances=False):
        return [i]            # only the query item itself: the loops drop element 0


PROFILES = {
    # the round-1 set: 60 short sessions, 80 aids, rows of <= 6 neighbours
    "reference_candidates.json": dict(seed=7, n_aids=80, k=6, full_rows=False,
                                      lengths=([1, 2, 2, 3, 3, 4, 5, 6, 8, 10, 12, 15, 19, 20, 21, 22, 25, 30, 40, 60] * 3)),
    # the wide set: full K = 15 rows, sessions up to OTTO's longest test session (458 events),
    # hundreds of distinct candidates so that most_common(100) truncates, long histories for the recency branch
    "reference_candidates_wide.json": dict(seed=11, n_aids=900, k=15, full_rows=True,
                                           lengths=[1, 2, 3, 5, 8, 13, 19, 20, 21, 30, 45, 60, 80, 81, 82, 100, 128, 200, 256,
                                                    300, 400, 458, 458, 257, 129, 83, 64, 33, 32, 31]),
}


def make_inputs(seed=7, n_aids=80, k=6, full_rows=False, lengths=()):
    rng = np.random.default_rng(seed)
    n_sessions = len(lengths)
    stems = ("time_weighted", "click_weighted", "cart_weighted", "order_weighted", "click_cart", "click_order", "cart_order")
    tables = {}
    for stem in stems:
        rows = {}
        for x in range(n_aids):
            if rng.random() < 0.8:
                n = k if full_rows and rng.random() < 0.9 else int(rng.integers(1, k + 1))
                rows[x] = [int(y) for y in rng.choice(n_aids, size=n, replace=False)]
        tables[stem] = rows
    sessions = []
    for s, L in enumerate(lengths[:n_sessions]):
        pool = n_aids if s % 3 else max(3, L // 2)
        aids = [int(a) for a in rng.integers(0, pool, size=L)]
        types = [int(t) for t in rng.choice([0, 1, 2], size=L, p=[0.7, 0.2, 0.1])]
        click = int(rng.integers(0, n_aids)) if rng.random() < 0.7 else []
        carts = [int(a) for a in rng.choice(n_aids, size=int(rng.integers(0, 4)), replace=False)]
        orders = [int(a) for a in rng.choice(n_aids, size=int(rng.integers(0, 3)), replace=False)]
        sessions.append({"session": 1000 + s, "aid": aids, "type": types, "click_labels": click, "cart_labels": carts,
                         "order_labels": orders})
    popular = json.load(open(OUT / "popular.json"))
    return n_aids, tables, sessions, popular


def namespace(tables, popular, coefficient):
    ns = {"np": np, "itertools": itertools, "Counter": Counter, "annoy_index": StubAnnoy(),
          "aid_idx": collections.defaultdict(int), "idx_aid": {0: None}, "event_type_coefficient": coefficient}
    for stem, rows in tables.items():
        ns[f"top_{stem}_covisitation"] = rows
    for event in ("click", "cart", "order"):
        ns[f"most_frequent_{event}_aids"] = popular[event]
    return ns


def coefficient_of(path: pathlib.Path) -> dict:
    m = re.search(r"event_type_coefficient = (\{[^}]*\})", path.read_text())
    return eval(m.group(1))           # a dict literal such as {0: 1, 1: 9, 2: 6}


def run(body: str, ns: dict, sessions, style: str, only=None, frame_name: str = "df_val"):
    code = compile(body, "<reference loop body>", "exec")
    rec = Recorder()
    ns = dict(ns, **{frame_name: rec})
    T = collections.namedtuple("T", ["Index", "session", "aid", "type", "click_labels", "cart_labels", "order_labels"])
    for i, s in enumerate(sessions):
        if only is not None and not only(s):
            continue
        if style == "itertuples":
            ns["t"] = T(i, s["session"], s["aid"], s["type"], s["click_labels"], s["cart_labels"], s["order_labels"])
        else:
            ns["idx"], ns["row"] = i, s
        exec(code, ns)
    return rec.cells


def main():
    for name, profile in PROFILES.items():
        make(name, profile)


def make(name, profile):
    n_aids, tables, sessions, popular = make_inputs(**profile)
    out = {"n_aids": n_aids, "table_k": profile["k"], "tables": {st: {str(x): ys for x, ys in rows.items()} for st, rows in tables.items()},
           "sessions": sessions, "popular": popular}

    p = REF / "ranker" / "covisitation_candidate_generation.py"
    cells = run(loop_body(p, "for t in tqdm(df_val.itertuples()"), namespace(tables, popular, None), sessions, "itertuples")
    out["ranker"] = [{e: [cells[i][f"{e}_candidates"], cells[i][f"{e}_candidate_scores"], cells[i][f"{e}_candidate_labels"]]
                      for e in ("click", "cart", "order")} for i in range(len(sessions))]

    p = REF / "covisitation" / "inference.py"
    ns = namespace(tables, popular, coefficient_of(p))
    is_long = lambda s: len(set(s["aid"])) >= 20                  # :127-131 of the script
    cells = run(loop_body(p, "for t in tqdm(df_val.loc[recency_weight_predictions_idx].itertuples()"), ns, sessions, "itertuples", is_long)
    cells2 = run(loop_body(p, "for t in tqdm(df_val.loc[covisitation_predictions_idx].itertuples()"), ns, sessions, "itertuples",
                 lambda s: not is_long(s))
    cells.update(cells2)
    out["standalone"] = [{e: cells[i][f"{e}_predictions"] for e in ("click", "cart", "order")} for i in range(len(sessions))]
    out["standalone_coefficient"] = {str(k): v for k, v in coefficient_of(p).items()}

    p = REF / "ranker" / "recency_weighted_candidate_generator.py"
    cells = run(loop_body(p, "for idx, row in tqdm(df_val.iterrows()"), namespace(tables, popular, coefficient_of(p)), sessions, "iterrows")
    out["recency"] = [{e: [cells[i][f"{e}_candidates"], [float(w) for w in cells[i][f"{e}_candidate_scores"]],
                           cells[i][f"{e}_candidate_labels"]] for e in ("click", "cart", "order")} for i in range(len(sessions))]

    p = REF / "ranker" / "regular_candidate_generation.py"
    cells = run(loop_body(p, "for t in tqdm(df_val.itertuples()"), namespace(tables, popular, None), sessions, "itertuples")
    out["regular"] = [{e: [cells[i][f"{e}_candidates"], cells[i][f"{e}_candidate_scores"], cells[i][f"{e}_candidate_labels"]]
                       for e in ("click", "cart", "order")} for i in range(len(sessions))]

    # the submission-mode loops of the four scripts (df_test: no labels) must give the same lists on the same sessions
    def same(cells, key, family, pick=lambda v: v):
        for i in range(len(sessions)):
            for e in ("click", "cart", "order"):
                want = out[family][i][e] if family == "standalone" else out[family][i][e][0 if key.endswith("candidates") else 1]
                assert pick(cells[i][f"{e}_{key}"]) == want, (family, key, i, e)
    p = REF / "ranker" / "covisitation_candidate_generation.py"
    cells = run(loop_body(p, "for t in tqdm(df_test.itertuples()"), namespace(tables, popular, None), sessions, "itertuples", frame_name="df_test")
    same(cells, "candidates", "ranker")
    same(cells, "candidate_scores", "ranker")
    p = REF / "covisitation" / "inference.py"
    ns = namespace(tables, popular, coefficient_of(p))
    ns["test_predictions"] = []          # the submission loops append {'session_type': '<session>_<event>s', 'labels': 'a b c'} (:387-391, :437-441)
    run(loop_body(p, "for t in tqdm(df_test.loc[recency_weight_predictions_idx].itertuples()"), ns, sessions, "itertuples", is_long, "df_test")
    run(loop_body(p, "for t in tqdm(df_test.loc[covisitation_predictions_idx].itertuples()"), ns, sessions, "itertuples",
        lambda s: not is_long(s), "df_test")
    index_of = {s["session"]: i for i, s in enumerate(sessions)}
    cells = collections.defaultdict(dict)
    for row in ns["test_predictions"]:
        sid, event = row["session_type"].rsplit("_", 1)
        cells[index_of[int(sid)]][f"{event[:-1]}_predictions"] = [int(a) for a in row["labels"].split()]
    assert len(ns["test_predictions"]) == 3 * len(sessions)
    same(cells, "predictions", "standalone")
    p = REF / "ranker" / "recency_weighted_candidate_generator.py"
    cells = run(loop_body(p, "for idx, row in tqdm(df_test.iterrows()"), namespace(tables, popular, coefficient_of(p)), sessions, "iterrows",
                frame_name="df_test")
    same(cells, "candidates", "recency")
    same(cells, "candidate_scores", "recency", lambda v: [float(w) for w in v])
    p = REF / "ranker" / "regular_candidate_generation.py"
    cells = run(loop_body(p, "for t in tqdm(df_test.itertuples()"), namespace(tables, popular, None), sessions, "itertuples", frame_name="df_test")
    same(cells, "candidates", "regular")
    same(cells, "candidate_scores", "regular")
    out["submission_loops_identical"] = True

    def plain(o):
        if isinstance(o, (np.integer,)):
            return int(o)
        if isinstance(o, (np.floating,)):
            return float(o)
        raise TypeError(type(o))
    json.dump(out, open(OUT / name, "w"), default=plain, separators=(",", ":"))
    most = max(len(out["ranker"][i]["click"][0]) for i in range(len(sessions)))
    print(name, "sessions", len(sessions), "long", sum(is_long(s) for s in sessions), "longest", max(len(s["aid"]) for s in sessions),
          "most ranker candidates", most, "bytes", (OUT / name).stat().st_size)


if __name__ == "__main__":
    main()
