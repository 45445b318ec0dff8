"""A second, independent witness for the build-half oracle: SURVEY.md Appendix A written as plain Python loops (no
merge / drop_duplicates / groupby), compared with the pandas oracle on random small frames (hypothesis).  The build
half has no vectors from the reference (it ships no builder); this guards the oracle against depending on the
behaviour of one pandas version."""
import numpy as np
import pandas as pd
from hypothesis import given, settings, strategies as st

from oracle import covisit_oracle as co


def brute_force(df: pd.DataFrame, spec: co.OracleSpec):
    """-> {(aid_x, aid_y): (count, sum(ts_x - ts_min), float32 weight sum in session order)} and the top-K rows."""
    rows = [tuple(int(v) for v in r) for r in df[["session", "aid", "ts", "type"]].itertuples(index=False)]
    if set(spec.event_types) != {0, 1, 2}:
        rows = [r for r in rows if r[3] in spec.event_types]                       # step 1
    sessions = {}
    for pos, (s, a, t, y) in enumerate(rows):
        sessions.setdefault(s, []).append((a, t, y, pos))
    acc = {}
    for s in sorted(sessions):
        ev = sorted(sessions[s], key=lambda e: (-e[1], e[3]))[:spec.tail_n]        # steps 2-3: ts desc, ties in row order
        seen = set()
        for (ax, tx, yx, _) in ev:                                                 # step 4: i-major ...
            for (ay, ty, yy, _) in ev:                                             # ... j-minor
                if abs(tx - ty) >= spec.window_s or ax == ay:
                    continue
                if yx not in spec.x_types or yy not in spec.y_types:
                    continue
                if (ax, ay) in seen:                                               # step 5: the first row wins
                    continue
                seen.add((ax, ay))
                if spec.weight_mode == co.WEIGHT_TIME:                             # step 6
                    w = np.float32(1.0 + 3.0 * (float(tx) - spec.ts_min) / float(spec.ts_max - spec.ts_min))
                elif spec.weight_mode == co.WEIGHT_TYPE:
                    w = np.float32(spec.type_weight[yy])
                else:
                    w = np.float32(1.0)
                c, ts, ws = acc.get((ax, ay), (0, 0, []))
                acc[(ax, ay)] = (c + 1, ts + tx - spec.ts_min, ws + [w])
    table = []
    for x in sorted({k[0] for k in acc}):                                          # steps 7-8
        # pandas' groupby sum of float32 runs in float64 and rounds once; the sums here are short
        ent = [(y, np.float32(np.sum(np.array(ws, dtype=np.float64)))) for (xx, y), (_, _, ws) in acc.items() if xx == x]
        ent.sort(key=lambda e: (-float(e[1]), e[0]))
        table += [(x, y, w) for y, w in ent[:spec.k]]
    return acc, table


frames = st.lists(
    st.tuples(st.integers(0, 5), st.integers(0, 7), st.integers(0, 40), st.integers(0, 2)), min_size=0, max_size=70)
specs = st.sampled_from([
    co.OracleSpec(co.WEIGHT_UNIT, window_s=7, tail_n=5, k=3, ts_min=0, ts_max=40),
    co.OracleSpec(co.WEIGHT_TYPE, window_s=12, tail_n=30, k=4, ts_min=0, ts_max=40),
    co.OracleSpec(co.WEIGHT_TIME, window_s=9, tail_n=8, k=20, ts_min=0, ts_max=40),
    co.OracleSpec(co.WEIGHT_UNIT, event_types=(1, 2), window_s=25, tail_n=30, k=15, ts_min=0, ts_max=40),
    co.OracleSpec(co.WEIGHT_TYPE, x_types=(0,), y_types=(1, 2), window_s=40, tail_n=6, k=5, ts_min=0, ts_max=40),
])


@settings(max_examples=120, deadline=None)
@given(frames, specs)
def test_pandas_oracle_equals_plain_loops(rows, spec):
    df = pd.DataFrame(rows, columns=["session", "aid", "ts", "type"]).astype(
        {"session": np.int32, "aid": np.int32, "ts": np.int32, "type": np.int8})
    acc, table = brute_force(df, spec)
    got = co.accumulate(df, spec, exact=True)
    assert len(got) == len(acc)
    for r in got.itertuples(index=False):
        c, ts, ws = acc[(int(r.aid_x), int(r.aid_y))]
        assert (int(r.cnt), int(r.tsum)) == (c, ts)
        if spec.weight_mode != co.WEIGHT_TIME:
            assert float(r.wgt) == float(sum(float(w) for w in ws))               # small integers: exact
        else:
            assert abs(float(r.wgt) - float(np.sum(np.array(ws, dtype=np.float64)))) <= 1e-5 * max(1.0, abs(float(r.wgt)))
    top = co.build(df, spec)
    if spec.weight_mode != co.WEIGHT_TIME:
        assert [(int(a), int(b), float(w)) for a, b, w in top.itertuples(index=False)] == [(x, y, float(w)) for x, y, w in table]
    else:
        # float sums may differ in the last bit between the two summation orders: compare rows and lengths per aid_x
        assert top.groupby("aid_x").size().to_dict() == pd.Series([x for x, _, _ in table]).value_counts().sort_index().to_dict() \
            if len(table) else len(top) == 0
