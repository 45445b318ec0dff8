"""GPU parity of the interaction features (csrc/features.cu) against the pandas restatement of
ranker/interaction_feature_engineering.py (oracle/interaction_oracle.py), on the regular candidate form of a
synthetic validation frame (history aids in the candidates exercise the occurrence / position features) and through
the CLI twin.  Integer features bit-exact; float features within 1e-6 relative (fp64 -> fp32 casts on both sides)."""
import numpy as np
import pandas as pd
import pytest
import torch

from oracle import candidates_oracle as oc
from oracle import interaction_oracle as ioc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods(native_lib):
    from otto_multi_objective_recommender_system_b200 import candidates, covisit, interaction_feature_engineering, synth
    return covisit, candidates, interaction_feature_engineering, synth


def _tables(cv, n_aids, seed):
    rng = np.random.default_rng(seed)
    tables = {}
    for stem, k in (("time_weighted", 20), ("cart_weighted", 15), ("cart_order", 15)):
        rows = []
        for x in range(n_aids):
            n = int(rng.integers(0, k + 1))
            rows += [(x, int(y), float(n - i)) for i, y in enumerate(rng.choice(n_aids, size=n, replace=False))]
        df = pd.DataFrame(rows, columns=["aid_x", "aid_y", "wgt"])
        tables[stem] = cv.TopKTable.from_rows(*(torch.tensor(df[c].to_numpy(), device="cuda:0") for c in ("aid_x", "aid_y")),
                                              torch.tensor(df["wgt"].to_numpy(), dtype=torch.float32, device="cuda:0"), n_aids, k)
    return tables


def _compare(got: pd.DataFrame, want: pd.DataFrame):
    assert len(got) == len(want)
    g = got.set_index(["session", "candidates"]).sort_index()
    w = want.set_index(["session", "candidates"]).sort_index()
    assert g.index.equals(w.index)
    for col in ioc.ROW_FEATURES + ioc.SESSION_FEATURES + ioc.AID_FEATURES:
        a, b = g[col].to_numpy(), w[col].to_numpy()
        assert a.dtype == b.dtype, (col, a.dtype, b.dtype)
        if a.dtype.kind == "f":
            assert np.array_equal(np.isnan(a), np.isnan(b)), col
            np.testing.assert_allclose(a[~np.isnan(a)], b[~np.isnan(b)], rtol=1e-6, atol=0, err_msg=col)
        else:
            assert np.array_equal(a, b), col


def test_interaction_features_match_oracle(mods):
    cv, cand_mod, ife, synth = mods
    n_aids = 300
    frame = synth.generate(synth.SynthSpec("test", 1500, n_aids, seed=41, first_session=5000))
    df = frame.to_pandas()
    # a few long sessions with heavy repeats so that positions beyond 255 and counts above 1 appear
    rng = np.random.default_rng(3)
    extra = pd.DataFrame({"session": np.repeat([9000, 9001], [300, 40]), "aid": rng.integers(0, 12, 340),
                          "ts": np.arange(340), "type": rng.choice([0, 1, 2], 340).astype(np.int8)})
    df = pd.concat([df, extra], ignore_index=True)
    sess = cv.ingest(synth.EventFrame.from_pandas(df, n_aids), "asc", device="cuda:0")
    frames = cand_mod.regular_candidates(sess, _tables(cv, n_aids, 5), 100)
    for event in ("click", "order"):
        got = ife.interaction_features(sess, frames[event])
        want = ioc.interaction_features(frames[event], df)
        _compare(got, want)
        assert int(got["session_candidate_occurrence_count"].max()) > 1 and int(got["session_candidate_cumcount_last"].max()) > 255


def test_scores_with_fraction_and_duplicate_rows(mods):
    cv, _, ife, synth = mods
    df = pd.DataFrame({"session": [1, 1, 1, 2, 4, 4], "aid": [3, 4, 3, 3, 9, 9], "ts": [1, 2, 3, 1, 1, 2], "type": [0, 1, 2, 0, 1, 1]})
    sess = cv.ingest(synth.EventFrame.from_pandas(df, 16), "asc", device="cuda:0")
    cand = pd.DataFrame({"session": [4, 1, 1, 1, 2, 2, 4, 1], "candidates": np.uint64([9, 3, 4, 5, 3, 11, 3, 3]),
                         "candidate_scores": np.float32([0.5, 2.25, 1, 7, 3, 0.75, 2, 2.25])})      # one duplicate row
    got = ife.interaction_features(sess, cand)
    want = ioc.interaction_features(cand, df)
    _compare(got, want)
    assert len(got) == 7
