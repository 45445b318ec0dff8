"""The interaction-feature oracle (oracle/interaction_oracle.py) against a hand-computed frame: the row, session and aid
features of ranker/interaction_feature_engineering.py:50-111 on six candidate rows."""
import numpy as np
import pandas as pd

from oracle import interaction_oracle as io_


def test_hand_computed_features():
    ev = pd.DataFrame({"session": [1, 1, 1, 1, 2, 2], "aid": [7, 8, 7, 9, 8, 8], "ts": [10, 11, 12, 13, 5, 6], "type": [0, 1, 2, 0, 0, 0]})
    cand = pd.DataFrame({"session": [1, 1, 1, 2, 2, 2], "candidates": np.uint64([7, 8, 5, 8, 7, 5]),
                         "candidate_scores": np.float32([4, 2, 1, 3, 3, 5])})
    f = io_.interaction_features(cand, ev).set_index(["session", "candidates"])
    # aid 7 in session 1: events at positions 1 and 3 (1-based), types click and order
    r = f.loc[(1, 7)]
    assert (r.session_candidate_occurrence_count, r.session_candidate_cumcount_last) == (2, 3)
    assert (r.session_candidate_click_occurrence_count, r.session_candidate_cart_occurrence_count, r.session_candidate_order_occurrence_count) == (1, 0, 1)
    # aid 5 never occurs: counts 0, the script's null position shown as 0
    r = f.loc[(1, 5)]
    assert (r.session_candidate_occurrence_count, r.session_candidate_cumcount_last) == (0, 0)
    # session 1 aggregates over its rows (scores 4, 2, 1; occurrence counts 2, 1, 0; last positions 3, 2, null)
    assert np.isclose(r.session_candidate_score_mean, 7 / 3) and np.isclose(r.session_candidate_score_std, np.std([4, 2, 1], ddof=1))
    assert (r.session_candidate_score_min, r.session_candidate_score_max) == (1, 4)
    assert np.isclose(r.session_candidate_occurrence_count_mean, 1.0) and r.session_candidate_occurrence_count_sum == 3
    assert r.session_candidate_occurrence_count_max == 2
    assert np.isclose(r.session_candidate_cumcount_last_mean, 2.5) and r.session_candidate_cumcount_last_sum == 5 and r.session_candidate_cumcount_last_max == 3
    # aid 5 over both sessions: scores 1 and 5, never present -> null mean (NaN), sum 0, max 0
    assert np.isclose(r.aid_candidate_score_mean, 3.0) and np.isclose(r.aid_candidate_score_std, np.std([1, 5], ddof=1)) and r.aid_candidate_score_max == 5
    assert np.isnan(r.aid_session_candidate_cumcount_last_mean) and r.aid_session_candidate_cumcount_last_sum == 0 and r.aid_session_candidate_cumcount_last_max == 0
    # aid 8 over both sessions: occurrence counts 1 and 2, last positions 2 and 2
    r = f.loc[(2, 8)]
    assert r.aid_session_candidate_occurrence_count_sum == 3 and r.aid_session_candidate_occurrence_count_max == 2
    assert np.isclose(r.aid_session_candidate_cumcount_last_mean, 2.0)
    assert f.dtypes["session_candidate_occurrence_count"] == np.uint16 and f.dtypes["aid_candidate_score_std"] == np.float32


def test_duplicates_are_dropped_and_single_rows_have_no_std():
    ev = pd.DataFrame({"session": [3], "aid": [1], "ts": [0], "type": [1]})
    cand = pd.DataFrame({"session": [3, 3], "candidates": np.uint64([1, 1]), "candidate_scores": np.float32([2, 2])})
    f = io_.interaction_features(cand, ev)
    assert len(f) == 1 and np.isnan(f["session_candidate_score_std"].iloc[0]) and f["session_candidate_cart_occurrence_count"].iloc[0] == 1
