"""CPU oracle for the interaction features over a candidate frame.  TEST INFRASTRUCTURE ONLY.

Only tests/ (and __graft_entry__.smoke / bench.py's CPU legs) may import this module; the product package never does.

Restates src/ranker/interaction_feature_engineering.py:31-113 in pandas.  PARITY UNPINNED: the reference script is a
polars 0.15.1 program (requirements.txt:88) whose library is absent here, it holds no test or golden vector, and its
logic sits under `if __name__ == '__main__'` behind `import settings`, so it cannot be executed.  Two behaviours of
that polars version are therefore CHOSEN here and stated, not observed:
  * `df_candidate.unique()` (:31, :84) and `.sort('session')` (:32) leave the row order inside a session unspecified;
    the oracle (and the CUDA path) keep the input order, and tests compare per (session, candidate);
  * `session_candidate_cumcount_last` is null for a candidate that never occurs in its session (left join, no
    fill_null: :71-79).  polars aggregates skip nulls: mean over no values is null (NaN here), sum over no values is
    0 and max over no values is null (0 here, the column is cast to an unsigned type next).
Everything else follows the script line by line:
  :47     events of the candidate sessions, sorted (session, ts) ascending
  :50-54  session_aid_cumcount = 1-based position of the event inside its session
  :55-57  session_candidate_cumcount_last = that position at the LAST occurrence of (session, aid)
  :59     session_candidate_occurrence_count = events of (session, aid)
  :60     session_candidate_{click,cart,order}_occurrence_count = events of (session, aid, type)
  :63-84  left joins onto the candidate rows on (session, candidates); counts fill_null(0), cast UInt16
  :86-97  per-session aggregates of candidate_scores (mean, std [ddof 1], min, max), of the occurrence count
          (mean, sum, max) and of cumcount_last (mean, sum, max), joined back on session
  :101-111 the same per candidate aid (without the score minimum), joined back on candidates
"""
from __future__ import annotations

import numpy as np
import pandas as pd

SESSION_FEATURES = ("session_candidate_score_mean", "session_candidate_score_std", "session_candidate_score_min",
                    "session_candidate_score_max", "session_candidate_occurrence_count_mean",
                    "session_candidate_occurrence_count_sum", "session_candidate_occurrence_count_max",
                    "session_candidate_cumcount_last_mean", "session_candidate_cumcount_last_sum",
                    "session_candidate_cumcount_last_max")
AID_FEATURES = ("aid_candidate_score_mean", "aid_candidate_score_std", "aid_candidate_score_max",
                "aid_session_candidate_occurrence_count_mean", "aid_session_candidate_occurrence_count_sum",
                "aid_session_candidate_occurrence_count_max", "aid_session_candidate_cumcount_last_mean",
                "aid_session_candidate_cumcount_last_sum", "aid_session_candidate_cumcount_last_max")
ROW_FEATURES = ("session_candidate_occurrence_count", "session_candidate_cumcount_last",
                "session_candidate_click_occurrence_count", "session_candidate_cart_occurrence_count",
                "session_candidate_order_occurrence_count")


def interaction_features(df_candidate: pd.DataFrame, df_events: pd.DataFrame) -> pd.DataFrame:
    """df_candidate: session, candidates, candidate_scores [, candidate_labels]; df_events: session, aid, ts, type.
    Returns the candidate frame (input row order, duplicates dropped) with the feature columns appended."""
    cand = df_candidate.drop_duplicates().reset_index(drop=True)                                   # :31
    cand = cand.assign(session=cand["session"].astype(np.int32), candidates=cand["candidates"].astype(np.int32))   # :33
    cand = cand.sort_values("session", kind="stable").reset_index(drop=True)                      # :32 (order inside a session kept)
    ev = df_events[df_events["session"].isin(cand["session"])]                                    # :47
    ev = ev.sort_values(["session", "ts"], kind="stable").reset_index(drop=True)                  # :48
    ev = ev.assign(session_aid_cumcount=ev.groupby("session").cumcount() + 1)                      # :50-54
    g = ev.groupby(["session", "aid"])
    per_aid = pd.DataFrame({"session_candidate_cumcount_last": g["session_aid_cumcount"].last(),          # :55-57
                            "session_candidate_occurrence_count": g["aid"].count()}).reset_index()     # :59
    per_aid = per_aid.rename(columns={"aid": "candidates"})
    out = cand.merge(per_aid, on=["session", "candidates"], how="left")                            # :63-69
    out["session_candidate_occurrence_count"] = out["session_candidate_occurrence_count"].fillna(0).astype(np.uint16)   # :70
    for value, name in enumerate(("click", "cart", "order")):                                      # :72-83
        col = f"session_candidate_{name}_occurrence_count"
        t = ev[ev["type"] == value].groupby(["session", "aid"])["aid"].count().rename(col).reset_index()
        t = t.rename(columns={"aid": "candidates"})
        out = out.merge(t, on=["session", "candidates"], how="left")
        out[col] = out[col].fillna(0).astype(np.uint16)
    last = out["session_candidate_cumcount_last"]                                                  # float with NaN = null

    def aggregates(key: str, prefix_score: str, prefix_occ: str, with_min: bool) -> pd.DataFrame:
        grp = out.groupby(key, sort=False)
        f = pd.DataFrame({f"{prefix_score}_mean": grp["candidate_scores"].mean().astype(np.float32),
                          f"{prefix_score}_std": grp["candidate_scores"].std(ddof=1).astype(np.float32)})
        if with_min:
            f[f"{prefix_score}_min"] = grp["candidate_scores"].min().astype(np.float32)
        f[f"{prefix_score}_max"] = grp["candidate_scores"].max().astype(np.float32)
        occ = grp["session_candidate_occurrence_count"]
        f[f"{prefix_occ}_occurrence_count_mean"] = occ.mean().astype(np.float32)
        f[f"{prefix_occ}_occurrence_count_sum"] = occ.sum().astype(np.uint32)
        f[f"{prefix_occ}_occurrence_count_max"] = occ.max().astype(np.uint16)
        lg = grp["session_candidate_cumcount_last"]
        f[f"{prefix_occ}_cumcount_last_mean"] = lg.mean().astype(np.float32)                      # NaN when no candidate of the group occurs
        f[f"{prefix_occ}_cumcount_last_sum"] = lg.sum().astype(np.uint32)                         # sum over no values = 0
        f[f"{prefix_occ}_cumcount_last_max"] = lg.max().fillna(0).astype(np.uint16)               # max over no values: null -> 0
        return f.reset_index()

    out = out.merge(aggregates("session", "session_candidate_score", "session_candidate", True), on="session", how="left")      # :86-98
    out = out.merge(aggregates("candidates", "aid_candidate_score", "aid_session_candidate", False), on="candidates", how="left")  # :101-112
    out["session_candidate_cumcount_last"] = last.fillna(0).astype(np.uint16)                      # null -> 0 (position 0 does not exist)
    return out
