"""Interaction features over a candidate frame on one B200 - twin of src/ranker/interaction_feature_engineering.py.

    python -m otto_multi_objective_recommender_system_b200.interaction_feature_engineering {validation|submission} --data DIR

  reads  DIR/candidate/{click,cart,order}_{validation,test}.pkl   (ranker/regular_candidate_generation.py:225-257;
         the files our regular_candidate_generation twin writes)
         DIR/splits/train.parquet + val.parquet (validation) | DIR/test.pkl or splits/test.parquet (submission)
  writes DIR/feature_engineering/{train,test}_{event}_interaction_features.pkl  (reference :113-118)
  any other mode raises ValueError('Invalid mode').

Columns appended to the candidate frame (reference names, :55-111): session_candidate_occurrence_count,
session_candidate_cumcount_last, session_candidate_{click,cart,order}_occurrence_count, ten session_candidate_*
aggregates and nine aid_* aggregates.  All counting and reducing runs in otto_interaction_features (csrc/features.cu).
Stated differences: rows keep the input order inside a session (polars' unique() / sort('session') leave it
unspecified); the script's nulls appear as cumcount_last = 0 and NaN means (oracle/interaction_oracle.py header).
"""
from __future__ import annotations

import argparse
import ctypes as C
import logging
import pathlib

import numpy as np
import torch

from . import _native as N
from . import covisit, io
from .covisit import EventCSR, _require_cuda, _stream_ptr

# short device column name -> the script's column name
COLUMN_NAMES = {
    "occurrence_count": "session_candidate_occurrence_count",
    "cumcount_last": "session_candidate_cumcount_last",
    "click_occurrence_count": "session_candidate_click_occurrence_count",
    "cart_occurrence_count": "session_candidate_cart_occurrence_count",
    "order_occurrence_count": "session_candidate_order_occurrence_count",
    "session_score_mean": "session_candidate_score_mean", "session_score_std": "session_candidate_score_std",
    "session_score_min": "session_candidate_score_min", "session_score_max": "session_candidate_score_max",
    "session_occurrence_count_mean": "session_candidate_occurrence_count_mean",
    "session_occurrence_count_sum": "session_candidate_occurrence_count_sum",
    "session_occurrence_count_max": "session_candidate_occurrence_count_max",
    "session_cumcount_last_mean": "session_candidate_cumcount_last_mean",
    "session_cumcount_last_sum": "session_candidate_cumcount_last_sum",
    "session_cumcount_last_max": "session_candidate_cumcount_last_max",
    "aid_score_mean": "aid_candidate_score_mean", "aid_score_std": "aid_candidate_score_std",
    "aid_score_max": "aid_candidate_score_max",
    "aid_occurrence_count_mean": "aid_session_candidate_occurrence_count_mean",
    "aid_occurrence_count_sum": "aid_session_candidate_occurrence_count_sum",
    "aid_occurrence_count_max": "aid_session_candidate_occurrence_count_max",
    "aid_cumcount_last_mean": "aid_session_candidate_cumcount_last_mean",
    "aid_cumcount_last_sum": "aid_session_candidate_cumcount_last_sum",
    "aid_cumcount_last_max": "aid_session_candidate_cumcount_last_max",
}
_TORCH = {"uint16": torch.int16, "uint32": torch.int32, "float32": torch.float32}      # storage dtypes (same width)
_NUMPY = {"uint16": np.uint16, "uint32": np.uint32, "float32": np.float32}


def interaction_features_device(sessions: EventCSR, session: torch.Tensor, candidates: torch.Tensor, scores: torch.Tensor) -> dict:
    """Device columns (session int32, candidates uint64 in int64 storage, candidate_scores float32; rows sorted by
    session) -> {short column name: device tensor [n_rows]} (unsigned columns in same-width signed storage)."""
    lib = N.lib()
    if sessions.order != "asc":
        raise ValueError("interaction features need the file-order CSR (ingest(..., order='asc'))")
    for name, t in (("session", session), ("candidates", candidates), ("candidate_scores", scores)):
        _require_cuda(t, name)
    dev = sessions.aid.device
    R = int(session.numel())
    session = session.to(torch.int32).contiguous()
    candidates = candidates.to(torch.int64).contiguous()
    scores = scores.to(torch.float32).contiguous()
    cols = {name: torch.empty(R, dtype=_TORCH[dt], device=dev) for name, dt in N.INTERACTION_COLUMNS}
    out = N.OttoInteractionFeatures(*[cols[name].data_ptr() for name, _ in N.INTERACTION_COLUMNS])
    frame = N.OttoCandidateFrame(R, session.data_ptr(), candidates.data_ptr(), scores.data_ptr())
    ss = N.OttoSessions(sessions.n_sessions, sessions.n_events, sessions.offsets.data_ptr(), sessions.aid.data_ptr(), sessions.type.data_ptr())
    need = int(lib.otto_interaction_scratch_bytes(sessions.n_sessions, R, sessions.n_aids))
    scratch = torch.empty(need, dtype=torch.uint8, device=dev)
    sid = sessions.session_ids.to(torch.int32).contiguous()
    with torch.cuda.device(dev):
        N.check(lib.otto_interaction_features(C.byref(ss), sid.data_ptr(), C.byref(frame), sessions.n_aids, C.byref(out),
                                              scratch.data_ptr(), need, _stream_ptr(dev)))
        torch.cuda.current_stream(dev).synchronize()
    return cols


def interaction_features(sessions: EventCSR, df_candidate):
    """ranker/interaction_feature_engineering.py:31-113 for one event type: candidate frame (pandas: session, candidates,
    candidate_scores [, candidate_labels]) -> the same frame (duplicates dropped, sorted by session, stable) with the
    24 feature columns under the script's names."""
    import pandas as pd
    dev = sessions.aid.device
    cand = df_candidate.drop_duplicates().reset_index(drop=True)
    cand = cand.sort_values("session", kind="stable").reset_index(drop=True)
    cols = interaction_features_device(sessions,
                                       torch.from_numpy(cand["session"].to_numpy().astype(np.int32)).to(dev),
                                       torch.from_numpy(cand["candidates"].to_numpy().astype(np.int64)).to(dev),
                                       torch.from_numpy(cand["candidate_scores"].to_numpy().astype(np.float32)).to(dev))
    out = cand.assign(session=cand["session"].astype(np.int32), candidates=cand["candidates"].astype(np.int32))      # :33
    for name, dt in N.INTERACTION_COLUMNS:
        out[COLUMN_NAMES[name]] = cols[name].cpu().numpy().view(_NUMPY[dt])
    return out


def main(argv=None) -> dict:
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", type=str)
    ap.add_argument("--data", type=pathlib.Path, required=True)
    ap.add_argument("--n-aids", type=int, default=None)
    ap.add_argument("--device", default="cuda:0")
    args = ap.parse_args(argv)
    if args.mode not in ("validation", "submission"):
        raise ValueError("Invalid mode")
    import pandas as pd
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(message)s")
    data, dev = args.data, torch.device(args.device)
    tag, prefix = ("validation", "train") if args.mode == "validation" else ("test", "test")
    if args.mode == "validation":          # :36-40: train ∪ val events; only the candidate sessions matter (:47)
        frame = io.read_event_frame(data / "splits" / "train.parquet", data / "splits" / "val.parquet", n_aids=args.n_aids)
    else:
        from .inference import _first_existing
        frame = io.read_event_frame(_first_existing(data / "test.pkl", data / "splits" / "test.parquet"), n_aids=args.n_aids)
    sessions = covisit.ingest(frame, "asc", device=dev)
    out_dir = data / "feature_engineering"
    out_dir.mkdir(parents=True, exist_ok=True)
    result = {"paths": []}
    for event in ("click", "cart", "order"):
        logging.info(f"Running {event} interaction feature engineering in {args.mode} mode")
        df_candidate = pd.read_pickle(data / "candidate" / f"{event}_{tag}.pkl")
        feats = interaction_features(sessions, df_candidate)
        path = out_dir / f"{prefix}_{event}_interaction_features.pkl"
        feats.to_pickle(path)
        logging.info(f"Saved {path.name} to {out_dir}: {feats.shape}")
        result["paths"].append(path)
        result[event] = feats
    return result


if __name__ == "__main__":
    main()
