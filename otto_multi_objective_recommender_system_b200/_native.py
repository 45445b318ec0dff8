"""ctypes binding of libotto_covisit.so (include/otto_covisit.h).

There is no CPU fallback: if the CUDA library is missing, loading raises and every product entry point
fails loudly.  Build it with `python -c "import __graft_entry__ as g; g.build()"` (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes as C
import os
import pathlib

_HERE = pathlib.Path(__file__).resolve().parent
LIB_PATH = _HERE / "libotto_covisit.so"

OTTO_OK = 0
OTTO_EINVAL, OTTO_ENOSPC, OTTO_ECUDA, OTTO_EOVERFLOW, OTTO_EUNSORTED = -22, -28, -5, -75, -71
WEIGHT_UNIT, WEIGHT_TYPE, WEIGHT_TIME = 0, 1, 2
MAX_TAIL, MAX_K, MAX_SEGMENTS, MAX_TABLES, MAX_SOURCES, MAX_TARGETS = 32, 32, 8, 8, 8, 4
HIST_RECENCY, HIST_TYPE_LE1, HIST_TYPE_GE1, HIST_TYPE_EQ0 = 0, 1, 2, 3

vp = C.c_void_p
i32, i64, u32 = C.c_int32, C.c_int64, C.c_uint32


class OttoEvents(C.Structure):
    _fields_ = [("n_sessions", i64), ("n_events", i64), ("session_offsets", vp), ("aid", vp), ("ts", vp),
                ("type", vp)]


class OttoCovisitSpec(C.Structure):
    _fields_ = [("n_aids", i32), ("weight_mode", i32), ("type_weight", i32 * 3), ("event_type_mask", u32),
                ("x_type_mask", u32), ("y_type_mask", u32), ("window_s", i32), ("tail_n", i32), ("k", i32),
                ("ts_min", i32), ("ts_max", i32), ("split_ub", i32), ("global_events", i64)]


class OttoBuildSizes(C.Structure):
    _fields_ = [("tail_capacity", i64), ("max_bins", i64), ("workspace_bytes", i64)]


class OttoBuildStats(C.Structure):
    _fields_ = [("tail_events", i64), ("pairs", i64), ("bins", i64), ("split_rows", i64), ("distinct", i64),
                ("pair_checksum", i64), ("table_overflow", i64), ("tier_records", i64 * 4), ("hot_pairs", i64),
                ("owner_records_max", i64), ("owner_bin_cuts", i64 * 9)]

    def as_dict(self) -> dict:
        d = {n: int(getattr(self, n)) for n, _ in self._fields_ if n not in ("tier_records", "owner_bin_cuts")}
        d["tier_records"] = [int(x) for x in self.tier_records]
        d["owner_bin_cuts"] = [int(x) for x in self.owner_bin_cuts]
        return d


class OttoPairSegment(C.Structure):
    _fields_ = [("records", vp), ("offsets", vp)]


MAX_OWNERS = 8


class OttoOwnerPlan(C.Structure):
    _fields_ = [("n_owners", i32), ("rank", i32), ("aid_cuts", i32 * (MAX_OWNERS + 1)), ("owner_records", vp * MAX_OWNERS)]


class OttoTopK(C.Structure):
    _fields_ = [("n_aids", i32), ("k", i32), ("aid_y", vp), ("wgt", vp), ("len", vp), ("cnt", vp), ("tsum", vp)]


class OttoSessions(C.Structure):
    _fields_ = [("n_sessions", i64), ("n_events", i64), ("session_offsets", vp), ("aid", vp), ("type", vp)]


class OttoCandidateSpec(C.Structure):
    _fields_ = [("n_tables", i32), ("table_aid_y", vp * MAX_TABLES), ("table_len", vp * MAX_TABLES),
                ("table_k", i32 * MAX_TABLES), ("n_aids", i32), ("n_sources", i32),
                ("source_table", i32 * MAX_SOURCES), ("source_hist", i32 * MAX_SOURCES), ("n_targets", i32),
                ("target_n_sources", i32 * MAX_TARGETS), ("target_sources", (i32 * MAX_SOURCES) * MAX_TARGETS),
                ("top_n", i32), ("drop_history", i32)]


class OttoRecencySpec(C.Structure):
    _fields_ = [("n_aids", i32), ("n", i32), ("table_aid_y", vp * 3), ("table_len", vp * 3), ("table_k", i32 * 3),
                ("hist", i32 * 3), ("bonus", C.c_double * 3), ("type_coefficient", C.c_double * 3), ("w_click", vp),
                ("w_cart", vp), ("w_offset", vp)]


class OttoLabels(C.Structure):
    _fields_ = [("offsets", vp), ("aid", vp)]


class OttoCandidateFrame(C.Structure):
    _fields_ = [("n_rows", i64), ("session", vp), ("candidates", vp), ("candidate_scores", vp)]


INTERACTION_COLUMNS = (
    ("occurrence_count", "uint16"), ("cumcount_last", "uint16"), ("click_occurrence_count", "uint16"),
    ("cart_occurrence_count", "uint16"), ("order_occurrence_count", "uint16"),
    ("session_score_mean", "float32"), ("session_score_std", "float32"), ("session_score_min", "float32"),
    ("session_score_max", "float32"), ("session_occurrence_count_mean", "float32"), ("session_occurrence_count_sum", "uint32"),
    ("session_occurrence_count_max", "uint16"), ("session_cumcount_last_mean", "float32"), ("session_cumcount_last_sum", "uint32"),
    ("session_cumcount_last_max", "uint16"),
    ("aid_score_mean", "float32"), ("aid_score_std", "float32"), ("aid_score_max", "float32"),
    ("aid_occurrence_count_mean", "float32"), ("aid_occurrence_count_sum", "uint32"), ("aid_occurrence_count_max", "uint16"),
    ("aid_cumcount_last_mean", "float32"), ("aid_cumcount_last_sum", "uint32"), ("aid_cumcount_last_max", "uint16"))


class OttoInteractionFeatures(C.Structure):
    _fields_ = [(name, vp) for name, _ in INTERACTION_COLUMNS]


class OttoCandidates(C.Structure):
    _fields_ = [("aid", vp), ("score", vp), ("len", vp)]


P = C.POINTER
_SIGNATURES = {
    "otto_last_error": (C.c_char_p, []),
    "otto_version": (C.c_int, []),
    "otto_launch_count": (C.c_uint64, []),
    "otto_profile_enable": (C.c_int, [C.c_int]),
    "otto_profile_reduce_ms": (C.c_int, [P(C.c_float)]),
    "otto_profile_scatter_ms": (C.c_int, [P(C.c_float)]),
    "otto_frame_is_sorted": (C.c_int, [vp, vp, i64, vp, P(i32), vp]),
    "otto_frame_check": (C.c_int, [vp, vp, i64, i32, vp, P(i64), vp]),
    "otto_ingest_scratch_bytes": (i64, [i64]),
    "otto_ingest_scan": (C.c_int, [vp, vp, vp, vp, i64, i32, vp, i64, P(i64), vp]),
    "otto_ingest_offsets": (C.c_int, [vp, i64, i64, vp, i64, vp, vp, vp, vp]),
    "otto_ingest_desc": (C.c_int, [vp, i64, vp, vp, vp, i64, vp, vp, vp, vp]),
    "otto_covisit_sizes": (C.c_int, [i64, i64, P(OttoCovisitSpec), P(OttoBuildSizes)]),
    "otto_covisit_count_begin": (C.c_int, [P(OttoEvents), P(OttoCovisitSpec), vp, i64, vp]),
    "otto_covisit_count_begin_asc": (C.c_int, [P(OttoEvents), P(OttoCovisitSpec), vp, i64, vp]),
    "otto_covisit_count_finish": (C.c_int, [P(OttoEvents), P(OttoCovisitSpec), vp, i64, P(OttoBuildStats), vp]),
    "otto_covisit_count": (C.c_int, [P(OttoEvents), P(OttoCovisitSpec), vp, i64, P(OttoBuildStats), vp]),
    "otto_covisit_views": (C.c_int, [P(OttoEvents), P(OttoCovisitSpec), vp, i64, P(vp), P(vp), P(vp), P(vp)]),
    "otto_covisit_scatter": (C.c_int, [P(OttoEvents), P(OttoCovisitSpec), vp, i64, vp, i64, vp]),
    "otto_covisit_count_finish_owned": (C.c_int, [P(OttoEvents), P(OttoCovisitSpec), vp, i64, P(OttoOwnerPlan), vp,
                                                  P(OttoBuildStats), vp]),
    "otto_covisit_plan_scratch_bytes": (i64, [i32]),
    "otto_covisit_plan_owners": (C.c_int, [vp, i32, i32, i32, vp, vp, vp, i64, P(i32), vp]),
    "otto_covisit_scatter_owned": (C.c_int, [P(OttoEvents), P(OttoCovisitSpec), vp, i64, P(OttoOwnerPlan), vp]),
    "otto_covisit_stage_plan_bytes": (i64, [P(OttoCovisitSpec), i64, i64, i32]),
    "otto_covisit_stage_plan": (C.c_int, [P(OttoEvents), P(OttoCovisitSpec), vp, i64, vp, i32, i32, vp, i64, P(i64), vp]),
    "otto_covisit_stage_totals": (C.c_int, [P(OttoCovisitSpec), i64, i64, vp, i32, P(i64), vp]),
    "otto_covisit_scatter_staged": (C.c_int, [P(OttoEvents), P(OttoCovisitSpec), vp, i64, vp, i32, vp, vp]),
    "otto_covisit_place_staged": (C.c_int, [P(OttoEvents), P(OttoCovisitSpec), vp, i64, vp, i32, P(vp), i32, i32, vp, i64, vp]),
    "otto_covisit_partition": (C.c_int, [P(OttoEvents), P(OttoCovisitSpec), vp, i64, vp, i64, vp]),
    "otto_covisit_reduce_scratch_bytes": (i64, [P(OttoCovisitSpec), i64, i64]),
    "otto_covisit_reduce": (C.c_int, [P(OttoCovisitSpec), vp, vp, i64, i64, i32, i32, P(OttoPairSegment), i32, vp,
                                      i64, P(OttoTopK), P(OttoBuildStats), vp]),
    "otto_covisit_merge_scratch_bytes": (i64, [i64]),
    "otto_covisit_merge_segments": (C.c_int, [P(OttoPairSegment), i32, i64, vp, i64, vp, vp, i64, P(i64), vp]),
    "otto_peer_alloc": (C.c_int, [i64, P(vp)]),
    "otto_peer_free": (C.c_int, [vp]),
    "otto_peer_get_handle": (C.c_int, [vp, C.c_char_p]),
    "otto_peer_open": (C.c_int, [C.c_char_p, P(vp)]),
    "otto_peer_close": (C.c_int, [vp]),
    "otto_covisit_build": (C.c_int, [P(OttoEvents), P(OttoCovisitSpec), vp, i64, P(OttoTopK), P(OttoBuildStats), vp]),
    "otto_covisit_build_bytes": (i64, [i64, i64, P(OttoCovisitSpec), i64, i64]),
    "otto_topk_row_offsets": (C.c_int, [P(OttoTopK), vp, P(i64), vp, i64, vp]),
    "otto_topk_to_rows": (C.c_int, [P(OttoTopK), vp, vp, vp, vp, vp]),
    "otto_rows_to_topk": (C.c_int, [vp, vp, vp, i64, P(OttoTopK), vp]),
    "otto_candidates_scratch_bytes": (i64, [i64, i32, P(OttoCandidateSpec)]),
    "otto_candidates": (C.c_int, [P(OttoSessions), i32, P(OttoCandidateSpec), vp, i64, P(OttoCandidates), vp]),
    "otto_row_offsets_scratch_bytes": (i64, [i64]),
    "otto_row_offsets": (C.c_int, [vp, i64, vp, P(i64), vp, i64, vp]),
    "otto_explode_candidates": (C.c_int, [vp, vp, vp, i64, i32, vp, vp, P(OttoLabels), vp, vp, vp, vp, vp]),
    "otto_recall_counts": (C.c_int, [vp, i64, i32, P(OttoLabels), i32, vp, vp]),
    "otto_regular_row_counts": (C.c_int, [P(OttoSessions), vp, vp, vp, i32, vp, vp]),
    "otto_regular_rows": (C.c_int, [P(OttoSessions), vp, vp, vp, i32, vp, vp, P(OttoLabels), vp, vp, vp, vp, vp]),
    "otto_interaction_scratch_bytes": (i64, [i64, i64, i32]),
    "otto_interaction_features": (C.c_int, [P(OttoSessions), vp, P(OttoCandidateFrame), i32, P(OttoInteractionFeatures), vp, i64, vp]),
    "otto_recency_scratch_bytes": (i64, [i32, i32]),
    "otto_recency_long": (C.c_int, [P(OttoSessions), vp, i32, i32, P(OttoRecencySpec), vp, i64, vp, vp]),
    "otto_recency_scored": (C.c_int, [P(OttoSessions), vp, i32, i32, P(OttoRecencySpec), vp, i64, i32, vp, vp, vp, vp]),
    "otto_assemble_predictions": (C.c_int, [P(OttoSessions), P(OttoCandidates), i32, i32, vp, i32, i32, vp, vp, vp]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


class OttoError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libotto_covisit error {code}: {message}")
        self.code = code


def lib() -> C.CDLL:
    """The loaded library; raises if it was never built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: the CUDA library has not been built. "
                "Run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc); there is no CPU fallback.")
        l = C.CDLL(os.fspath(LIB_PATH))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(code: int) -> None:
    if code != OTTO_OK:
        raise OttoError(code, lib().otto_last_error().decode("utf-8", "replace"))
