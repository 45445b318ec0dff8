"""Covisitation candidate generation on one B200: host side of otto_candidates / otto_assemble_predictions.

Mirrors the reference consumers of the top-K tables:
  * src/ranker/covisitation_candidate_generation.py:108-157 / :248-288 - ranker form, most_common(100),
    output frames (session, candidates uint64, candidate_scores float32[, candidate_labels uint8])
    written to candidate/{click,cart,order}_covisitation_{validation,test}.pkl (:177-197, :290-307)
  * src/covisitation/inference.py:204-247 / :396-441 - standalone form, most_common(20) + history +
    popular fill (the fastText/Annoy neighbour term is out of scope: SURVEY.md §2)
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _native as N
from .covisit import EventCSR, TopKTable, _require_cuda, _stream_ptr

STEMS = ("time_weighted", "click_weighted", "cart_weighted", "order_weighted", "click_cart", "click_order", "cart_order")


@dataclass(frozen=True)
class CandidateSpec:
    """sources: (table stem, history set); targets: per event type the ordered sources to concatenate."""
    sources: tuple
    targets: dict
    top_n: int = 100
    drop_history: bool = True


def reference_spec(stems_available, top_n: int = 100) -> CandidateSpec:
    """The reference's list recipe (ranker/covisitation_candidate_generation.py:119-138) restricted to the
    tables that exist: clicks = time + click_w + cart_w + click_cart + cart_order, carts = orders =
    time + cart_w + cart_order; time over the history in recency order, everything else over the sorted
    unique aids with type <= 1 (note: cart_order too, :124).  Absent stems contribute nothing, like the
    reference's `if aid in table` guards."""
    have = set(stems_available)
    recipe = {"time_weighted": N.HIST_RECENCY, "click_weighted": N.HIST_TYPE_LE1, "cart_weighted": N.HIST_TYPE_LE1,
              "click_cart": N.HIST_TYPE_LE1, "cart_order": N.HIST_TYPE_LE1}
    order = {"click": ["time_weighted", "click_weighted", "cart_weighted", "click_cart", "cart_order"],
             "cart": ["time_weighted", "cart_weighted", "cart_order"],
             "order": ["time_weighted", "cart_weighted", "cart_order"]}
    sources = tuple((s, recipe[s]) for s in recipe if s in have)
    index = {s: i for i, (s, _) in enumerate(sources)}
    targets = {t: tuple(index[s] for s in lst if s in index) for t, lst in order.items()}
    return CandidateSpec(sources, targets, top_n, True)


@dataclass
class Candidates:
    """Fixed-stride candidate lists on the device: [target, session, rank]."""
    targets: tuple
    aid: torch.Tensor     # int32 [T, S, N], -1 padded
    score: torch.Tensor   # int32 [T, S, N]
    len: torch.Tensor     # int32 [T, S]
    session_ids: torch.Tensor

    def to_frames(self, labels: dict | None = None) -> dict:
        """The exploded frames the ranker script pickles (:177-197): session, candidates uint64,
        candidate_scores float32 (+ candidate_labels uint8 when labels = {target: {session: set}})."""
        import pandas as pd
        out = {}
        n = self.aid.shape[2]
        sid = self.session_ids.cpu().numpy()
        for ti, t in enumerate(self.targets):
            ln = self.len[ti].cpu().numpy()
            mask = np.arange(n)[None, :] < ln[:, None]
            f = pd.DataFrame({"session": np.repeat(sid, ln),
                              "candidates": self.aid[ti].cpu().numpy()[mask].astype(np.uint64),
                              "candidate_scores": self.score[ti].cpu().numpy()[mask].astype(np.float32)})
            if labels is not None:
                lab = labels.get(t, {})
                f["candidate_labels"] = np.fromiter(
                    (int(int(a) in lab.get(int(s), ())) for s, a in zip(f["session"], f["candidates"])),
                    dtype=np.uint8, count=len(f))
            out[t] = f
        return out


def _sessions_struct(csr: EventCSR) -> N.OttoSessions:
    return N.OttoSessions(csr.n_sessions, csr.n_events, csr.offsets.data_ptr(), csr.aid.data_ptr(), csr.type.data_ptr())


def max_session_len(csr: EventCSR) -> int:
    if csr.n_sessions == 0:
        return 1
    return int((csr.offsets[1:] - csr.offsets[:-1]).max().item())


class CandidateGenerator:
    """Keeps the scratch and output buffers across calls."""

    def __init__(self, tables: dict, spec: CandidateSpec, n_aids: int):
        self.lib = N.lib()
        self.tables, self.spec, self.n_aids = tables, spec, n_aids
        stems = sorted({s for s, _ in spec.sources})
        self.stems = stems
        for s in stems:
            _require_cuda(tables[s].aid_y, f"table {s}")
        if len(stems) > N.MAX_TABLES or len(spec.sources) > N.MAX_SOURCES or len(spec.targets) > N.MAX_TARGETS:
            raise ValueError("too many tables / sources / targets")
        cs = N.OttoCandidateSpec()
        cs.n_tables = len(stems)
        for i, s in enumerate(stems):
            cs.table_aid_y[i] = tables[s].aid_y.data_ptr()
            cs.table_len[i] = tables[s].len.data_ptr()
            cs.table_k[i] = tables[s].k
        cs.n_aids = n_aids
        cs.n_sources = len(spec.sources)
        for i, (s, h) in enumerate(spec.sources):
            cs.source_table[i] = stems.index(s)
            cs.source_hist[i] = h
        self.target_names = tuple(spec.targets)
        cs.n_targets = len(self.target_names)
        for ti, t in enumerate(self.target_names):
            cs.target_n_sources[ti] = len(spec.targets[t])
            for j, src in enumerate(spec.targets[t]):
                cs.target_sources[ti][j] = src
        cs.top_n = spec.top_n
        cs.drop_history = 1 if spec.drop_history else 0
        self.cspec = cs
        self.scratch = None
        self.out = None

    def __call__(self, sessions: EventCSR, max_len: int | None = None) -> Candidates:
        if sessions.order != "asc":
            raise ValueError("candidate generation needs sessions in file order (ingest(..., order='asc'))")
        _require_cuda(sessions.aid, "sessions")
        dev = sessions.aid.device
        S, T, n = sessions.n_sessions, len(self.target_names), self.spec.top_n
        max_len = max_session_len(sessions) if max_len is None else max_len
        need = int(self.lib.otto_candidates_scratch_bytes(S, max_len, C.byref(self.cspec)))
        if need < 0:
            N.check(N.OTTO_EINVAL)
        if self.scratch is None or self.scratch.numel() < need:
            self.scratch = torch.empty(need, dtype=torch.uint8, device=dev)
        if self.out is None or self.out[0].shape != (T, S, n):
            self.out = (torch.empty((T, S, n), dtype=torch.int32, device=dev),
                        torch.empty((T, S, n), dtype=torch.int32, device=dev),
                        torch.empty((T, S), dtype=torch.int32, device=dev))
        aid, score, ln = self.out
        oc = N.OttoCandidates(aid.data_ptr(), score.data_ptr(), ln.data_ptr())
        ss = _sessions_struct(sessions)
        with torch.cuda.device(dev):
            N.check(self.lib.otto_candidates(C.byref(ss), max_len, C.byref(self.cspec), self.scratch.data_ptr(),
                                             self.scratch.numel(), C.byref(oc), _stream_ptr(dev)))
        return Candidates(self.target_names, aid, score, ln, sessions.session_ids)


def generate_candidates(sessions: EventCSR, tables: dict, spec: CandidateSpec | None = None) -> Candidates:
    spec = reference_spec(tables.keys()) if spec is None else spec
    return CandidateGenerator(tables, spec, sessions.n_aids)(sessions)


def assemble_predictions(sessions: EventCSR, cand: Candidates, popular: dict, n: int = 20):
    """covisitation/inference.py:238-243 -> (pred int32 [T, S, n] with -1 padding, long_session bool [S]).
    popular = {target: most frequent aids} (data/aid_frequencies/*_20_most_frequent_*_aids.json, :76-83)."""
    lib = N.lib()
    dev = sessions.aid.device
    T, S = len(cand.targets), sessions.n_sessions
    n_pop = max(len(popular[t]) for t in cand.targets)
    pop = torch.full((T, n_pop), -1, dtype=torch.int32)
    for ti, t in enumerate(cand.targets):
        pop[ti, :len(popular[t])] = torch.tensor(list(popular[t]), dtype=torch.int32)
    pop = pop.to(dev)
    pred = torch.empty((T, S, n), dtype=torch.int32, device=dev)
    long_session = torch.zeros(S, dtype=torch.uint8, device=dev)
    ss = _sessions_struct(sessions)
    oc = N.OttoCandidates(cand.aid.data_ptr(), cand.score.data_ptr(), cand.len.data_ptr())
    with torch.cuda.device(dev):
        N.check(lib.otto_assemble_predictions(C.byref(ss), C.byref(oc), T, cand.aid.shape[2], pop.data_ptr(), n_pop, n,
                                              pred.data_ptr(), long_session.data_ptr(), _stream_ptr(dev)))
    return pred, long_session.bool()


def regular_candidates(sessions: EventCSR, tables: dict, n: int = 100, labels: dict | None = None, n_chunks: int = 15) -> dict:
    """ranker/regular_candidate_generation.py:139-180,225-257: per session its unique aids (most recent first, scores
    |H| .. 1, :163) followed by the ranker-form votes (most_common(n), history dropped) -> the exploded frames the
    script pickles as candidate/{event}_{validation,test}.pkl.  Like the script, sessions are processed in chunks
    (:218: 15) so the dense [target, session, |H| + n] device rows stay small.  The fastText / Annoy term (:155-156)
    is not on this path."""
    import pandas as pd
    lib = N.lib()
    dev = sessions.aid.device
    cand = generate_candidates(sessions, tables, reference_spec(tables.keys(), n))
    T, S = len(cand.targets), sessions.n_sessions
    sid_all = sessions.session_ids.cpu().numpy()
    parts = {t: [] for t in cand.targets}
    dummy = torch.zeros(1, dtype=torch.int32, device=dev)
    n_chunks = max(1, min(n_chunks, S))
    for c in range(n_chunks):
        lo, hi = c * S // n_chunks, (c + 1) * S // n_chunks
        if hi <= lo:
            continue
        sub = sessions.slice_sessions(lo, hi)
        W = max_session_len(sub) + n
        c_aid, c_score, c_len = (x[:, lo:hi].contiguous() for x in (cand.aid, cand.score, cand.len))
        pred = torch.empty((T, hi - lo, W), dtype=torch.int32, device=dev)
        ss = _sessions_struct(sub)
        oc = N.OttoCandidates(c_aid.data_ptr(), c_score.data_ptr(), c_len.data_ptr())
        with torch.cuda.device(dev):
            # no popular fill, rows wide enough for the whole history: row = history + votes
            N.check(lib.otto_assemble_predictions(C.byref(ss), C.byref(oc), T, n, dummy.data_ptr(), 0, W, pred.data_ptr(), None,
                                                  _stream_ptr(dev)))
        row_len = (pred >= 0).sum(dim=2)
        hist_len = row_len - c_len
        j = torch.arange(W, device=dev)[None, None, :]
        vote = torch.gather(c_score, 2, (j - hist_len[:, :, None]).clamp_(0, n - 1).expand(T, hi - lo, W))
        score = torch.where(j < hist_len[:, :, None], hist_len[:, :, None] - j, vote)
        mask = j < row_len[:, :, None]
        for ti, t in enumerate(cand.targets):
            m = mask[ti]
            parts[t].append((np.repeat(sid_all[lo:hi], row_len[ti].cpu().numpy()), pred[ti][m].cpu().numpy(), score[ti][m].cpu().numpy()))
    out = {}
    for t, chunks in parts.items():
        f = pd.DataFrame({"session": np.concatenate([c[0] for c in chunks]) if chunks else np.zeros(0, sid_all.dtype),
                          "candidates": (np.concatenate([c[1] for c in chunks]) if chunks else np.zeros(0, np.int32)).astype(np.uint64),
                          "candidate_scores": (np.concatenate([c[2] for c in chunks]) if chunks else np.zeros(0, np.int32)).astype(np.float32)})
        if labels is not None:
            lab = labels.get(t, {})
            f["candidate_labels"] = np.fromiter((int(int(a) in lab.get(int(s), ())) for s, a in zip(f["session"], f["candidates"])),
                                                dtype=np.uint8, count=len(f))
        out[t] = f
    return out


_WEIGHT_CACHE: dict = {}


def recency_weights(max_len: int):
    """np.logspace(0.1 | 0.5, 1, L, base=2) - 1 for every session length L <= max_len, concatenated
    (covisitation/inference.py:152-154); made with numpy on the host so the fp64 values are the reference's."""
    if max_len in _WEIGHT_CACHE:
        return _WEIGHT_CACHE[max_len]
    offs = np.zeros(max_len + 2, dtype=np.int64)
    offs[1:] = np.cumsum(np.arange(0, max_len + 1))
    wc = np.zeros(int(offs[-1]) + max_len + 1, dtype=np.float64)
    wk = np.zeros_like(wc)
    for L in range(1, max_len + 1):
        wc[offs[L]:offs[L] + L] = np.logspace(0.1, 1, L, base=2, endpoint=True) - 1
        wk[offs[L]:offs[L] + L] = np.logspace(0.5, 1, L, base=2, endpoint=True) - 1
    _WEIGHT_CACHE[max_len] = (wc, wk, offs[:max_len + 1])
    return _WEIGHT_CACHE[max_len]


def recency_long_predictions(sessions: EventCSR, tables: dict, pred: torch.Tensor, long_session: torch.Tensor,
                             n: int = 20) -> torch.Tensor:
    """covisitation/inference.py:142-199: overwrites the rows of `pred` [3, S, n] (clicks, carts, orders) that
    belong to long sessions (>= n unique aids) with the recency-weighted ranking."""
    lib = N.lib()
    dev = sessions.aid.device
    idx = torch.nonzero(long_session).flatten().to(torch.int32)
    if idx.numel() == 0:
        return pred
    if pred.shape[0] != 3 or pred.shape[2] != n:
        raise ValueError("pred must be [3, sessions, n] for the targets click, cart, order")
    max_len = max_session_len(sessions)
    dkey = (max_len, str(dev))
    if dkey not in _WEIGHT_CACHE:
        _WEIGHT_CACHE[dkey] = tuple(torch.from_numpy(a).to(dev) for a in recency_weights(max_len))
    wc_d, wk_d, off_d = _WEIGHT_CACHE[dkey]
    spec = N.OttoRecencySpec()
    spec.n_aids, spec.n = sessions.n_aids, n
    max_k = 1
    for t, stem in enumerate(("time_weighted", "cart_weighted", "cart_order")):
        tb = tables.get(stem)
        if tb is not None:
            _require_cuda(tb.aid_y, f"table {stem}")
            spec.table_aid_y[t], spec.table_len[t], spec.table_k[t] = tb.aid_y.data_ptr(), tb.len.data_ptr(), tb.k
            max_k = max(max_k, tb.k)
    for t, (hsel, bonus, coef) in enumerate(zip((N.HIST_TYPE_EQ0, N.HIST_TYPE_LE1, N.HIST_TYPE_GE1), (0.05, 0.05, 0.15), (1.0, 9.0, 6.0))):
        spec.hist[t], spec.bonus[t], spec.type_coefficient[t] = hsel, bonus, coef
    spec.w_click, spec.w_cart, spec.w_offset = wc_d.data_ptr(), wk_d.data_ptr(), off_d.data_ptr()
    need = int(lib.otto_recency_scratch_bytes(max_len, max_k))
    scratch = torch.empty(need, dtype=torch.uint8, device=dev)
    ss = _sessions_struct(sessions)
    with torch.cuda.device(dev):
        N.check(lib.otto_recency_long(C.byref(ss), idx.data_ptr(), idx.numel(), max_len, C.byref(spec), scratch.data_ptr(), need,
                                      pred.data_ptr(), _stream_ptr(dev)))
        torch.cuda.current_stream(dev).synchronize()      # wc_d / wk_d / scratch must outlive the kernel
    return pred


def recency_weighted_candidates(sessions: EventCSR, labels: dict | None = None, keep_f64: bool = False) -> dict:
    """ranker/recency_weighted_candidate_generator.py:61-144 (validation) / :169-236 (test): every unique aid of a
    session ranked by its recency-weighted event score (type coefficients {0: 1, 1: 6, 2: 1}), for clicks, carts and
    orders -> the exploded frames the script pickles as {event}_recency_weighted_{validation,test}.pkl (columns
    session, candidates uint64, candidate_scores float32 [, candidate_labels uint8]).  The scores are the script's
    fp64 Counter values bit for bit before the float32 cast (otto_recency_scored).  Sessions are processed in two
    groups (<= 32 events, longer) so that the dense device outputs stay small."""
    import pandas as pd
    lib = N.lib()
    dev = sessions.aid.device
    _require_cuda(sessions.aid, "sessions")
    if sessions.order != "asc":
        raise ValueError("candidate generation needs the file-order CSR (ingest(..., order='asc'))")
    max_len = max_session_len(sessions)
    dkey = (max_len, str(dev))
    if dkey not in _WEIGHT_CACHE:
        _WEIGHT_CACHE[dkey] = tuple(torch.from_numpy(a).to(dev) for a in recency_weights(max_len))
    wc_d, wk_d, off_d = _WEIGHT_CACHE[dkey]
    lens = sessions.offsets[1:] - sessions.offsets[:-1]
    sid = sessions.session_ids.cpu().numpy()
    ss = _sessions_struct(sessions)
    need = int(lib.otto_recency_scratch_bytes(max_len, 1))
    scratch = torch.empty(need, dtype=torch.uint8, device=dev)
    parts = {t: [] for t in ("click", "cart", "order")}
    for lo, hi in ((0, 32), (32, max(max_len, 32))):
        idx = torch.nonzero((lens > lo) & (lens <= hi)).flatten().to(torch.int32)
        if idx.numel() == 0:
            continue
        n = hi
        spec = N.OttoRecencySpec()
        spec.n_aids, spec.n = sessions.n_aids, n
        for t, coef in enumerate((1.0, 6.0, 1.0)):                                          # :25
            spec.hist[t], spec.bonus[t], spec.type_coefficient[t] = N.HIST_TYPE_EQ0, 0.0, coef
        spec.w_click, spec.w_cart, spec.w_offset = wc_d.data_ptr(), wk_d.data_ptr(), off_d.data_ptr()
        aid = torch.empty((3, idx.numel(), n), dtype=torch.int32, device=dev)
        score = torch.empty((3, idx.numel(), n), dtype=torch.float64, device=dev)
        ln = torch.empty((3, idx.numel()), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            N.check(lib.otto_recency_scored(C.byref(ss), idx.data_ptr(), idx.numel(), max_len, C.byref(spec), scratch.data_ptr(),
                                            need, 1, aid.data_ptr(), score.data_ptr(), ln.data_ptr(), _stream_ptr(dev)))
            torch.cuda.current_stream(dev).synchronize()
        rows = idx.cpu().numpy()
        for ti, t in enumerate(parts):
            l = ln[ti].cpu().numpy()
            mask = np.arange(n)[None, :] < l[:, None]
            parts[t].append((np.repeat(rows, l), aid[ti].cpu().numpy()[mask], score[ti].cpu().numpy()[mask]))
    out = {}
    for t, chunks in parts.items():
        if chunks:
            row = np.concatenate([c[0] for c in chunks])
            a = np.concatenate([c[1] for c in chunks])
            sc = np.concatenate([c[2] for c in chunks])
            order = np.argsort(row, kind="stable")                 # back to session order; ranks stay in order
            row, a, sc = row[order], a[order], sc[order]
        else:
            row, a, sc = np.zeros(0, np.int64), np.zeros(0, np.int32), np.zeros(0, np.float64)
        f = pd.DataFrame({"session": sid[row], "candidates": a.astype(np.uint64), "candidate_scores": sc.astype(np.float32)})
        if keep_f64:
            f["candidate_scores_f64"] = sc
        if labels is not None:
            lab = labels.get(t, {})
            f["candidate_labels"] = np.fromiter((int(int(x) in lab.get(int(s), ())) for s, x in zip(f["session"], f["candidates"])),
                                                dtype=np.uint8, count=len(f))
        out[t] = f
    return out


def recall_at_20(pred: torch.Tensor, labels: list) -> float:
    """covisitation/inference.py:251-257 on device predictions: sum |pred ∩ label| / sum min(|label|, 20)."""
    p = pred.cpu().numpy()
    hits = sum(len(set(int(a) for a in row if a >= 0).intersection(l)) for row, l in zip(p, labels))
    denom = sum(min(len(l), 20) for l in labels)
    return hits / denom if denom else 0.0
