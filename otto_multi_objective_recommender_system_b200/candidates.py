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
class LabelCSR:
    """Ground truth of a frame's sessions on the device (OttoLabels): aids of session i (row order of the CSR's
    session_ids) at aid[offsets[i] : offsets[i + 1]], unique and ascending."""
    offsets: torch.Tensor   # int64 [S + 1]
    aid: torch.Tensor       # int32

    def to_c(self) -> N.OttoLabels:
        return N.OttoLabels(self.offsets.data_ptr(), self.aid.data_ptr())

    @staticmethod
    def build(session_ids, labels, device) -> "LabelCSR":
        """labels: {session id: iterable of aids} (sessions without an entry have no labels) or a sequence of
        iterables aligned with session_ids.  One pass over the labelled sessions on the host, then a lexsort."""
        sid = np.asarray(session_ids.cpu() if hasattr(session_ids, "cpu") else session_ids)
        S = int(sid.size)
        if isinstance(labels, dict):
            index = {int(x): i for i, x in enumerate(sid.tolist())}
            items = [(index[int(k)], v) for k, v in labels.items() if int(k) in index]
        else:
            items = list(enumerate(labels))
        rows = np.fromiter((i for i, v in items for _ in v), dtype=np.int64)
        aids = np.fromiter((int(a) for _, v in items for a in v), dtype=np.int64)
        if rows.size:
            key = np.unique(rows << 32 | aids)                     # unique + ascending (row, aid)
            rows, aids = key >> 32, key & 0xFFFFFFFF
        offsets = np.zeros(S + 1, dtype=np.int64)
        np.cumsum(np.bincount(rows, minlength=S), out=offsets[1:])
        return LabelCSR(torch.from_numpy(offsets).to(device), torch.from_numpy(aids.astype(np.int32)).to(device))


def _row_offsets(lib, length: torch.Tensor):
    """Exclusive scan of int32 lengths on the device -> (int64 offsets [n + 1], total)."""
    dev = length.device
    n = int(length.numel())
    off = torch.empty(n + 1, dtype=torch.int64, device=dev)
    need = int(lib.otto_row_offsets_scratch_bytes(n))
    scratch = torch.empty(need, dtype=torch.uint8, device=dev)
    total = C.c_int64(0)
    with torch.cuda.device(dev):
        N.check(lib.otto_row_offsets(length.data_ptr(), n, off.data_ptr(), C.byref(total), scratch.data_ptr(), need, _stream_ptr(dev)))
    return off, int(total.value)


def _frame_from_device(session, cand, score, label=None):
    """Flat device columns -> the pickled frame layout (one D2H copy per column, no per-row host work)."""
    import pandas as pd
    cols = {"session": session.cpu().numpy(), "candidates": cand.cpu().numpy().view(np.uint64),
            "candidate_scores": score.cpu().numpy()}
    if label is not None:
        cols["candidate_labels"] = label.cpu().numpy()
    return pd.DataFrame(cols)


@dataclass
class Candidates:
    """Fixed-stride candidate lists on the device: [target, session, rank]."""
    targets: tuple
    aid: torch.Tensor     # int32 [T, S, N], -1 padded
    score: torch.Tensor   # int32 [T, S, N]
    len: torch.Tensor     # int32 [T, S]
    session_ids: torch.Tensor

    def to_device_columns(self, ti: int, labels: "LabelCSR | None" = None):
        """One target exploded on the device (otto_explode_candidates): session int32, candidates uint64 (int64
        storage), candidate_scores float32 [, candidate_labels uint8]."""
        lib = N.lib()
        dev = self.aid.device
        S, n = int(self.aid.shape[1]), int(self.aid.shape[2])
        off, total = _row_offsets(lib, self.len[ti])
        session = torch.empty(total, dtype=torch.int32, device=dev)
        cand = torch.empty(total, dtype=torch.int64, device=dev)
        score = torch.empty(total, dtype=torch.float32, device=dev)
        label = torch.empty(total, dtype=torch.uint8, device=dev) if labels is not None else None
        lc = labels.to_c() if labels is not None else None
        sid = self.session_ids.to(torch.int32)
        with torch.cuda.device(dev):
            N.check(lib.otto_explode_candidates(self.aid[ti].data_ptr(), self.score[ti].data_ptr(), self.len[ti].data_ptr(), S, n,
                                                off.data_ptr(), sid.data_ptr(), C.byref(lc) if lc is not None else None,
                                                session.data_ptr(), cand.data_ptr(), score.data_ptr(),
                                                label.data_ptr() if label is not None else None, _stream_ptr(dev)))
        return session, cand, score, label

    def to_frames(self, labels: dict | None = None) -> dict:
        """The exploded frames the ranker script pickles (:177-197): session, candidates uint64,
        candidate_scores float32 (+ candidate_labels uint8 when labels = {target: {session: set}} or
        {target: LabelCSR}).  Explode, casts and label marking run on the device."""
        out = {}
        for ti, t in enumerate(self.targets):
            lab = None
            if labels is not None:
                lab = labels.get(t, {})
                if not isinstance(lab, LabelCSR):
                    lab = LabelCSR.build(self.session_ids, lab, self.aid.device)
            out[t] = _frame_from_device(*self.to_device_columns(ti, lab))
        return out


def _sessions_struct(csr: EventCSR) -> N.OttoSessions:
    return N.OttoSessions(csr.n_sessions, csr.n_events, csr.offsets.data_ptr(), csr.aid.data_ptr(), csr.type.data_ptr())


def max_session_len(csr: EventCSR) -> int:
    if csr.n_sessions == 0:
        return 1
    if getattr(csr, "max_len_dev", None) is not None:
        return max(1, int(csr.max_len_dev.item()))          # computed by otto_ingest_offsets
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


def regular_candidates(sessions: EventCSR, tables: dict, n: int = 100, labels: dict | None = None) -> dict:
    """ranker/regular_candidate_generation.py:139-180,225-257: per session its unique aids (most recent first, scores
    |H| .. 1, :163) followed by the ranker-form votes (most_common(n), history dropped) -> the exploded frames the
    script pickles as candidate/{event}_{validation,test}.pkl.  Rows, scores and labels are written by
    otto_regular_rows; the reference's 15 session chunks (:218) only bound its host memory and are not reproduced -
    including their side effect: `df_val.loc[start:end]` is label-inclusive, so the reference's concatenated pickle holds
    every chunk-boundary session twice (14 duplicated sessions); this frame holds each session once.
    The fastText / Annoy term (:155-156) is not on this path."""
    lib = N.lib()
    dev = sessions.aid.device
    cand = generate_candidates(sessions, tables, reference_spec(tables.keys(), n))
    S = sessions.n_sessions
    ss = _sessions_struct(sessions)
    sid = sessions.session_ids.to(torch.int32)
    out = {}
    for ti, t in enumerate(cand.targets):
        lab = None
        if labels is not None:
            lab = labels.get(t, {})
            if not isinstance(lab, LabelCSR):
                lab = LabelCSR.build(sessions.session_ids, lab, dev)
        rows = torch.empty(S, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            N.check(lib.otto_regular_row_counts(C.byref(ss), cand.aid[ti].data_ptr(), cand.score[ti].data_ptr(),
                                                cand.len[ti].data_ptr(), n, rows.data_ptr(), _stream_ptr(dev)))
        off, total = _row_offsets(lib, rows)
        session = torch.empty(total, dtype=torch.int32, device=dev)
        cnd = torch.empty(total, dtype=torch.int64, device=dev)
        score = torch.empty(total, dtype=torch.float32, device=dev)
        label = torch.empty(total, dtype=torch.uint8, device=dev) if lab is not None else None
        lc = lab.to_c() if lab is not None else None
        with torch.cuda.device(dev):
            N.check(lib.otto_regular_rows(C.byref(ss), cand.aid[ti].data_ptr(), cand.score[ti].data_ptr(), cand.len[ti].data_ptr(), n,
                                          off.data_ptr(), sid.data_ptr(), C.byref(lc) if lc is not None else None,
                                          session.data_ptr(), cnd.data_ptr(), score.data_ptr(),
                                          label.data_ptr() if label is not None else None, _stream_ptr(dev)))
        out[t] = _frame_from_device(session, cnd, score, label)
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
            # membership of (row, aid) in the sorted label keys: one vectorised np.isin, no per-row Python
            lab = labels.get(t, {})
            if not isinstance(lab, LabelCSR):
                lab = LabelCSR.build(sessions.session_ids, lab, "cpu")
            lo = lab.offsets.cpu().numpy()
            lkey = (np.repeat(np.arange(lo.size - 1, dtype=np.int64), np.diff(lo)) << 32) | lab.aid.cpu().numpy().astype(np.int64)
            f["candidate_labels"] = np.isin((row.astype(np.int64) << 32) | a.astype(np.int64), lkey).astype(np.uint8)
        out[t] = f
    return out


def recall_counts(pred: torch.Tensor, labels: "LabelCSR", k: int = 20):
    """(hits, denominator) of covisitation/inference.py:251-252 on the device (otto_recall_counts)."""
    lib = N.lib()
    dev = pred.device
    _require_cuda(pred, "pred")
    out = torch.zeros(2, dtype=torch.int64, device=dev)
    p = pred.contiguous()
    lc = labels.to_c()
    with torch.cuda.device(dev):
        N.check(lib.otto_recall_counts(p.data_ptr(), int(p.shape[0]), int(p.shape[1]), C.byref(lc), k, out.data_ptr(), _stream_ptr(dev)))
    h, d = out.tolist()
    return int(h), int(d)


def recall_at_20(pred: torch.Tensor, labels) -> float:
    """covisitation/inference.py:251-257 on device predictions [S, n]: sum |set(pred) ∩ label| / sum min(|label|, 20).
    labels: LabelCSR, or a sequence of label collections aligned with the rows of pred."""
    if not pred.is_cuda:
        pred = pred.to("cuda")           # the counting runs on the device; there is no host implementation
    pred = pred.to(torch.int32)
    if not isinstance(labels, LabelCSR):
        labels = LabelCSR.build(np.arange(int(pred.shape[0])), list(labels), pred.device)
    hits, denom = recall_counts(pred, labels, 20)
    return hits / denom if denom else 0.0
