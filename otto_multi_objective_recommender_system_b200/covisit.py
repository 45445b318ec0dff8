"""Covisitation-matrix build on one B200: host side of the C ABI (include/otto_covisit.h).

This is the builder the reference ships without: it turns `(session, aid, ts, type)`
frames into the per-aid top-K tables that src/covisitation/inference.py:87-111 and
src/ranker/covisitation_candidate_generation.py:49-73 read as `top_15_<stem>_<part>.pqt`.
torch is plumbing only (device buffers, streams); every step of the path runs in libotto_covisit.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field, replace

import numpy as np
import torch

from . import _native as N
from .synth import EventFrame

TS_MIN = 1659304800   # dataset min / max ts (EDA notebook cell 6); constants of the time weight
TS_MAX = 1662328791


@dataclass(frozen=True)
class CovisitSpec:
    """Recipe of one matrix variant.  Presets below are the three graded variants."""
    weight_mode: int = N.WEIGHT_TIME
    type_weight: tuple = (1, 6, 3)
    event_types: tuple = (0, 1, 2)     # pre-filter before the tail cut
    x_types: tuple = (0, 1, 2)         # pair-level filters
    y_types: tuple = (0, 1, 2)
    window_s: int = 86400
    tail_n: int = 30
    k: int = 20
    ts_min: int = TS_MIN
    ts_max: int = TS_MAX
    split_ub: int = 0                  # 0 = library default
    global_events: int = 0             # multi-GPU: events of all ranks (sizes the bin arrays); 0 = single frame

    def to_c(self, n_aids: int) -> N.OttoCovisitSpec:
        mask = lambda ts: sum(1 << int(t) for t in set(ts))
        return N.OttoCovisitSpec(n_aids, self.weight_mode, (C.c_int32 * 3)(*[int(w) for w in self.type_weight]),
                                 mask(self.event_types), mask(self.x_types), mask(self.y_types), self.window_s,
                                 self.tail_n, self.k, self.ts_min, self.ts_max, self.split_ub, self.global_events)


    _MODES = {"time": N.WEIGHT_TIME, "type": N.WEIGHT_TYPE, "unit": N.WEIGHT_UNIT}

    @classmethod
    def from_dict(cls, d: dict) -> "CovisitSpec":
        """Recipe from a plain mapping (configs/*.json): weight = time | type | unit plus any field of the spec."""
        d = dict(d)
        mode = d.pop("weight", "time")
        if mode not in cls._MODES:
            raise ValueError(f"weight must be one of {sorted(cls._MODES)}, got {mode!r}")
        allowed = {"type_weight", "event_types", "x_types", "y_types", "window_s", "tail_n", "k", "ts_min", "ts_max", "split_ub"}
        unknown = set(d) - allowed
        if unknown:
            raise ValueError(f"unknown recipe keys: {sorted(unknown)}")
        for key in ("type_weight", "event_types", "x_types", "y_types"):
            if key in d:
                d[key] = tuple(int(v) for v in d[key])
                if key != "type_weight" and not set(d[key]) <= {0, 1, 2}:
                    raise ValueError(f"{key} must be a subset of [0, 1, 2]")
        if "type_weight" in d and len(d["type_weight"]) != 3:
            raise ValueError("type_weight needs three entries: clicks, carts, orders")
        return cls(weight_mode=cls._MODES[mode], **d)


def load_stem_recipes(path) -> dict:
    """{stem: CovisitSpec} from a JSON file (configs/unpinned_stems.example.json); keys starting with '_' are comments.
    Stems must be file stems the readers know (covisitation/inference.py:87-111)."""
    import json
    from .candidates import STEMS
    with open(path) as f:
        raw = json.load(f)
    out = {}
    for stem, recipe in raw.items():
        if stem.startswith("_"):
            continue
        if stem not in STEMS:
            raise ValueError(f"unknown matrix stem {stem!r}; the readers load {STEMS}")
        out[stem] = CovisitSpec.from_dict(recipe)
    return out


CLICKS = CovisitSpec(N.WEIGHT_TIME, k=20)                                  # stem "time_weighted"
CARTS_ORDERS = CovisitSpec(N.WEIGHT_TYPE, type_weight=(1, 6, 3), k=15)     # stem "cart_weighted"
BUY2BUY = CovisitSpec(N.WEIGHT_UNIT, event_types=(1, 2), window_s=14 * 86400, k=15)   # stem "cart_order"
VARIANTS = {"time_weighted": CLICKS, "cart_weighted": CARTS_ORDERS, "cart_order": BUY2BUY}


def _stream_ptr(device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must live on a CUDA device: this path has no CPU implementation")


@dataclass
class EventCSR:
    """Session-sorted CSR on the device.  order='desc': most recent first (builder layout);
    order='asc': file order (candidate-generation layout)."""
    session_ids: torch.Tensor   # int32 [S]
    offsets: torch.Tensor       # int32 [S + 1]
    aid: torch.Tensor           # int32 [E]
    ts: torch.Tensor            # int32 [E]
    type: torch.Tensor          # uint8 [E]
    n_aids: int
    order: str
    max_len_dev: torch.Tensor | None = None   # int32 [1] longest session (filled by ingest)

    @property
    def n_sessions(self) -> int:
        return int(self.session_ids.numel())

    @property
    def n_events(self) -> int:
        return int(self.aid.numel())

    def slice_sessions(self, lo: int, hi: int) -> "EventCSR":
        """Sessions [lo, hi) as their own CSR (the multi-GPU shard of a rank)."""
        lo, hi = max(0, lo), min(self.n_sessions, hi)
        e0, e1 = int(self.offsets[lo].item()), int(self.offsets[hi].item())
        # the whole frame's longest session stays a valid upper bound for the slice
        return EventCSR(self.session_ids[lo:hi], (self.offsets[lo:hi + 1] - e0).contiguous(), self.aid[e0:e1],
                        self.ts[e0:e1], self.type[e0:e1], self.n_aids, self.order, self.max_len_dev)


def _zero_copy_ok(frame: EventFrame) -> bool:
    return all((not t.is_cuda) and t.is_pinned() and t.is_contiguous() for t in (frame.session, frame.aid, frame.ts, frame.type)) \
        and frame.aid.dtype == torch.int32 and frame.type.dtype == torch.uint8


def ingest(frame: EventFrame, order: str = "desc", device=None, zero_copy: bool = False) -> EventCSR:
    """Frame columns -> CSR.  Replaces the sort + 100k-session chunk writers
    (utilities/split_dataset_writer_parquet.py:13-33) and builder step 2 (ts-descending stable sort).

    zero_copy (builder input from a PINNED host frame in file order): only `session` and `ts` are copied to the
    device; `aid` and `type` stay in pinned host memory and the result is an order='asc' CSR whose tails the build
    reads over PCIe (otto_covisit_count_begin_asc) - the build only ever touches the <= 30 most recent events of a
    session, 63 % of OTTO's events.  The frame must be sorted by (session, ts); aid / type are validated where they are
    read (the tail kernels), not here."""
    if order not in ("asc", "desc"):
        raise ValueError("Invalid order")
    if zero_copy and not _zero_copy_ok(frame):
        raise ValueError("zero_copy needs contiguous pinned host columns (aid int32, type uint8)")
    device = torch.device(device if device is not None else (frame.aid.device if frame.aid.is_cuda else "cuda"))
    lib = N.lib()
    sess = frame.session.to(device=device, dtype=torch.int32, non_blocking=zero_copy).contiguous()
    ts = frame.ts.to(device=device, dtype=torch.int32, non_blocking=zero_copy).contiguous()
    if zero_copy:
        aid, typ = frame.aid, frame.type
    else:
        aid = frame.aid.to(device=device, dtype=torch.int32).contiguous()
        typ = frame.type.to(device=device, dtype=torch.uint8).contiguous()
    E = int(sess.numel())
    if E >= 2 ** 31:
        raise ValueError("frames are limited to 2^31 - 1 events per device")
    with torch.cuda.device(device):
        st = _stream_ptr(device)
        need = int(lib.otto_ingest_scratch_bytes(E))
        scratch = torch.empty(need, dtype=torch.uint8, device=device)
        info = (C.c_int64 * 3)()

        def scan():
            # one read of the frame: session starts per tile, (session, ts) order, aid in [0, n_aids) and type in {0, 1, 2}
            N.check(lib.otto_ingest_scan(sess.data_ptr(), None if zero_copy else aid.data_ptr(), ts.data_ptr(),
                                         None if zero_copy else typ.data_ptr(), E, int(frame.n_aids), scratch.data_ptr(), need, info, st))
            if info[2]:
                raise N.OttoError(N.OTTO_EINVAL, f"{info[2]} events have an aid outside [0, {frame.n_aids}) or a type above 2")
        scan()
        if not info[1] and zero_copy:
            raise N.OttoError(N.OTTO_EUNSORTED, "zero_copy ingest needs a frame sorted by (session, ts)")
        if not info[1]:
            # unsorted input: sort by (session, ts), stable, like df.sort_values(['session', 'ts']).  The reference's
            # frames are written sorted (utilities/split_dataset_writer_parquet.py:17), so this is the rare path and the
            # only place where a torch op does the work.
            o = torch.sort(ts, stable=True).indices
            o = o[torch.sort(sess[o], stable=True).indices]
            sess, aid, ts, typ = sess[o].contiguous(), aid[o].contiguous(), ts[o].contiguous(), typ[o].contiguous()
            scan()
        S = int(info[0])
        ids = torch.empty(S, dtype=torch.int32, device=device)
        offsets = torch.empty(S + 1, dtype=torch.int32, device=device)
        max_len = torch.zeros(1, dtype=torch.int32, device=device)
        N.check(lib.otto_ingest_offsets(sess.data_ptr(), E, S, scratch.data_ptr(), need, ids.data_ptr(), offsets.data_ptr(),
                                        max_len.data_ptr(), st))
        if order == "asc" or zero_copy:
            return EventCSR(ids, offsets, aid, ts, typ, frame.n_aids, "asc", max_len)
        aid_d, ts_d, typ_d = torch.empty_like(aid), torch.empty_like(ts), torch.empty_like(typ)
        N.check(lib.otto_ingest_desc(offsets.data_ptr(), ids.numel(), aid.data_ptr(), ts.data_ptr(), typ.data_ptr(), E,
                                     aid_d.data_ptr(), ts_d.data_ptr(), typ_d.data_ptr(), st))
    return EventCSR(ids, offsets, aid_d, ts_d, typ_d, frame.n_aids, "desc", max_len)


@dataclass
class TopKTable:
    """Fixed-stride per-aid top-K table on the device (OttoTopK)."""
    aid_y: torch.Tensor   # int32 [A, K], -1 padded
    wgt: torch.Tensor     # float32 [A, K]
    len: torch.Tensor     # int32 [A]
    cnt: torch.Tensor | None = None    # uint32 as int32 storage [A, K]
    tsum: torch.Tensor | None = None   # uint64 as int64 storage [A, K]

    @property
    def n_aids(self) -> int:
        return int(self.aid_y.shape[0])

    @property
    def k(self) -> int:
        return int(self.aid_y.shape[1])

    @staticmethod
    def empty(n_aids: int, k: int, device, exact: bool = False) -> "TopKTable":
        return TopKTable(torch.empty((n_aids, k), dtype=torch.int32, device=device),
                         torch.empty((n_aids, k), dtype=torch.float32, device=device),
                         torch.empty((n_aids,), dtype=torch.int32, device=device),
                         torch.empty((n_aids, k), dtype=torch.int32, device=device) if exact else None,
                         torch.empty((n_aids, k), dtype=torch.int64, device=device) if exact else None)

    def to_c(self) -> N.OttoTopK:
        return N.OttoTopK(self.n_aids, self.k, self.aid_y.data_ptr(), self.wgt.data_ptr(), self.len.data_ptr(),
                          self.cnt.data_ptr() if self.cnt is not None else None,
                          self.tsum.data_ptr() if self.tsum is not None else None)

    def to_rows(self):
        """File rows (aid_x, aid_y, wgt): aid_x ascending, best first - the `top_<k>_<stem>` layout."""
        lib = N.lib()
        dev = self.aid_y.device
        with torch.cuda.device(dev):
            st = _stream_ptr(dev)
            A = self.n_aids
            row_off = torch.empty(A + 1, dtype=torch.int64, device=dev)
            scratch = torch.empty(A // 2048 + 8, dtype=torch.int64, device=dev)
            n_rows = C.c_int64(0)
            tc = self.to_c()
            N.check(lib.otto_topk_row_offsets(C.byref(tc), row_off.data_ptr(), C.byref(n_rows), scratch.data_ptr(),
                                              scratch.numel() * 8, st))
            n = int(n_rows.value)
            ax = torch.empty(n, dtype=torch.int32, device=dev)
            ay = torch.empty(n, dtype=torch.int32, device=dev)
            w = torch.empty(n, dtype=torch.float32, device=dev)
            N.check(lib.otto_topk_to_rows(C.byref(tc), row_off.data_ptr(), ax.data_ptr(), ay.data_ptr(), w.data_ptr(), st))
        return ax, ay, w

    def to_host_rows(self, buffers: list | None = None):
        """to_rows() copied into pinned host memory (a pageable .cpu() runs at ~2 GB/s, pinned at PCIe speed).
        `buffers`: pinned tensors from an earlier call, reused when large enough.  Returns (ax, ay, w, buffers)."""
        rows = self.to_rows()
        n = rows[0].numel()
        if buffers is None or buffers[0].numel() < n:
            buffers = [torch.empty(max(n, 1), dtype=r.dtype, pin_memory=True) for r in rows]
        out = []
        for r, b in zip(rows, buffers):
            b[:n].copy_(r, non_blocking=True)
            out.append(b[:n])
        torch.cuda.current_stream(self.aid_y.device).synchronize()
        return out[0], out[1], out[2], buffers

    def to_pandas(self):
        import pandas as pd
        ax, ay, w = self.to_rows()
        return pd.DataFrame({"aid_x": ax.cpu().numpy(), "aid_y": ay.cpu().numpy(), "wgt": w.cpu().numpy()})

    @staticmethod
    def from_rows(aid_x: torch.Tensor, aid_y: torch.Tensor, wgt: torch.Tensor | None, n_aids: int, k: int) -> "TopKTable":
        """Rows grouped by aid_x in rank order -> table; the device form of covisitation_df_to_dict
        (covisitation/inference.py:19-35)."""
        lib = N.lib()
        dev = aid_x.device
        _require_cuda(aid_x, "aid_x")
        t = TopKTable.empty(n_aids, k, dev)
        ax = aid_x.to(torch.int32).contiguous()
        ay = aid_y.to(torch.int32).contiguous()
        w = wgt.to(torch.float32).contiguous() if wgt is not None else None
        with torch.cuda.device(dev):
            tc = t.to_c()
            N.check(lib.otto_rows_to_topk(ax.data_ptr(), ay.data_ptr(), w.data_ptr() if w is not None else None,
                                          ax.numel(), C.byref(tc), _stream_ptr(dev)))
        return t


class CovisitBuilder:
    """Phased build for one rank; buffers are kept so that repeated builds (benchmarks, the seven stems of
    one pipeline run) do not re-allocate."""

    def __init__(self, csr: EventCSR, spec: CovisitSpec, exact: bool = False):
        # order='desc': the most-recent-first CSR of otto_ingest_desc; order='asc': file order, reversed by the tail kernels
        # (then aid / type may be pinned host tensors, read over PCIe)
        _require_cuda(csr.ts, "csr.ts")
        _require_cuda(csr.offsets, "csr.offsets")
        for name, t in (("aid", csr.aid), ("type", csr.type)):
            if not t.is_cuda and not (csr.order == "asc" and t.is_pinned()):
                raise RuntimeError(f"csr.{name} must live on a CUDA device (or, for an order='asc' CSR, in pinned host memory)")
        self.lib = N.lib()
        self.csr, self.spec, self.exact = csr, spec, exact
        self.device = csr.ts.device
        self.cspec = spec.to_c(csr.n_aids)
        self.ev = N.OttoEvents(csr.n_sessions, csr.n_events, csr.offsets.data_ptr(), csr.aid.data_ptr(),
                               csr.ts.data_ptr(), csr.type.data_ptr())
        sizes = N.OttoBuildSizes()
        N.check(self.lib.otto_covisit_sizes(csr.n_sessions, csr.n_events, C.byref(self.cspec), C.byref(sizes)))
        self.sizes = sizes
        self.workspace = torch.empty(sizes.workspace_bytes, dtype=torch.uint8, device=self.device)
        self.records = None
        self.merged = None
        self.scratch = None
        self.stats = N.OttoBuildStats()
        self.table = None

    # -- phases -------------------------------------------------------------------------------
    def _st(self) -> int:
        return _stream_ptr(self.device)

    def count_begin(self) -> None:
        with torch.cuda.device(self.device):
            begin = self.lib.otto_covisit_count_begin_asc if self.csr.order == "asc" else self.lib.otto_covisit_count_begin
            N.check(begin(C.byref(self.ev), C.byref(self.cspec), self.workspace.data_ptr(), self.workspace.numel(), self._st()))

    def count_finish(self) -> dict:
        with torch.cuda.device(self.device):
            N.check(self.lib.otto_covisit_count_finish(C.byref(self.ev), C.byref(self.cspec), self.workspace.data_ptr(),
                                                       self.workspace.numel(), C.byref(self.stats), self._st()))
        return self.stats.as_dict()

    def views(self) -> dict:
        """Device views into the workspace as tensors (no copies)."""
        ptrs = [C.c_void_p() for _ in range(4)]
        N.check(self.lib.otto_covisit_views(C.byref(self.ev), C.byref(self.cspec), self.workspace.data_ptr(),
                                            self.workspace.numel(), *[C.byref(p) for p in ptrs]))
        base = self.workspace.data_ptr()
        A, B = self.csr.n_aids, int(self.stats.bins)

        def view(ptr, dtype, n):
            off = ptr.value - base
            return self.workspace[off: off + n * torch.empty(0, dtype=dtype).element_size()].view(dtype)
        return {"bin_offsets": view(ptrs[0], torch.int64, B + 1), "bin_base": view(ptrs[1], torch.int32, A + 1),
                "bin_x": view(ptrs[2], torch.int32, max(B, 1)), "row_total": view(ptrs[3], torch.int32, A)}

    def scatter(self) -> torch.Tensor:
        """Records grouped by bin; bin offsets (views()["bin_offsets"]) are valid afterwards.  The buffer holds the
        P final records followed by the staging area of the hot rows."""
        P = int(self.stats.pairs) + int(self.stats.hot_pairs)
        if self.records is None or self.records.numel() < max(P, 1):
            self.records = torch.empty(max(P, 1), dtype=torch.int64, device=self.device)   # 8-byte {aid_y, v}
        with torch.cuda.device(self.device):
            N.check(self.lib.otto_covisit_scatter(C.byref(self.ev), C.byref(self.cspec), self.workspace.data_ptr(),
                                                  self.workspace.numel(), self.records.data_ptr(), P, self._st()))
        return self.records

    # -- owner-direct scatter (multi-GPU; protocol in include/otto_covisit.h) ------------------
    def owner_plan(self, aid_cuts, rank: int, owner_records=None) -> "N.OttoOwnerPlan":
        G = len(aid_cuts) - 1
        if not 1 <= G <= N.MAX_OWNERS:
            raise ValueError(f"owner-direct scatter supports 1..{N.MAX_OWNERS} owners")
        plan = N.OttoOwnerPlan()
        plan.n_owners, plan.rank = G, rank
        for i in range(N.MAX_OWNERS + 1):
            plan.aid_cuts[i] = int(aid_cuts[min(i, G)])
        for i in range(N.MAX_OWNERS):
            plan.owner_records[i] = int(owner_records[i]) if owner_records is not None and i < G else None
        return plan

    def count_finish_owned(self, aid_cuts, rank: int, row_before: torch.Tensor) -> dict:
        """views()["row_total"] must hold the totals over all ranks; row_before the pairs of lower ranks per row."""
        _require_cuda(row_before, "row_before")
        if row_before.numel() < self.csr.n_aids or row_before.element_size() != 4:
            raise ValueError("row_before must hold n_aids 32-bit counts")
        plan = self.owner_plan(aid_cuts, rank)
        with torch.cuda.device(self.device):
            N.check(self.lib.otto_covisit_count_finish_owned(C.byref(self.ev), C.byref(self.cspec), self.workspace.data_ptr(),
                                                             self.workspace.numel(), C.byref(plan), row_before.data_ptr(),
                                                             C.byref(self.stats), self._st()))
        return self.stats.as_dict()

    def scatter_owned(self, aid_cuts, rank: int, owner_records) -> None:
        plan = self.owner_plan(aid_cuts, rank, owner_records)
        with torch.cuda.device(self.device):
            N.check(self.lib.otto_covisit_scatter_owned(C.byref(self.ev), C.byref(self.cspec), self.workspace.data_ptr(),
                                                        self.workspace.numel(), C.byref(plan), self._st()))

    # -- staged scatter (multi-GPU, sender-side combining; protocol in include/otto_covisit.h) ----
    def stage_plan(self, counts_all: torch.Tensor | None, world: int, rank: int, sync: bool = True):
        """Bucket plan from counts_all: [world, n_aids] per-row counts of every rank (None with one rank = the
        workspace's own, after count_finish).  sync=True -> records every rank will stage (the same list on every
        rank; one synchronisation); sync=False enqueues only - read the list with stage_totals() later, e.g. behind
        the synchronisation of count_finish_owned."""
        need = int(self.lib.otto_covisit_stage_plan_bytes(C.byref(self.cspec), self.csr.n_sessions, self.csr.n_events, world))
        if need < 0:
            raise ValueError(f"staged scatter supports 1..{N.MAX_OWNERS} ranks")
        if getattr(self, "_stage_scratch", None) is None or self._stage_scratch.numel() < need:
            self._stage_scratch = torch.empty(need, dtype=torch.uint8, device=self.device)
        totals = (C.c_int64 * N.MAX_OWNERS)()
        with torch.cuda.device(self.device):
            N.check(self.lib.otto_covisit_stage_plan(C.byref(self.ev), C.byref(self.cspec), self.workspace.data_ptr(),
                                                     self.workspace.numel(), counts_all.data_ptr() if counts_all is not None else None,
                                                     world, rank, self._stage_scratch.data_ptr(), self._stage_scratch.numel(),
                                                     totals if sync else None, self._st()))
        return [int(totals[g]) for g in range(world)] if sync else None

    def stage_totals(self, world: int) -> list:
        if getattr(self, "_stage_scratch", None) is None:
            raise RuntimeError("stage_totals() before stage_plan()")
        totals = (C.c_int64 * N.MAX_OWNERS)()
        with torch.cuda.device(self.device):
            N.check(self.lib.otto_covisit_stage_totals(C.byref(self.cspec), self.csr.n_sessions, self.csr.n_events,
                                                       self._stage_scratch.data_ptr(), world, totals, self._st()))
        return [int(totals[g]) for g in range(world)]

    def scatter_staged(self, world: int, staged_ptr: int) -> None:
        """Pass A: this rank's pairs into the coarse buckets of its staging buffer (device pointer; at least
        stage_totals()[rank] records - the library trusts the caller's allocation, as it does for the record buffer)."""
        if getattr(self, "_stage_scratch", None) is None:
            raise RuntimeError("scatter_staged() before stage_plan()")
        with torch.cuda.device(self.device):
            N.check(self.lib.otto_covisit_scatter_staged(C.byref(self.ev), C.byref(self.cspec), self.workspace.data_ptr(),
                                                         self.workspace.numel(), self._stage_scratch.data_ptr(), world, staged_ptr,
                                                         self._st()))

    def place_staged(self, staged_ptrs, aid_lo: int, aid_hi: int) -> torch.Tensor:
        """Pass B at the owner of rows [aid_lo, aid_hi): every rank's staging buffer (as mapped here) -> self.records."""
        P = int(self.stats.pairs) + int(self.stats.hot_pairs)
        if self.records is None or self.records.numel() < max(P, 1):
            self.records = torch.empty(max(P, 1), dtype=torch.int64, device=self.device)
        world = len(staged_ptrs)
        ptrs = (C.c_void_p * world)(*[int(q) for q in staged_ptrs])
        with torch.cuda.device(self.device):
            N.check(self.lib.otto_covisit_place_staged(C.byref(self.ev), C.byref(self.cspec), self.workspace.data_ptr(),
                                                       self.workspace.numel(), self._stage_scratch.data_ptr(), world, ptrs,
                                                       aid_lo, aid_hi, self.records.data_ptr(), max(P, 1), self._st()))
        return self.records

    def partition(self) -> torch.Tensor:
        """Second half of scatter() on self.records (already filled, here by every rank of the box)."""
        P = int(self.stats.pairs) + int(self.stats.hot_pairs)
        with torch.cuda.device(self.device):
            N.check(self.lib.otto_covisit_partition(C.byref(self.ev), C.byref(self.cspec), self.workspace.data_ptr(),
                                                    self.workspace.numel(), self.records.data_ptr(), P, self._st()))
        return self.records

    def reduce(self, segments=None, bin_lo: int = 0, bin_hi: int | None = None, aid_lo: int = 0,
               aid_hi: int | None = None, table: TopKTable | None = None, sync: bool = True) -> TopKTable:
        """segments: list of (records tensor, offsets tensor int64 [bins + 1]); default = this rank's own."""
        A = self.csr.n_aids
        bin_hi = int(self.stats.bins) if bin_hi is None else bin_hi
        aid_hi = A if aid_hi is None else aid_hi
        v = self.views()
        if segments is None:
            segments = [(self.records, v["bin_offsets"])]
        if table is None:
            if self.table is None:
                self.table = TopKTable.empty(A, self.spec.k, self.device, self.exact)
            table = self.table
        need = int(self.lib.otto_covisit_reduce_scratch_bytes(C.byref(self.cspec), bin_hi - bin_lo, aid_hi - aid_lo))
        if self.scratch is None or self.scratch.numel() < need:
            self.scratch = torch.empty(need, dtype=torch.uint8, device=self.device)
        # a segment's records are a tensor, or (device pointer, n_records) for a peer's slab mapped over NVLink
        ptr = lambda r: r[0] if isinstance(r, tuple) else r.data_ptr()
        segs = (N.OttoPairSegment * len(segments))(*[N.OttoPairSegment(ptr(r), o.data_ptr()) for r, o in segments])
        tc = table.to_c()
        with torch.cuda.device(self.device):
            N.check(self.lib.otto_covisit_reduce(C.byref(self.cspec), v["bin_base"].data_ptr(), v["bin_x"].data_ptr(),
                                                 bin_lo, bin_hi, aid_lo, aid_hi, segs, len(segments),
                                                 self.scratch.data_ptr(), self.scratch.numel(), C.byref(tc),
                                                 C.byref(self.stats) if sync else None, self._st()))
        return table

    def merge_segments(self, segments, n_bins: int):
        """Multi-GPU owner side: the G received (records, offsets) segments of my bins -> one bin-contiguous
        segment, so that the reduce kernels see a single run per bin."""
        # a segment's records are a tensor, or (device pointer, n_records) for a peer's slab mapped over NVLink
        ptr = lambda r: r[0] if isinstance(r, tuple) else r.data_ptr()
        cnt = lambda r: int(r[1]) if isinstance(r, tuple) else int(r.numel())
        total_cap = sum(cnt(r) for r, _ in segments)
        if self.merged is None or self.merged.numel() < max(total_cap, 1):
            self.merged = torch.empty(max(total_cap, 1), dtype=torch.int64, device=self.device)
        merged = self.merged
        offsets = torch.empty(n_bins + 1, dtype=torch.int64, device=self.device)
        need = int(self.lib.otto_covisit_merge_scratch_bytes(n_bins))
        scratch = torch.empty(need, dtype=torch.uint8, device=self.device)
        segs = (N.OttoPairSegment * len(segments))(*[N.OttoPairSegment(ptr(r), o.data_ptr()) for r, o in segments])
        n = C.c_int64(0)
        with torch.cuda.device(self.device):
            N.check(self.lib.otto_covisit_merge_segments(segs, len(segments), n_bins, merged.data_ptr(), merged.numel(),
                                                         offsets.data_ptr(), scratch.data_ptr(), need, C.byref(n), self._st()))
        return merged, offsets

    # -- whole build --------------------------------------------------------------------------
    def build(self, sync: bool = True) -> TopKTable:
        """count -> scatter -> reduce on the current stream.  The pair count has to reach the host once to
        size the record buffer, so the first call synchronises; later calls reuse the buffers."""
        self.count_begin()
        self.count_finish()
        self.scatter()
        return self.reduce(sync=sync)


def build_topk(csr: EventCSR, spec: CovisitSpec, exact: bool = False):
    """One matrix variant on one GPU -> (TopKTable, stats dict)."""
    b = CovisitBuilder(csr, spec, exact=exact)
    t = b.build()
    return t, b.stats.as_dict()


class FramePipeline:
    """Builds one matrix per frame from a STREAM of pinned host frames, two frames in flight: while frame i is being
    ingested and built on the compute stream, frame i + 1 is uploaded on a copy stream and the rows of frame i - 1
    travel back (PCIe is full duplex).  A step of the pipeline costs max(upload, build + rows back) instead of their
    sum; every frame is still uploaded, built and read back in full.  Use: `for ax, ay, w in FramePipeline(spec,
    device).run(frames): ...` - the yielded tensors are pinned host views that stay valid until two frames later.
    The reference has nothing comparable (its builder is absent); the case it serves is a sequence of different
    frames (train + validation, train + test, the chunked frames of a larger catalogue)."""

    def __init__(self, spec: CovisitSpec, device, exact: bool = False):
        self.spec, self.device, self.exact = spec, torch.device(device), exact
        self.copy_stream = torch.cuda.Stream(self.device)
        self.out_stream = torch.cuda.Stream(self.device)
        self.slots = [dict(cols=None, uploaded=None, consumed=None, out=None, out_done=None, builder=None) for _ in range(2)]

    def _upload(self, slot: dict, frame: EventFrame) -> None:
        if not all(t.is_pinned() for t in (frame.session, frame.aid, frame.ts, frame.type)):
            raise ValueError("FramePipeline needs pinned host frames")
        n = len(frame)
        if slot["cols"] is None or slot["cols"][0].numel() < n:
            slot["cols"] = [torch.empty(n, dtype=d, device=self.device) for d in (torch.int32, torch.int32, torch.int32, torch.uint8)]
        with torch.cuda.stream(self.copy_stream):
            if slot["consumed"] is not None:
                self.copy_stream.wait_event(slot["consumed"])     # the build that read this slot's columns is done
            for dst, src in zip(slot["cols"], (frame.session, frame.aid, frame.ts, frame.type)):
                dst[:n].copy_(src, non_blocking=True)
            slot["uploaded"] = torch.cuda.Event()
            slot["uploaded"].record(self.copy_stream)
        slot["n"], slot["n_aids"] = n, frame.n_aids

    def _build(self, slot: dict) -> None:
        st = torch.cuda.current_stream(self.device)
        st.wait_event(slot["uploaded"])
        n = slot["n"]
        f = EventFrame(*(c[:n] for c in slot["cols"]), n_aids=slot["n_aids"])
        csr = ingest(f, "asc", device=self.device)
        b = CovisitBuilder(csr, self.spec, exact=self.exact)
        old = slot["builder"]
        if old is not None and old.workspace.numel() >= b.workspace.numel():
            b.workspace, b.records, b.scratch, b.table = old.workspace, old.records, old.scratch, old.table
        slot["builder"] = b
        rows = b.build().to_rows()
        slot["consumed"] = torch.cuda.Event()
        slot["consumed"].record(st)
        if slot["out_done"] is not None:
            slot["out_done"].synchronize()                        # the previous rows of this slot have left
        if slot["out"] is None or any(o.numel() < r.numel() for o, r in zip(slot["out"], rows)):
            slot["out"] = [torch.empty(max(1, r.numel()), dtype=r.dtype, pin_memory=True) for r in rows]
        self.out_stream.wait_event(slot["consumed"])
        with torch.cuda.stream(self.out_stream):
            views = []
            for o, r in zip(slot["out"], rows):
                o[:r.numel()].copy_(r, non_blocking=True)
                r.record_stream(self.out_stream)
                views.append(o[:r.numel()])
            slot["out_done"] = torch.cuda.Event()
            slot["out_done"].record(self.out_stream)
        slot["views"] = views

    def run(self, frames):
        it = iter(frames)
        nxt = next(it, None)
        if nxt is None:
            return
        self._upload(self.slots[0], nxt)
        i, pending = 0, None
        while nxt is not None:
            cur = self.slots[i & 1]
            nxt = next(it, None)
            if nxt is not None:
                self._upload(self.slots[(i + 1) & 1], nxt)         # enqueued before the host starts driving the build
            self._build(cur)
            if pending is not None:
                pending["out_done"].synchronize()
                yield tuple(pending["views"])
            pending = cur
            i += 1
        pending["out_done"].synchronize()
        yield tuple(pending["views"])
