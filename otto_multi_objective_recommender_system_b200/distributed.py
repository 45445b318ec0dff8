"""Multi-GPU covisitation build: one process per GPU, sessions sharded by contiguous chunk, pair records
exchanged by one all-to-all keyed on aid_x ownership.

    rank r                                   collective (torch.distributed, NCCL over NVLink)
    ------                                   -----------------------------------------------
    count_begin  (tails, in-session dedupe, local pairs per aid_x row)
                                             all-reduce  row totals [A] uint32 -> identical bins everywhere
    count_finish (bins, local record offsets)
    scatter      (records grouped by bin; an owner's bins are one contiguous slab)
                                             all-reduce  bin counts [B] int64 -> balanced aid_x ranges
                                             all-to-all  record slabs + their bin offsets
    reduce       (accumulate + top-k over the G received segments of my aid range)
                                             broadcast   each owner's rows (gather_table) for candidate-gen

Owner-direct variant (default on one NVLink box, `GpuRankBackend(peer=...)`): the exchange is fused into the
scatter kernel.  The per-row pair counts are all-gathered, so every rank knows the owner of each aid_x row and the
position of its own run inside the owner's record buffer; the scatter kernel stores each record straight into that
buffer through the peer mapping (NVLink writes), and everything behind it - partition of the hot rows, accumulate,
top-k - is local to the owner and identical to a single-GPU build of the owner's rows:

    count_begin                              all-gather  row counts [G, A] uint32 -> totals, counts of lower ranks
    count_finish_owned (bins, my layout, cursors into the owners' buffers)
    scatter_owned (records -> owners' HBM)   all-reduce  1 element: "every scatter has landed"
    partition + reduce (one local segment)

Staged variant of the same protocol (default, `GpuRankBackend.staged`): the scatter kernel's short runs cross NVLink
badly (~63-byte store packets), so a rank first combines them in coarse aid_x buckets of its OWN peer-mapped staging
buffer and the owner pulls its buckets in large reads and places the records (include/otto_covisit.h, "Staged scatter"):

    count_begin                              all-gather  row counts [G, A] uint32 -> totals, counts of lower ranks
    stage_plan (bucket table of every rank)  - enqueued before count_finish_owned, whose synchronisation covers it
    count_finish_owned (bins, my layout)
    scatter_staged (pass A, local)           all-reduce  1 element: "every rank has staged"
    place_staged (pass B: my buckets out of every rank's staging buffer -> my record buffer)
    partition + reduce (one local segment)

Every rank forms the same bins (the row totals are reduced before bins exist) and the accumulators are
integers, so the G-GPU table equals the 1-GPU table byte for byte.  The reference has no distributed code
(SURVEY.md §2.1); its only partitioning is the aid_x-range "part" files, which is what ownership mirrors.

The rank-local work sits behind a small backend interface so that the exchange logic runs in CPU tests
(world_size 2, gloo) with a numpy stand-in; the product backend is covisit.CovisitBuilder.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist

import ctypes as C
import os

from . import _native as N
from .covisit import CovisitBuilder, CovisitSpec, EventCSR, TopKTable


STAGED_DEFAULT = "1"      # staged wins on one NVLink box: 15.1 vs 15.6 ms at 2 GPUs, 9.1 vs 10.0 at 4 (profiles/r02_staged_*.json)


class _RawCuda:
    """__cuda_array_interface__ view of a raw device allocation (int64 elements)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}


class PeerRecords:
    """The rank's pair-record buffer, mapped by every other rank of the box (CUDA IPC over NVLink).

    The owner of an aid_x range never receives a copy of the senders' slabs: its merge kernel reads them in
    place from the senders' HBM, so the exchange of SURVEY.md §8e is fused into the merge
    (otto_covisit_merge_segments with peer pointers) and NCCL only carries the small control arrays."""

    def __init__(self, device, group=None):
        self.lib, self.device, self.group = N.lib(), device, group
        self.capacity, self.ptr, self.local, self.peers = 0, None, None, []

    def ensure(self, n_records: int) -> torch.Tensor:
        """Collective.  Grows every rank's buffer to the largest need and re-exchanges the handles."""
        need = torch.tensor([max(int(n_records), 1)], dtype=torch.int64, device=self.device)
        dist.all_reduce(need, op=dist.ReduceOp.MAX, group=self.group)
        need = int(need.item())
        if need <= self.capacity:
            return self.local
        self.close()
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.capacity = need + need // 8
        with torch.cuda.device(self.device):
            p = C.c_void_p()
            N.check(self.lib.otto_peer_alloc(self.capacity * 8, C.byref(p)))
            self.ptr = p.value
            handle = C.create_string_buffer(64)
            N.check(self.lib.otto_peer_get_handle(self.ptr, handle))
            mine = torch.tensor(list(handle.raw), dtype=torch.uint8, device=self.device)
            every = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(every, mine, group=self.group)
            self.peers = []
            for g in range(world):
                if g == rank:
                    self.peers.append(self.ptr)
                    continue
                q = C.c_void_p()
                N.check(self.lib.otto_peer_open(bytes(every[g].cpu().tolist()), C.byref(q)))
                self.peers.append(q.value)
        self.local = torch.as_tensor(_RawCuda(self.ptr, self.capacity), device=self.device)
        return self.local

    def ensure_same(self, n_records: int) -> torch.Tensor:
        """Like ensure() when every rank passes the SAME n_records (the owner-direct layout is computed identically
        everywhere): no collective unless the buffers have to grow."""
        if int(n_records) <= self.capacity:
            return self.local
        return self.ensure(n_records)

    def close(self) -> None:
        if self.ptr is None:
            return
        torch.cuda.synchronize(self.device)
        if dist.is_initialized():
            dist.barrier(group=self.group)      # nobody may still be reading my buffer
        with torch.cuda.device(self.device):
            for q in self.peers:
                if q != self.ptr:
                    self.lib.otto_peer_close(q)
            self.lib.otto_peer_free(self.ptr)
        self.capacity, self.ptr, self.local, self.peers = 0, None, None, []


class GpuRankBackend:
    """Rank-local phases on one B200 (the product backend)."""

    def __init__(self, csr: EventCSR, spec: CovisitSpec, exact: bool = False, peer: PeerRecords | None = None):
        if dist.is_initialized() and dist.get_world_size() > 1 and spec.global_events == 0:
            # collective: the bins come from the all-reduced bounds, so size the bin arrays for the global frame
            total = torch.tensor([csr.n_events], dtype=torch.int64, device=csr.ts.device)
            dist.all_reduce(total)
            from dataclasses import replace
            spec = replace(spec, global_events=int(total.item()))
        self.b = CovisitBuilder(csr, spec, exact=exact)
        self.n_aids, self.k, self.device = csr.n_aids, spec.k, csr.ts.device
        self.peer = peer          # set: records are exchanged through peer memory instead of NCCL
        self._plan_scratch = self._row_before = self._gathered = None

    def count_begin(self) -> torch.Tensor:
        self.b.count_begin()
        self.b.stats.bins = 0
        return self.b.views()["row_total"]          # int32 view of the uint32 pair counts per row, reduced in place

    def count_finish(self):
        stats = self.b.count_finish()
        return stats, self.b.views()["bin_base"]

    def scatter(self):
        """-> (records, bin_offsets); the offsets exist only now (the sub-bins of hot rows are sized while scattering)."""
        if self.peer is not None:
            self.b.records = self.peer.ensure(int(self.b.stats.pairs) + int(self.b.stats.hot_pairs))     # collective
        records = self.b.scatter()
        return records, self.b.views()["bin_offsets"]

    # -- owner-direct scatter: records go straight into the owner's peer-mapped buffer -----------
    @property
    def owner_direct(self) -> bool:
        return (self.peer is not None and os.environ.get("OTTO_OWNER_DIRECT", "1") != "0"
                and dist.is_initialized() and 1 < dist.get_world_size(self.peer.group) <= N.MAX_OWNERS)

    def plan_owners(self, gathered: torch.Tensor, world: int, rank: int):
        """gathered: [world, A] per-row pair counts of every rank (uint32 bit patterns).  One kernel pass
        (otto_covisit_plan_owners) writes the totals into the workspace's row_total, the counts of the lower ranks
        into row_before and finds the balanced aid cuts; one synchronisation."""
        A = self.n_aids
        if self._plan_scratch is None:
            need = int(self.b.lib.otto_covisit_plan_scratch_bytes(A))
            self._plan_scratch = torch.empty(need, dtype=torch.uint8, device=self.device)
            self._row_before = torch.empty(A, dtype=torch.int32, device=self.device)
        cuts = (C.c_int32 * (N.MAX_OWNERS + 1))()
        with torch.cuda.device(self.device):
            N.check(self.b.lib.otto_covisit_plan_owners(gathered.data_ptr(), world, rank, A, self.b.views()["row_total"].data_ptr(),
                                                        self._row_before.data_ptr(), self._plan_scratch.data_ptr(),
                                                        self._plan_scratch.numel(), cuts, self.b._st()))
        return [int(cuts[i]) for i in range(world + 1)], self._row_before

    def count_finish_owned(self, aid_cuts, rank, row_before):
        stats = self.b.count_finish_owned(aid_cuts, rank, row_before)
        return stats, self.b.views()["bin_base"]

    def owner_bin_cuts(self, aid_cuts, bin_base) -> list:
        return [int(self.b.stats.owner_bin_cuts[i]) for i in range(len(aid_cuts))]

    def scatter_owned(self, aid_cuts, rank) -> None:
        # every rank computed the same layout, so the largest buffer any owner needs is known without a collective
        self.b.records = self.peer.ensure_same(int(self.b.stats.owner_records_max))
        self.b.scatter_owned(aid_cuts, rank, self.peer.peers)

    # -- staged scatter: sender-side combining, the owner pulls its buckets over NVLink in large reads --------
    @property
    def staged(self) -> bool:
        """Transport of the owner-direct protocol: False = the scatter kernel stores every ~63-byte run straight into the
        owner's buffer, True = runs are combined in coarse buckets of the sender's own (peer-mapped) staging buffer and
        the owner places them (include/otto_covisit.h, "Staged scatter").  OTTO_STAGED=0/1 overrides the default."""
        return self.owner_direct and os.environ.get("OTTO_STAGED", STAGED_DEFAULT) != "0"

    def stage_plan(self, gathered: torch.Tensor, world: int, rank: int) -> None:
        """Enqueued before count_finish_owned, whose synchronisation then covers the plan's kernels."""
        self.b.stage_plan(gathered, world, rank, sync=False)

    def scatter_staged(self, world: int) -> None:
        totals = self.b.stage_totals(world)                            # the same list on every rank
        self.peer.ensure_same(max(totals))                             # the peer-mapped buffer is the STAGING buffer here
        self.b.scatter_staged(world, self.peer.ptr)

    def place_staged(self, aid_cuts, rank: int) -> None:
        if self.b.records is not None and self.b.records.data_ptr() == self.peer.ptr:
            self.b.records = None                 # a direct-scatter build left the peer buffer here: it is the staging buffer now
        self.b.place_staged(self.peer.peers, aid_cuts[rank], aid_cuts[rank + 1])

    def partition(self):
        records = self.b.partition()
        return records, self.b.views()["bin_offsets"]

    def reduce(self, segments, bin_lo, bin_hi, aid_lo, aid_hi) -> TopKTable:
        # the reduce kernels stream each bin as ONE run: received segments are merged first (the owner-direct scatter
        # never gets here with more than one)
        if len(segments) >= 2:
            segments = [self.b.merge_segments(segments, bin_hi - bin_lo)]
        return self.b.reduce(segments, bin_lo, bin_hi, aid_lo, aid_hi)

    def stats(self) -> dict:
        return self.b.stats.as_dict()


@dataclass
class OwnerPlan:
    aid_cuts: list      # [G + 1] aid_x range of every owner
    bin_cuts: list      # [G + 1] the same cuts in bin ids


def plan_owners(bin_counts: torch.Tensor, bin_base: torch.Tensor, world: int) -> OwnerPlan:
    """Contiguous aid_x ranges with (nearly) equal pair counts: Zipf skew makes equal-width ranges useless.
    bin_counts: global records per bin [B]; bin_base: first bin of every aid [A + 1]."""
    A = bin_base.numel() - 1
    cum = torch.zeros(bin_counts.numel() + 1, dtype=torch.int64, device=bin_counts.device)
    torch.cumsum(bin_counts.to(torch.int64), 0, out=cum[1:])
    before_aid = cum[bin_base.to(torch.int64)]                   # records in rows < x, for x = 0..A
    total = int(cum[-1].item())
    targets = torch.tensor([total * g // world for g in range(1, world)], dtype=torch.int64, device=cum.device)
    inner = torch.searchsorted(before_aid, targets, right=False).clamp_(0, A).tolist() if world > 1 else []
    aid_cuts = [0] + [int(x) for x in inner] + [A]
    for i in range(1, len(aid_cuts)):                              # keep the cuts monotone
        aid_cuts[i] = max(aid_cuts[i], aid_cuts[i - 1])
    bb = bin_base.to(torch.int64)
    bin_cuts = [int(bb[x].item()) for x in aid_cuts]
    return OwnerPlan(aid_cuts, bin_cuts)


def plan_rows(row_total: torch.Tensor, world: int) -> list:
    """aid cuts [G + 1] from the pairs per aid_x row (int64): contiguous ranges with (nearly) equal pair counts."""
    A = row_total.numel()
    before_aid = torch.zeros(A + 1, dtype=torch.int64, device=row_total.device)
    torch.cumsum(row_total, 0, out=before_aid[1:])
    total = int(before_aid[-1].item())
    targets = torch.tensor([total * g // world for g in range(1, world)], dtype=torch.int64, device=row_total.device)
    inner = torch.searchsorted(before_aid, targets, right=False).clamp_(0, A).tolist() if world > 1 else []
    aid_cuts = [0] + [int(x) for x in inner] + [A]
    for i in range(1, len(aid_cuts)):                              # keep the cuts monotone
        aid_cuts[i] = max(aid_cuts[i], aid_cuts[i - 1])
    return aid_cuts


def plan_owners_host(gathered: torch.Tensor, world: int, rank: int):
    """Host-side twin of otto_covisit_plan_owners (used by the CPU stand-in backend of the gloo tests):
    -> (aid cuts, totals, counts of the lower ranks)."""
    counts = gathered.to(torch.int64) & 0xFFFFFFFF                 # [G, A]
    total = counts.sum(0)
    if int(total.max().item()) >= 2 ** 32:
        raise ValueError("an aid_x row holds 2^32 or more pairs")
    before = counts[:rank].sum(0).to(torch.int32)                  # wraps into the uint32 bit pattern
    return plan_rows(total, world), total, before


def _build_owner_direct(backend, group, mark, world: int, rank: int):
    """The exchange fused into the scatter: records are stored straight into the owner's buffer (module docstring).
    Host synchronisations per build: the aid cuts (plan), the buffer sizes (count_finish_owned), the final stats."""
    local = backend.count_begin()                                   # int32 view of this rank's uint32 row counts
    mark("count_begin")
    gathered = getattr(backend, "_gathered", None)
    if gathered is None or gathered.shape != (world, local.numel()) or gathered.device != local.device:
        gathered = torch.empty((world, local.numel()), dtype=local.dtype, device=local.device)
        backend._gathered = gathered
    dist.all_gather(list(gathered.unbind(0)), local, group=group)
    aid_cuts, before = backend.plan_owners(gathered, world, rank)  # workspace row_total := totals over all ranks
    mark("allgather_rows+plan")
    staged = getattr(backend, "staged", False)
    if staged:
        backend.stage_plan(gathered, world, rank)                    # bucket table of every rank (no synchronisation)
    stats, bin_base = backend.count_finish_owned(aid_cuts, rank, before)
    mark("count_finish")
    if staged:
        backend.scatter_staged(world)                                # pass A: my pairs into my own staging buffer
    else:
        backend.scatter_owned(aid_cuts, rank)
    # "every rank's records have landed" (staged: "every rank has staged"): a rank's part of this collective is ordered
    # behind its scatter kernel.  The staging buffers are safe to overwrite in the next build because that build's
    # all-gather of the row counts sits between a peer's place pass and my next pass A.
    landed = torch.zeros(1, dtype=torch.int32, device=local.device)
    dist.all_reduce(landed, group=group)
    if staged:
        backend.place_staged(aid_cuts, rank)                          # pass B: my buckets out of every rank's staging buffer
    mark("scatter")
    records, bin_off = backend.partition()
    mark("partition")
    bin_cuts = backend.owner_bin_cuts(aid_cuts, bin_base)
    plan = OwnerPlan(aid_cuts, bin_cuts)
    lo, hi = bin_cuts[rank], bin_cuts[rank + 1]
    table = backend.reduce([(records, bin_off[lo:hi + 1])], lo, hi, aid_cuts[rank], aid_cuts[rank + 1])
    mark("merge+reduce")
    out_stats = dict(backend.stats())
    out_stats["owned_aids"] = aid_cuts[rank + 1] - aid_cuts[rank]
    return table, (aid_cuts[rank], aid_cuts[rank + 1]), out_stats, plan


def build_topk_distributed(backend, group=None, timing: dict | None = None):
    """All ranks call this with their own session shard.  Returns (table, (aid_lo, aid_hi), stats, plan): rows
    [aid_lo, aid_hi) of `table` are final on this rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    import time
    t_last = [time.perf_counter()]
    use_events = timing is not None and timing.get("__events__")

    def mark(name):     # phase timing for profiling runs only
        if timing is None:
            return
        if use_events:
            # CUDA events on the build's stream, no synchronisation: read with phase_ms_from_events() after the build
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            timing.setdefault("__marks__", []).append((name, e))
            return
        torch.cuda.synchronize()
        now = time.perf_counter()
        timing[name] = timing.get(name, 0.0) + (now - t_last[0]) * 1e3
        t_last[0] = now
    mark("start")
    if world > 1 and getattr(backend, "owner_direct", False):
        return _build_owner_direct(backend, group, mark, world, rank)
    ub = backend.count_begin()
    mark("count_begin")
    if world > 1:
        dist.all_reduce(ub, group=group)
    mark("allreduce_ub")
    stats, bin_base = backend.count_finish()
    mark("count_finish")
    records, bin_off = backend.scatter()
    mark("scatter")
    B = int(stats["bins"])
    counts = (bin_off[1:B + 1] - bin_off[:B]).clone()
    if world > 1:
        dist.all_reduce(counts, group=group)
    plan = plan_owners(counts, bin_base, world)
    mark("allreduce_counts+plan")
    lo, hi = plan.bin_cuts[rank], plan.bin_cuts[rank + 1]
    peer = getattr(backend, "peer", None)
    if world == 1:
        segments = [(records, bin_off)]
    elif peer is not None:
        # slab (sender g -> owner o) = records_g[cuts[g][o] : cuts[g][o + 1]], read in place over NVLink
        cut_off = bin_off[torch.tensor(plan.bin_cuts, device=bin_off.device)]
        every = [torch.empty_like(cut_off) for _ in range(world)]
        dist.all_gather(every, cut_off, group=group)
        cuts = torch.stack(every).tolist()
        n_mine = hi - lo + 1
        send_off = torch.cat([bin_off[plan.bin_cuts[o]:plan.bin_cuts[o + 1] + 1] for o in range(world)])
        off_send_sizes = [plan.bin_cuts[o + 1] - plan.bin_cuts[o] + 1 for o in range(world)]
        recv_off = torch.empty(world * n_mine, dtype=torch.int64, device=records.device)
        # also the barrier "every sender's scatter is complete": a rank's part of this collective is ordered
        # behind its scatter kernel on its stream
        dist.all_to_all_single(recv_off, send_off, [n_mine] * world, off_send_sizes, group=group)
        segments = [((peer.peers[g] + 8 * cuts[g][rank], cuts[g][rank + 1] - cuts[g][rank]), recv_off[g * n_mine:(g + 1) * n_mine])
                    for g in range(world)]
        backend._keepalive = (recv_off,)
    else:
        # slab of owner o = records[bin_off[cut_o] : bin_off[cut_o + 1]]; offsets travel with their slab
        cut_off = bin_off[torch.tensor(plan.bin_cuts, device=bin_off.device)]
        send_counts = (cut_off[1:] - cut_off[:-1])
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts, group=group)
        send_list, recv_list = send_counts.tolist(), recv_counts.tolist()
        first = int(cut_off[0].item())
        recv_records = torch.empty(max(1, sum(recv_list)), dtype=records.dtype, device=records.device)
        dist.all_to_all_single(recv_records[:sum(recv_list)], records[first:first + sum(send_list)], recv_list,
                               send_list, group=group)
        n_mine = hi - lo + 1
        send_off = torch.cat([bin_off[plan.bin_cuts[o]:plan.bin_cuts[o + 1] + 1] for o in range(world)])
        off_send_sizes = [plan.bin_cuts[o + 1] - plan.bin_cuts[o] + 1 for o in range(world)]
        recv_off = torch.empty(world * n_mine, dtype=torch.int64, device=records.device)
        dist.all_to_all_single(recv_off, send_off, [n_mine] * world, off_send_sizes, group=group)
        segments, at = [], 0
        for g in range(world):
            segments.append((recv_records[at:at + max(1, recv_list[g])] if recv_list[g] else recv_records[:1],
                             recv_off[g * n_mine:(g + 1) * n_mine]))
            at += recv_list[g]
        backend._keepalive = (recv_records, recv_off)
    mark("all_to_all")
    table = backend.reduce(segments, lo, hi, plan.aid_cuts[rank], plan.aid_cuts[rank + 1])
    mark("merge+reduce")
    out_stats = dict(backend.stats())
    out_stats["owned_aids"] = plan.aid_cuts[rank + 1] - plan.aid_cuts[rank]
    out_stats["sent_records"] = int(stats["pairs"])
    return table, (plan.aid_cuts[rank], plan.aid_cuts[rank + 1]), out_stats, plan


def phase_ms_from_events(timing: dict) -> dict:
    """{phase: ms} from the CUDA events a build recorded with timing = {"__events__": True} (call after a synchronise)."""
    marks = timing.get("__marks__", [])
    out = {}
    for (_, a), (name, b) in zip(marks, marks[1:]):
        if name != "start":
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
    return out


def gather_table(table: TopKTable, plan: OwnerPlan, group=None) -> TopKTable:
    """Every rank ends up with all rows (one broadcast per owner range); needed before candidate generation."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return table
    world = dist.get_world_size(group)
    for o in range(world):
        lo, hi = plan.aid_cuts[o], plan.aid_cuts[o + 1]
        if hi > lo:
            src = dist.get_global_rank(group, o) if group is not None else o
            for t in (table.aid_y, table.wgt, table.len):
                dist.broadcast(t[lo:hi], src=src, group=group)
    return table
