"""World-size-2 gloo tests of the multi-GPU exchange logic (owner planning, slab all-to-all, segment
assembly) with a numpy stand-in for the rank-local CUDA phases.  The stand-in follows the same interface as
distributed.GpuRankBackend and forms the same bins (row totals, split_ub, aid_y-hash slices)."""
import os
import socket

import numpy as np
import pandas as pd
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import covisit_oracle as co


def _hash32(x):
    x = x.astype(np.uint64)
    m = np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16); x = (x * np.uint64(0x7feb352d)) & m
    x ^= x >> np.uint64(15); x = (x * np.uint64(0x846ca68b)) & m
    x ^= x >> np.uint64(16)
    return x


def _sub_bin(y, nb):
    return (_hash32(y) * nb.astype(np.uint64)) >> np.uint64(32)


class NumpyRankBackend:
    """Same phases as the CUDA backend, computed with pandas/numpy from the oracle's deduplicated pairs."""

    def __init__(self, df, spec, n_aids, split_ub):
        self.df, self.spec, self.n_aids, self.split_ub = df, spec, n_aids, split_ub

    def count_begin(self):
        # local pairs per aid_x row (after the in-session dedupe); the driver all-reduces this tensor in place
        self.pairs = co.dedup_pairs(self.df, self.spec)
        self.local = np.bincount(self.pairs["aid_x"].to_numpy(), minlength=self.n_aids).astype(np.int64)
        self.total = torch.from_numpy(self.local.astype(np.int32).copy())
        return self.total

    def count_finish(self):
        tot = self.total.numpy().astype(np.int64)
        target = max(1, self.split_ub // 2)
        self.nb = np.where(tot > self.split_ub, np.maximum(1, -(-tot // target)), 1).astype(np.int64)
        self.bin_base = np.concatenate([[0], np.cumsum(self.nb)]).astype(np.int64)
        self.bin_x = np.repeat(np.arange(self.n_aids), self.nb)
        stats = {"bins": int(self.bin_base[-1]), "pairs": len(self.pairs)}
        return stats, torch.from_numpy(self.bin_base.astype(np.int32))

    def scatter(self):
        x, y = self.pairs["aid_x"].to_numpy().astype(np.int64), self.pairs["aid_y"].to_numpy().astype(np.int64)
        nbx = self.nb[x]
        b = self.bin_base[x] + np.where(nbx > 1, _sub_bin(y, nbx).astype(np.int64), 0)
        order = np.argsort(b, kind="stable")
        self.rec = (y[order] | (np.ones_like(y) << 32)).astype(np.int64)       # v = 1 (unit weights)
        B = int(self.bin_base[-1])
        self.bin_off = np.concatenate([[0], np.cumsum(np.bincount(b, minlength=B))]).astype(np.int64)
        return torch.from_numpy(self.rec), torch.from_numpy(self.bin_off)

    def reduce(self, segments, bin_lo, bin_hi, aid_lo, aid_hi):
        rows = []
        for rec, off in segments:
            rec, off = rec.numpy(), off.numpy()
            off = off - off[0]
            for i, b in enumerate(range(bin_lo, bin_hi)):
                r = rec[off[i]:off[i + 1]]
                rows.append(pd.DataFrame({"aid_x": self.bin_x[b], "aid_y": r & 0xFFFFFFFF, "wgt": (r >> 32).astype(np.float32)}))
        acc = pd.concat(rows).groupby(["aid_x", "aid_y"])["wgt"].sum().reset_index() if rows else pd.DataFrame(columns=["aid_x", "aid_y", "wgt"])
        return co.topk(acc, self.spec.k)

    def stats(self):
        return {}


class NumpyOwnerBackend(NumpyRankBackend):
    """Stand-in for the owner-direct scatter: "storing into the owner's buffer" is an object all-gather of
    (position, record) lists that every owner applies to its own array."""
    owner_direct = True

    def plan_owners(self, gathered, world, rank):
        from otto_multi_objective_recommender_system_b200 import distributed
        cuts, total, before = distributed.plan_owners_host(gathered, world, rank)
        self.total.copy_(total.to(torch.int32))                     # the stand-in's "workspace row_total"
        return cuts, before

    def owner_bin_cuts(self, aid_cuts, bin_base):
        return [int(bin_base[x]) for x in aid_cuts]

    def count_finish_owned(self, aid_cuts, rank, row_before):
        stats, bin_base = self.count_finish()                       # bins from the totals (self.total was summed in place)
        tot = self.total.numpy().astype(np.int64)
        cuts = np.asarray(aid_cuts)
        self.owner = np.searchsorted(cuts, np.arange(self.n_aids), side="right") - 1
        self.owner = np.minimum(self.owner, len(aid_cuts) - 2)
        hot = self.nb > 1
        E = np.concatenate([[0], np.cumsum(tot)])
        Hs = np.concatenate([[0], np.cumsum(np.where(hot, tot, 0))])
        P = E[cuts[1:]] - E[cuts[:-1]]
        H = Hs[cuts[1:]] - Hs[cuts[:-1]]
        o = self.owner
        pos = np.where(hot, P[o] + Hs[:-1] - Hs[cuts[o]], E[:-1] - E[cuts[o]])
        self.cursor = pos + row_before.numpy().astype(np.int64)
        self.rank, self.P_mine, self.H_mine = rank, int(P[rank]), int(H[rank])
        mine = o == rank
        self.lay = np.where(mine, tot, 0)
        self.lay_off = np.concatenate([[0], np.cumsum(self.lay)])
        self.lay_hot = np.concatenate([[0], np.cumsum(np.where(mine & hot, tot, 0))])
        stats["pairs"], stats["hot_pairs"] = self.P_mine, self.H_mine
        return stats, bin_base

    def scatter_owned(self, aid_cuts, rank):
        x, y = self.pairs["aid_x"].to_numpy().astype(np.int64), self.pairs["aid_y"].to_numpy().astype(np.int64)
        order = np.argsort(x, kind="stable")
        x, y = x[order], y[order]
        first = np.concatenate([[0], np.cumsum(np.bincount(x, minlength=self.n_aids))])
        at = self.cursor[x] + (np.arange(len(x)) - first[x])        # my run of a row is contiguous behind the lower ranks'
        rec = (y | (np.ones_like(y) << 32)).astype(np.int64)
        world = len(aid_cuts) - 1
        outbox = [(at[self.owner[x] == o], rec[self.owner[x] == o]) for o in range(world)]
        inbox = [None] * world
        dist.all_gather_object(inbox, outbox)
        self.buf = np.full(self.P_mine + self.H_mine, -1, dtype=np.int64)
        for sender in inbox:
            pos, r = sender[rank]
            assert (self.buf[pos] == -1).all(), "two ranks wrote the same slot"
            self.buf[pos] = r
        # the only empty slots are the final positions of the hot rows (filled from the staging area by partition)
        assert int((self.buf == -1).sum()) == self.H_mine, "the owners' layout has holes"

    def partition(self):
        B = int(self.bin_base[-1])
        cnt = np.zeros(B, dtype=np.int64)
        parts = {}
        for xrow in np.nonzero(self.lay)[0]:
            n, b0, nb = self.lay[xrow], self.bin_base[xrow], self.nb[xrow]
            if nb == 1:
                cnt[b0] = n
                parts[b0] = self.buf[self.lay_off[xrow]:self.lay_off[xrow] + n]
            else:
                st = self.P_mine + self.lay_hot[xrow]
                r = self.buf[st:st + n]
                sub = _sub_bin(r & 0xFFFFFFFF, np.full(n, nb)).astype(np.int64)
                for j in range(nb):
                    parts[b0 + j] = r[sub == j]
                    cnt[b0 + j] = len(parts[b0 + j])
        self.bin_off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
        self.rec = np.concatenate([parts[b] for b in sorted(parts)]) if parts else np.zeros(0, dtype=np.int64)
        assert len(self.rec) == self.P_mine
        return torch.from_numpy(self.rec), torch.from_numpy(self.bin_off)


class NumpyStagedBackend(NumpyOwnerBackend):
    """Stand-in for the staged scatter (include/otto_covisit.h, "Staged scatter"): pass A appends my pairs to coarse
    aid_x buckets of my own staging array (segments start on even slots, like the library's), "peers map each other's
    staging buffers" is an object all-gather, pass B walks the segments of my buckets in every rank's staging array and
    places the records with the cursors of my own layout."""
    staged = True
    LOGA = 3                                                        # 8 rows per bucket (the library: 2048)

    def stage_plan(self, gathered, world, rank):
        counts = (gathered.numpy().astype(np.int64) & 0xFFFFFFFF)   # [G, A]
        nb = -(-self.n_aids // (1 << self.LOGA))
        pad = np.zeros((world, nb << self.LOGA), dtype=np.int64)
        pad[:, :self.n_aids] = counts
        self.bcnt = pad.reshape(world, nb, 1 << self.LOGA).sum(2)
        even = (self.bcnt + 1) & ~1
        self.bo = np.concatenate([np.zeros((world, 1), dtype=np.int64), np.cumsum(even, 1)], 1)

    def scatter_staged(self, world):
        x, y = self.pairs["aid_x"].to_numpy().astype(np.int64), self.pairs["aid_y"].to_numpy().astype(np.int64)
        me = dist.get_rank()
        assert len(x) == int(self.bcnt[me].sum())
        self.staging = np.full(int(self.bo[me, -1]), -1, dtype=np.int64)
        cur = self.bo[me, :-1].copy()
        for xi, yi in zip(x, y):                                    # runs in any order: the place pass must not care
            b = xi >> self.LOGA
            self.staging[cur[b]] = yi | (1 << 32) | ((xi & ((1 << self.LOGA) - 1)) << 40)
            cur[b] += 1

    def place_staged(self, aid_cuts, rank):
        world = len(aid_cuts) - 1
        every = [None] * world
        dist.all_gather_object(every, self.staging)
        lo, hi = aid_cuts[rank], aid_cuts[rank + 1]
        self.buf = np.full(self.P_mine + self.H_mine, -1, dtype=np.int64)
        hot = self.nb > 1
        cursor = np.where(hot, self.P_mine + self.lay_hot[:-1], self.lay_off[:-1])
        if hi > lo:
            for b in range(lo >> self.LOGA, ((hi - 1) >> self.LOGA) + 1):
                for g in range(world):
                    seg = every[g][self.bo[g, b]:self.bo[g, b] + self.bcnt[g, b]]
                    assert (seg >= 0).all(), "a staged segment has holes"
                    for r in seg:
                        xrow = (b << self.LOGA) + (int(r) >> 40)
                        if lo <= xrow < hi:                             # a bucket across an owner cut is read by both owners
                            self.buf[cursor[xrow]] = int(r) & ((1 << 40) - 1)
                            cursor[xrow] += 1
        assert int((self.buf == -1).sum()) == self.H_mine, "the owners' layout has holes"


def _worker(rank, world, port, split_ub, out, owner_direct=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from otto_multi_objective_recommender_system_b200 import distributed, synth
    frame = synth.generate(synth.SynthSpec("train", 600, 80, seed=17))
    df = frame.to_pandas()
    sessions = np.sort(df["session"].unique())
    mine = sessions[rank * len(sessions) // world:(rank + 1) * len(sessions) // world]     # contiguous session chunk
    spec = co.OracleSpec(co.WEIGHT_UNIT, k=7)
    kind = {False: NumpyRankBackend, True: NumpyOwnerBackend, "staged": NumpyStagedBackend}[owner_direct]
    backend = kind(df.loc[df["session"].isin(mine)], spec, 80, split_ub)
    table, (lo, hi), stats, plan = distributed.build_topk_distributed(backend)
    assert table["aid_x"].between(lo, hi - 1).all()
    gathered = [None] * world
    dist.all_gather_object(gathered, (table, plan.aid_cuts))
    if rank == 0:
        got = pd.concat([g[0] for g in gathered], ignore_index=True)
        want = co.build(df, spec)
        out.put((got.astype({"aid_x": np.int32, "aid_y": np.int32}).to_dict("list"), want.to_dict("list"),
                 [g[1] for g in gathered]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,split_ub,owner_direct", [(2, 1 << 30, False), (2, 40, False), (2, 1 << 30, True), (2, 40, True),
                                                        (3, 40, True), (2, 1 << 30, "staged"), (2, 40, "staged"),
                                                        (3, 40, "staged")])
def test_multi_rank_exchange_equals_single_process(world, split_ub, owner_direct):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, split_ub, out, owner_direct)) for r in range(world)]
    for p in procs:
        p.start()
    got, want, cuts = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(c == cuts[0] for c in cuts), "every rank must derive the same ownership plan"
    assert got["aid_x"] == want["aid_x"] and got["aid_y"] == want["aid_y"] and got["wgt"] == want["wgt"]


def test_plan_rows_balances_skewed_rows():
    from otto_multi_objective_recommender_system_b200 import distributed
    rows = torch.tensor([1000, 1, 1, 1, 500, 500, 1, 1], dtype=torch.int64)
    assert distributed.plan_rows(rows, 1) == [0, 8]
    cuts = distributed.plan_rows(rows, 2)
    assert cuts[0] == 0 and cuts[-1] == 8 and cuts == sorted(cuts)
    left = int(rows[:cuts[1]].sum())
    assert abs(left - 1002) <= 1000                                  # a cut never splits a row
    c8 = distributed.plan_rows(rows, 8)
    assert len(c8) == 9 and c8 == sorted(c8) and c8[0] == 0 and c8[-1] == 8
    assert distributed.plan_rows(torch.zeros(5, dtype=torch.int64), 3) == [0, 0, 0, 5]


def test_plan_owners_balances_skewed_rows():
    from otto_multi_objective_recommender_system_b200 import distributed
    counts = torch.tensor([1000, 1, 1, 1, 500, 500, 1, 1], dtype=torch.int64)      # one hot row split in 3 bins
    bin_base = torch.tensor([0, 3, 4, 5, 6, 7, 8], dtype=torch.int32)              # aid 0 owns bins 0..2
    plan = distributed.plan_owners(counts, bin_base, 2)
    assert plan.aid_cuts[0] == 0 and plan.aid_cuts[-1] == 6 and plan.aid_cuts == sorted(plan.aid_cuts)
    # cuts fall on aid boundaries, never inside a split row
    assert plan.bin_cuts == [int(bin_base[a]) for a in plan.aid_cuts]
    assert distributed.plan_owners(counts, bin_base, 1).aid_cuts == [0, 6]
    p8 = distributed.plan_owners(counts, bin_base, 8)
    assert len(p8.aid_cuts) == 9 and p8.aid_cuts == sorted(p8.aid_cuts)
