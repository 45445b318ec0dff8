"""Staged scatter (sender-side combining, include/otto_covisit.h) == single-GPU build, byte for byte.

The protocol is driven here with G simulated ranks on ONE device (every rank's builder, staging buffer and record buffer
live in the same HBM; "peer" pointers are plain device pointers), so the bucket plan, the packed records, the place
pass at the owner, owner cuts through buckets and hot rows are covered on a 1-GPU box.  The NCCL / CUDA-IPC form of the
same calls runs in tests/test_distributed_gpu.py (>= 2 GPUs)."""
from dataclasses import replace

import pytest
import torch

pytestmark = pytest.mark.gpu


def _staged_build(csr, spec, world, dev):
    from otto_multi_objective_recommender_system_b200 import distributed
    S = csr.n_sessions
    spec = replace(spec, global_events=csr.n_events)
    ranks = [distributed.GpuRankBackend(csr.slice_sessions(g * S // world, (g + 1) * S // world), spec, exact=True)
             for g in range(world)]
    counts = torch.stack([r.count_begin().clone() for r in ranks])            # what the all-gather delivers: [G, A]
    cuts = None
    for g, r in enumerate(ranks):
        aid_cuts, before = r.plan_owners(counts, world, g)
        assert cuts is None or aid_cuts == cuts                                # every rank finds the same cuts
        cuts = aid_cuts
        r.count_finish_owned(aid_cuts, g, before)
    staged = []
    for g, r in enumerate(ranks):
        if g % 2:                                                              # both forms of the plan call
            assert r.b.stage_plan(counts, world, g, sync=False) is None
            totals = r.b.stage_totals(world)
        else:
            totals = r.b.stage_plan(counts, world, g)
        assert sum(totals) >= int(counts.to(torch.int64).bitwise_and(0xFFFFFFFF).sum())      # segments are padded to even
        staged.append(torch.empty(max(totals[g], 1), dtype=torch.int64, device=dev))
        r.b.scatter_staged(world, staged[g].data_ptr())
    tables, stats = [], []
    for g, r in enumerate(ranks):
        r.b.place_staged([t.data_ptr() for t in staged], cuts[g], cuts[g + 1])
        records, bin_off = r.partition()
        bin_cuts = r.owner_bin_cuts(cuts, None)
        lo, hi = bin_cuts[g], bin_cuts[g + 1]
        tables.append(r.reduce([(records, bin_off[lo:hi + 1])], lo, hi, cuts[g], cuts[g + 1]))
        stats.append(r.stats())
    return tables, cuts, stats


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("variant,split_ub", [("CLICKS", 0), ("CLICKS", 96), ("CARTS_ORDERS", 64), ("BUY2BUY", 0)])
def test_staged_scatter_equals_single_build(native_lib, variant, split_ub, world):
    from otto_multi_objective_recommender_system_b200 import covisit, synth
    dev = torch.device("cuda", 0)
    spec = replace(getattr(covisit, variant), split_ub=split_ub)
    frame = synth.generate(synth.SynthSpec("train", 20000, 2500, seed=31), device=dev)
    csr = covisit.ingest(frame, "desc", device=dev)
    single, sstats = covisit.build_topk(csr, spec, exact=True)
    tables, cuts, stats = _staged_build(csr, spec, world, dev)
    for g, t in enumerate(tables):
        lo, hi = cuts[g], cuts[g + 1]
        for name in ("aid_y", "wgt", "len", "cnt", "tsum"):
            assert torch.equal(getattr(t, name)[lo:hi], getattr(single, name)[lo:hi]), (world, g, name)
    assert sum(s["pair_checksum"] for s in stats) == sstats["pair_checksum"]
    assert sum(s["distinct"] for s in stats) == sstats["distinct"]


def test_staged_scatter_tiny_frame(native_lib):
    """A dozen sessions over four ranks: buckets, tiles and owner ranges of a handful of records."""
    from otto_multi_objective_recommender_system_b200 import covisit, synth
    dev = torch.device("cuda", 0)
    frame = synth.generate(synth.SynthSpec("train", 12, 40, seed=3), device=dev)
    csr = covisit.ingest(frame, "desc", device=dev)
    single, sstats = covisit.build_topk(csr, covisit.CLICKS, exact=True)
    tables, cuts, stats = _staged_build(csr, covisit.CLICKS, 4, dev)
    for g, t in enumerate(tables):
        lo, hi = cuts[g], cuts[g + 1]
        for name in ("aid_y", "wgt", "len"):
            assert torch.equal(getattr(t, name)[lo:hi], getattr(single, name)[lo:hi]), (g, name)
    assert sum(s["distinct"] for s in stats) == sstats["distinct"]


def test_staged_scatter_argument_errors(native_lib):
    """The error behaviour the header promises: -1 for an impossible rank count, OTTO_ENOSPC for a plan scratch that is too
    small (nothing launched), OTTO_EINVAL for a rank outside the box."""
    import ctypes as C
    from otto_multi_objective_recommender_system_b200 import _native as N, covisit, synth
    dev = torch.device("cuda", 0)
    frame = synth.generate(synth.SynthSpec("train", 200, 60, seed=5), device=dev)
    csr = covisit.ingest(frame, "desc", device=dev)
    b = covisit.CovisitBuilder(csr, covisit.CLICKS)
    b.count_begin()
    b.count_finish()
    lib = N.lib()
    assert lib.otto_covisit_stage_plan_bytes(C.byref(b.cspec), csr.n_sessions, csr.n_events, 0) == -1
    assert lib.otto_covisit_stage_plan_bytes(C.byref(b.cspec), csr.n_sessions, csr.n_events, N.MAX_OWNERS + 1) == -1
    need = lib.otto_covisit_stage_plan_bytes(C.byref(b.cspec), csr.n_sessions, csr.n_events, 1)
    small = torch.empty(need - 1, dtype=torch.uint8, device=dev)
    totals = (C.c_int64 * N.MAX_OWNERS)()
    args = (C.byref(b.ev), C.byref(b.cspec), b.workspace.data_ptr(), b.workspace.numel(), None)
    assert lib.otto_covisit_stage_plan(*args, 1, 0, small.data_ptr(), small.numel(), totals, b._st()) == N.OTTO_ENOSPC
    full = torch.empty(need, dtype=torch.uint8, device=dev)
    assert lib.otto_covisit_stage_plan(*args, 1, 1, full.data_ptr(), need, totals, b._st()) == N.OTTO_EINVAL      # rank >= n_ranks
    assert lib.otto_covisit_stage_plan(*args, 2, 0, full.data_ptr(), need, totals, b._st()) in (N.OTTO_EINVAL, N.OTTO_ENOSPC)  # counts_all NULL with 2 ranks
    assert lib.otto_covisit_stage_plan(*args, 1, 0, full.data_ptr(), need, totals, b._st()) == N.OTTO_OK
    assert totals[0] >= b.stats.pairs and totals[0] - b.stats.pairs <= 2 * ((csr.n_aids >> 11) + 1)                # even-padded segments
