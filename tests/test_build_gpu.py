"""GPU parity tests of the covisitation build: CUDA path (through the C ABI) vs the pandas oracle."""
import json
import pathlib

import numpy as np
import pandas as pd
import pytest
import torch

from oracle import covisit_oracle as co
import parity_helpers as H

pytestmark = pytest.mark.gpu
GOLDEN = pathlib.Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def cv(native_lib):
    from otto_multi_objective_recommender_system_b200 import covisit
    return covisit


def synth_frame(n_sessions, n_aids, seed=42, kind="train"):
    from otto_multi_objective_recommender_system_b200 import synth
    return synth.generate(synth.SynthSpec(kind, n_sessions, n_aids, seed=seed))


def frame_from_df(df, n_aids=None):
    from otto_multi_objective_recommender_system_b200.synth import EventFrame
    return EventFrame.from_pandas(df, n_aids)


def run_build(cv, frame, spec, exact=True):
    csr = cv.ingest(frame, "desc", device="cuda:0")
    table, stats = cv.build_topk(csr, spec, exact=exact)
    got = table.to_pandas()
    if exact:
        ax = got["aid_x"].to_numpy()
        # exact integer side outputs, same row order as to_rows()
        mask = torch.arange(table.k, device="cuda:0")[None, :] < table.len[:, None]
        got["cnt"] = table.cnt[mask].cpu().numpy().astype(np.int64) & 0xFFFFFFFF
        got["tsum"] = table.tsum[mask].cpu().numpy()
        assert len(ax) == int(mask.sum())
    return got, stats, table


def check_against_oracle(cv, frame, spec, what):
    df = frame.to_pandas()
    ospec = H.oracle_spec(spec)
    got, stats, table = run_build(cv, frame, spec)
    acc = co.accumulate(df, ospec, exact=True)
    assert stats["pairs"] == int(acc["cnt"].sum()), f"{what}: pair count P"
    assert stats["distinct"] == len(acc), f"{what}: distinct count D"
    assert stats["table_overflow"] == 0
    if spec.weight_mode == cv.N.WEIGHT_TIME:
        # bit-exact on the integer accumulators and on the GPU's own weight definition ...
        want = H.gpu_formula_topk(acc, ospec)
        H.assert_int_table_equal(got, want, what + " (integer form)")
        assert np.array_equal(got["cnt"].to_numpy(), want["cnt"].to_numpy()), what
        assert np.array_equal(got["tsum"].to_numpy(), want["tsum"].to_numpy()), what
        # ... and within 1e-5 of pandas' float32 running sums, top-K identical up to near-ties
        H.assert_time_table_close(got, acc, spec.k, what)
    else:
        H.assert_int_table_equal(got, co.topk(acc, spec.k), what)
    # table invariants
    ln = table.len.cpu().numpy()
    ay = table.aid_y.cpu().numpy()
    assert ((ay >= 0).sum(axis=1) == ln).all(), f"{what}: padding"
    return got, stats


VARIANTS = ["CLICKS", "CARTS_ORDERS", "BUY2BUY"]


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("shape", [(64, 40, 1), (3000, 500, 7), (20000, 3000, 42)])
def test_build_matches_oracle(cv, variant, shape):
    frame = synth_frame(shape[0], shape[1], seed=shape[2])
    check_against_oracle(cv, frame, getattr(cv, variant), f"{variant} {shape}")


@pytest.mark.parametrize("variant", VARIANTS)
def test_build_matches_committed_golden_vectors(cv, variant):
    g = json.load(open(GOLDEN / "oracle_small.json"))
    df = pd.DataFrame(g["frame"]).astype({"session": np.int32, "aid": np.int32, "ts": np.int32, "type": np.int8})
    spec = getattr(cv, variant)
    got, _, _ = run_build(cv, frame_from_df(df, 120), spec, exact=False)
    w = g["tables"][{"CLICKS": "clicks", "CARTS_ORDERS": "carts_orders", "BUY2BUY": "buy2buy"}[variant]]
    assert got["aid_x"].tolist() == w["aid_x"]
    if variant == "CLICKS":
        np.testing.assert_allclose(got["wgt"].to_numpy(), np.array(w["wgt"], np.float32), rtol=H.RTOL_TIME)
    else:
        assert got["aid_y"].tolist() == w["aid_y"]
        assert got["wgt"].tolist() == w["wgt"]


def test_session_747_fixture(cv):
    ev = json.load(open(GOLDEN / "session_747.json"))["events"]
    df = pd.DataFrame({"session": 747, "aid": [e["aid"] for e in ev], "ts": [e["ts"] for e in ev],
                       "type": [e["type"] for e in ev]})
    # remap the 7-digit aids to a dense range so the table stays small
    ids = {a: i for i, a in enumerate(sorted(df["aid"].unique()))}
    df["aid"] = df["aid"].map(ids)
    frame = frame_from_df(df, len(ids))
    for variant in VARIANTS:
        got, stats = check_against_oracle(cv, frame, getattr(cv, variant), f"747 {variant}")
    got, stats, _ = run_build(cv, frame, cv.CLICKS)
    assert stats["pairs"] == 204          # hand-derived in tests/test_oracle.py
    got, stats, _ = run_build(cv, frame, cv.BUY2BUY)
    assert stats["pairs"] == 6 and (got["wgt"] == 1.0).all()


@pytest.mark.parametrize("split_ub", [16, 64, 1000])
@pytest.mark.parametrize("variant", VARIANTS)
def test_split_rows_and_merge(cv, variant, split_ub):
    """Small split_ub forces hot rows into aid_y-hash sub-bins and through the merge kernel."""
    from dataclasses import replace
    frame = synth_frame(4000, 300, seed=11)
    spec = replace(getattr(cv, variant), split_ub=split_ub)
    got, stats = check_against_oracle(cv, frame, spec, f"{variant} split_ub={split_ub}")
    if split_ub <= 64 and variant != "BUY2BUY":
        assert stats["split_rows"] > 0


@pytest.mark.parametrize("variant", VARIANTS)
def test_hot_rows_large_bins_multipass(cv, variant):
    """Few aids and many sessions: every row holds tens of thousands of records, which exercises the
    block kernels and the multi-pass path (no split: split_ub is huge)."""
    from dataclasses import replace
    frame = synth_frame(30000, 24, seed=5)
    spec = replace(getattr(cv, variant), split_ub=1 << 30)
    check_against_oracle(cv, frame, spec, f"{variant} hot rows")


@pytest.mark.parametrize("variant", ["CLICKS", "CARTS_ORDERS"])
def test_tied_entries_hand_over_to_the_hash_table_kernel(cv, variant):
    """More than 64 entries tie around the K-th weight of a row (one aid_x seen with hundreds of partners, every
    partner once, every event at one timestamp): the owner-table tiers cannot hold the candidates and hand the bin
    to the hash-table kernel, whose exact K-round selection breaks the ties by aid_y."""
    rows, ts = [], 1660000000
    partner = 10
    for s in range(4):            # aid 0: 4 x 29 = 116 partners -> warp tier (<= 384 records)
        rows += [(s, 0, ts, 0)] + [(s, partner + j, ts, 0) for j in range(29)]
        partner += 29
    for s in range(4, 30):        # aid 1: 26 x 29 = 754 partners -> 128-thread tier
        rows += [(s, 1, ts, 0)] + [(s, partner + j, ts, 0) for j in range(29)]
        partner += 29
    df = pd.DataFrame(rows, columns=["session", "aid", "ts", "type"])
    frame = frame_from_df(df, partner + 5)
    got, stats = check_against_oracle(cv, frame, getattr(cv, variant), f"{variant} tied entries")
    if variant == "CLICKS":
        # every bin is far below 6144 records, so records counted by the last tier were handed over (type / unit
        # weights carry aid_y bits in the 32-bit selection key: their ties never overflow the candidate list)
        assert stats["tier_records"][3] >= 116 + 754, stats["tier_records"]
    top0 = got[got["aid_x"] == 0]["aid_y"].tolist()
    assert top0 == list(range(10, 10 + getattr(cv, variant).k)), top0


@pytest.mark.parametrize("variant", ["CLICKS", "CARTS_ORDERS"])
def test_bins_at_the_tier_bounds(cv, variant):
    """Rows of exactly 31 .. 6145 records around every bound of the reduce tiers (32: no table; 384: one warp; 1536 /
    3072 / 6144: 128 / 256 / 512 threads; 6145: split into sub-bins) - once with all-distinct partners (a table at its
    three-quarters fill) and once with 50 partners (every owner followed by ~n / 50 records)."""
    sizes = [31, 32, 33, 383, 384, 385, 1535, 1536, 1537, 3071, 3072, 3073, 6143, 6144, 6145]
    rng = np.random.default_rng(17)
    rows, s, ts = [], 0, 1660000000
    first_partner = 100
    for i, n in enumerate(sizes):
        for x, pool in ((i, n), (50 + i, 50)):           # aid x: n two-event sessions [x, partner]
            for j in range(n):
                y = first_partner + (j if pool == n else int(rng.integers(0, pool)))
                ty = rng.choice([0, 1, 2], size=2, p=[0.6, 0.25, 0.15])
                rows += [(s, x, ts, int(ty[0])), (s, y, ts + 1, int(ty[1]))]
                s += 1
                ts += int(rng.integers(1, 60))
    df = pd.DataFrame(rows, columns=["session", "aid", "ts", "type"])
    frame = frame_from_df(df, first_partner + max(sizes) + 5)
    got, stats = check_against_oracle(cv, frame, getattr(cv, variant), f"{variant} tier bounds")
    assert stats["split_rows"] == 2 and all(r > 0 for r in stats["tier_records"][:3])


def test_accumulator_width_guard(cv):
    """Bins beyond the owner-table tiers go to the hash-table kernel (24-bit counts, 40-bit time sums): a layout in which
    an entry could exceed them is refused instead of wrapping silently."""
    from dataclasses import replace
    from otto_multi_objective_recommender_system_b200 import _native as N
    frame = synth_frame(70000, 24, seed=5)
    csr = cv.ingest(frame, "desc", device="cuda:0")
    # 70,000 sessions over 24 aids: rows of about a million pairs, unsplit; the real time range passes ...
    cv.build_topk(csr, replace(cv.CLICKS, split_ub=1 << 30))
    # ... a range of 2^24 - 1 seconds times 70,000 possible sessions per entry does not fit 40 bits
    wide = replace(cv.CLICKS, split_ub=1 << 30, ts_min=cv.CLICKS.ts_max - (1 << 24) + 1)
    with pytest.raises(N.OttoError, match="40 bits"):
        cv.build_topk(csr, wide)
    # with the default split_ub no bin leaves the owner-table tiers and the same range is fine
    cv.build_topk(csr, replace(wide, split_ub=0))


def test_edge_sessions(cv):
    """Lengths 1, 2, 30, 31, 32, 33, 500; ts ties across the tail boundary; window edge |dt| == W."""
    rng = np.random.default_rng(0)
    rows = []
    sid = 0
    for length in [1, 2, 30, 31, 32, 33, 64, 500, 1, 29, 30, 31]:
        ts = np.sort(rng.integers(1659304800, 1659304800 + 3 * 86400, size=length))
        if length >= 31:
            ts[-31:-28] = ts[-30]        # a tie run straddling the 30-event cut
        aid = rng.integers(0, 50, size=length)
        typ = rng.choice([0, 1, 2], size=length, p=[0.6, 0.25, 0.15])
        rows += [(sid, a, t, y) for a, t, y in zip(aid, ts, typ)]
        sid += 1
    # window edge: two events exactly 86400 s apart (out) and 86399 s apart (in)
    rows += [(sid, 1, 1659400000, 0), (sid, 2, 1659400000 + 86400, 0)]
    rows += [(sid + 1, 3, 1659400000, 1), (sid + 1, 4, 1659400000 + 86399, 2)]
    # all events share one ts and one aid appears many times
    rows += [(sid + 2, a, 1659500000, 1) for a in [5, 6, 5, 7, 5, 6, 8] * 6]
    df = pd.DataFrame(rows, columns=["session", "aid", "ts", "type"])
    frame = frame_from_df(df, 50)
    for variant in VARIANTS:
        check_against_oracle(cv, frame, getattr(cv, variant), f"edge {variant}")


def test_generic_spec_type_masks(cv):
    """x_types / y_types pair filters of the generic CovisitSpec (unpinned stems such as click_cart)."""
    from dataclasses import replace
    frame = synth_frame(3000, 200, seed=21)
    spec = replace(cv.CARTS_ORDERS, x_types=(0,), y_types=(1, 2), k=10)
    check_against_oracle(cv, frame, spec, "x=clicks y=carts/orders")
    spec = replace(cv.CLICKS, event_types=(0, 1), tail_n=12, window_s=3600, k=7)
    check_against_oracle(cv, frame, spec, "clicks+carts tail 12 1h")


def test_unpinned_stem_recipes_build_like_the_oracle(cv):
    """configs/unpinned_stems.example.json: the four stems without a known recipe go through the same kernels."""
    recipes = cv.load_stem_recipes(pathlib.Path(__file__).resolve().parents[1] / "configs" / "unpinned_stems.example.json")
    assert sorted(recipes) == ["click_cart", "click_order", "click_weighted", "order_weighted"]
    frame = synth_frame(3000, 200, seed=23)
    for stem, spec in recipes.items():
        check_against_oracle(cv, frame, spec, f"recipe {stem}")


def test_empty_and_tiny_inputs(cv):
    df = pd.DataFrame({"session": [5], "aid": [3], "ts": [1659304800], "type": [0]})
    got, stats, table = run_build(cv, frame_from_df(df, 8), cv.CLICKS)
    assert len(got) == 0 and stats["pairs"] == 0 and int(table.len.sum()) == 0
    # buy2buy on a frame without carts/orders: every tail is empty
    df = pd.DataFrame({"session": [1, 1, 2, 2], "aid": [1, 2, 3, 4], "ts": [1659304800 + i for i in range(4)], "type": [0] * 4})
    got, stats, _ = run_build(cv, frame_from_df(df, 8), cv.BUY2BUY)
    assert len(got) == 0 and stats["tail_events"] == 0


def test_ingest_desc_order_matches_stable_sort(cv):
    frame = synth_frame(5000, 400, seed=9)
    df = frame.to_pandas()
    want = df.sort_values(["session", "ts"], ascending=[True, False], kind="stable").reset_index(drop=True)
    csr = cv.ingest(frame, "desc", device="cuda:0")
    assert np.array_equal(csr.aid.cpu().numpy(), want["aid"].to_numpy())
    assert np.array_equal(csr.ts.cpu().numpy(), want["ts"].to_numpy())
    assert np.array_equal(csr.type.cpu().numpy(), want["type"].to_numpy().astype(np.uint8))
    # shuffled input is sorted first (stable by (session, ts)): same CSR as sorting on the host
    perm = np.random.default_rng(1).permutation(len(df))
    sh = df.iloc[perm].reset_index(drop=True)
    want2 = sh.sort_values(["session", "ts"], kind="stable").sort_values(["session", "ts"], ascending=[True, False], kind="stable")
    csr2 = cv.ingest(frame_from_df(sh, 400), "desc", device="cuda:0")
    assert np.array_equal(csr2.aid.cpu().numpy(), want2["aid"].to_numpy())


def _tables_equal(a, b):
    return all(torch.equal(x, y) for x, y in ((a.aid_y, b.aid_y), (a.wgt.view(torch.int32), b.wgt.view(torch.int32)), (a.len, b.len)))


@pytest.mark.parametrize("variant", VARIANTS)
def test_ascending_and_zero_copy_frames_build_the_same_table(cv, variant):
    """otto_covisit_count_begin_asc: a CSR in file order (ts ascending) gives the table of the most-recent-first CSR -
    the tail kernels apply the stable ts-descending sort while they copy - with aid / type on the device or left in
    pinned host memory (ingest(..., zero_copy=True), the end-to-end path of bench.py)."""
    from otto_multi_objective_recommender_system_b200 import synth
    spec = getattr(cv, variant)
    frame = synth_frame(6000, 700, seed=13)
    want, wstats = cv.build_topk(cv.ingest(frame, "desc", device="cuda:0"), spec, exact=True)
    got, gstats = cv.build_topk(cv.ingest(frame, "asc", device="cuda:0"), spec, exact=True)
    assert _tables_equal(got, want) and torch.equal(got.cnt, want.cnt) and torch.equal(got.tsum, want.tsum)
    assert gstats["pairs"] == wstats["pairs"] and gstats["tail_events"] == wstats["tail_events"]
    pinned = synth.EventFrame(*(t.cpu().contiguous().pin_memory() for t in (frame.session, frame.aid, frame.ts, frame.type)),
                              n_aids=frame.n_aids)
    csr = cv.ingest(pinned, "asc", device="cuda:0", zero_copy=True)
    assert not csr.aid.is_cuda and not csr.type.is_cuda and csr.ts.is_cuda
    got, gstats = cv.build_topk(csr, spec, exact=True)
    assert _tables_equal(got, want) and gstats["pairs"] == wstats["pairs"]
    # zero-copy columns are validated where they are read (buy2buy never reads the clicks it filters out)
    if variant == "BUY2BUY":
        return
    bad = synth.EventFrame(pinned.session, pinned.aid.clone().pin_memory(), pinned.ts, pinned.type, frame.n_aids)
    last = int(cv.ingest(frame, "asc", device="cuda:0").offsets[1].item()) - 1     # most recent event of the first session
    bad.aid[last] = frame.n_aids + 7
    with pytest.raises(cv.N.OttoError):
        cv.build_topk(cv.ingest(bad, "asc", device="cuda:0", zero_copy=True), spec)


@pytest.mark.parametrize("variant", VARIANTS)
def test_ascending_frames_with_runs_of_equal_ts_across_the_tail_cut(cv, variant):
    """Sessions of 45-70 events in runs of equal ts (1-40 events per second): the 30-event tail cuts through runs, where
    the stable descending sort keeps the FIRST events of the run in file order.  Checked against the oracle."""
    rng = np.random.default_rng(3)
    rows = []
    for s in range(300):
        n, t = int(rng.integers(45, 71)), 1660000000 + int(rng.integers(0, 1000000))
        while n > 0:
            run = min(n, int(rng.integers(1, 41)))
            rows += [(s, int(rng.integers(0, 150)), t, int(rng.choice([0, 1, 2], p=[0.5, 0.3, 0.2]))) for _ in range(run)]
            n -= run
            t += int(rng.integers(1, 5000))
    df = pd.DataFrame(rows, columns=["session", "aid", "ts", "type"])
    frame = frame_from_df(df, 150)
    spec = getattr(cv, variant)
    check_against_oracle(cv, frame, spec, f"{variant} tie runs (desc CSR)")
    want, _ = cv.build_topk(cv.ingest(frame, "desc", device="cuda:0"), spec, exact=True)
    got, _ = cv.build_topk(cv.ingest(frame, "asc", device="cuda:0"), spec, exact=True)
    assert _tables_equal(got, want) and torch.equal(got.cnt, want.cnt) and torch.equal(got.tsum, want.tsum)


def test_frame_pipeline_equals_sequential_builds(cv):
    """FramePipeline (upload of frame i + 1 on a copy stream while frame i is built, rows of frame i - 1 on their way
    back): the rows of every frame equal a plain ingest + build of that frame, for frames of different sizes."""
    from otto_multi_objective_recommender_system_b200 import synth
    frames = [synth_frame(n, 300, seed=s) for n, s in ((2000, 1), (3500, 2), (800, 3), (2600, 4), (2000, 1))]
    pinned = [synth.EventFrame(*(t.cpu().contiguous().pin_memory() for t in (f.session, f.aid, f.ts, f.type)), n_aids=f.n_aids)
              for f in frames]
    got = [tuple(x.clone() for x in rows) for rows in cv.FramePipeline(cv.CARTS_ORDERS, "cuda:0").run(pinned)]
    assert len(got) == len(frames)
    for f, (ax, ay, w) in zip(frames, got):
        t, _ = cv.build_topk(cv.ingest(f, "desc", device="cuda:0"), cv.CARTS_ORDERS)
        wx, wy, ww = (x.cpu() for x in t.to_rows())
        assert torch.equal(ax, wx) and torch.equal(ay, wy) and torch.equal(w.view(torch.int32), ww.view(torch.int32))


def test_build_is_deterministic(cv):
    frame = synth_frame(8000, 800, seed=2)
    csr = cv.ingest(frame, "desc", device="cuda:0")
    a, _ = cv.build_topk(csr, cv.CLICKS, exact=True)
    b, _ = cv.build_topk(csr, cv.CLICKS, exact=True)
    for f in ("aid_y", "wgt", "len"):
        assert torch.equal(getattr(a, f), getattr(b, f)), f


def test_rows_round_trip(cv):
    frame = synth_frame(2000, 300, seed=4)
    csr = cv.ingest(frame, "desc", device="cuda:0")
    t, _ = cv.build_topk(csr, cv.CARTS_ORDERS)
    ax, ay, w = t.to_rows()
    t2 = cv.TopKTable.from_rows(ax, ay, w, t.n_aids, t.k)
    assert torch.equal(t.len, t2.len) and torch.equal(t.aid_y, t2.aid_y)
    m = t.aid_y >= 0
    assert torch.equal(t.wgt[m], t2.wgt[m])


def test_linearity_over_session_halves(cv):
    """P(all) = P(first half) + P(second half): sessions are independent (size-independent property)."""
    frame = synth_frame(6000, 500, seed=13)
    csr = cv.ingest(frame, "desc", device="cuda:0")
    half = csr.n_sessions // 2
    stats = []
    for part in (csr, csr.slice_sessions(0, half), csr.slice_sessions(half, csr.n_sessions)):
        b = cv.CovisitBuilder(part, cv.CARTS_ORDERS)
        b.build()
        stats.append(b.stats.as_dict())
    assert stats[0]["pairs"] == stats[1]["pairs"] + stats[2]["pairs"]
    assert stats[0]["pair_checksum"] == stats[1]["pair_checksum"] + stats[2]["pair_checksum"]


def test_bin_arrays_guard_refuses_more_bins_than_sized(cv):
    """Multi-GPU contract: bins come from the all-reduced bounds, so a rank whose workspace was sized for its own
    shard only (global_events = 0) must be refused before anything is written past the bin arrays."""
    from dataclasses import replace
    frame = synth_frame(400, 60, seed=3)
    csr = cv.ingest(frame, "desc", device="cuda:0")
    b = cv.CovisitBuilder(csr, replace(cv.CLICKS, split_ub=16))
    b.count_begin()
    b.stats.bins = 0
    ub = b.views()["row_total"]
    ub.mul_(1000)                      # what an all-reduce over many more ranks would do to the row totals
    with pytest.raises(cv.N.OttoError) as e:
        b.count_finish()
    assert e.value.code == cv.N.OTTO_ENOSPC and "global_events" in str(e.value)
    # sized for the global frame, the same bounds are fine
    b2 = cv.CovisitBuilder(csr, replace(cv.CLICKS, split_ub=16, global_events=1000 * csr.n_events))
    b2.count_begin()
    b2.stats.bins = 0
    b2.views()["row_total"].mul_(1000)
    assert b2.count_finish()["bins"] > csr.n_aids


def test_one_shot_c_entry_point_with_workspace_retry(cv):
    """otto_covisit_build through raw ctypes, the way INTEGRATION.md shows it: a first call with a workspace that is
    too small answers OTTO_ENOSPC with the pair count filled in, otto_covisit_build_bytes sizes the retry."""
    import ctypes as C
    N = cv.N
    lib = N.lib()
    frame = synth_frame(3000, 500, seed=7)
    csr = cv.ingest(frame, "desc", device="cuda:0")
    spec = cv.CARTS_ORDERS
    cspec = spec.to_c(csr.n_aids)
    ev = N.OttoEvents(csr.n_sessions, csr.n_events, csr.offsets.data_ptr(), csr.aid.data_ptr(), csr.ts.data_ptr(), csr.type.data_ptr())
    sizes = N.OttoBuildSizes()
    N.check(lib.otto_covisit_sizes(csr.n_sessions, csr.n_events, C.byref(cspec), C.byref(sizes)))
    table = cv.TopKTable.empty(csr.n_aids, spec.k, "cuda:0")
    tc = table.to_c()
    stats = N.OttoBuildStats()
    st = torch.cuda.current_stream().cuda_stream
    ws = torch.empty(sizes.workspace_bytes, dtype=torch.uint8, device="cuda:0")
    rc = lib.otto_covisit_build(C.byref(ev), C.byref(cspec), ws.data_ptr(), ws.numel(), C.byref(tc), C.byref(stats), st)
    assert rc == N.OTTO_ENOSPC and stats.pairs > 0 and b"workspace too small" in lib.otto_last_error()
    assert stats.hot_pairs == 0
    need = lib.otto_covisit_build_bytes(csr.n_sessions, csr.n_events, C.byref(cspec), stats.pairs + stats.hot_pairs, stats.bins)
    ws = torch.empty(need, dtype=torch.uint8, device="cuda:0")
    N.check(lib.otto_covisit_build(C.byref(ev), C.byref(cspec), ws.data_ptr(), ws.numel(), C.byref(tc), C.byref(stats), st))
    torch.cuda.synchronize()
    want = co.build(frame.to_pandas(), H.oracle_spec(spec))
    H.assert_int_table_equal(table.to_pandas(), want, "one-shot build")
    assert stats.distinct > 0 and stats.table_overflow == 0 and sum(stats.tier_records) == stats.pairs
    # the same with hot (split) rows: their records pass through the staging area behind the final ones, so the retry
    # has to be sized with pairs + hot_pairs; a workspace sized with pairs alone is refused again, not overrun
    from dataclasses import replace
    cspec = replace(spec, split_ub=64).to_c(csr.n_aids)
    N.check(lib.otto_covisit_sizes(csr.n_sessions, csr.n_events, C.byref(cspec), C.byref(sizes)))   # more bins: larger fixed part
    ws = torch.empty(sizes.workspace_bytes, dtype=torch.uint8, device="cuda:0")
    rc = lib.otto_covisit_build(C.byref(ev), C.byref(cspec), ws.data_ptr(), ws.numel(), C.byref(tc), C.byref(stats), st)
    assert rc == N.OTTO_ENOSPC and stats.hot_pairs > 0 and stats.split_rows > 0
    short = lib.otto_covisit_build_bytes(csr.n_sessions, csr.n_events, C.byref(cspec), stats.pairs, stats.bins)
    ws = torch.empty(short, dtype=torch.uint8, device="cuda:0")
    assert lib.otto_covisit_build(C.byref(ev), C.byref(cspec), ws.data_ptr(), ws.numel(), C.byref(tc), C.byref(stats), st) == N.OTTO_ENOSPC
    need = lib.otto_covisit_build_bytes(csr.n_sessions, csr.n_events, C.byref(cspec), stats.pairs + stats.hot_pairs, stats.bins)
    ws = torch.empty(need, dtype=torch.uint8, device="cuda:0")
    N.check(lib.otto_covisit_build(C.byref(ev), C.byref(cspec), ws.data_ptr(), ws.numel(), C.byref(tc), C.byref(stats), st))
    torch.cuda.synchronize()
    H.assert_int_table_equal(table.to_pandas(), want, "one-shot build with split rows")


def test_event_contents_are_validated(cv):
    """aids index device arrays (in owner-direct mode a peer's), types are packed into two bits, ts - ts_min is stored as
    u32: out-of-range events must be refused, not scattered."""
    N = cv.N
    good = pd.DataFrame({"session": [1, 1, 2, 2], "aid": [1, 2, 3, 4], "ts": [1659304800 + i for i in range(4)], "type": [0, 1, 2, 0]})
    for col, val in (("aid", 8), ("aid", -1), ("type", 3)):
        bad = good.copy()
        bad.loc[2, col] = val
        frame = cv.EventFrame(*(torch.from_numpy(bad[c].to_numpy().astype(np.int64)) for c in ("session", "aid", "ts", "type")), n_aids=8)
        with pytest.raises(N.OttoError, match="outside"):
            cv.ingest(frame, "desc", device="cuda:0")
    # millisecond timestamps (the pickles' unit) in time mode: refused at count_finish, nothing written out of bounds
    ms = good.assign(ts=good["ts"].astype(np.int64) % 2_000_000 + 1_700_000_000)
    csr = cv.ingest(frame_from_df(ms, 8), "desc", device="cuda:0")
    b = cv.CovisitBuilder(csr, cv.CLICKS)
    b.count_begin()
    with pytest.raises(N.OttoError, match="ts outside"):
        b.count_finish()
    # the same frame is fine for the variants that do not weight by time
    table, stats = cv.build_topk(csr, cv.CARTS_ORDERS)
    assert stats["pairs"] == 4
    # a CSR whose aids were valid at ingest but exceed the spec's n_aids (caller passes a smaller n_aids to the builder)
    csr_small = cv.EventCSR(csr.session_ids, csr.offsets, csr.aid, csr.ts, csr.type, 3, "desc")
    b = cv.CovisitBuilder(csr_small, cv.CARTS_ORDERS)
    b.count_begin()
    with pytest.raises(N.OttoError, match="aid outside"):
        b.count_finish()
