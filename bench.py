#!/usr/bin/env python
"""Benchmark of the covisitation hot path (contract: see README / DESIGN.md §Measurement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--scale F] [--variant clicks|carts_orders|buy2buy]
  python bench.py --impl reference ...     # the CPU path (pandas oracle port) on the box's host cores

A step = one full build of one covisitation matrix (tail CSR -> pair-gen -> scatter -> accumulate ->
top-K) over the whole synthetic OTTO-shaped event CSR, already resident in HBM.  `value` = events / s.
`e2e` = the same metric through the public API from pinned HOST frame columns: H2D copy + ingest +
build + D2H of the top-K rows inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import pathlib
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "covisit_build_events_per_s"
UNIT = "events/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of full OTTO scale (sessions and aids)")
    ap.add_argument("--variant", default="clicks", choices=["clicks", "carts_orders", "buy2buy"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-sample", type=float, default=0.01, help="fraction of full scale for the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-candidates", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true", help="skip the three-matrices + candidates pipeline line")
    ap.add_argument("--pipeline-steps", type=int, default=2)
    ap.add_argument("--recall-sample", type=int, default=10000, help="sessions of the recall@20 check against the oracle")
    ap.add_argument("--write-digest", action="store_true", help="N = 1: store this run's parity digest under tests/golden/")
    ap.add_argument("--split-ub", type=int, default=0)
    ap.add_argument("--nccl-exchange", action="store_true", help="N > 1: exchange records with an NCCL all-to-all instead of NVLink peer memory")
    ap.add_argument("--peer-read", action="store_true", help="N > 1: owners read the senders' slabs over NVLink (the variant before the owner-direct scatter)")
    ap.add_argument("--staged", type=int, default=None, choices=[0, 1],
                    help="N > 1: 1 = staged scatter (sender-side combining, owners pull their buckets over NVLink), 0 = direct "
                         "scatter (NVLink stores); default: distributed.STAGED_DEFAULT / OTTO_STAGED")
    ap.add_argument("--dist-timing", action="store_true", help="N > 1: synchronise between phases and print their times (stderr)")
    return ap.parse_args()


def workload_name(args) -> str:
    return (f"synthetic OTTO-shaped train frame at scale {args.scale:g} "
            f"({args.variant} covisitation, tail 30, top-k per aid)")


# ----------------------------------------------------------------------------- clocks

class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = tempfile.NamedTemporaryFile(prefix="clocks_", suffix=".csv", delete=False).name
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU arms

def _oracle_spec(variant):
    from oracle import covisit_oracle as co
    return {"clicks": co.CLICKS, "carts_orders": co.CARTS_ORDERS, "buy2buy": co.BUY2BUY}[variant]


_CPU_DF = None      # inherited by the forked workers: no frame is pickled
_CPU_DIR = None     # /dev/shm scratch of one cpu_build call


def _cpu_chunk(args):
    """Stage 1: accumulate one contiguous session range, write the partial sums split by aid_x bucket."""
    i, lo, hi, variant, edges = args
    import numpy as np
    from oracle import covisit_oracle as co
    acc = co.accumulate(_CPU_DF.iloc[lo:hi], _oracle_spec(variant))
    ax = acc["aid_x"].to_numpy()
    ay, w = acc["aid_y"].to_numpy(), acc["wgt"].to_numpy()
    cut = np.searchsorted(ax, edges)               # accumulate() returns rows sorted by (aid_x, aid_y)
    for r in range(len(edges) - 1):
        sl = slice(cut[r], cut[r + 1])
        np.save(f"{_CPU_DIR}/p{i}_{r}_x.npy", ax[sl]); np.save(f"{_CPU_DIR}/p{i}_{r}_y.npy", ay[sl]); np.save(f"{_CPU_DIR}/p{i}_{r}_w.npy", w[sl])
    return len(acc)


def _cpu_bucket(args):
    """Stage 2: one aid_x bucket - sum the partial sums of every session range, top-K."""
    r, n_parts, variant = args
    import numpy as np
    import pandas as pd
    from oracle import covisit_oracle as co
    cols = {c: np.concatenate([np.load(f"{_CPU_DIR}/p{i}_{r}_{c}.npy") for i in range(n_parts)]) for c in "xyw"}
    acc = pd.DataFrame({"aid_x": cols["x"], "aid_y": cols["y"], "wgt": cols["w"]})
    acc = acc.groupby(["aid_x", "aid_y"], as_index=False)["wgt"].sum()
    return co.topk(acc.astype({"wgt": "float32"}), _oracle_spec(variant).k)


def cpu_build(df, variant: str, workers: int):
    """The pandas restatement (oracle port).  workers > 1: what a chunked CPU builder does with every core of the box,
    both stages in a fork pool - (1) contiguous session ranges (the reference's chunk files are 100k consecutive
    sessions) accumulated independently, their partial sums exchanged through /dev/shm split by aid_x bucket;
    (2) every aid_x bucket summed over the ranges and cut to its top K.  Round 1 merged the partial sums with one serial
    concat + groupby in the parent, which capped 16-32 processes at 2.1x of one."""
    global _CPU_DF, _CPU_DIR
    from oracle import covisit_oracle as co
    spec = _oracle_spec(variant)
    if workers <= 1:
        return co.build(df, spec)
    import multiprocessing as mp
    import shutil
    import numpy as np
    import pandas as pd
    sess = df["session"].to_numpy()
    ids = np.unique(sess)
    n_parts = min(len(ids), workers * 4)          # a few parts per worker: session lengths are skewed
    cuts = [int(np.searchsorted(sess, ids[i * len(ids) // n_parts])) for i in range(n_parts)] + [len(df)]
    n_aids = int(df["aid"].max()) + 1
    n_buckets = workers * 2
    edges = np.linspace(0, n_aids, n_buckets + 1).astype(np.int64)
    _CPU_DF = df
    _CPU_DIR = tempfile.mkdtemp(prefix="otto_cpu_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        with mp.get_context("fork").Pool(workers) as pool:
            pool.map(_cpu_chunk, [(i, cuts[i], cuts[i + 1], variant, edges) for i in range(n_parts)], chunksize=1)
            tops = pool.map(_cpu_bucket, [(r, n_parts, variant) for r in range(n_buckets)], chunksize=1)
    finally:
        shutil.rmtree(_CPU_DIR, ignore_errors=True)
    return pd.concat(tops, ignore_index=True)


def cpu_baseline(args, workers: int) -> dict:
    from otto_multi_objective_recommender_system_b200 import synth
    frame = synth.generate(synth.SynthSpec.scaled("train", args.cpu_sample))
    df = frame.to_pandas()
    t0 = time.perf_counter()
    cpu_build(df, args.variant, workers)
    dt = time.perf_counter() - t0
    out = {"value": len(df) / dt, "unit": UNIT, "cores": workers, "kind": "port", "seconds": dt,
           "sample": f"pandas oracle (oracle/covisit_oracle.py), {args.variant}, {args.cpu_sample:g} of full scale = "
                     f"{len(df)} events, one build"}
    try:
        # context beside the pandas figure: the plain-C restatement of the same recipe (oracle/covisit_oracle.c), one core
        from oracle import covisit_oracle_c as cc
        cc.lib()
        t0 = time.perf_counter()
        cc.build(df, _oracle_spec(args.variant))
        dtc = time.perf_counter() - t0
        out["c_port"] = {"value": len(df) / dtc, "unit": UNIT, "cores": 1, "seconds": dtc,
                         "sample": "plain-C oracle (oracle/covisit_oracle.c) on the same frame, one build"}
    except Exception as e:              # no compiler on the box: the pandas figure stands alone
        out["c_port"] = {"unavailable": str(e)[:200]}
    return out


def _c_range(args):
    from oracle import covisit_oracle_c as cc
    variant, lo, hi = args
    return len(cc.build_c(_CPU_DF, _oracle_spec(variant), x_range=(lo, hi)))


def c_port_rate(df, variant: str, workers: int) -> dict:
    """Context beside the pandas figure: the plain-C restatement of the same recipe (oracle/covisit_oracle.c) on the same
    frame, one aid_x range per process.  Never the line's `value`: the reference's idiom is pandas."""
    global _CPU_DF
    try:
        import multiprocessing as mp
        import numpy as np
        from oracle import covisit_oracle_c as cc
        cc.lib()
        _CPU_DF = df
        edges = np.linspace(0, int(df["aid"].max()) + 1, max(1, workers) + 1).astype(np.int64)
        jobs = [(variant, int(edges[r]), int(edges[r + 1])) for r in range(len(edges) - 1)]
        t0 = time.perf_counter()
        if workers > 1:
            with mp.get_context("fork").Pool(workers) as pool:
                pool.map(_c_range, jobs, chunksize=1)
        else:
            _c_range(jobs[0])
        dt = time.perf_counter() - t0
        return {"value": len(df) / dt, "unit": UNIT, "cores": workers, "seconds": dt,
                "sample": f"plain-C oracle (oracle/covisit_oracle.c) on the same frame, {len(jobs)} aid_x ranges in {workers} processes, one build"}
    except Exception as e:              # no compiler on the box: the pandas figure stands alone
        return {"unavailable": str(e)[:200]}


def run_reference(args):
    """--impl reference: the CPU path on the host cores.  The reference repo has no builder to run (SURVEY.md
    §0.1), so this is the oracle port, with every host core, on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from otto_multi_objective_recommender_system_b200 import synth
    workers = os.cpu_count() or 1
    frame = synth.generate(synth.SynthSpec.scaled("train", args.cpu_sample))
    df = frame.to_pandas()
    steps, warmup = max(1, min(args.steps, 3)), max(0, min(args.warmup, 1))
    for _ in range(warmup):
        cpu_build(df, args.variant, workers)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_build(df, args.variant, workers)
    dt = (time.perf_counter() - t0) / steps
    value = len(df) / dt
    sample = (f"pandas oracle port, {workers} worker processes (session ranges accumulated in parallel, aid_x buckets "
              f"merged in parallel), {args.variant}, {args.cpu_sample:g} of full scale = {len(df)} events per step")
    c_port = c_port_rate(df, args.variant, workers)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "cpu_sample": args.cpu_sample, "events_per_step": len(df),
                   "note": "same generator, recipe and metric as the b200 arm; the CPU arm times a bounded sample of the "
                           "frame (cpu_sample of full scale), the b200 arm the whole frame"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample, "c_port": c_port},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ----------------------------------------------------------------------------- parity digest

DIGEST_FILE = ROOT / "tests" / "golden" / "bench_digest.json"


def table_digest(torch, table, lo: int, hi: int):
    """Order-independent 64-bit digest of rows [lo, hi) of a top-K table: wrapping int64 sums over the valid entries
    of mix(slot index, aid_y, float bits of wgt) and of the row lengths.  Per-rank digests add up (mod 2^64) to the
    digest of the whole table, so the N-GPU result can be compared with the committed single-GPU one."""
    k = table.k
    ay = table.aid_y[lo:hi].to(torch.int64)
    wb = table.wgt[lo:hi].contiguous().view(torch.int32).to(torch.int64)
    ln = table.len[lo:hi].to(torch.int64)
    slot = (torch.arange(lo, hi, device=ay.device, dtype=torch.int64) * k)[:, None] + torch.arange(k, device=ay.device, dtype=torch.int64)[None, :]
    valid = torch.arange(k, device=ay.device)[None, :] < ln[:, None]
    h = (slot * -7046029254386353131 + ay) * -4658895280553007687          # golden-ratio / murmur multipliers, wrapping
    h = (h ^ (h >> 29)) * -7723592293110705685 + wb * 2654435761
    h = h ^ (h >> 32)
    return torch.stack([torch.where(valid, h, torch.zeros_like(h)).sum(), ln.sum()])


def digest_key(args) -> str:
    return f"{args.variant}@{args.scale:g}"


# ----------------------------------------------------------------------------- pipeline (north_star target)

def run_pipeline(args, torch, dist, mods, dev, world, rank, peer, csr, barrier, allmax, allsum):
    """Three matrices (clicks, carts-orders, buy2buy) + top-20 candidates for every test session, inputs resident in
    HBM: the north_star target (< 10 s on 8 B200).  N > 1: every matrix is built by all ranks (sessions sharded, rows
    owned by aid_x range), all-gathered, and the test sessions are sharded over the ranks.
    Returns the `pipeline` object of the JSON line."""
    candidates, covisit, distributed, synth = mods
    A = csr.n_aids
    ev = lambda: torch.cuda.Event(enable_timing=True)
    test = synth.generate(synth.SynthSpec.scaled("test", args.scale), device=dev)
    sess_all = covisit.ingest(test, "asc", device=dev)
    T_all = sess_all.n_sessions
    lo_s, hi_s = rank * T_all // world, (rank + 1) * T_all // world
    sess = sess_all.slice_sessions(lo_s, hi_s) if world > 1 else sess_all
    popular = {t: list(range(20)) for t in ("click", "cart", "order")}
    state = {}

    def once(timed: bool):
        marks = {}
        tables = {}
        barrier()
        w0 = time.perf_counter()
        for stem, vspec in covisit.VARIANTS.items():
            e0, e1 = ev(), ev()
            e0.record()
            if world > 1:
                be = distributed.GpuRankBackend(csr, vspec, peer=peer)
                table, _, _, plan = distributed.build_topk_distributed(be)
                distributed.gather_table(table, plan)
            else:
                table, _ = covisit.build_topk(csr, vspec)
            e1.record()
            tables[stem] = table
            marks[stem] = (e0, e1)
        c0, c1 = ev(), ev()
        c0.record()
        gen = candidates.CandidateGenerator(tables, candidates.reference_spec(tables.keys(), 20), A)
        mlen = candidates.max_session_len(sess)
        cand = gen(sess, mlen)
        pred, long_s = candidates.assemble_predictions(sess, cand, popular, 20)
        candidates.recency_long_predictions(sess, tables, pred, long_s, 20)
        c1.record()
        barrier()
        wall = time.perf_counter() - w0
        state.update(tables=tables, pred=pred, long_s=long_s)
        return wall, {k: a.elapsed_time(b) for k, (a, b) in marks.items()}, c0.elapsed_time(c1)

    once(False)                                    # allocations, first-touch, side streams
    walls, per_variant, cand_ms = [], {}, []
    for _ in range(max(1, args.pipeline_steps)):
        w, pv, cm = once(True)
        walls.append(allmax(w))
        for k, v in pv.items():
            per_variant.setdefault(k, []).append(allmax(v))
        cand_ms.append(allmax(cm))
    out = {"seconds": min(walls), "what": "3 builds (+ all-gather of the tables at N > 1) + otto_candidates + assemble + recency branch; "
                                          "inputs resident in HBM; host wall clock between barriers, max over ranks, best of "
                                          f"{len(walls)}",
           "per_variant_ms": {k: min(v) for k, v in per_variant.items()}, "candidates_ms": min(cand_ms),
           "test_sessions": T_all, "test_sessions_rank0": sess.n_sessions, "long_sessions_all_ranks": allsum(int(state["long_s"].sum())),
           "target_seconds": 10.0}
    # ---- recall@20 of the GPU lists against the oracle on a fixed sample (held-out tails of rank 0's first sessions)
    if rank == 0 and args.recall_sample > 0:
        out.update(recall_check(args, torch, mods, dev, sess, state["tables"], popular))
    return out


def recall_check(args, torch, mods, dev, sess, tables, popular) -> dict:
    """recall@20 of the standalone model on held-out tails (validation.py:73-83 semantics: history up to a cutoff, the
    rest is ground truth) for the first `--recall-sample` sessions: CUDA path vs the restated reference loop
    (oracle/candidates_oracle.py) fed the same table rows.  Both the lists and the three recalls must be equal."""
    import numpy as np
    import pandas as pd
    from oracle import candidates_oracle as oc
    candidates, covisit, distributed, synth = mods
    n = min(args.recall_sample, sess.n_sessions)
    off = sess.offsets[:n + 1].cpu().numpy().astype(np.int64)
    aid = sess.aid[:off[-1]].cpu().numpy()
    typ = sess.type[:off[-1]].cpu().numpy()
    sid = sess.session_ids[:n].cpu().numpy()
    rng = np.random.default_rng(7)
    rows, labels = [], {"click": [], "cart": [], "order": []}
    kept = 0
    for i in range(n):
        a, t = aid[off[i]:off[i + 1]].tolist(), typ[off[i]:off[i + 1]].tolist()
        if len(a) < 2:
            continue
        cut = int(rng.integers(0, len(a) - 1))
        (ha, ht), (c, k, o) = oc.split_for_recall(a, t, cut)
        rows.append(pd.DataFrame({"session": int(sid[i]), "aid": ha, "ts": np.arange(len(ha)), "type": ht}))
        labels["click"].append(c)
        labels["cart"].append(k)
        labels["order"].append(o)
        kept += 1
    if not rows:
        return {"recall_equal": None}
    hdf = pd.concat(rows, ignore_index=True)
    hs = covisit.ingest(synth.EventFrame.from_pandas(hdf, sess.n_aids), "asc", device=dev)
    cand = candidates.generate_candidates(hs, tables, candidates.reference_spec(tables.keys(), 20))
    pred, long_s = candidates.assemble_predictions(hs, cand, popular, 20)
    candidates.recency_long_predictions(hs, tables, pred, long_s, 20)
    # the table rows the sampled sessions can touch, as the reference's dicts
    need = torch.unique(hs.aid.to(torch.int64))
    otables = {}
    for stem, tb in tables.items():
        ay = tb.aid_y[need].cpu().numpy()
        ln = tb.len[need].cpu().numpy()
        otables[stem] = {int(x): [int(v) for v in ay[j, :ln[j]]] for j, x in enumerate(need.cpu().tolist()) if ln[j] > 0}
    hl = oc.session_lists(hdf)
    want = {"click": [], "cart": [], "order": []}
    pops = [popular["click"], popular["cart"], popular["order"]]
    for t in hl.itertuples():
        if len(set(t.aid)) >= 20:
            w = oc.recency_predictions(t.aid, t.type, otables, 20)
        else:
            w = oc.standalone_predictions(t.aid, t.type, otables, pops, 20)
        for ti, name in enumerate(("click", "cart", "order")):
            want[name].append(list(w[ti]))
    p = pred.cpu().numpy()
    lists_equal = all([int(a) for a in p[ti, i] if a >= 0] == want[name][i]
                      for ti, name in enumerate(("click", "cart", "order")) for i in range(len(hl)))
    got_r, want_r = {}, {}
    for ti, name in enumerate(("click", "cart", "order")):
        csr_l = candidates.LabelCSR.build(np.arange(len(hl)), [set([x]) if not isinstance(x, (list, set, tuple)) else set(x) for x in labels[name]], dev)
        got_r[name] = candidates.recall_at_20(pred[ti], csr_l)
        want_r[name] = oc.recall_at_20(want[name], [[x] if not isinstance(x, (list, set, tuple)) else list(x) for x in labels[name]])
    weighted = 0.1 * got_r["click"] + 0.3 * got_r["cart"] + 0.6 * got_r["order"]
    return {"recall_sample_sessions": kept, "recall_at_20": {**got_r, "weighted": weighted},
            "recall_at_20_oracle": want_r, "lists_equal": bool(lists_equal),
            "recall_equal": bool(lists_equal and all(got_r[k] == want_r[k] for k in got_r))}


# ----------------------------------------------------------------------------- B200 arm

def run_b200(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    g.build()
    from otto_multi_objective_recommender_system_b200 import _native as N
    from otto_multi_objective_recommender_system_b200 import candidates, covisit, distributed, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"        # the version banner goes to stdout, where the one JSON line belongs
        dist.init_process_group("nccl", device_id=dev)
    lib = N.lib()
    spec = {"clicks": covisit.CLICKS, "carts_orders": covisit.CARTS_ORDERS, "buy2buy": covisit.BUY2BUY}[args.variant]
    if args.split_ub:
        from dataclasses import replace
        spec = replace(spec, split_ub=args.split_ub)

    # ---- synthetic frame on the device; N > 1: strong scaling, the same frame sharded by session chunk ----
    sspec = synth.SynthSpec.scaled("train", args.scale, seed=42)
    frame = synth.generate(sspec, device=dev)
    if world > 1:
        full = covisit.ingest(frame, "asc", device=dev)
        S_all = full.n_sessions
        lo_s, hi_s = rank * S_all // world, (rank + 1) * S_all // world
        e0, e1 = int(full.offsets[lo_s].item()), int(full.offsets[hi_s].item())
        frame = synth.EventFrame(frame.session[e0:e1].clone(), frame.aid[e0:e1].clone(), frame.ts[e0:e1].clone(),
                                 frame.type[e0:e1].clone(), frame.n_aids)
        del full
        torch.cuda.empty_cache()
    csr = covisit.ingest(frame, "desc", device=dev)
    E, S, A = csr.n_events, csr.n_sessions, csr.n_aids
    if args.peer_read:
        os.environ["OTTO_OWNER_DIRECT"] = "0"
    if args.staged is not None:
        os.environ["OTTO_STAGED"] = str(args.staged)
    peer = distributed.PeerRecords(dev) if world > 1 and not args.nccl_exchange else None
    backend = distributed.GpuRankBackend(csr, spec, peer=peer)
    builder = backend.b

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def allsum(x: int) -> int:
        if world == 1:
            return int(x)
        t = torch.tensor([int(x)], device=dev, dtype=torch.int64)
        dist.all_reduce(t)
        return int(t.item())

    def allmax(x: float) -> float:
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ev = lambda: torch.cuda.Event(enable_timing=True)
    import ctypes as C
    phases = ["count_begin", "count_finish", "scatter", "reduce"]
    reduce_ms, scatter_ms = [], []
    last = {}
    dist_timing = {}

    def step(marks=None):
        if world > 1:
            # sessions sharded by chunk; all-reduce of bounds / counts; all-to-all of pair slabs; owner reduce
            tb, rng_, st, _ = distributed.build_topk_distributed(backend, timing=dist_timing if args.dist_timing else None)
            last.update(st)
            last["table"], last["range"] = tb, rng_
            return
        m = [ev() for _ in range(5)] if marks is not None else None
        if m: m[0].record()
        builder.count_begin()
        if m: m[1].record()
        builder.count_finish()
        if m: m[2].record()
        builder.scatter()
        if m: m[3].record()
        last["table"], last["range"] = builder.reduce(sync=True), (0, csr.n_aids)
        if m: m[4].record()
        if marks is not None:
            marks.append(m)

    for _ in range(max(args.warmup, 3)):
        step()
    stats = builder.stats.as_dict()
    dist_timing.clear()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.otto_launch_count()
    marks = [] if world == 1 else None
    t_start, t_end = ev(), ev()
    barrier()
    t_start.record()
    for _ in range(args.steps):
        step(marks)
    t_end.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches_total = lib.otto_launch_count() - launches0          # this rank's kernels inside the timed region
    launches = launches_total // max(1, args.steps)
    dist_phase_ms, nvlink = None, None
    if world > 1:
        # phase times at N > 1: three extra steps OUTSIDE the timed region with CUDA events recorded between the phases
        # (no synchronisation inside a build), and the bytes this rank's scatter stored into OTHER owners' HBM
        acc = {}
        barrier()                                      # rank 0 just spent ~0.2 s stopping the clock sampler
        for _ in range(3):
            tm = {"__events__": True}
            distributed.build_topk_distributed(backend, timing=tm)
            torch.cuda.synchronize(dev)
            for k, v in distributed.phase_ms_from_events(tm).items():
                acc.setdefault(k, []).append(v)
        dist_phase_ms = {k: allmax(statistics.mean(v)) for k, v in acc.items()}
        if backend.owner_direct and getattr(backend, "_gathered", None) is not None:
            lo_o, hi_o = last["range"]
            mine = backend._gathered[rank].to(torch.int64) & 0xFFFFFFFF
            remote = int((mine.sum() - mine[lo_o:hi_o].sum()).item()) * 8
            remote_max = allmax(float(remote))
            sc = dist_phase_ms.get("scatter")
            nvlink = {"off_rank_bytes_per_gpu_max": remote_max, "scatter_ms": sc,
                      "achieved_gbs": remote_max / (sc * 1e-3) / 1e9 if sc else None,
                      "peak_gbs": 770.0, "peak_source": "measured peer copy per direction (B200_PROFILING.md); nominal 900",
                      "what": ("bytes the place pass reads out of the other ranks' staging buffers (large NVLink reads) / scatter "
                               "phase time (bucket plan + pass A into the local staging buffer + the 4-byte all-reduce 'every "
                               "rank has staged' + place pass)") if backend.staged else
                              ("bytes the scatter kernel stores into other owners' HBM (NVLink stores) / scatter phase time "
                               "(kernel + the 4-byte all-reduce that orders 'all scatters have landed')")}
    if world == 1:
        # per-kernel times of the reduce phase: three extra steps OUTSIDE the timed region with the library's event
        # bracketing on (profiled calls run the block kernels back to back instead of concurrently)
        N.check(lib.otto_profile_enable(1))
        for _ in range(3):
            step()
            r5 = (C.c_float * 5)()
            N.check(lib.otto_profile_reduce_ms(r5))
            reduce_ms.append(list(r5))
            r3 = (C.c_float * 3)()
            N.check(lib.otto_profile_scatter_ms(r3))
            scatter_ms.append(list(r3))
        N.check(lib.otto_profile_enable(0))
    ms_step = allmax(t_start.elapsed_time(t_end)) / args.steps
    events_all = allsum(E)
    value = events_all / (ms_step * 1e-3)

    # ---- parity: digest of the rows this rank owns, summed over the ranks, against the committed single-GPU digest ----
    lo_a, hi_a = last["range"]
    dg = table_digest(torch, last["table"], lo_a, hi_a)
    run_stats = stats if world == 1 else last
    tot = torch.tensor([int(run_stats["pairs"]), int(run_stats["distinct"]), int(run_stats["pair_checksum"])], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(dg)
        dist.all_reduce(tot)
    digest = {"rows": f"{int(dg[0].item()) & 0xFFFFFFFFFFFFFFFF:016x}", "row_len_sum": int(dg[1].item()), "pairs": int(tot[0].item()),
              "distinct": int(tot[1].item()), "pair_checksum": int(tot[2].item())}
    golden = {}
    try:
        golden = json.load(open(DIGEST_FILE))
    except (OSError, ValueError):
        pass
    want_digest = golden.get(digest_key(args))
    parity = {"digest": digest, "golden": want_digest,
              "matches_n1": (digest == want_digest) if want_digest is not None else None,
              "what": "wrapping 64-bit sum over the owned rows of mix(slot, aid_y, wgt bits) + row lengths, all-reduced; pairs / "
                      "distinct / pair_checksum summed over ranks; golden = tests/golden/bench_digest.json (written by a 1-GPU run)"}
    if args.write_digest and world == 1 and rank == 0:
        golden[digest_key(args)] = digest
        json.dump(golden, open(DIGEST_FILE, "w"), indent=1, sort_keys=True)
        parity["golden"], parity["matches_n1"] = digest, True

    # ---- roofline (algorithmic bytes: DESIGN.md, "Kernels") ----
    E30, P, B, D = stats["tail_events"], stats["pairs"], stats["bins"], stats["distinct"]
    K = spec.k
    alg = {
        # tail CSR (read offsets + tail events, write 8 B / tail event) + dedupe pass (read tail CSR, write row masks)
        "count_begin": 4 * (S + 1) + 9 * E30 + 4 * (S + 1) + 8 * E30 + 4 * (S + 1) + 8 * E30 + 4 * E30,
        "count_finish": 3 * 4 * A + 8 * A,                                # bins, offsets, cursors from the row counts
        "scatter": 4 * (S + 1) + 12 * E30 + 8 * P,                        # read tail CSR + masks, write records (staging of hot rows not counted)
        "reduce": 8 * P + 8 * (B + 1) + 8 * A * K + 4 * A,                # read records + offsets, write table
    }
    kernel_of = {"count_begin": "tail_copy_kernel", "count_finish": "pairgen_kernel<count>",
                 "scatter": "pairgen_kernel<scatter>", "reduce": "reduce_{small,block}_kernel"}
    peaks = {}
    try:
        peaks = json.load(open(ROOT / "MEASURED_PEAKS.json"))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    bytes_all = allsum(sum(alg.values()))
    whole = {"algorithmic_bytes": bytes_all, "achieved_per_gpu": bytes_all / world / (ms_step * 1e-3) / 1e9,
             "frac": bytes_all / world / (ms_step * 1e-3) / 1e9 / peak}
    if world == 1:
        phase_ms = {p: statistics.mean(m[i].elapsed_time(m[i + 1]) for m in marks) for i, p in enumerate(phases)}
        # per-kernel view: the count / scatter phases are one hot kernel each (plus scans of a few us); the reduce
        # phase is five launches that the library brackets with CUDA events on this stream (otto_profile_reduce_ms)
        tiers = ["reduce_classify + otable_warp_kernel (bins <= 384 records)",
                 "otable_block_kernel<512 threads> (bins <= 6144)",
                 "otable_block_kernel<256 threads> (bins <= 3072)",
                 "otable_block_kernel<128 threads> (bins <= 1536)",
                 "reduce_block_kernel<hash table> (hand-overs, bins > 6144) + merge_split_rows_kernel"]
        tier_of = [0, 3, 2, 1]       # launch order (warp, 512, 256, 128) -> index into stats.tier_records
        red_ms = [statistics.mean(r[i] for r in reduce_ms) for i in range(5)] if reduce_ms else [0.0] * 5
        tr = stats.get("tier_records", [0, 0, 0, 0])
        sc_ms = [statistics.mean(r[i] for r in scatter_ms) for i in range(3)] if scatter_ms else [0.0] * 3
        Ph = stats.get("hot_pairs", 0)
        kernels = {
            "tail_copy_all_kernel + pairgen_kernel<count>": (phase_ms["count_begin"], alg["count_begin"]),
            "pairgen_kernel<scatter>": (sc_ms[0], alg["scatter"]),
            "partition_kernel<count> (+ bin offsets scan)": (sc_ms[1], 8 * Ph + 12 * B),
            "partition_kernel<move>": (sc_ms[2], 16 * Ph),
        }
        for i in range(4):       # a tier reads its records once and writes the rows of its bins
            n_rec = tr[tier_of[i]]
            kernels[tiers[i]] = (red_ms[i], 8 * n_rec + (8 * A * K * n_rec) // max(1, sum(tr)))
        kernels[tiers[4]] = (red_ms[4], 20 * K * 2 * stats["split_rows"])
        dom = max(kernels, key=lambda k: kernels[k][0])
        dom_ms, dom_bytes = kernels[dom]
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
        traffic = None
        try:      # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
            tj = json.load(open(ROOT / "profiles" / "r02_dram_traffic.json"))
            if abs(args.scale - tj.get("scale", -1)) < 1e-9 and args.variant == tj.get("variant"):
                traffic = tj["kernels"].get(dom)
        except (OSError, ValueError, KeyError):
            pass
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes": dom_bytes,
                    "kernel_ms": dom_ms,
                    "timing": "CUDA events on the launching stream: phase events over the timed steps for the tail / pairgen "
                              "kernels; the reduce kernels from 3 extra serialised steps bracketed inside the library",
                    "kernels": {k: {"ms": round(v[0], 4), "algorithmic_bytes": int(v[1]),
                                    "gbs": round(v[1] / (v[0] * 1e-3) / 1e9, 1) if v[0] > 0 else None,
                                    "frac": round(v[1] / (v[0] * 1e-3) / 1e9 / peak, 4) if v[0] > 0 else None}
                                for k, v in kernels.items()},
                    "phase_ms": phase_ms, "whole_build": whole}
    else:
        sent = allsum(P * 8)
        roofline = {"bound": "hbm", "kernel": "whole build (phases interleave with collectives)", "achieved":
                    whole["achieved_per_gpu"], "peak": peak, "unit": "GB/s", "frac": whole["frac"], "traffic": None,
                    "peak_source": peak_src, "algorithmic_bytes": bytes_all,
                    "exchange": {"transport": ("nccl all_to_all" if args.nccl_exchange else "peer read (owners read the senders' slabs over NVLink)"
                                               if not backend.owner_direct else
                                               "staged scatter (runs combined in coarse buckets of the sender's staging buffer, "
                                               "owners pull their buckets over NVLink and place the records)" if backend.staged else
                                               "owner-direct scatter (records stored into the owner's buffer over NVLink)"),
                                 "record_bytes_total": sent, "per_gpu_per_step": sent / world,
                                 "note": "record_bytes_total includes the records a rank keeps for itself; nvlink = off-rank only",
                                 "nvlink": nvlink},
                    "phase_ms": dist_phase_ms}

    # ---- end to end from pinned host columns ----
    e2e = None
    if args.e2e_steps > 0:
        host = synth.EventFrame(*(t.cpu().pin_memory() for t in (frame.session, frame.aid, frame.ts, frame.type)),
                                n_aids=A)
        del frame
        d2h = [0]
        pinned = [None]

        def copy_out(tensors):
            """device -> pinned host buffers that live across steps (a serving loop would keep them too)"""
            if pinned[0] is None or any(b.numel() < t.numel() for b, t in zip(pinned[0], tensors)):
                pinned[0] = [torch.empty(max(1, t.numel()), dtype=t.dtype, pin_memory=True) for t in tensors]
            out = []
            for b, t in zip(pinned[0], tensors):
                b[:t.numel()].copy_(t.reshape(-1), non_blocking=True)
                out.append(b[:t.numel()])
            torch.cuda.current_stream(dev).synchronize()
            return out

        # All four columns are copied and the CSR stays in file order: the tail kernels apply the builder's descending
        # sort while they copy (otto_covisit_count_begin_asc), so the whole-frame reversal of round 1 is gone.  Leaving
        # aid / type in pinned host memory and reading only the tails over PCIe (ingest(..., zero_copy=True)) moves 14 %
        # fewer bytes but was slower on this box: SM loads from host memory reach 25 GB/s against 55 GB/s for the copy
        # engine (tools/time_e2e.py, profiles/r02_experiments.md).
        h2d = sum(t.numel() * t.element_size() for t in (host.session, host.aid, host.ts, host.type))

        def e2e_step():
            f = synth.EventFrame(host.session.to(dev, non_blocking=True), host.aid.to(dev, non_blocking=True),
                                 host.ts.to(dev, non_blocking=True), host.type.to(dev, non_blocking=True), A)
            c = covisit.ingest(f, "asc", device=dev)
            be = distributed.GpuRankBackend(c, spec, peer=peer)
            b = be.b
            b.workspace = builder.workspace          # reuse device buffers, as a long-running service would
            b.records, b.scratch, b.table = builder.records, builder.scratch, builder.table
            if world > 1:
                t, (lo, hi), _, _ = distributed.build_topk_distributed(be)
                out = copy_out([x[lo:hi] for x in (t.aid_y, t.wgt, t.len)])      # the rows this rank owns
            else:
                t = b.build()
                out = copy_out(list(t.to_rows()))
            d2h[0] = sum(x.numel() * x.element_size() for x in out)
            return out

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier()
        dt = allmax((time.perf_counter() - t0) / args.e2e_steps)
        e2e = {"value": events_all / dt, "unit": UNIT, "h2d_bytes_per_step": allsum(h2d),
               "d2h_bytes_per_step": allsum(d2h[0]), "ms_per_step": dt * 1e3, "steps": args.e2e_steps}
        if world == 1:
            # context, not the headline: the same per-step work with two frames in flight (covisit.FramePipeline: the upload
            # of frame i + 1 overlaps the build of frame i and the rows of frame i - 1 on their way back)
            n_pipe = max(4, args.e2e_steps + 2)
            pipe = covisit.FramePipeline(spec, dev)
            for _ in pipe.run([host] * 2):
                pass
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for _ in pipe.run([host] * n_pipe):
                pass
            torch.cuda.synchronize(dev)
            dtp = (time.perf_counter() - t0) / n_pipe
            e2e["pipelined"] = {"value": events_all / dtp, "unit": UNIT, "ms_per_step": dtp * 1e3, "frames_in_flight": 2,
                                "steps": n_pipe, "what": "every frame uploaded, built and read back in full; uploads on a copy "
                                "stream, rows back on a third stream (covisit.FramePipeline)"}
            del pipe

    # ---- the north_star pipeline: three matrices + candidates for every test session (configs 3 / 4 / 5) ----
    pipeline, cand_info = None, None
    if not args.no_pipeline and not args.no_candidates:
        # free the build buffers of the timed loop first (the pipeline allocates one builder per variant)
        last.pop("table", None)
        pipeline = run_pipeline(args, torch, dist, (candidates, covisit, distributed, synth), dev, world, rank, peer, csr, barrier,
                                allmax, allsum)
        cand_info = {"metric": "candidate_gen_sessions_per_s", "value": pipeline["test_sessions"] / (pipeline["candidates_ms"] * 1e-3),
                     "unit": "sessions/s", "ms": pipeline["candidates_ms"], "sessions": pipeline["test_sessions"],
                     "tables": list(covisit.VARIANTS), "top_n": 20, "targets": 3,
                     "what": "otto_candidates + otto_assemble_predictions + otto_recency_long over the rank's shard of the test "
                             "sessions (CUDA events, max over ranks)"}

    if args.dist_timing and world > 1:
        n_calls = args.steps
        print(f"rank {rank} phase ms/step: " + json.dumps({k: round(v / n_calls, 3) for k, v in dist_timing.items()}), file=sys.stderr)
    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cpu = cpu_baseline(args, 1)
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32/u64", "data": "synthetic",
            "config": {"workload": workload_name(args), "sessions_rank0": S, "events_rank0": E, "aids": A,
                       "tail_events_rank0": E30, "pairs_rank0": P, "distinct_pairs_rank0": D, "bins": B,
                       "split_rows": stats["split_rows"], "k": K, "events_all_ranks": events_all,
                       "parallelism": "1 GPU" if world == 1 else f"sessions sharded over {world} GPUs, rows owned by aid_x range (transport: roofline.exchange)",
                       "l2": "inputs larger than L2 (event CSR and pair records are GBs)",
                       "cpu_sample": args.cpu_sample,
                       "note": "the reference arm (--impl reference) and cpu_baseline time the same recipe on cpu_sample of this frame"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "candidates": cand_info, "pipeline": pipeline,
            "parity": parity, "gpu_launches": int(launches_total), "gpu_launches_per_step": int(launches), "clocks": clocks}))
    if world > 1:
        if peer is not None:
            peer.close()
        dist.destroy_process_group()
    if parity["matches_n1"] is False and os.environ.get("OTTO_BENCH_STRICT", "1") != "0":
        sys.exit(f"parity digest differs from the committed single-GPU digest for {digest_key(args)}: {digest} != {want_digest}")


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
