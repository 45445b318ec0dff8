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
    ap.add_argument("--split-ub", type=int, default=0)
    ap.add_argument("--nccl-exchange", action="store_true", help="N > 1: exchange records with an NCCL all-to-all instead of NVLink peer memory")
    ap.add_argument("--peer-read", action="store_true", help="N > 1: owners read the senders' slabs over NVLink (the variant before the owner-direct scatter)")
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


def _cpu_chunk(args):
    lo, hi, variant = args
    from oracle import covisit_oracle as co
    return co.accumulate(_CPU_DF.iloc[lo:hi], _oracle_spec(variant))


def cpu_build(df, variant: str, workers: int):
    """The pandas restatement (oracle port).  workers > 1: contiguous session ranges (the reference's chunk files are
    100k consecutive sessions) accumulated in a fork pool, partial sums combined with one groupby, then top-K - what a
    chunked CPU builder does with every core of the box."""
    global _CPU_DF
    from oracle import covisit_oracle as co
    spec = _oracle_spec(variant)
    if workers <= 1:
        return co.build(df, spec)
    import multiprocessing as mp
    import numpy as np
    import pandas as pd
    sess = df["session"].to_numpy()
    ids = np.unique(sess)
    n_parts = min(len(ids), workers * 4)          # a few parts per worker: session lengths are skewed
    cuts = [int(np.searchsorted(sess, ids[i * len(ids) // n_parts])) for i in range(n_parts)] + [len(df)]
    _CPU_DF = df
    with mp.get_context("fork").Pool(workers) as pool:
        accs = pool.map(_cpu_chunk, [(cuts[i], cuts[i + 1], variant) for i in range(n_parts)], chunksize=1)
    acc = pd.concat(accs, ignore_index=True).groupby(["aid_x", "aid_y"], as_index=False)["wgt"].sum()
    return co.topk(acc.astype({"wgt": "float32"}), spec.k)


def cpu_baseline(args, workers: int) -> dict:
    from otto_multi_objective_recommender_system_b200 import synth
    frame = synth.generate(synth.SynthSpec.scaled("train", args.cpu_sample))
    df = frame.to_pandas()
    t0 = time.perf_counter()
    cpu_build(df, args.variant, workers)
    dt = time.perf_counter() - t0
    return {"value": len(df) / dt, "unit": UNIT, "cores": workers, "kind": "port", "seconds": dt,
            "sample": f"pandas oracle (oracle/covisit_oracle.py), {args.variant}, {args.cpu_sample:g} of full scale = "
                      f"{len(df)} events, one build"}


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
    sample = (f"pandas oracle port, {workers} worker processes, {args.variant}, {args.cpu_sample:g} of full scale = "
              f"{len(df)} events per step")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": workload_name(args), "cpu_sample": args.cpu_sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


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
            _, _, st, _ = distributed.build_topk_distributed(backend, timing=dist_timing if args.dist_timing else None)
            last.update(st)
            return
        m = [ev() for _ in range(5)] if marks is not None else None
        if m: m[0].record()
        builder.count_begin()
        if m: m[1].record()
        builder.count_finish()
        if m: m[2].record()
        builder.scatter()
        if m: m[3].record()
        builder.reduce(sync=True)
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
    launches = (lib.otto_launch_count() - launches0) // max(1, args.steps)
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
        tiers = ["reduce_classify + reduce_warp_kernel<512 slots> (bins <= 256 records)",
                 "reduce_block_kernel<512 threads, 8192 slots> (bins > 3072)",
                 "reduce_block_kernel<256 threads, 4096 slots> (bins <= 3072)",
                 "reduce_block_kernel<128 threads, 2048 slots> (bins <= 1024)", "merge_split_rows_kernel"]
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
            tj = json.load(open(ROOT / "profiles" / "r01_dram_traffic.json"))
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
                                               "owner-direct scatter (records stored into the owner's buffer over NVLink)"),
                                 "record_bytes_total": sent, "per_gpu_per_step": sent / world,
                                 "note": "upper bound: includes the records a rank keeps for itself"}}

    # ---- end to end from pinned host columns ----
    e2e = None
    if args.e2e_steps > 0:
        host = synth.EventFrame(*(t.cpu().pin_memory() for t in (frame.session, frame.aid, frame.ts, frame.type)),
                                n_aids=A)
        h2d = sum(t.numel() * t.element_size() for t in (host.session, host.aid, host.ts, host.type))
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

        def e2e_step():
            f = synth.EventFrame(host.session.to(dev, non_blocking=True), host.aid.to(dev, non_blocking=True),
                                 host.ts.to(dev, non_blocking=True), host.type.to(dev, non_blocking=True), A)
            c = covisit.ingest(f, "desc", device=dev)
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

    # ---- candidate generation from the three graded matrices (second metric of BASELINE.json) ----
    cand_info = None
    if world == 1 and not args.no_candidates:
        tables = {}
        for stem, vspec in covisit.VARIANTS.items():
            tables[stem], _ = covisit.build_topk(csr, vspec)
        test = synth.generate(synth.SynthSpec.scaled("test", args.scale), device=dev)
        sess = covisit.ingest(test, "asc", device=dev)
        gen = candidates.CandidateGenerator(tables, candidates.reference_spec(tables.keys(), 20), A)
        mlen = candidates.max_session_len(sess)
        for _ in range(2):
            gen(sess, mlen)
        torch.cuda.synchronize(dev)
        c0, c1 = ev(), ev()
        c0.record()
        reps = 3
        for _ in range(reps):
            gen(sess, mlen)
        c1.record()
        torch.cuda.synchronize(dev)
        cms = c0.elapsed_time(c1) / reps
        # the whole standalone model: votes + history / popular assembly + recency branch of the long sessions
        popular = {t: list(range(20)) for t in ("click", "cart", "order")}

        def full_model():
            cand = gen(sess, mlen)
            pred, long_s = candidates.assemble_predictions(sess, cand, popular, 20)
            candidates.recency_long_predictions(sess, tables, pred, long_s, 20)
            return long_s
        long_s = full_model()
        torch.cuda.synchronize(dev)
        f0 = time.perf_counter()
        for _ in range(reps):
            full_model()
        torch.cuda.synchronize(dev)
        fms = (time.perf_counter() - f0) / reps * 1e3
        cand_info = {"metric": "candidate_gen_sessions_per_s", "value": sess.n_sessions / (cms * 1e-3),
                     "unit": "sessions/s", "ms": cms, "sessions": sess.n_sessions, "events": sess.n_events,
                     "tables": list(tables), "top_n": 20, "targets": 3,
                     "full_model": {"ms": fms, "sessions_per_s": sess.n_sessions / (fms * 1e-3),
                                    "long_sessions": int(long_s.sum()),
                                    "what": "otto_candidates + otto_assemble_predictions + otto_recency_long, host wall clock"}}

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
                       "l2": "inputs larger than L2 (event CSR and pair records are GBs)"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "candidates": cand_info,
            "gpu_launches": int(launches), "clocks": clocks}))
    if world > 1:
        if peer is not None:
            peer.close()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
