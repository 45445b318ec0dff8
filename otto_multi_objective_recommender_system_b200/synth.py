"""Synthetic OTTO-shaped event frames (SURVEY.md Appendix C calibration targets).

The Kaggle dataset is not available offline, so every test and benchmark runs on frames made
here.  The generator is counter-based (splitmix64 over event / session indices) and uses only
integer tensor ops plus ``searchsorted`` against integer threshold tables that are computed once
on the host in float64.  The same call therefore yields bit-identical frames on ``cpu`` and on
``cuda`` devices, which lets the GPU parity tests and the CPU oracle look at the same events.

Calibration sources (all reference artefacts, relative to /root/reference):
  * session length train 16.80 / 6 / 33.58 / [2, 500], test 4.14 / 2 / 8.22 / [1, 458]
    (eda/session_count_distribution.png)
  * type mix train 89.85 / 7.80 / 2.35 %, test 90.83 / 8.23 / 0.95 %  (EDA notebook cell 5)
  * ts span 1659304800 .. 1661723999 (train), .. 1662328791 (test)   (EDA notebook cell 6)
  * 12,899,779 sessions / 216,716,096 events / 1,855,603 aids          (EDA notebook cell 5)
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

TS_TRAIN_MIN = 1659304800
TS_TRAIN_MAX = 1661723999
TS_TEST_MIN = 1661724000
TS_TEST_MAX = 1662328791

FULL_TRAIN_SESSIONS = 12_899_779
FULL_TEST_SESSIONS = 1_671_803
FULL_AIDS = 1_855_603

_M64 = (1 << 64) - 1


def _s64(x: int) -> int:
    """Python int (mod 2^64) -> the same bit pattern as a signed int64."""
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


_C1 = _s64(0xBF58476D1CE4E5B9)
_C2 = _s64(0x94D049BB133111EB)
_GOLD = _s64(0x9E3779B97F4A7C15)


def _lsr(x: torch.Tensor, k: int) -> torch.Tensor:
    """Logical shift right on int64 tensors (torch's >> is arithmetic)."""
    return (x >> k) & ((1 << (64 - k)) - 1)


def _mix(x: torch.Tensor) -> torch.Tensor:
    """splitmix64 finaliser; int64 multiplication wraps, which is what we want."""
    x = (x ^ _lsr(x, 30)) * _C1
    x = (x ^ _lsr(x, 27)) * _C2
    return x ^ _lsr(x, 31)


def _stream(idx: torch.Tensor, seed: int, stream: int) -> torch.Tensor:
    """64 random bits per index for a named stream; pure function of (seed, stream, idx)."""
    base = _s64((seed * 0x632BE59BD9B4E019 + stream * 0xD1342543DE82EF95) & _M64)
    return _mix(_mix(idx * _GOLD + base))


def _u32(bits: torch.Tensor) -> torch.Tensor:
    return _lsr(bits, 32)


def _u40(bits: torch.Tensor) -> torch.Tensor:
    return _lsr(bits, 24)


def _cdf_thresholds(p: np.ndarray, bits: int) -> np.ndarray:
    """Integer CDF thresholds: value v is drawn when thr[v-1] <= u < thr[v] for u uniform in [0, 2^bits)."""
    c = np.cumsum(p.astype(np.float64))
    c /= c[-1]
    thr = np.floor(c * float(1 << bits)).astype(np.int64)
    thr[-1] = 1 << bits
    return thr


def _lognormal_length_pmf(mu: float, sigma: float, lo: int, hi: int) -> np.ndarray:
    """pmf of clip(round(LogNormal(mu, sigma)), lo, hi) over lo..hi."""
    def cdf(x: float) -> float:
        if x <= 0:
            return 0.0
        return 0.5 * (1.0 + math.erf((math.log(x) - mu) / (sigma * math.sqrt(2.0))))
    ks = np.arange(lo, hi + 1)
    p = np.array([cdf(k + 0.5) - cdf(k - 0.5) for k in ks])
    p[0] = cdf(lo + 0.5)
    p[-1] = 1.0 - cdf(hi - 0.5)
    return p


def _exp_quantiles(mean_s: float, n: int) -> np.ndarray:
    """n-point integer quantile table of Exp(mean_s) seconds (floor)."""
    q = (np.arange(n, dtype=np.float64) + 0.5) / n
    return np.floor(-mean_s * np.log1p(-q)).astype(np.int64)


@dataclass
class SynthSpec:
    kind: str = "train"            # "train" | "test"
    n_sessions: int = 129_000
    n_aids: int = 18_556
    seed: int = 42
    first_session: int = 0
    zipf_s: float = 0.8
    zipf_q: float = 12.0
    p_repeat: float = 0.18         # event repeats an earlier aid of its session
    p_local: float = 0.60          # event drawn near the session centre (rank space)

    @staticmethod
    def scaled(kind: str, fraction: float, seed: int | None = None) -> "SynthSpec":
        """OTTO-shaped spec at a fraction of full scale (sessions and aids scale together)."""
        if kind == "train":
            return SynthSpec("train", max(2, round(FULL_TRAIN_SESSIONS * fraction)),
                             max(64, round(FULL_AIDS * fraction)), 42 if seed is None else seed, 0)
        if kind == "test":
            return SynthSpec("test", max(2, round(FULL_TEST_SESSIONS * fraction)),
                             max(64, round(FULL_AIDS * fraction)), 43 if seed is None else seed,
                             FULL_TRAIN_SESSIONS)
        raise ValueError("Invalid kind")


@dataclass
class EventFrame:
    """(session, aid, ts, type) columns, sorted by (session, ts) like the reference's chunk files
    (utilities/split_dataset_writer_parquet.py:17); ts in seconds."""
    session: torch.Tensor   # int32
    aid: torch.Tensor       # int32
    ts: torch.Tensor        # int32
    type: torch.Tensor      # uint8
    n_aids: int

    def __len__(self) -> int:
        return int(self.session.numel())

    def to(self, device) -> "EventFrame":
        return EventFrame(self.session.to(device), self.aid.to(device), self.ts.to(device),
                          self.type.to(device), self.n_aids)

    def to_pandas(self):
        import pandas as pd
        return pd.DataFrame({
            "session": self.session.cpu().numpy(), "aid": self.aid.cpu().numpy(),
            "ts": self.ts.cpu().numpy(), "type": self.type.cpu().numpy().astype(np.int8)})

    @staticmethod
    def from_pandas(df, n_aids: int | None = None) -> "EventFrame":
        n_aids = int(df["aid"].max()) + 1 if n_aids is None else n_aids
        return EventFrame(torch.from_numpy(df["session"].to_numpy().astype(np.int32)),
                          torch.from_numpy(df["aid"].to_numpy().astype(np.int32)),
                          torch.from_numpy(df["ts"].to_numpy().astype(np.int32)),
                          torch.from_numpy(df["type"].to_numpy().astype(np.uint8)), n_aids)


def generate(spec: SynthSpec, device: str | torch.device = "cpu") -> EventFrame:
    """Make one frame.  All randomness is a pure function of (spec.seed, index)."""
    dev = torch.device(device)
    train = spec.kind == "train"
    if spec.kind not in ("train", "test"):
        raise ValueError("Invalid kind")
    S, A, seed = spec.n_sessions, spec.n_aids, spec.seed

    # ---- host-side tables (float64 once, then integers) ----
    if train:
        len_lo, len_hi = 2, 500
        pmf = _lognormal_length_pmf(1.79, 1.43, len_lo, len_hi)
        type_p = np.array([0.8985, 0.0780, 0.0235])
        ts_lo, ts_hi = TS_TRAIN_MIN, TS_TRAIN_MAX
    else:
        len_lo, len_hi = 1, 458
        pmf = _lognormal_length_pmf(0.62, 1.22, len_lo, len_hi)
        type_p = np.array([0.9083, 0.0823, 0.0094])
        ts_lo, ts_hi = TS_TEST_MIN, TS_TEST_MAX
    len_thr = torch.from_numpy(_cdf_thresholds(pmf, 32)).to(dev)
    type_thr = torch.from_numpy(_cdf_thresholds(type_p, 32)[:2].copy()).to(dev)
    ranks = np.arange(A, dtype=np.float64)
    zm_thr = torch.from_numpy(_cdf_thresholds(1.0 / np.power(ranks + spec.zipf_q, spec.zipf_s), 40)).to(dev)
    gap_short = torch.from_numpy(_exp_quantiles(60.0, 4096)).to(dev)
    gap_long = torch.from_numpy(_exp_quantiles(3 * 86400.0, 4096)).to(dev)
    # offset magnitude scale g: P(g = k) = 2^-(k+1); thresholds on a 32-bit uniform
    geo_thr = torch.tensor([(1 << 32) - (1 << (31 - k)) for k in range(20)], dtype=torch.int64, device=dev)
    gperm = torch.Generator(device="cpu").manual_seed(seed * 7919 + 11)
    perm = torch.randperm(A, generator=gperm).to(dev)            # rank -> aid id

    # ---- sessions ----
    sidx = torch.arange(S, dtype=torch.int64, device=dev)
    length = len_lo + torch.searchsorted(len_thr, _u32(_stream(sidx, seed, 1)), right=True)
    offsets = torch.zeros(S + 1, dtype=torch.int64, device=dev)
    torch.cumsum(length, 0, out=offsets[1:])
    E = int(offsets[-1].item())
    centre = torch.searchsorted(zm_thr, _u40(_stream(sidx, seed, 2)), right=True).clamp_(max=A - 1)

    # ---- events ----
    eidx = torch.arange(E, dtype=torch.int64, device=dev)
    sess = torch.repeat_interleave(sidx, length, output_size=E)
    j = eidx - offsets[sess]                                      # position inside the session

    def base_rank(idx: torch.Tensor, s: torch.Tensor) -> torch.Tensor:
        """Rank drawn for event idx ignoring repeats: local (around the centre) or global."""
        sel = _u32(_stream(idx, seed, 4))
        glob = torch.searchsorted(zm_thr, _u40(_stream(idx, seed, 5)), right=True).clamp_(max=A - 1)
        g = torch.searchsorted(geo_thr, _u32(_stream(idx, seed, 6)), right=True)
        raw = _stream(idx, seed, 7)
        mag = _u32(raw) & ((8 << g) - 1)                          # heavy-tailed offset, P(off > m) ~ 1/m
        sign = 1 - 2 * (raw & 1)
        loc = torch.remainder(centre[s] + sign * (mag + 1), A)
        is_local = sel < int(spec.p_local * (1 << 32))
        return torch.where(is_local, loc, glob)

    rep = _u32(_stream(eidx, seed, 8)) < int(spec.p_repeat * (1 << 32))
    rep &= j > 0
    k = _u32(_stream(eidx, seed, 9)) % j.clamp(min=1)             # earlier event to copy
    src = torch.where(rep, offsets[sess] + k, eidx)
    rank = base_rank(src, sess)
    aid = perm[rank]

    tu = _u32(_stream(eidx, seed, 10))
    etype = (tu >= type_thr[0]).to(torch.uint8) + (tu >= type_thr[1]).to(torch.uint8)

    gu = _stream(eidx, seed, 11)
    long_gap = (_u32(gu) % 10) == 0
    q = _lsr(gu, 8) & 4095
    gap = torch.where(long_gap, gap_long[q], gap_short[q])
    gap = torch.where(j == 0, torch.zeros_like(gap), gap)
    cg = torch.cumsum(gap, 0)
    span = cg[offsets[1:] - 1] - cg[offsets[:-1]]                 # seconds from first to last event
    cg = cg - cg[offsets[sess]]
    # sessions whose span exceeds the data range are compressed into it; all start where they still fit
    R = ts_hi - ts_lo
    too_long = span > R
    cg = torch.where(too_long[sess], (cg * R) // span.clamp(min=1)[sess], cg)
    span = torch.where(too_long, torch.full_like(span, R), span)
    room = (R + 1 - span).clamp_(min=1)
    start = ts_lo + _u32(_stream(sidx, seed, 3)) % room
    ts = (start[sess] + cg).clamp_(max=ts_hi)

    return EventFrame((sess + spec.first_session).to(torch.int32), aid.to(torch.int32),
                      ts.to(torch.int32), etype, A)


def frame_stats(frame: EventFrame) -> dict:
    """Achieved statistics to publish next to each run (SURVEY.md Appendix C asks for this)."""
    sess = frame.session.cpu().numpy()
    _, counts = np.unique(sess, return_counts=True)
    aid_counts = np.bincount(frame.aid.cpu().numpy(), minlength=frame.n_aids)
    t = np.bincount(frame.type.cpu().numpy(), minlength=3) / max(1, len(sess))
    return {
        "sessions": int(counts.size), "events": int(sess.size), "aids_seen": int((aid_counts > 0).sum()),
        "len_mean": float(counts.mean()), "len_median": float(np.median(counts)), "len_std": float(counts.std()),
        "len_min": int(counts.min()), "len_max": int(counts.max()),
        "tail30_events": int(np.minimum(counts, 30).sum()),
        "raw_join_rows": int((np.minimum(counts, 30).astype(np.int64) ** 2).sum()),
        "type_mix": [float(x) for x in t],
        "top_aid_share": float(aid_counts.max() / max(1, sess.size)),
    }
