"""Shared comparison helpers for the parity tests (oracle = checker, never the thing under test)."""
import numpy as np
import pandas as pd

from oracle import covisit_oracle as co

RTOL_TIME = 1e-5   # north_star: time-weighted sums agree within 1e-5 relative


def oracle_spec(spec) -> co.OracleSpec:
    """Product CovisitSpec -> OracleSpec (same fields, float type weights)."""
    return co.OracleSpec(weight_mode=spec.weight_mode, type_weight=tuple(float(w) for w in spec.type_weight),
                         event_types=tuple(spec.event_types), x_types=tuple(spec.x_types), y_types=tuple(spec.y_types),
                         window_s=spec.window_s, tail_n=spec.tail_n, k=spec.k, ts_min=spec.ts_min, ts_max=spec.ts_max)


def gpu_formula_topk(acc: pd.DataFrame, spec) -> pd.DataFrame:
    """Top-K of the oracle's exact integer accumulators under the GPU's weight definition
    wgt = float32(cnt + 3 * tsum / (ts_max - ts_min)); ties by aid_y ascending."""
    w = (acc["cnt"].to_numpy().astype(np.float64)
         + (3.0 / float(spec.ts_max - spec.ts_min)) * acc["tsum"].to_numpy().astype(np.float64)).astype(np.float32)
    t = acc.assign(wgt=w)
    return co.topk(t[["aid_x", "aid_y", "wgt", "cnt", "tsum"]], spec.k)


def assert_int_table_equal(got: pd.DataFrame, want: pd.DataFrame, what: str) -> None:
    assert len(got) == len(want), f"{what}: {len(got)} rows vs oracle {len(want)}"
    for c in ("aid_x", "aid_y"):
        assert np.array_equal(got[c].to_numpy(), want[c].to_numpy()), f"{what}: column {c} differs"
    assert np.array_equal(got["wgt"].to_numpy().view(np.uint32), want["wgt"].to_numpy().view(np.uint32)), \
        f"{what}: weights differ"


def assert_time_table_close(got: pd.DataFrame, acc: pd.DataFrame, k: int, what: str) -> None:
    """got = GPU rows; acc = oracle accumulate() rows (float32 pandas sums).  Every GPU pair must carry the
    oracle's weight within RTOL_TIME, and the GPU row set may differ from the oracle's top-K only where the
    oracle's weights tie with the K-th weight within that tolerance."""
    key = lambda d: d["aid_x"].to_numpy().astype(np.int64) << 32 | d["aid_y"].to_numpy().astype(np.int64)
    ow = pd.Series(acc["wgt"].to_numpy(), index=key(acc))
    gk = key(got)
    assert ow.index.is_unique
    missing = ~np.isin(gk, ow.index.to_numpy())
    assert not missing.any(), f"{what}: GPU emitted pairs the oracle never saw"
    w_or = ow.loc[gk].to_numpy()
    np.testing.assert_allclose(got["wgt"].to_numpy(), w_or, rtol=RTOL_TIME, atol=0, err_msg=what)
    want = co.topk(acc, k)
    assert len(got) == len(want), f"{what}: {len(got)} rows vs oracle {len(want)}"
    assert np.array_equal(got["aid_x"].to_numpy(), want["aid_x"].to_numpy()), what
    diff = np.nonzero(got["aid_y"].to_numpy() != want["aid_y"].to_numpy())[0]
    if diff.size:
        # any disagreement must be a swap among near-tied weights
        w_want = want["wgt"].to_numpy()[diff]
        w_got = w_or[diff]
        np.testing.assert_allclose(w_got, w_want, rtol=2 * RTOL_TIME, atol=0,
                                   err_msg=f"{what}: top-K differs beyond weight ties")
