"""Synthetic market data with the shapes the reference's envs consume.

There is no network on the build or GPU boxes, so every test, fixture and bench line
uses seeded synthetic tables:

* ``close[T, D]``  = 100 * exp(cumsum(N(0, 0.01^2)))               (f64)
* ``tech[K, T, D]``: indicator k ~ N(k + 2, 10^2)                   (f64, continuous,
  so the "disable" flag of env_stocktrading.py:105,174 never fires unless planted)
* ``turbulence[T]`` ~ Gamma(shape 2, scale 20)   (P(>= 99) ~ 4 %)
* OHLV columns for the cash-penalty env: close * (1 + N(0, 0.002^2)), volume U[1e5, 1e7]

The long DataFrame layout matches what ``data_split`` produces
(/root/reference/finrl/meta/preprocessor/preprocessors.py:24-33): rows sorted by
(date, tic), index = day number repeated D times.
"""
from __future__ import annotations

import numpy as np

INDICATORS = [
    "macd",
    "boll_ub",
    "boll_lb",
    "rsi_30",
    "cci_30",
    "dx_30",
    "close_30_sma",
    "close_60_sma",
]


def make_tables(T: int, D: int, K: int, seed: int = 0):
    """Return (close[T,D], tech[K,T,D], turbulence[T]) float64 arrays."""
    rng = np.random.default_rng(seed)
    close = 100.0 * np.exp(np.cumsum(rng.normal(0.0, 0.01, size=(T, D)), axis=0))
    tech = np.stack([rng.normal(k + 2.0, 10.0, size=(T, D)) for k in range(K)], axis=0) if K else np.zeros((0, T, D))
    turb = rng.gamma(2.0, 20.0, size=T)
    return close, tech, turb


def make_ohlv(close: np.ndarray, seed: int = 0):
    """open/high/low/volume columns for the cash-penalty env, each [T, D] f64."""
    rng = np.random.default_rng(seed + 7919)
    o = close * (1.0 + rng.normal(0.0, 0.002, size=close.shape))
    h = close * (1.0 + np.abs(rng.normal(0.0, 0.002, size=close.shape)))
    l = close * (1.0 - np.abs(rng.normal(0.0, 0.002, size=close.shape)))
    v = rng.uniform(1e5, 1e7, size=close.shape)
    return o, h, l, v


def tickers(D: int):
    return [f"S{j:03d}" for j in range(D)]


def dates(T: int):
    # ISO strings sort like dates; business-day spacing is irrelevant to the step path.
    base = np.datetime64("2010-01-04")
    return [str(base + np.timedelta64(t, "D")) for t in range(T)]


def make_frame(close, tech, turb, tech_names=None, risk_col="turbulence", extra_cols=None):
    """Long DataFrame in the layout ``data_split`` produces (index = day number)."""
    import pandas as pd

    T, D = close.shape
    K = tech.shape[0]
    tech_names = list(tech_names) if tech_names is not None else INDICATORS[:K]
    assert len(tech_names) == K
    tic = tickers(D)
    ds = dates(T)
    cols = {
        "date": np.repeat(np.array(ds, dtype=object), D),
        "tic": np.tile(np.array(tic, dtype=object), T),
        "close": close.reshape(-1),
    }
    for k, name in enumerate(tech_names):
        cols[name] = tech[k].reshape(-1)
    cols[risk_col] = np.repeat(turb, D)
    if extra_cols:
        for name, arr in extra_cols.items():
            cols[name] = np.asarray(arr).reshape(-1)
    df = pd.DataFrame(cols)
    df.index = np.repeat(np.arange(T), D)
    return df


def make_np_arrays(close, tech, turb):
    """(price_array[T,D], tech_array[T,D*K] stock-major, turbulence_array[T]) as produced by
    ``df_to_array`` (/root/reference/finrl/meta/data_processors/processor_yahoofinance.py:293-318)."""
    K, T, D = tech.shape
    tech_array = np.transpose(tech, (1, 2, 0)).reshape(T, D * K)  # per stock: its K indicators
    return close.copy(), tech_array.copy(), turb.copy()


def make_cov_table(close, lookback: int = 252):
    """Rolling covariance of daily returns, as the portfolio tutorial builds ``cov_list``
    (/root/reference/tutorials/2-Advance/FinRL_PortfolioAllocation_Explainable_DRL.py:157-174):
    for day i >= lookback the 253 closes [i-lookback, i] give 252 pct-change returns whose
    sample covariance (ddof=1) is attached to day i.  Returns (cov[T-lookback, D, D], first_day)."""
    T, D = close.shape
    out = np.empty((T - lookback, D, D))
    for i in range(lookback, T):
        win = close[i - lookback : i + 1]
        ret = win[1:] / win[:-1] - 1.0
        out[i - lookback] = np.cov(ret, rowvar=False, ddof=1)
    return out, lookback


def make_actions(shape, seed: int = 1, low=-1.0, high=1.0, dtype=np.float32):
    rng = np.random.default_rng(seed)
    return rng.uniform(low, high, size=shape).astype(dtype)
