"""Table precompute on the GPU: rolling covariance and the turbulence index.

These feed the envs' tables; in the reference they are per-day pandas loops:

* ``cov_list`` of the portfolio tutorial
  (/root/reference/tutorials/2-Advance/FinRL_PortfolioAllocation_Explainable_DRL.py:157-174): for day
  i >= lookback the 253 closes [i-lookback, i] give 252 pct-change returns whose ``.cov()`` is attached to i.
* ``FeatureEngineer.calculate_turbulence`` (/root/reference/finrl/meta/preprocessor/preprocessors.py:215-267):
  Mahalanobis distance of day i's returns from the mean of the previous 252 days under their covariance
  (``np.linalg.pinv``), zero for the first year and until the third positive value.

Both run in ``csrc/preprocess.cu``: the covariance / mean windows (one block per window) and the D x D
pseudo-inverse quadratic form (one block per day: cyclic Jacobi eigen-solve in shared memory with
``np.linalg.pinv``'s ``rcond`` cut-off on the eigenvalues).  Floating point: pandas' ``.cov`` is a one-pass
Welford update and LAPACK's SVD is not Jacobi, so agreement is ~1e-10 relative for well-conditioned windows,
not bit-exact; rank-deficient windows agree because the same eigenvalues fall under the cut-off.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi


def _returns(close, device):
    import torch

    c = torch.as_tensor(np.asarray(close, dtype=np.float64), device=device)
    ret = torch.full_like(c, float("nan"))
    ret[1:] = c[1:] / c[:-1] - 1.0  # DataFrame.pct_change()
    return ret.contiguous()


def _rolling_cov(ret, first_row, n_rows, n_out, want_mean=False):
    import torch

    T, D = ret.shape
    cov = torch.empty((n_out, D, D), dtype=torch.float64, device=ret.device)
    mean = torch.empty((n_out, D), dtype=torch.float64, device=ret.device) if want_mean else None
    with torch.cuda.device(ret.device):
        _cabi.check(
            _cabi.lib().frl_rolling_cov(_cabi.ptr(ret), T, D, int(first_row), int(n_rows), int(n_out), _cabi.ptr(cov),
                                        _cabi.ptr(mean), _cabi.current_stream(ret.device)),
            "frl_rolling_cov",
        )
    return cov, mean


def rolling_covariance(close, lookback: int = 252, device="cuda"):
    """``cov_list``: tensor [T - lookback, D, D] (f64, on ``device``); entry k belongs to day lookback + k."""
    ret = _returns(close, device)
    T = ret.shape[0]
    if T <= lookback:
        raise ValueError(f"need more than lookback={lookback} days, got {T}")
    cov, _ = _rolling_cov(ret, first_row=1, n_rows=lookback, n_out=T - lookback)
    return cov


def turbulence_index(close, start: int = 252, device="cuda", rcond: float = 1e-15):
    """``calculate_turbulence`` for a complete (NaN-free) close matrix [T, D]: tensor [T] f64.  Window
    statistics by ``frl_rolling_cov``, pseudo-inverse quadratic form by ``frl_turbulence`` (batched Jacobi
    eigen-solve in csrc/preprocess.cu; ``rcond`` is ``np.linalg.pinv``'s default cut-off)."""
    import torch

    ret = _returns(close, device)
    T, D = ret.shape
    out = torch.zeros(T, dtype=torch.float64, device=ret.device)
    if T <= start:
        return out
    # day i uses returns of days [i - start, i); for i == start the first of them (day 0) is the NaN row of
    # pct_change and is dropped by the reference (hist_price.iloc[isna().sum().min():]) -> start - 1 rows
    cov, mean = _rolling_cov(ret, first_row=1, n_rows=start, n_out=T - start - 1, want_mean=True) if T > start + 1 else (None, None)
    cov0, mean0 = _rolling_cov(ret, first_row=1, n_rows=start - 1, n_out=1, want_mean=True)
    covs = (cov0 if cov is None else torch.cat([cov0, cov], dim=0)).contiguous()      # window of day start + k
    means = (mean0 if mean is None else torch.cat([mean0, mean], dim=0)).contiguous()
    scratch = torch.empty(T - start, dtype=torch.float64, device=ret.device)
    with torch.cuda.device(ret.device):
        _cabi.check(
            _cabi.lib().frl_turbulence(_cabi.ptr(ret), T, D, int(start), _cabi.ptr(covs), _cabi.ptr(means), float(rcond),
                                       _cabi.ptr(scratch), _cabi.ptr(out), _cabi.current_stream(ret.device)),
            "frl_turbulence",
        )
    return out
