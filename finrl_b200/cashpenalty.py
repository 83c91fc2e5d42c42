"""Batched StockTradingEnvCashpenalty: N copies of the reference cash-penalty env on one GPU.

Reference: /root/reference/finrl/meta/env_stock_trading/env_stocktrading_cashpenalty.py.  Fractional
(or discretised) share trading with ``hmax`` in currency, reward = cash-penalised gain per elapsed
step computed BEFORE trading, CASH SHORTAGE termination or ``patient`` mode (incl. quirk Q9), and
turbulence liquidation.  One thread per env, one streaming pass over the assets with ping-pong holdings buffers (finrl_b200/csrc/cashpenalty.cu).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _cabi
from ._base import BatchedEnvBase


def frame_to_cashpenalty_arrays(df, daily_information_cols: Sequence[str], date_col_name="date", stock_col="tic"):
    """The reference's view of the frame: ``dates = df[date].sort_values().unique()``, ``assets =
    df.tic.unique()`` (order of first appearance) and ``get_date_vector`` = for each asset its columns
    (:73-74, :160-173).  Returns close[T,D], info[T,D,C], turbulence[T] or None, dates, assets."""
    assets = list(df[stock_col].unique())
    dates = np.sort(df[date_col_name].unique())
    T, D = len(dates), len(assets)
    if len(df) != T * D:
        raise ValueError(f"frame has {len(df)} rows, expected n_dates*n_assets = {T}*{D}")
    piv = df.set_index([date_col_name, stock_col]).sort_index()

    def col(name):
        return np.ascontiguousarray(piv[name].unstack(stock_col)[assets].to_numpy(dtype=np.float64))

    close = col("close")
    info = np.stack([col(c) for c in daily_information_cols], axis=2) if len(daily_information_cols) else np.zeros((T, D, 0))
    turb = col("turbulence")[:, 0] if "turbulence" in df.columns else None
    return close, info, turb, dates, assets


@dataclass
class CashPenaltyTables:
    n_days: int
    stock_dim: int
    n_cols: int
    close: "torch.Tensor"     # [T, D] f64
    turb: "torch.Tensor"      # [T] f64
    obs_tmpl: "torch.Tensor"  # [T, O] f32
    close_rc: "torch.Tensor" = None  # [T, D, 2] f64: (close, correctly rounded 1 / close) pairs

    @property
    def obs_dim(self) -> int:
        return 1 + self.stock_dim + self.stock_dim * self.n_cols

    @staticmethod
    def from_arrays(close, info, turb, device) -> "CashPenaltyTables":
        import torch

        close = np.array(close, dtype=np.float64, order="C")  # own, writable copy
        T, D = close.shape
        if not 1 <= D <= 128:
            raise ValueError(f"number of assets must be in 1..128 for the cash-penalty kernel (got {D})")
        info = np.ascontiguousarray(info, dtype=np.float64).reshape(T, D, -1)
        Cc = info.shape[2]
        O = 1 + D + D * Cc
        tmpl = np.zeros((T, O), dtype=np.float32)
        tmpl[:, 1 + D :] = info.reshape(T, D * Cc).astype(np.float32)
        turb = np.zeros(T) if turb is None else np.ascontiguousarray(turb, dtype=np.float64)
        dev = torch.device(device)
        with np.errstate(divide="ignore", invalid="ignore"):
            close_rc = np.stack([close, 1.0 / close], axis=2)  # IEEE division: correctly rounded reciprocals (1/0 = inf)
        return CashPenaltyTables(n_days=T, stock_dim=D, n_cols=Cc, close=torch.from_numpy(close).to(dev),
                                 turb=torch.from_numpy(turb).to(dev), obs_tmpl=torch.from_numpy(tmpl).to(dev),
                                 close_rc=torch.from_numpy(np.ascontiguousarray(close_rc)).to(dev))


class BatchedStockTradingEnvCashpenalty(BatchedEnvBase):
    """Keywords mirror ``StockTradingEnvCashpenalty.__init__`` (:52-70); extra: ``n_envs``, ``device``,
    ``tables``.  ``random_start`` draws per-env starting points with Python's ``random`` like the
    reference when N is small, on the device otherwise; parity runs use ``random_start=False``.  ``step`` /
    ``rollout`` / ``observe`` come from :class:`BatchedEnvBase`."""

    _PREFIX = "frl_cashpenalty"
    _STAT7_NAME = "shortage_count"
    _ACTION_NAME = "n_assets"

    def __init__(self, df=None, buy_cost_pct=3e-3, sell_cost_pct=3e-3, date_col_name="date", hmax=10,
                 discrete_actions=False, shares_increment=1, turbulence_threshold=None, print_verbosity=10,
                 initial_amount=1e6, daily_information_cols=("open", "close", "high", "low", "volume"),
                 cache_indicator_data=True, cash_penalty_proportion=0.1, random_start=True, patient=False, currency="$",
                 *, n_envs=1, device="cuda", tables: Optional[CashPenaltyTables] = None):
        torch = self._bind_device(device)
        self.df = df
        self.dates = self.assets = None
        if tables is None:
            if df is None:
                raise ValueError("either df or tables is required")
            close, info, turb, self.dates, self.assets = frame_to_cashpenalty_arrays(df, list(daily_information_cols), date_col_name)
            if turbulence_threshold is not None and turb is None:
                raise KeyError("turbulence")
            tables = CashPenaltyTables.from_arrays(close, info, turb, self.device)
        self.tables = tables
        D, T, O = tables.stock_dim, tables.n_days, tables.obs_dim
        N = int(n_envs)
        self.n_envs, self.stock_dim, self.n_days, self.state_space = N, D, T, O
        self.random_start, self.patient, self.discrete_actions = bool(random_start), bool(patient), bool(discrete_actions)
        self.hmax, self.initial_amount = hmax, initial_amount
        self.buy_cost_pct, self.sell_cost_pct = buy_cost_pct, sell_cost_pct
        self.turbulence_threshold, self.cash_penalty_proportion = turbulence_threshold, cash_penalty_proportion
        self.shares_increment = shares_increment
        self.daily_information_cols = list(daily_information_cols)
        dev = self.device
        self.cash = torch.empty(N, dtype=torch.float64, device=dev)
        self.hold = torch.empty((D, N), dtype=torch.float64, device=dev)      # stock-major, buffer 0
        self.hold_alt = torch.zeros((D, N), dtype=torch.float64, device=dev)  # buffer 1 (ping-pong, see the kernel)
        self.date_index = torch.empty(N, dtype=torch.int32, device=dev)
        self.starting_point = torch.empty(N, dtype=torch.int32, device=dev)
        self.fresh = torch.empty(N, dtype=torch.uint8, device=dev)
        self.last_cash = torch.empty(N, dtype=torch.float64, device=dev)
        self.last_total = torch.empty(N, dtype=torch.float64, device=dev)
        self.sum_trades = torch.empty(N, dtype=torch.float64, device=dev)
        self._stats_block = _cabi.new_stats_block(torch, dev)
        self.stats = self._stats_block[:_cabi.N_STATS]
        self._obs = torch.empty((N, O), dtype=torch.float32, device=dev)
        self._rew = torch.empty(N, dtype=torch.float64, device=dev)
        self._flags = torch.empty(N, dtype=torch.uint8, device=dev)
        p = _cabi.CashPenaltyParams()
        p.n_envs, p.stock_dim, p.n_cols, p.n_days, p.obs_dim, p.env_stride = N, D, tables.n_cols, T, O, N
        p.discrete_actions, p.shares_increment = int(bool(discrete_actions)), int(shares_increment)
        p.use_turbulence, p.patient = int(turbulence_threshold is not None), int(bool(patient))
        p.buy_cost_pct, p.sell_cost_pct = float(buy_cost_pct), float(sell_cost_pct)
        # scalar hmax (weak Python float: the product keeps the action dtype) or a per-asset array (numpy
        # array-array promotion: float32 actions * float64 hmax -> float64) — `actions * self.hmax` (:268)
        self._hmax_vec = None
        if np.isscalar(hmax):
            p.hmax = float(hmax)
        else:
            hv = np.asarray(hmax)
            if hv.shape != (D,):
                raise ValueError(f"hmax array must have shape ({D},), got {hv.shape}")
            self._hmax_vec = torch.as_tensor(hv.astype(np.float64), device=dev)
            p.hmax, p.hmax_vec, p.hmax_vec_f32 = 0.0, self._hmax_vec.data_ptr(), int(hv.dtype == np.float32)
        p.turbulence_threshold = float(turbulence_threshold) if turbulence_threshold is not None else 0.0
        p.initial_amount, p.cash_penalty_proportion = float(initial_amount), float(cash_penalty_proportion)
        p.close, p.turb, p.obs_tmpl = tables.close.data_ptr(), tables.turb.data_ptr(), tables.obs_tmpl.data_ptr()
        import os

        use_rc = tables.close_rc is not None and os.environ.get("FRL_CP_NO_RCLOSE") != "1"  # A/B switch
        p.close_rc = tables.close_rc.data_ptr() if use_rc else None
        p.cash, p.hold, p.date_index, p.start = self.cash.data_ptr(), self.hold.data_ptr(), self.date_index.data_ptr(), self.starting_point.data_ptr()
        p.hold_alt = self.hold_alt.data_ptr()
        p.fresh, p.last_cash, p.last_total, p.sum_trades = self.fresh.data_ptr(), self.last_cash.data_ptr(), self.last_total.data_ptr(), self.sum_trades.data_ptr()
        p.random_start = int(self.random_start)  # in-kernel auto-reset redraws the starting point (reset_seed per launch)
        self._p = p
        self.reset()

    @property
    def holdings(self):
        """Holdings in the natural [N, D] layout: per env the current one of the two stock-major buffers."""
        use_alt = (self.fresh & 2).bool()
        return self._torch.where(use_alt[None, :], self.hold_alt, self.hold).t()

    def reset(self, mask=None, start_points=None, out=None):
        torch = self._torch
        out = self._obs if out is None else out
        if start_points is None and self.random_start:
            hi = int(self.n_days * 0.5)  # random.choice(range(int(len(self.dates) * 0.5))) (:135-137)
            if self.n_envs <= 4096:
                import random

                start_points = [random.choice(range(hi)) for _ in range(self.n_envs)]
            else:
                start_points = torch.randint(0, max(hi, 1), (self.n_envs,), device=self.device, dtype=torch.int32)
        sp = None
        if start_points is not None:
            sp = torch.as_tensor(start_points, device=self.device).to(torch.int32).contiguous()
        mask = self._mask(mask)
        with torch.cuda.device(self.device):
            _cabi.check(
                _cabi.lib().frl_cashpenalty_reset(C.byref(self._p), _cabi.ptr(mask), _cabi.ptr(sp), _cabi.ptr(out), self._stream()),
                "frl_cashpenalty_reset",
            )
        self.launches += 2
        return out
