"""Batched StockTradingEnvStopLoss: N copies of the reference stop-loss env on one GPU.

Reference: /root/reference/finrl/meta/env_stock_trading/env_stocktrading_stoploss.py.  The cash-penalty
env plus: an incrementally tracked average buy price per asset (:416-428), forced liquidation of assets
whose close fell below ``stoploss_penalty`` x that average while cash is above ``stoploss_penalty`` x the
initial amount (:354-361), and a reward (:255-290) that subtracts the stop-loss and low-profit penalties
and adds the profits of sells above ``min_profit_penalty`` x the average.  Kernel: finrl_b200/csrc/stoploss.cu.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _cabi
from ._base import BatchedEnvBase
from .cashpenalty import CashPenaltyTables, frame_to_cashpenalty_arrays


class BatchedStockTradingEnvStopLoss(BatchedEnvBase):
    """Keywords mirror ``StockTradingEnvStopLoss.__init__`` (:64-84); extra: ``n_envs``, ``device``, ``tables``
    (the same :class:`CashPenaltyTables` the cash-penalty env uses)."""

    _PREFIX = "frl_stoploss"
    _STAT7_NAME = "shortage_count"
    _ACTION_NAME = "n_assets"

    def __init__(self, df=None, buy_cost_pct=3e-3, sell_cost_pct=3e-3, date_col_name="date", hmax=10,
                 discrete_actions=False, shares_increment=1, stoploss_penalty=0.9, profit_loss_ratio=2,
                 turbulence_threshold=None, print_verbosity=10, initial_amount=1e6,
                 daily_information_cols=("open", "close", "high", "low", "volume"), cache_indicator_data=True,
                 cash_penalty_proportion=0.1, random_start=True, patient=False, currency="$",
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
        self.stoploss_penalty = stoploss_penalty
        self.min_profit_penalty = 1 + profit_loss_ratio * (1 - stoploss_penalty)  # (:101)
        self.daily_information_cols = list(daily_information_cols)
        dev = self.device
        f64 = dict(dtype=torch.float64, device=dev)
        self.cash = torch.empty(N, **f64)
        # two buffers of the six stock-major per-asset arrays (holdings, previous holdings, average buy price,
        # buy counts, closing_diff_avg_buy, profit_sell_diff_avg_buy); bit 1 of `fresh` names an env's current one
        self._assets = torch.zeros((2, 6, D, N), **f64)
        self.date_index = torch.empty(N, dtype=torch.int32, device=dev)
        self.starting_point = torch.empty(N, dtype=torch.int32, device=dev)
        self.fresh = torch.empty(N, dtype=torch.uint8, device=dev)
        self.last_cash = torch.empty(N, **f64)
        self.last_total = torch.empty(N, **f64)
        self.sum_trades = torch.empty(N, **f64)
        self._stats_block = _cabi.new_stats_block(torch, self.device)
        self.stats = self._stats_block[:_cabi.N_STATS]
        self._obs = torch.empty((N, O), dtype=torch.float32, device=dev)
        self._rew = torch.empty(N, **f64)
        self._flags = torch.empty(N, dtype=torch.uint8, device=dev)
        p = _cabi.StopLossParams()
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
        p.stoploss_penalty, p.min_profit_penalty = float(stoploss_penalty), float(self.min_profit_penalty)
        p.close, p.turb, p.obs_tmpl = tables.close.data_ptr(), tables.turb.data_ptr(), tables.obs_tmpl.data_ptr()
        p.cash, p.date_index, p.start = self.cash.data_ptr(), self.date_index.data_ptr(), self.starting_point.data_ptr()
        p.assets = self._assets.data_ptr()
        p.fresh, p.last_cash, p.last_total, p.sum_trades = self.fresh.data_ptr(), self.last_cash.data_ptr(), self.last_total.data_ptr(), self.sum_trades.data_ptr()
        p.random_start = int(self.random_start)  # in-kernel auto-reset redraws the starting point (reset_seed per launch)
        self._p = p
        self.reset()

    def _asset_array(self, idx):
        """Array ``idx`` of every env's current buffer, stock-major [D, N]."""
        use_alt = (self.fresh & 2).bool()
        return self._torch.where(use_alt[None, :], self._assets[1, idx], self._assets[0, idx])

    hold = property(lambda self: self._asset_array(0))
    prev_hold = property(lambda self: self._asset_array(1))
    avg_buy = property(lambda self: self._asset_array(2))
    n_buys = property(lambda self: self._asset_array(3))
    cdiff = property(lambda self: self._asset_array(4))
    pdiff = property(lambda self: self._asset_array(5))

    @property
    def holdings(self):
        """Holdings in the natural [N, D] layout."""
        return self.hold.t()

    def reset(self, mask=None, start_points=None, out=None):
        torch = self._torch
        out = self._obs if out is None else out
        if start_points is None and self.random_start:
            hi = int(self.n_days * 0.5)  # random.choice(range(int(len(self.dates) * 0.5))) (:138-140)
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
                _cabi.lib().frl_stoploss_reset(C.byref(self._p), _cabi.ptr(mask), _cabi.ptr(sp), _cabi.ptr(out), self._stream()),
                "frl_stoploss_reset",
            )
        self.launches += 2
        return out
