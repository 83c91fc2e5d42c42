"""Batched StockTradingEnvCashpenalty: N copies of the reference cash-penalty env on one GPU.

Reference: /root/reference/finrl/meta/env_stock_trading/env_stocktrading_cashpenalty.py.  Fractional
(or discretised) share trading with ``hmax`` in currency, reward = cash-penalised gain per elapsed
step computed BEFORE trading, CASH SHORTAGE termination or ``patient`` mode (incl. quirk Q9), and
turbulence liquidation.  One thread per env, two streaming passes over the assets (finrl_b200/csrc/cashpenalty.cu).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _cabi


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

    @property
    def obs_dim(self) -> int:
        return 1 + self.stock_dim + self.stock_dim * self.n_cols

    @staticmethod
    def from_arrays(close, info, turb, device) -> "CashPenaltyTables":
        import torch

        close = np.ascontiguousarray(close, dtype=np.float64)
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
        return CashPenaltyTables(n_days=T, stock_dim=D, n_cols=Cc, close=torch.from_numpy(close).to(dev),
                                 turb=torch.from_numpy(turb).to(dev), obs_tmpl=torch.from_numpy(tmpl).to(dev))


class BatchedStockTradingEnvCashpenalty:
    """Keywords mirror ``StockTradingEnvCashpenalty.__init__`` (:52-70); extra: ``n_envs``, ``device``,
    ``tables``.  ``random_start`` draws per-env starting points with Python's ``random`` like the
    reference when N is small, on the device otherwise; parity runs use ``random_start=False``."""

    def __init__(self, df=None, buy_cost_pct=3e-3, sell_cost_pct=3e-3, date_col_name="date", hmax=10,
                 discrete_actions=False, shares_increment=1, turbulence_threshold=None, print_verbosity=10,
                 initial_amount=1e6, daily_information_cols=("open", "close", "high", "low", "volume"),
                 cache_indicator_data=True, cash_penalty_proportion=0.1, random_start=True, patient=False, currency="$",
                 *, n_envs=1, device="cuda", tables: Optional[CashPenaltyTables] = None):
        import torch

        self._torch = torch
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _cabi.EngineError("finrl_b200 runs on CUDA devices only (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        _cabi.lib()
        if not np.isscalar(hmax):
            raise NotImplementedError("per-asset hmax arrays are not supported yet (scalar hmax only)")
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
        self.stats = torch.zeros(_cabi.N_STATS, dtype=torch.float64, device=dev)
        self._obs = torch.empty((N, O), dtype=torch.float32, device=dev)
        self._rew = torch.empty(N, dtype=torch.float64, device=dev)
        self._flags = torch.empty(N, dtype=torch.uint8, device=dev)
        p = _cabi.CashPenaltyParams()
        p.n_envs, p.stock_dim, p.n_cols, p.n_days, p.obs_dim, p.env_stride = N, D, tables.n_cols, T, O, N
        p.discrete_actions, p.shares_increment = int(bool(discrete_actions)), int(shares_increment)
        p.use_turbulence, p.patient = int(turbulence_threshold is not None), int(bool(patient))
        p.buy_cost_pct, p.sell_cost_pct, p.hmax = float(buy_cost_pct), float(sell_cost_pct), float(hmax)
        p.turbulence_threshold = float(turbulence_threshold) if turbulence_threshold is not None else 0.0
        p.initial_amount, p.cash_penalty_proportion = float(initial_amount), float(cash_penalty_proportion)
        p.close, p.turb, p.obs_tmpl = tables.close.data_ptr(), tables.turb.data_ptr(), tables.obs_tmpl.data_ptr()
        p.cash, p.hold, p.date_index, p.start = self.cash.data_ptr(), self.hold.data_ptr(), self.date_index.data_ptr(), self.starting_point.data_ptr()
        p.hold_alt = self.hold_alt.data_ptr()
        p.fresh, p.last_cash, p.last_total, p.sum_trades = self.fresh.data_ptr(), self.last_cash.data_ptr(), self.last_total.data_ptr(), self.sum_trades.data_ptr()
        self._p = p
        self.launches = 0
        self.kernel_events = None
        self.reset()

    def _stream(self):
        return _cabi.current_stream(self.device)

    @property
    def holdings(self):
        """Holdings in the natural [N, D] layout: per env the current one of the two stock-major buffers."""
        use_alt = (self.fresh & 2).bool()
        return self._torch.where(use_alt[None, :], self.hold_alt, self.hold).t()

    def _as_actions(self, actions, ndim):
        torch = self._torch
        if not isinstance(actions, torch.Tensor):
            actions = torch.as_tensor(np.asarray(actions))
        if actions.dtype not in (torch.float32, torch.float64):
            actions = actions.to(torch.float32)
        if actions.device != self.device:
            actions = actions.to(self.device, non_blocking=True)
        if actions.dim() != ndim or actions.shape[-1] != self.stock_dim:
            raise ValueError(f"actions must have {ndim} dims ending in n_assets={self.stock_dim}, got {tuple(actions.shape)}")
        return actions

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
        if mask is not None:
            mask = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
        with torch.cuda.device(self.device):
            _cabi.check(
                _cabi.lib().frl_cashpenalty_reset(C.byref(self._p), _cabi.ptr(mask), _cabi.ptr(sp), _cabi.ptr(out), self._stream()),
                "frl_cashpenalty_reset",
            )
        self.launches += 2
        return out

    def observe(self, out=None):
        out = self._obs if out is None else out
        with self._torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().frl_cashpenalty_observe(C.byref(self._p), _cabi.ptr(out), self._stream()), "frl_cashpenalty_observe")
        self.launches += 1
        return out

    def step(self, actions, auto_reset=False, want_obs=True, accumulate_stats=False, want_done=True):
        """One ``step`` -> (state[N,O] f32, reward[N] f64, done[N] bool, flags[N] u8)."""
        a = self._as_actions(actions, 2)
        if a.shape[0] != self.n_envs:
            raise ValueError(f"actions must have n_envs={self.n_envs} rows")
        a = a.contiguous()
        obs = self._obs if want_obs else None
        ev = self.kernel_events
        with self._torch.cuda.device(self.device):
            if ev is not None:
                e0, e1 = self._torch.cuda.Event(enable_timing=True), self._torch.cuda.Event(enable_timing=True)
                e0.record()
            rc = _cabi.lib().frl_cashpenalty_step(
                C.byref(self._p), _cabi.ptr(a), int(a.dtype == self._torch.float64), _cabi.ptr(self._rew),
                _cabi.ptr(self._flags), _cabi.ptr(obs), int(auto_reset),
                _cabi.ptr(self.stats) if accumulate_stats else None, self._stream(),
            )
            if ev is not None:
                e1.record()
                ev.append((e0, e1))
        _cabi.check(rc, "frl_cashpenalty_step")
        self.launches += 1
        done = (self._flags & _cabi.FLAG_DONE).bool() if want_done else None
        return obs, self._rew, done, self._flags

    def rollout(self, actions, layout="KND", obs_mode="last", auto_reset=True, accumulate_stats=True):
        torch = self._torch
        a = self._as_actions(actions, 3)
        D, N = self.stock_dim, self.n_envs
        if layout == "KND":
            K, ok = a.shape[0], a.shape[1] == N
        elif layout == "NKD":
            K, ok = a.shape[1], a.shape[0] == N
        else:
            raise ValueError("layout must be 'KND' or 'NKD'")
        if not ok:
            raise ValueError(f"actions shape {tuple(a.shape)} does not match n_envs={N} for layout {layout}")
        a = a.contiguous()
        step_stride, env_stride = (N * D, D) if layout == "KND" else (D, K * D)
        mode = {"none": _cabi.OBS_NONE, "last": _cabi.OBS_LAST, "all": _cabi.OBS_ALL}[obs_mode]
        rewards = torch.empty((K, N), dtype=torch.float64, device=self.device)
        flags = torch.empty((K, N), dtype=torch.uint8, device=self.device)
        obs = None
        if mode == _cabi.OBS_LAST:
            obs = self._obs
        elif mode == _cabi.OBS_ALL:
            obs = torch.empty((K, N, self.state_space), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(
                _cabi.lib().frl_cashpenalty_rollout(
                    C.byref(self._p), _cabi.ptr(a), int(a.dtype == torch.float64), step_stride, env_stride, int(K),
                    _cabi.ptr(rewards), _cabi.ptr(flags), _cabi.ptr(obs), mode, int(auto_reset),
                    _cabi.ptr(self.stats) if accumulate_stats else None, self._stream(),
                ),
                "frl_cashpenalty_rollout",
            )
        self.launches += 1
        return obs, rewards, flags

    def read_stats(self, reset=False):
        vals = self.stats.tolist()
        if reset:
            self.stats.zero_()
        names = list(_cabi.STAT_NAMES)
        names[7] = "shortage_count"
        return dict(zip(names, vals))
