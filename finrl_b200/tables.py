"""Market tables: the reference's DataFrame / array inputs turned into device-resident tables.

The reference re-queries a pandas frame on every step (``self.df.loc[self.day, :]``,
/root/reference/finrl/meta/env_stock_trading/env_stocktrading.py:336; 61 % of its step time is
``.unique()`` on that frame).  Here the frame is read ONCE into dense per-day rows that live in HBM
(5.4 MB for DOW-30 x 2500 days — L2-resident) and are shared by all envs of a GPU.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np


def frame_to_arrays(df, stock_dim: int, tech_indicator_list: Sequence[str], risk_indicator_col: Optional[str],
                    price_col: str = "close"):
    """Long FinRL frame (index = day number, rows sorted by (date, tic); what ``data_split`` returns,
    /root/reference/finrl/meta/preprocessor/preprocessors.py:24-33) -> (close[T,D], tech[K,T,D], risk[T]|None).

    Mirrors how the env reads it: ``df.loc[day].close.values`` is the D prices of day ``day`` in frame
    order, ``df.loc[day][tech].values`` one indicator for the D stocks, and the risk column's first
    entry of the day (env_stocktrading.py:341)."""
    days = df.index.unique()
    T = len(days)
    D = int(stock_dim)
    if len(df) != T * D:
        raise ValueError(f"frame has {len(df)} rows, expected n_days*stock_dim = {T}*{D}")
    idx = np.asarray(df.index)
    if not np.array_equal(idx, np.repeat(np.arange(T), D)):
        raise ValueError("frame index must be the day number 0..T-1 repeated stock_dim times (use data_split)")
    close = np.ascontiguousarray(df[price_col].to_numpy(dtype=np.float64).reshape(T, D))
    K = len(tech_indicator_list)
    tech = np.empty((K, T, D), dtype=np.float64)
    for k, name in enumerate(tech_indicator_list):
        tech[k] = df[name].to_numpy(dtype=np.float64).reshape(T, D)
    risk = None
    if risk_indicator_col is not None and risk_indicator_col in df.columns:
        risk = np.ascontiguousarray(df[risk_indicator_col].to_numpy(dtype=np.float64).reshape(T, D)[:, 0])
    return close, tech, risk


@dataclass
class TradingTables:
    """Device tables of the StockTradingEnv path (layout: include/finrl_b200.h, frl_trading_params)."""

    n_days: int
    stock_dim: int
    n_tech: int
    close: "torch.Tensor"         # [T, pitch] f64, rows zero-padded (pitch 32 or 128)
    disable_mask: "torch.Tensor"  # [T, pitch/32] int32 bit mask (first indicator == 1.0)
    risk: "torch.Tensor"          # [T] f64
    obs_tmpl: "torch.Tensor"      # [T, O] f32
    obs_tmpl4: Optional["torch.Tensor"]  # [T, 4*O] f32: every row four times (bulk-copy image), or None
    host_close: np.ndarray
    host_tech: np.ndarray
    host_risk: np.ndarray
    host_tmpl: Optional[np.ndarray] = None  # [T, O] f32 copy of obs_tmpl for the factored host observation

    @property
    def obs_dim(self) -> int:
        return 1 + 2 * self.stock_dim + self.n_tech * self.stock_dim

    @staticmethod
    def from_arrays(close, tech, risk, device, allow_nonpositive_close: bool = False) -> "TradingTables":
        import torch

        close = np.ascontiguousarray(close, dtype=np.float64)
        T, D = close.shape
        if not 1 <= D <= 128:
            raise ValueError(f"stock_dim must be in 1..128 for the StockTradingEnv kernels (got {D})")
        tech = np.ascontiguousarray(tech, dtype=np.float64).reshape(-1, T, D)
        K = tech.shape[0]
        # A close <= 0 (or NaN) on a tradable stock has no defined behaviour in the reference: its buy path divides
        # by the price (`state[0] // (price * (1 + cost))`, env_stocktrading.py:178-180: ZeroDivisionError while the
        # cash is still a Python number, inf shares "available" once it is a numpy scalar).  Refuse such tables unless
        # the stock carries the disable flag that day (first indicator == 1.0: never traded) or the caller insists.
        tradable = np.ones((T, D), dtype=bool) if K == 0 else tech[0] != 1.0
        bad = ~(close > 0) & tradable
        if bad.any() and not allow_nonpositive_close:
            t, j = np.argwhere(bad)[0]
            raise ValueError(f"close[{t}, {j}] = {close[t, j]} is not a positive price (and the stock is not disabled that day); "
                             "clean the frame or pass allow_nonpositive_close=True")
        risk = np.zeros(T) if risk is None else np.ascontiguousarray(risk, dtype=np.float64)
        if risk.shape != (T,):
            raise ValueError(f"risk must have shape ({T},), got {risk.shape}")
        O = 1 + 2 * D + K * D
        pitch = 32 if D <= 32 else 128
        close32 = np.zeros((T, pitch), dtype=np.float64)
        close32[:, :D] = close
        mask = np.zeros((T, pitch // 32), dtype=np.uint32)
        if K > 0:
            bits = np.zeros((T, pitch), dtype=np.uint32)
            bits[:, :D] = tech[0] == 1.0  # `state[index + 2D + 1] != True`
            words = bits.reshape(T, pitch // 32, 32) << np.arange(32, dtype=np.uint32)[None, None, :]
            mask = words.sum(axis=2).astype(np.uint32)
        tmpl = np.zeros((T, O), dtype=np.float32)
        tmpl[:, 1 : 1 + D] = close.astype(np.float32)
        if K > 0:
            tmpl[:, 1 + 2 * D :] = np.transpose(tech, (1, 0, 2)).reshape(T, K * D).astype(np.float32)
        dev = torch.device(device)
        # 4-row image for the bulk-copy observation writer (thread-per-env kernel, D <= 32, image <= 8 KB)
        tmpl4 = torch.from_numpy(np.tile(tmpl, (1, 4))).to(dev) if (D <= 32 and 16 * O <= 8192) else None
        return TradingTables(
            n_days=T, stock_dim=D, n_tech=K,
            close=torch.from_numpy(close32).to(dev),
            disable_mask=torch.from_numpy(np.ascontiguousarray(mask).view(np.int32)).to(dev),
            risk=torch.from_numpy(risk).to(dev),
            obs_tmpl=torch.from_numpy(tmpl).to(dev), obs_tmpl4=tmpl4,
            host_close=close, host_tech=tech, host_risk=risk, host_tmpl=tmpl,
        )

    @staticmethod
    def from_frame(df, stock_dim, tech_indicator_list, risk_indicator_col, device) -> "TradingTables":
        close, tech, risk = frame_to_arrays(df, stock_dim, tech_indicator_list, risk_indicator_col)
        return TradingTables.from_arrays(close, tech, risk, device)
