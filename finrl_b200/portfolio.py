"""Batched StockPortfolioEnv: N copies of the reference portfolio-allocation env on one GPU.

Reference: /root/reference/finrl/meta/env_portfolio_allocation/env_portfolio.py.  ``step`` applies
softmax weights (no max-subtraction, quirk Q8), the weighted one-day return and the portfolio-value
update; the reward is the new portfolio value.  The (D+K) x D observation (covariance matrix stacked
on the indicator rows) depends on the day only, so ``obs_table[day]`` is exposed next to a
materialising path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _cabi
from ._base import BatchedEnvBase


def frame_to_portfolio_arrays(df, stock_dim: int, tech_indicator_list: Sequence[str]):
    """Long frame with a per-row ``cov_list`` column (as the portfolio tutorial builds it,
    tutorials/2-Advance/FinRL_PortfolioAllocation_Explainable_DRL.py:157-174) -> close[T,D], cov[T,D,D], tech[K,T,D]."""
    days = df.index.unique()
    T, D = len(days), int(stock_dim)
    if len(df) != T * D:
        raise ValueError(f"frame has {len(df)} rows, expected n_days*stock_dim = {T}*{D}")
    close = np.ascontiguousarray(df["close"].to_numpy(dtype=np.float64).reshape(T, D))
    covs = df["cov_list"].values
    cov = np.stack([np.asarray(covs[t * D], dtype=np.float64) for t in range(T)])  # self.data["cov_list"].values[0]
    tech = np.stack([df[name].to_numpy(dtype=np.float64).reshape(T, D) for name in tech_indicator_list]) \
        if len(tech_indicator_list) else np.zeros((0, T, D))
    return close, cov, tech


@dataclass
class PortfolioTables:
    n_days: int
    stock_dim: int
    n_tech: int
    ret: "torch.Tensor"        # [T, pitch] f64 (pitch 32 for D <= 32, else 128)
    obs_table: "torch.Tensor"  # [T, (D+K)*D] f32
    host_obs: np.ndarray       # [T, D+K, D] f64 (the reference's state matrices)

    @property
    def obs_dim(self) -> int:
        return (self.stock_dim + self.n_tech) * self.stock_dim

    @staticmethod
    def from_arrays(close, cov, tech, device) -> "PortfolioTables":
        import torch

        close = np.ascontiguousarray(close, dtype=np.float64)
        T, D = close.shape
        if not 1 <= D <= 128:
            raise ValueError(f"stock_dim must be in 1..128 for the portfolio kernels (got {D})")
        cov = np.ascontiguousarray(cov, dtype=np.float64).reshape(T, D, D)
        tech = np.ascontiguousarray(tech, dtype=np.float64).reshape(-1, T, D)
        K = tech.shape[0]
        ret = np.zeros((T, 32 if D <= 32 else 128), dtype=np.float64)
        ret[1:, :D] = (close[1:] / close[:-1]) - 1  # (self.data.close.values / last_day_memory.close.values) - 1
        state = np.concatenate([cov, np.transpose(tech, (1, 0, 2))], axis=1)  # np.append(cov, tech rows, axis=0)
        dev = torch.device(device)
        return PortfolioTables(
            n_days=T, stock_dim=D, n_tech=K, ret=torch.from_numpy(ret).to(dev),
            obs_table=torch.from_numpy(state.reshape(T, -1).astype(np.float32)).to(dev), host_obs=state,
        )


class BatchedStockPortfolioEnv(BatchedEnvBase):
    """Keywords mirror ``StockPortfolioEnv.__init__`` (env_portfolio.py:66-80); ``hmax``,
    ``transaction_cost_pct``, ``reward_scaling``, ``turbulence_threshold`` and ``lookback`` are accepted
    and unused, exactly as in the reference's ``step``.  ``step`` / ``rollout`` / ``read_stats`` come from
    :class:`BatchedEnvBase`; observations are returned as [N, D+K, D] float32 (materialised lazily)."""

    _PREFIX = "frl_portfolio"

    def __init__(self, df=None, stock_dim=None, hmax=None, initial_amount=1_000_000, transaction_cost_pct=None,
                 reward_scaling=None, state_space=None, action_space=None, tech_indicator_list: Sequence[str] = (),
                 turbulence_threshold=None, lookback=252, day=0, *, n_envs=1, device="cuda",
                 tables: Optional[PortfolioTables] = None, track_weights: bool = False):
        torch = self._bind_device(device)
        if tables is None:
            if df is None:
                raise ValueError("either df or tables is required")
            tables = PortfolioTables.from_arrays(*frame_to_portfolio_arrays(df, stock_dim, list(tech_indicator_list)), self.device)
        self.tables, self.df = tables, df
        D, K, T = tables.stock_dim, tables.n_tech, tables.n_days
        if stock_dim is not None and int(stock_dim) != D:
            raise ValueError(f"stock_dim={stock_dim} but the tables hold {D} stocks")
        if state_space is not None and int(state_space) != D:
            raise ValueError(f"state_space={state_space} != stock_dim={D}")
        N = int(n_envs)
        self.n_envs, self.stock_dim, self.n_tech, self.n_days = N, D, K, T
        self.obs_shape = (D + K, D)
        self.initial_amount = initial_amount
        self.tech_indicator_list = list(tech_indicator_list)
        dev = self.device
        self.portfolio_value = torch.full((N,), float(initial_amount), dtype=torch.float64, device=dev)
        self.day = torch.full((N,), int(day), dtype=torch.int32, device=dev)
        self.reward = torch.zeros(N, dtype=torch.float64, device=dev)
        self._stats_block = _cabi.new_stats_block(torch, dev)
        self.stats = self._stats_block[:_cabi.N_STATS]
        self._obs = None  # materialised lazily: [N, (D+K)*D] f32 can be large
        self._rew = torch.empty(N, dtype=torch.float64, device=dev)
        self._flags = torch.empty(N, dtype=torch.uint8, device=dev)
        p = _cabi.PortfolioParams()
        p.n_envs, p.stock_dim, p.n_tech, p.n_days, p.obs_dim = N, D, K, T, tables.obs_dim
        p.initial_amount = float(initial_amount)
        p.ret, p.obs_table = tables.ret.data_ptr(), tables.obs_table.data_ptr()
        p.ret_pitch = int(tables.ret.shape[1])
        p.pv, p.day, p.reward = self.portfolio_value.data_ptr(), self.day.data_ptr(), self.reward.data_ptr()
        # optional: the step's portfolio_return and softmax weights (the reference's logging memories)
        self.last_return = torch.zeros(N, dtype=torch.float64, device=dev) if track_weights else None
        self.last_weights = torch.zeros((N, D), dtype=torch.float64, device=dev) if track_weights else None
        p.ret_out = self.last_return.data_ptr() if track_weights else None
        p.weights_out = self.last_weights.data_ptr() if track_weights else None
        self._p = p

    def _obs_buf(self):
        if self._obs is None:
            self._obs = self._torch.empty((self.n_envs, self.tables.obs_dim), dtype=self._torch.float32, device=self.device)
        return self._obs

    def _obs_out(self):
        return self._obs_buf()

    def _shape_obs(self, obs):
        return None if obs is None else obs.view(self.n_envs, *self.obs_shape)

    def observe_view(self):
        """The same observation without per-env copies: a gather of the per-day table by ``day``
        (a [N, D+K, D] tensor produced by torch indexing; callers in lock-step can use obs_table[day[0]])."""
        return self.tables.obs_table[self.day.long()].view(self.n_envs, *self.obs_shape)

    def reset(self, mask=None, want_obs=True):
        torch = self._torch
        mask = self._mask(mask)
        out = self._obs_buf() if want_obs else None
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().frl_portfolio_reset(C.byref(self._p), _cabi.ptr(mask), _cabi.ptr(out), self._stream()), "frl_portfolio_reset")
        self.launches += 2 if want_obs else 1
        return out.view(self.n_envs, *self.obs_shape) if want_obs else None

    def rollout(self, actions, layout="KND", obs_mode="none", auto_reset=True, accumulate_stats=True, **kw):
        """As :meth:`BatchedEnvBase.rollout`; the default is ``obs_mode="none"`` because the observation is a
        pure per-day table row (``tables.obs_table[day]``)."""
        return super().rollout(actions, layout=layout, obs_mode=obs_mode, auto_reset=auto_reset,
                               accumulate_stats=accumulate_stats, **kw)
