"""Batched numpy/ElegantRL StockTradingEnv: N copies of the reference env stepped by one CUDA kernel.

Reference: /root/reference/finrl/meta/env_stock_trading/env_stocktrading_np.py.  Same constructor
keywords (``config`` dict with ``price_array, tech_array, turbulence_array, if_train``), same
``reset`` / ``step`` semantics incl. the mixed f32/f64 arithmetic and quirk Q6; every
method acts on all N envs and returns device tensors, following ElegantRL's vectorised-env convention
(``env_num > 1``: ``reset() -> Tensor[env_num, state_dim]``, ``step(Tensor[env_num, action_dim])``).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _cabi
from ._base import BatchedEnvBase


@dataclass
class NpTables:
    """Device tables of the numpy env (layout: include/finrl_b200.h, frl_np_params)."""

    n_days: int
    stock_dim: int
    tech_dim: int
    price: "torch.Tensor"      # [T, pitch] f32 (pitch 32 for D <= 32, else 128)
    turb_bool: "torch.Tensor"  # [T] f32
    obs_tmpl: "torch.Tensor"   # [T, O] f32
    host_price: np.ndarray     # price_ary  [T, D] f32
    host_tech: np.ndarray      # tech_ary   [T, tech_dim] f32 (already * 2^-7)
    host_turb_bool: np.ndarray
    host_turb_ary: np.ndarray

    @property
    def obs_dim(self) -> int:
        return 3 + 3 * self.stock_dim + self.tech_dim

    @staticmethod
    def from_arrays(price_array, tech_array, turbulence_array, turbulence_thresh, device) -> "NpTables":
        """The constructor's table transforms (env_stocktrading_np.py:27-35, 164-169), done once on the
        host with numpy exactly as the reference does, then laid out as per-day rows."""
        import torch

        price = np.ascontiguousarray(np.asarray(price_array).astype(np.float32))
        T, D = price.shape
        if not 1 <= D <= 128:
            raise ValueError(f"stock_dim must be in 1..128 for the numpy-env kernels (got {D})")
        tech = np.ascontiguousarray(np.asarray(tech_array).astype(np.float32) * 2**-7).reshape(T, -1)
        turb = np.asarray(turbulence_array)
        turb_bool = (turb > turbulence_thresh).astype(np.float32)
        sig = 1 / (1 + np.exp(-(turb / turbulence_thresh) * np.e)) - 0.5
        turb_ary = (sig * turbulence_thresh * 2**-5).astype(np.float32)
        TD = tech.shape[1]
        O = 3 + 3 * D + TD
        price32 = np.zeros((T, 32 if D <= 32 else 128), dtype=np.float32)
        price32[:, :D] = price
        tmpl = np.zeros((T, O), dtype=np.float32)
        tmpl[:, 1] = turb_ary
        tmpl[:, 2] = turb_bool
        tmpl[:, 3 : 3 + D] = price * np.array(2**-6, dtype=np.float32)
        tmpl[:, 3 + 3 * D :] = tech
        dev = torch.device(device)
        return NpTables(
            n_days=T, stock_dim=D, tech_dim=TD,
            price=torch.from_numpy(price32).to(dev), turb_bool=torch.from_numpy(turb_bool).to(dev),
            obs_tmpl=torch.from_numpy(tmpl).to(dev),
            host_price=price, host_tech=tech, host_turb_bool=turb_bool, host_turb_ary=turb_ary,
        )


class BatchedNpStockTradingEnv(BatchedEnvBase):
    """N numpy-env instances on one GPU.  Keywords mirror ``StockTradingEnv.__init__``
    (env_stocktrading_np.py:9-22); extra: ``n_envs``, ``device``, ``tables``.  ``step`` / ``rollout`` /
    ``observe`` / ``read_stats`` come from :class:`BatchedEnvBase`; ``step`` returns (state[N,O] f32,
    reward[N] f64, done[N] bool, flags[N] u8) with the reward's numpy kind in flag bits 4-5."""

    _PREFIX = "frl_np"

    def __init__(self, config=None, initial_account=1e6, gamma=0.99, turbulence_thresh=99, min_stock_rate=0.1,
                 max_stock=1e2, initial_capital=1e6, buy_cost_pct=1e-3, sell_cost_pct=1e-3, reward_scaling=2**-11,
                 initial_stocks=None, *, n_envs=1, device="cuda", tables=None, if_train=None, obs_amount_floor=None):
        torch = self._bind_device(device)
        if tables is None:
            if config is None:
                raise ValueError("either config or tables is required")
            tables = NpTables.from_arrays(config["price_array"], config["tech_array"], config["turbulence_array"],
                                          turbulence_thresh, self.device)
        if if_train is None:
            if_train = bool(config["if_train"]) if config is not None else False
        self.tables = tables
        D, T, TD, O = tables.stock_dim, tables.n_days, tables.tech_dim, tables.obs_dim
        N = int(n_envs)
        self.n_envs = self.env_num = N
        self.stock_dim = self.action_dim = D
        self.state_dim = O
        self.max_step = T - 1
        self.if_train = bool(if_train)
        self.if_discrete = False
        self.env_name = "StockEnv"
        self.target_return = 10.0
        self.gamma, self.max_stock, self.min_stock_rate = gamma, max_stock, min_stock_rate
        self.buy_cost_pct, self.sell_cost_pct, self.reward_scaling = buy_cost_pct, sell_cost_pct, reward_scaling
        self.initial_capital = initial_capital
        init = np.zeros(D, dtype=np.float32) if initial_stocks is None else np.asarray(initial_stocks, dtype=np.float32)
        if init.shape != (D,):
            raise ValueError(f"initial_stocks must have shape ({D},)")
        self.initial_stocks = init
        dev = self.device
        self._init_stocks = torch.from_numpy(init.copy()).to(dev)
        self.amount = torch.empty(N, dtype=torch.float64, device=dev)
        self.kinds = torch.empty(N, dtype=torch.uint8, device=dev)
        self.stocks = torch.empty((D, N), dtype=torch.float32, device=dev)  # stock-major
        self.cool = torch.empty((D, N), dtype=torch.float32, device=dev)
        self.day = torch.empty(N, dtype=torch.int32, device=dev)
        self.total_asset = torch.empty(N, dtype=torch.float64, device=dev)
        self.gamma_reward = torch.empty(N, dtype=torch.float64, device=dev)
        self.initial_total_asset = torch.empty(N, dtype=torch.float64, device=dev)
        self.episode_return = torch.zeros(N, dtype=torch.float64, device=dev)
        self._stats_block = _cabi.new_stats_block(torch, dev)
        self.stats = self._stats_block[:_cabi.N_STATS]
        self._obs = torch.empty((N, O), dtype=torch.float32, device=dev)
        self._rew = torch.empty(N, dtype=torch.float64, device=dev)
        self._flags = torch.empty(N, dtype=torch.uint8, device=dev)
        p = _cabi.NpParams()
        p.n_envs, p.stock_dim, p.tech_dim, p.n_days, p.obs_dim, p.env_stride = N, D, TD, T, O, N
        p.gamma, p.max_stock, p.min_stock_rate = float(gamma), float(max_stock), float(min_stock_rate)
        p.buy_cost_pct, p.sell_cost_pct, p.reward_scaling = float(buy_cost_pct), float(sell_cost_pct), float(reward_scaling)
        p.initial_capital = float(initial_capital)
        # StockEnvNAS100's get_state shows max(amount, 1e4) (env_nas100_wrds.py:157)
        p.obs_amount_floor = float("-inf") if obs_amount_floor is None else float(obs_amount_floor)
        p.price, p.turb_bool, p.obs_tmpl = tables.price.data_ptr(), tables.turb_bool.data_ptr(), tables.obs_tmpl.data_ptr()
        p.price_pitch = int(tables.price.shape[1])
        p.train_reset = int(self.if_train)  # in-kernel auto-reset redraws the if_train position (reset_seed per launch)
        p.init_stocks = self._init_stocks.data_ptr()
        p.amount, p.kinds, p.stocks, p.cool = self.amount.data_ptr(), self.kinds.data_ptr(), self.stocks.data_ptr(), self.cool.data_ptr()
        p.day, p.total, p.gamma_reward = self.day.data_ptr(), self.total_asset.data_ptr(), self.gamma_reward.data_ptr()
        p.init_total, p.episode_return = self.initial_total_asset.data_ptr(), self.episode_return.data_ptr()
        self._p = p
        self.reset()

    # ------------------------------------------------------------------------------------------
    def draw_train_reset(self, rng=None):
        """The random draws of the reference's ``if_train`` reset (:85-92), per env in order:
        ``rd.randint(0, 64, D)`` then ``rd.uniform(0.95, 1.05)``.  ``rng`` defaults to numpy's GLOBAL
        RandomState like the reference, so a single env reproduces its stream exactly."""
        rd = np.random if rng is None else rng
        N, D = self.n_envs, self.stock_dim
        stocks0 = np.empty((N, D), dtype=np.float32)
        factor = np.empty(N, dtype=np.float64)
        for n in range(N):
            stocks0[n] = (self.initial_stocks + rd.randint(0, 64, size=self.initial_stocks.shape)).astype(np.float32)
            factor[n] = rd.uniform(0.95, 1.05)
        return stocks0, factor

    def reset(self, mask=None, stocks0=None, factor=None, out=None):
        """``reset()`` of all envs (or those with ``mask[n] != 0``).  With ``if_train`` the random
        initial position is drawn by :meth:`draw_train_reset` unless ``stocks0`` [N,D] / ``factor`` [N]
        are given."""
        torch = self._torch
        out = self._obs if out is None else out
        if self.if_train and stocks0 is None:
            if self.n_envs <= 4096:
                stocks0, factor = self.draw_train_reset()
            else:  # large batches: same distributions, drawn on the device
                stocks0 = self._init_stocks[None, :] + torch.randint(0, 64, (self.n_envs, self.stock_dim), device=self.device).float()
                factor = torch.empty(self.n_envs, dtype=torch.float64, device=self.device).uniform_(0.95, 1.05)
        s0 = f = None
        if stocks0 is not None:
            s0 = torch.as_tensor(stocks0, device=self.device).to(torch.float32).reshape(self.n_envs, self.stock_dim).t().contiguous()
            f = torch.as_tensor(factor, device=self.device).to(torch.float64).reshape(self.n_envs).contiguous()
        mask = self._mask(mask)
        with torch.cuda.device(self.device):
            _cabi.check(
                _cabi.lib().frl_np_reset(C.byref(self._p), _cabi.ptr(mask), _cabi.ptr(s0), _cabi.ptr(f), _cabi.ptr(out), self._stream()),
                "frl_np_reset",
            )
        self.launches += 2
        return out

    # ------------------------------------------------------------------------------------------
    def get_state(self):
        k = self.kinds
        return {
            "amount": self.amount.clone(), "amount_kind": k & 3, "stocks": self.stocks.t().contiguous(),
            "cool": self.cool.t().contiguous(), "day": self.day.clone(), "total": self.total_asset.clone(),
            "total_kind": (k >> 2) & 3, "gamma_reward": self.gamma_reward.clone(), "gr_kind": (k >> 4) & 3,
            "init_total": self.initial_total_asset.clone(), "episode_return": self.episode_return.clone(),
        }
