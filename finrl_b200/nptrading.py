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


@dataclass
class NpTables:
    """Device tables of the numpy env (layout: include/finrl_b200.h, frl_np_params)."""

    n_days: int
    stock_dim: int
    tech_dim: int
    price: "torch.Tensor"      # [T, 32] f32
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
        if not 1 <= D <= 32:
            raise ValueError(f"stock_dim must be in 1..32 for the numpy-env kernel (got {D})")
        tech = np.ascontiguousarray(np.asarray(tech_array).astype(np.float32) * 2**-7).reshape(T, -1)
        turb = np.asarray(turbulence_array)
        turb_bool = (turb > turbulence_thresh).astype(np.float32)
        sig = 1 / (1 + np.exp(-(turb / turbulence_thresh) * np.e)) - 0.5
        turb_ary = (sig * turbulence_thresh * 2**-5).astype(np.float32)
        TD = tech.shape[1]
        O = 3 + 3 * D + TD
        price32 = np.zeros((T, 32), dtype=np.float32)
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


class BatchedNpStockTradingEnv:
    """N numpy-env instances on one GPU.  Keywords mirror ``StockTradingEnv.__init__``
    (env_stocktrading_np.py:9-22); extra: ``n_envs``, ``device``, ``tables``."""

    def __init__(self, config=None, initial_account=1e6, gamma=0.99, turbulence_thresh=99, min_stock_rate=0.1,
                 max_stock=1e2, initial_capital=1e6, buy_cost_pct=1e-3, sell_cost_pct=1e-3, reward_scaling=2**-11,
                 initial_stocks=None, *, n_envs=1, device="cuda", tables=None, if_train=None, obs_amount_floor=None):
        import torch

        self._torch = torch
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _cabi.EngineError("finrl_b200 runs on CUDA devices only (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        _cabi.lib()
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
        self.stats = torch.zeros(_cabi.N_STATS, dtype=torch.float64, device=dev)
        self._obs = torch.empty((N, O), dtype=torch.float32, device=dev)
        self._reward = torch.empty(N, dtype=torch.float64, device=dev)
        self._flags = torch.empty(N, dtype=torch.uint8, device=dev)
        p = _cabi.NpParams()
        p.n_envs, p.stock_dim, p.tech_dim, p.n_days, p.obs_dim, p.env_stride = N, D, TD, T, O, N
        p.gamma, p.max_stock, p.min_stock_rate = float(gamma), float(max_stock), float(min_stock_rate)
        p.buy_cost_pct, p.sell_cost_pct, p.reward_scaling = float(buy_cost_pct), float(sell_cost_pct), float(reward_scaling)
        p.initial_capital = float(initial_capital)
        # StockEnvNAS100's get_state shows max(amount, 1e4) (env_nas100_wrds.py:157)
        p.obs_amount_floor = float("-inf") if obs_amount_floor is None else float(obs_amount_floor)
        p.price, p.turb_bool, p.obs_tmpl = tables.price.data_ptr(), tables.turb_bool.data_ptr(), tables.obs_tmpl.data_ptr()
        p.init_stocks = self._init_stocks.data_ptr()
        p.amount, p.kinds, p.stocks, p.cool = self.amount.data_ptr(), self.kinds.data_ptr(), self.stocks.data_ptr(), self.cool.data_ptr()
        p.day, p.total, p.gamma_reward = self.day.data_ptr(), self.total_asset.data_ptr(), self.gamma_reward.data_ptr()
        p.init_total, p.episode_return = self.initial_total_asset.data_ptr(), self.episode_return.data_ptr()
        self._p = p
        self.launches = 0
        self.kernel_events = None
        self.reset()

    def _stream(self):
        return _cabi.current_stream(self.device)

    def _as_actions(self, actions, ndim):
        torch = self._torch
        if not isinstance(actions, torch.Tensor):
            actions = torch.as_tensor(np.asarray(actions))
        if actions.dtype not in (torch.float32, torch.float64):
            actions = actions.to(torch.float32)
        if actions.device != self.device:
            actions = actions.to(self.device, non_blocking=True)
        if actions.dim() != ndim or actions.shape[-1] != self.stock_dim:
            raise ValueError(f"actions must have {ndim} dims ending in stock_dim={self.stock_dim}, got {tuple(actions.shape)}")
        return actions

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
        if mask is not None:
            mask = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
        with torch.cuda.device(self.device):
            _cabi.check(
                _cabi.lib().frl_np_reset(C.byref(self._p), _cabi.ptr(mask), _cabi.ptr(s0), _cabi.ptr(f), _cabi.ptr(out), self._stream()),
                "frl_np_reset",
            )
        self.launches += 2
        return out

    def observe(self, out=None):
        out = self._obs if out is None else out
        with self._torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().frl_np_observe(C.byref(self._p), _cabi.ptr(out), self._stream()), "frl_np_observe")
        self.launches += 1
        return out

    def step(self, actions, auto_reset: bool = False, want_obs: bool = True, accumulate_stats: bool = False,
             want_done: bool = True):
        """One ``step`` of every env -> (state[N,O] f32, reward[N] f64, done[N] bool, flags[N] u8).
        Buffers are engine-owned and overwritten by the next call."""
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
            rc = _cabi.lib().frl_np_step(
                C.byref(self._p), _cabi.ptr(a), int(a.dtype == self._torch.float64), _cabi.ptr(self._reward),
                _cabi.ptr(self._flags), _cabi.ptr(obs), int(auto_reset),
                _cabi.ptr(self.stats) if accumulate_stats else None, self._stream(),
            )
            if ev is not None:
                e1.record()
                ev.append((e0, e1))
        _cabi.check(rc, "frl_np_step")
        self.launches += 1
        done = (self._flags & _cabi.FLAG_DONE).bool() if want_done else None
        return obs, self._reward, done, self._flags

    def rollout(self, actions, layout: str = "KND", obs_mode: str = "last", auto_reset: bool = True,
                accumulate_stats: bool = True, rewards=None, flags=None, obs=None):
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
        if rewards is None:
            rewards = torch.empty((K, N), dtype=torch.float64, device=self.device)
        if flags is None:
            flags = torch.empty((K, N), dtype=torch.uint8, device=self.device)
        if mode == _cabi.OBS_LAST and obs is None:
            obs = self._obs
        elif mode == _cabi.OBS_ALL and obs is None:
            obs = torch.empty((K, N, self.state_dim), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(
                _cabi.lib().frl_np_rollout(
                    C.byref(self._p), _cabi.ptr(a), int(a.dtype == torch.float64), step_stride, env_stride, int(K),
                    _cabi.ptr(rewards), _cabi.ptr(flags), _cabi.ptr(obs) if mode else None, mode, int(auto_reset),
                    _cabi.ptr(self.stats) if accumulate_stats else None, self._stream(),
                ),
                "frl_np_rollout",
            )
        self.launches += 1
        return (obs if mode else None), rewards, flags

    # ------------------------------------------------------------------------------------------
    def get_state(self):
        k = self.kinds
        return {
            "amount": self.amount.clone(), "amount_kind": k & 3, "stocks": self.stocks.t().contiguous(),
            "cool": self.cool.t().contiguous(), "day": self.day.clone(), "total": self.total_asset.clone(),
            "total_kind": (k >> 2) & 3, "gamma_reward": self.gamma_reward.clone(), "gr_kind": (k >> 4) & 3,
            "init_total": self.initial_total_asset.clone(), "episode_return": self.episode_return.clone(),
        }

    def read_stats(self, reset: bool = False):
        vals = self.stats.tolist()
        if reset:
            self.stats.zero_()
        return dict(zip(_cabi.STAT_NAMES, vals))
