"""Batched CryptoEnv (sibling env) + the drop-in single-env class.

Reference: /root/reference/finrl/meta/env_cryptocurrency_trading/env_multiple_crypto.py.  Fractional
float32 positions, float64 prices and cash, per-coin action normaliser from the first day's price
magnitude, no turbulence, reward * 2**-16, ``gamma_return`` terminal reward.  Kernel: csrc/crypto.cu.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _cabi
from ._base import BatchedEnvBase
from .spaces import Box
from .vec_env import BatchedVecEnv


class BatchedCryptoEnv(BatchedEnvBase):
    _PREFIX = "frl_crypto"
    _ACTION_NAME = "crypto_num"

    def __init__(self, config, lookback=1, initial_capital=1e6, buy_cost_pct=1e-3, sell_cost_pct=1e-3, gamma=0.99, *,
                 n_envs=1, device="cuda"):
        torch = self._bind_device(device)
        price = np.ascontiguousarray(config["price_array"], dtype=np.float64)
        T, D = price.shape
        if not 1 <= D <= 32:
            raise ValueError(f"crypto_num must be in 1..32 (got {D})")
        tech = np.ascontiguousarray(config["tech_array"], dtype=np.float64).reshape(T, -1)
        TD, LB = tech.shape[1], int(lookback)
        # _generate_action_normalizer (:103-111)
        self.action_norm_vector = np.asarray([1 / (10 ** math.floor(math.log(p, 10))) for p in price[0]]) * 10000
        O = 1 + D + TD * LB
        tmpl = np.zeros((T, O), dtype=np.float32)
        for l in range(LB):  # get_state (:93-99): tech rows time, time-1, ... each * 2**-15, cast to float32
            rows = np.zeros((T, TD))
            rows[l:] = tech[: T - l]
            tmpl[:, 1 + D + l * TD : 1 + D + (l + 1) * TD] = (rows * 2**-15).astype(np.float32)
        price32 = np.zeros((T, 32))
        price32[:, :D] = price
        norm32 = np.ones(32)
        norm32[:D] = self.action_norm_vector
        dev = self.device
        N = int(n_envs)
        self.n_envs = self.env_num = N
        self.stock_dim = self.action_dim = self.crypto_num = D
        self.lookback, self.n_days, self.max_step = LB, T, T - LB - 1
        self.obs_dim = O
        self.state_dim = 1 + (D + TD) * LB  # the reference's (inconsistent for lookback > 1) formula (:36)
        self.env_name, self.if_discrete, self.target_return = "MulticryptoEnv", False, 10
        self.initial_cash = self.initial_capital = initial_capital
        self.price_array, self.tech_array = price, tech
        self._price = torch.from_numpy(price32).to(dev)
        self._norm = torch.from_numpy(norm32).to(dev)
        self._tmpl = torch.from_numpy(tmpl).to(dev)
        self.cash = torch.empty(N, dtype=torch.float64, device=dev)
        self.stocks = torch.empty((D, N), dtype=torch.float32, device=dev)
        self.time = torch.empty(N, dtype=torch.int32, device=dev)
        self.total_asset = torch.empty(N, dtype=torch.float64, device=dev)
        self.gamma_return = torch.zeros(N, dtype=torch.float64, device=dev)
        self.episode_return = torch.zeros(N, dtype=torch.float64, device=dev)
        self._stats_block = _cabi.new_stats_block(torch, dev)
        self.stats = self._stats_block[:_cabi.N_STATS]
        self._obs = torch.empty((N, O), dtype=torch.float32, device=dev)
        self._rew = torch.empty(N, dtype=torch.float64, device=dev)
        self._flags = torch.empty(N, dtype=torch.uint8, device=dev)
        p = _cabi.CryptoParams()
        p.n_envs, p.stock_dim, p.tech_dim, p.n_days, p.lookback, p.obs_dim, p.env_stride = N, D, TD, T, LB, O, N
        p.initial_capital, p.buy_cost_pct, p.sell_cost_pct, p.gamma = float(initial_capital), float(buy_cost_pct), float(sell_cost_pct), float(gamma)
        p.price, p.act_norm, p.obs_tmpl = self._price.data_ptr(), self._norm.data_ptr(), self._tmpl.data_ptr()
        p.cash, p.stocks, p.time = self.cash.data_ptr(), self.stocks.data_ptr(), self.time.data_ptr()
        p.total, p.gamma_return, p.episode_return = self.total_asset.data_ptr(), self.gamma_return.data_ptr(), self.episode_return.data_ptr()
        self._p = p
        self.reset()

    def reset(self, mask=None, out=None):
        torch = self._torch
        out = self._obs if out is None else out
        mask = self._mask(mask)
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().frl_crypto_reset(C.byref(self._p), _cabi.ptr(mask), _cabi.ptr(out), self._stream()), "frl_crypto_reset")
        self.launches += 2
        return out


class CryptoEnv:
    """Drop-in for ``env_multiple_crypto.CryptoEnv`` (gym protocol; ``step`` returns ``info = None`` like the reference)."""

    def __init__(self, config, lookback=1, initial_capital=1e6, buy_cost_pct=1e-3, sell_cost_pct=1e-3, gamma=0.99, device="cuda"):
        self._args = dict(lookback=lookback, initial_capital=initial_capital, buy_cost_pct=buy_cost_pct,
                          sell_cost_pct=sell_cost_pct, gamma=gamma)
        self._config, self._device = config, device
        self.engine = e = BatchedCryptoEnv(config, n_envs=1, device=device, **self._args)
        self.lookback, self.initial_cash, self.initial_total_asset = lookback, initial_capital, initial_capital
        self.buy_cost_pct, self.sell_cost_pct, self.gamma, self.max_stock = buy_cost_pct, sell_cost_pct, gamma, 1
        self.price_array, self.tech_array = e.price_array, e.tech_array
        self.action_norm_vector = e.action_norm_vector
        self.crypto_num, self.max_step = e.crypto_num, e.max_step
        self.env_name, self.state_dim, self.action_dim = e.env_name, e.state_dim, e.action_dim
        self.if_discrete, self.target_return = False, 10
        self.observation_space = Box(low=-3000, high=3000, shape=(self.state_dim,), dtype=np.float32)
        self.action_space = Box(low=-1, high=1, shape=(self.action_dim,), dtype=np.float32)
        self.episode_return = 0.0
        self._pull()

    def _pull(self):
        e = self.engine
        self.time = int(e.time[0].item())
        self.cash = np.float64(e.cash[0].item())
        self.stocks = e.stocks[:, 0].cpu().numpy()
        self.total_asset = np.float64(e.total_asset[0].item())
        self.gamma_return = float(e.gamma_return[0].item())
        self.current_price, self.current_tech = self.price_array[self.time], self.tech_array[self.time]

    def reset(self):
        state = self.engine.reset()
        self._pull()
        return state[0].cpu().numpy()

    def step(self, actions):
        import torch

        a = np.asarray(actions)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)
        state, reward, done, flags = self.engine.step(torch.as_tensor(a.reshape(1, -1)))
        # the reference scales the caller's array in place (:62-64)
        if isinstance(actions, np.ndarray) and actions.dtype in (np.float32, np.float64):
            actions *= self.action_norm_vector.astype(actions.dtype)
        self._pull()
        d = bool(done[0].item())
        self.cumu_return = self.total_asset / self.initial_cash
        if d:
            self.episode_return = float(self.engine.episode_return[0].item())
        return state[0].cpu().numpy(), np.float64(reward[0].item()), d, None

    def get_state(self):
        return self.engine.observe()[0].cpu().numpy()

    def close(self):
        pass

    def vectorized(self, n_envs, tensor_mode=True):
        return BatchedVecEnv(BatchedCryptoEnv(self._config, n_envs=n_envs, device=self._device, **self._args),
                             tensor_mode=tensor_mode, obs_shape=(self.engine.obs_dim,))
