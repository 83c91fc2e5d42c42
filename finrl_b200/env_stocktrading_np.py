"""Drop-in for ``finrl.meta.env_stock_trading.env_stocktrading_np.StockTradingEnv`` (the ElegantRL env).

Same constructor (``config`` dict + keywords), attributes and gym protocol as the reference class
(/root/reference/finrl/meta/env_stock_trading/env_stocktrading_np.py:8-169); ``reset`` / ``step`` run on
the GPU through a 1-env :class:`finrl_b200.nptrading.BatchedNpStockTradingEnv`.  What
``finrl/agents/elegantrl/models.py`` touches is preserved: the class is called as ``env_cls(config=...)``
(:58), ``env.env_num`` is assignable (:59), ``Arguments`` reads ``env_name, state_dim, action_dim, max_step,
if_discrete, target_return`` and the prediction loop reads ``amount, price_ary, day, stocks,
initial_total_asset, max_step`` (:105-127).  ``vectorized(n_envs)`` gives the batched tensor env.
"""
from __future__ import annotations

import numpy as np

from .nptrading import BatchedNpStockTradingEnv
from .spaces import Box
from .vec_env import BatchedVecEnv

_KINDS = {0: float, 1: np.float32, 2: np.float64}


class StockTradingEnv:
    def __init__(self, config, initial_account=1e6, gamma=0.99, turbulence_thresh=99, min_stock_rate=0.1,
                 max_stock=1e2, initial_capital=1e6, buy_cost_pct=1e-3, sell_cost_pct=1e-3, reward_scaling=2**-11,
                 initial_stocks=None, device="cuda", obs_amount_floor=None):
        self._kw = dict(initial_account=initial_account, gamma=gamma, turbulence_thresh=turbulence_thresh,
                        min_stock_rate=min_stock_rate, max_stock=max_stock, initial_capital=initial_capital,
                        buy_cost_pct=buy_cost_pct, sell_cost_pct=sell_cost_pct, reward_scaling=reward_scaling,
                        initial_stocks=initial_stocks, obs_amount_floor=obs_amount_floor)
        self._config, self._device = config, device
        self.engine = e = BatchedNpStockTradingEnv(config, n_envs=1, device=device, **self._kw)
        t = e.tables
        self.price_ary, self.tech_ary = t.host_price, t.host_tech
        self.turbulence_bool, self.turbulence_ary = t.host_turb_bool, t.host_turb_ary
        self.gamma, self.max_stock, self.min_stock_rate = gamma, max_stock, min_stock_rate
        self.buy_cost_pct, self.sell_cost_pct, self.reward_scaling = buy_cost_pct, sell_cost_pct, reward_scaling
        self.initial_capital = initial_capital
        self.initial_stocks = e.initial_stocks
        self.env_name = "StockEnv"
        self.env_num = 1
        self.state_dim, self.action_dim, self.max_step = e.state_dim, e.action_dim, e.max_step
        self.if_train = e.if_train
        self.if_discrete = False
        self.target_return = 10.0
        self.episode_return = 0.0
        self.observation_space = Box(low=-3000, high=3000, shape=(self.state_dim,), dtype=np.float32)
        self.action_space = Box(low=-1, high=1, shape=(self.action_dim,), dtype=np.float32)
        self._pull()

    def _pull(self):
        """Mirror the device state into the reference's attributes, with the numpy scalar type each
        one would have (the kind travels with the value)."""
        st = self.engine.get_state()
        self.day = int(st["day"][0].item())
        self.amount = _KINDS[int(st["amount_kind"][0].item())](st["amount"][0].item())
        self.stocks = st["stocks"][0].cpu().numpy()
        self.stocks_cool_down = st["cool"][0].cpu().numpy()
        self.total_asset = _KINDS[int(st["total_kind"][0].item())](st["total"][0].item())
        self.gamma_reward = _KINDS[int(st["gr_kind"][0].item())](st["gamma_reward"][0].item())
        self.initial_total_asset = np.float32(st["init_total"][0].item())

    def reset(self):
        state = self.engine.reset()  # if_train: draws rd.randint / rd.uniform from numpy's global RNG like the reference
        self._pull()
        return state[0].cpu().numpy()

    def step(self, actions):
        import torch

        a = np.asarray(actions)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)
        state, reward, done, flags = self.engine.step(torch.as_tensor(a.reshape(1, -1)))
        fl = int(flags[0].item())
        self._pull()
        done = bool(fl & 1)
        reward = _KINDS[(fl >> 4) & 3](reward[0].item())
        if done:
            self.episode_return = float(self.engine.episode_return[0].item())
        return state[0].cpu().numpy(), reward, done, dict()

    def get_state(self, price=None):
        return self.engine.observe()[0].cpu().numpy()

    @staticmethod
    def sigmoid_sign(ary, thresh):
        def sigmoid(x):
            return 1 / (1 + np.exp(-x * np.e)) - 0.5

        return sigmoid(ary / thresh) * thresh

    def vectorized(self, n_envs, tensor_mode=True):
        """N copies on the GPU following ElegantRL's vectorised convention (device tensors in/out)."""
        eng = BatchedNpStockTradingEnv(tables=self.engine.tables, n_envs=n_envs, device=self._device,
                                       if_train=self.if_train, **self._kw)
        return BatchedVecEnv(eng, tensor_mode=tensor_mode)
