"""Drop-in for ``finrl.meta.env_stock_trading.env_stocktrading_cashpenalty.StockTradingEnvCashpenalty``.

Same constructor and gym protocol as the reference class
(/root/reference/finrl/meta/env_stock_trading/env_stocktrading_cashpenalty.py:19-409); ``step`` / ``reset``
run on the GPU through a 1-env :class:`finrl_b200.cashpenalty.BatchedStockTradingEnvCashpenalty`.
``get_sb_env`` is the reference's ``DummyVecEnv`` over a deep copy of this object; ``get_multiproc_env(n)``
returns a batched GPU VecEnv of n envs instead of forking n processes.
"""
from __future__ import annotations

import numpy as np

from .cashpenalty import BatchedStockTradingEnvCashpenalty
from .spaces import Box, gym_env_base
from .vec_env import BatchedVecEnv, _DeepCopyVecMixin


class StockTradingEnvCashpenalty(_DeepCopyVecMixin, gym_env_base()):
    metadata = {"render.modes": ["human"]}

    def __init__(self, df, buy_cost_pct=3e-3, sell_cost_pct=3e-3, date_col_name="date", hmax=10, discrete_actions=False,
                 shares_increment=1, turbulence_threshold=None, print_verbosity=10, initial_amount=1e6,
                 daily_information_cols=["open", "close", "high", "low", "volume"], cache_indicator_data=True,
                 cash_penalty_proportion=0.1, random_start=True, patient=False, currency="$", device="cuda"):
        self._kw = dict(buy_cost_pct=buy_cost_pct, sell_cost_pct=sell_cost_pct, date_col_name=date_col_name, hmax=hmax,
                        discrete_actions=discrete_actions, shares_increment=shares_increment,
                        turbulence_threshold=turbulence_threshold, print_verbosity=print_verbosity,
                        initial_amount=initial_amount, daily_information_cols=list(daily_information_cols),
                        cache_indicator_data=cache_indicator_data, cash_penalty_proportion=cash_penalty_proportion,
                        random_start=random_start, patient=patient, currency=currency)
        self._device = device
        self.engine = e = BatchedStockTradingEnvCashpenalty(df, n_envs=1, device=device, **self._kw)
        self.assets, self.dates = e.assets, e.dates
        self.df = df.set_index(date_col_name)  # like the reference (:79): DRL_prediction counts df.index.unique()
        self.random_start, self.discrete_actions, self.patient, self.currency = random_start, discrete_actions, patient, currency
        self.shares_increment, self.hmax, self.initial_amount = shares_increment, hmax, initial_amount
        self.print_verbosity = print_verbosity
        self.buy_cost_pct, self.sell_cost_pct = buy_cost_pct, sell_cost_pct
        self.turbulence_threshold = turbulence_threshold
        self.daily_information_cols = list(daily_information_cols)
        self.cash_penalty_proportion = cash_penalty_proportion
        self.state_space = e.state_space
        self.action_space = Box(low=-1, high=1, shape=(e.stock_dim,))
        self.observation_space = Box(low=-np.inf, high=np.inf, shape=(self.state_space,))
        self.turbulence = 0
        self.episode = -1
        self.episode_history = []
        self.reset()

    def _pull(self):
        e = self.engine
        self.date_index = int(e.date_index[0].item())
        self.starting_point = int(e.starting_point[0].item())
        cash = float(e.cash[0].item())
        hold = e.holdings[0].cpu().numpy()
        info = e.tables.obs_tmpl[self.date_index, 1 + e.stock_dim :].double().cpu().numpy()
        return np.concatenate([[cash], hold, info])

    @property
    def current_step(self):
        return self.date_index - self.starting_point

    @property
    def cash_on_hand(self):
        return self.state_memory[-1][0]

    @property
    def holdings(self):
        return self.state_memory[-1][1 : len(self.assets or range(self.engine.stock_dim)) + 1]

    def seed(self, seed=None):
        import random
        import time

        random.seed(int(round(time.time() * 1000)) if seed is None else seed)

    def reset(self):
        self.seed()
        self.sum_trades = 0
        self.engine.reset()
        self.turbulence = 0
        self.episode += 1
        self.actions_memory, self.transaction_memory = [], []
        self.account_information = {"cash": [], "asset_value": [], "total_assets": [], "reward": []}
        init_state = self._pull()
        self.state_memory = [init_state]
        return init_state

    def step(self, actions):
        import torch

        a = np.asarray(actions)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)
        e = self.engine
        prev_hold = e.holdings[0].cpu().numpy()
        obs, reward, done, flags = e.step(torch.as_tensor(a.reshape(1, -1)), want_obs=False)
        self.sum_trades = float(e.sum_trades[0].item())
        reward = float(reward[0].item())
        if bool(done[0].item()):
            return self.state_memory[-1], reward, True, {}
        cash, total = float(e.last_cash[0].item()), float(e.last_total[0].item())
        self.account_information["cash"].append(cash)
        self.account_information["asset_value"].append(total - cash)
        self.account_information["total_assets"].append(total)
        self.account_information["reward"].append(reward)
        state = self._pull()
        self.actions_memory.append(a)
        self.transaction_memory.append(state[1 : 1 + e.stock_dim] - prev_hold)
        if self.turbulence_threshold is not None:
            self.turbulence = float(e.tables.turb[self.date_index].item())
        self.state_memory.append(state)
        return state, reward, False, {}

    def get_multiproc_env(self, n=10):
        e = self.get_vec_env(n)
        return e, e.reset()

    def get_vec_env(self, n_envs, tensor_mode=False):
        return BatchedVecEnv(self._make_engine(n_envs), tensor_mode=tensor_mode)

    def _make_engine(self, n_envs):
        kw = {k: v for k, v in self._kw.items() if k not in ("date_col_name",)}
        return BatchedStockTradingEnvCashpenalty(tables=self.engine.tables, n_envs=n_envs, device=self._device, **kw)

    def save_asset_memory(self):
        import pandas as pd

        if self.current_step == 0:
            return None
        info = dict(self.account_information)
        if self.dates is not None:
            info["date"] = self.dates[-len(info["cash"]):]
        return pd.DataFrame(info)

    def save_action_memory(self):
        import pandas as pd

        if self.current_step == 0:
            return None
        n = len(self.account_information["cash"])
        return pd.DataFrame({"date": None if self.dates is None else self.dates[-n:], "actions": self.actions_memory,
                             "transactions": self.transaction_memory})
