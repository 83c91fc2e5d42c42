"""Drop-in for ``finrl.meta.env_portfolio_allocation.env_portfolio.StockPortfolioEnv``.

Same constructor, attributes and gym protocol as the reference class
(/root/reference/finrl/meta/env_portfolio_allocation/env_portfolio.py:15-261); ``step`` / ``reset`` run on
the GPU through a 1-env :class:`finrl_b200.portfolio.BatchedStockPortfolioEnv`.  ``state`` is returned as
the float64 (D+K, D) matrix of the reference (it is a pure per-day table row).
"""
from __future__ import annotations

import numpy as np

from .portfolio import BatchedStockPortfolioEnv
from .spaces import Box, gym_env_base
from .vec_env import BatchedVecEnv, dummy_vec_env


class StockPortfolioEnv(gym_env_base()):
    metadata = {"render.modes": ["human"]}

    def __init__(self, df, stock_dim, hmax, initial_amount, transaction_cost_pct, reward_scaling, state_space,
                 action_space, tech_indicator_list, turbulence_threshold=None, lookback=252, day=0, device="cuda"):
        self.day, self.lookback, self.df = day, lookback, df
        self.stock_dim, self.hmax, self.initial_amount = stock_dim, hmax, initial_amount
        self.transaction_cost_pct, self.reward_scaling = transaction_cost_pct, reward_scaling
        self.state_space, self.tech_indicator_list = state_space, tech_indicator_list
        self.action_space = Box(low=0, high=1, shape=(action_space,))
        self.observation_space = Box(low=-np.inf, high=np.inf, shape=(state_space + len(tech_indicator_list), state_space))
        self._device = device
        self.engine = BatchedStockPortfolioEnv(df, stock_dim=stock_dim, initial_amount=initial_amount, state_space=state_space,
                                               tech_indicator_list=tech_indicator_list, day=day, n_envs=1, device=device,
                                               track_weights=True)
        self._dates = df.date.to_numpy().reshape(self.engine.n_days, stock_dim)[:, 0] if "date" in df.columns else None
        self.state = self.engine.tables.host_obs[self.day]
        self.covs = self.state[:stock_dim]
        self.terminal = False
        self.turbulence_threshold = turbulence_threshold
        self.portfolio_value = self.initial_amount
        self.asset_memory = [self.initial_amount]
        self.portfolio_return_memory = [0]
        self.actions_memory = [[1 / self.stock_dim] * self.stock_dim]
        self.date_memory = [self._date()]

    @property
    def data(self):
        return self.df.loc[self.day, :]

    def _date(self):
        return None if self._dates is None else self._dates[self.day]

    def step(self, actions):
        import torch

        a = np.asarray(actions)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)
        obs, reward, done, flags = self.engine.step(torch.as_tensor(a.reshape(1, -1)), want_obs=False)
        self.terminal = bool(done[0].item())
        if self.terminal:
            print("=================================")
            print(f"begin_total_asset:{self.asset_memory[0]}")
            print(f"end_total_asset:{self.portfolio_value}")
            r = np.asarray(self.portfolio_return_memory, dtype=np.float64)
            if r.size > 1 and r.std(ddof=1) != 0:
                print("Sharpe: ", (252**0.5) * r.mean() / r.std(ddof=1))
            print("=================================")
            return self.state, self.reward, self.terminal, {}
        self.day = int(self.engine.day[0].item())
        self.state = self.engine.tables.host_obs[self.day]
        self.covs = self.state[: self.stock_dim]
        self.portfolio_value = float(self.engine.portfolio_value[0].item())
        self.reward = float(reward[0].item())
        # logging memories straight from the kernel (portfolio_return and the softmax weights it used)
        self.portfolio_return_memory.append(float(self.engine.last_return[0].item()))
        self.actions_memory.append(self.engine.last_weights[0].cpu().numpy())
        self.date_memory.append(self._date())
        self.asset_memory.append(self.portfolio_value)
        return self.state, self.reward, self.terminal, {}

    def reset(self):
        self.engine.reset(want_obs=False)
        self.asset_memory = [self.initial_amount]
        self.day = 0
        self.state = self.engine.tables.host_obs[0]
        self.covs = self.state[: self.stock_dim]
        self.portfolio_value = self.initial_amount
        self.terminal = False
        self.portfolio_return_memory = [0]
        self.actions_memory = [[1 / self.stock_dim] * self.stock_dim]
        self.date_memory = [self._date()]
        return self.state

    def render(self, mode="human"):
        return self.state

    def softmax_normalization(self, actions):
        """The reference's public helper (:225-229).  ``step`` does not use it — the weights are computed
        on the device — it is kept for callers that invoke it directly."""
        e = np.exp(actions)
        return e / np.sum(e)

    def save_asset_memory(self):
        import pandas as pd

        return pd.DataFrame({"date": self.date_memory, "daily_return": self.portfolio_return_memory})

    def save_action_memory(self):
        import pandas as pd

        df_actions = pd.DataFrame(self.actions_memory)
        df_actions.columns = self.data.tic.values
        df_actions.index = pd.Index(self.date_memory, name="date")
        return df_actions

    def _seed(self, seed=None):
        self.np_random = np.random.RandomState(seed)
        return [seed]

    def get_sb_env(self):
        """``DummyVecEnv([lambda: self])`` + first observation, as the reference (env_portfolio.py:258-261)."""
        e = dummy_vec_env([lambda: self])
        obs = e.reset()
        return e, obs

    def get_vec_env(self, n_envs, tensor_mode=False):
        eng = BatchedStockPortfolioEnv(tables=self.engine.tables, n_envs=n_envs, device=self._device,
                                       initial_amount=self.initial_amount)
        return BatchedVecEnv(eng, action_low=0.0, action_high=1.0, tensor_mode=tensor_mode)
