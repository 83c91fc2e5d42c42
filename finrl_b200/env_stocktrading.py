"""Drop-in for ``finrl.meta.env_stock_trading.env_stocktrading.StockTradingEnv``.

Same constructor signature, attributes and gym protocol as the reference class
(/root/reference/finrl/meta/env_stock_trading/env_stocktrading.py:19-552); ``step`` / ``reset`` run on
the GPU through a 1-env :class:`finrl_b200.trading.BatchedStockTradingEnv` (there is no CPU path).  The
Python side only converts the device state back into the reference's list-of-floats ``state`` and keeps
the episode memories (``asset_memory`` etc.) that ``DRLAgent.DRL_prediction`` reads
(finrl/agents/stablebaselines3/models.py:112-125).  For throughput use ``get_vec_env(n_envs)``.
"""
from __future__ import annotations

import numpy as np

from .spaces import Box, gym_env_base
from .trading import BatchedStockTradingEnv
from .vec_env import BatchedVecEnv, dummy_vec_env


class StockTradingEnv(gym_env_base()):
    metadata = {"render.modes": ["human"]}

    def __init__(self, df, stock_dim, hmax, initial_amount, num_stock_shares, buy_cost_pct, sell_cost_pct,
                 reward_scaling, state_space, action_space, tech_indicator_list, turbulence_threshold=None,
                 risk_indicator_col="turbulence", make_plots=False, print_verbosity=10, day=0, initial=True,
                 previous_state=[], model_name="", mode="", iteration="", device="cuda"):
        self.day = day
        self.df = df
        self.stock_dim = stock_dim
        self.hmax = hmax
        self.num_stock_shares = num_stock_shares
        self.initial_amount = initial_amount
        self.buy_cost_pct = buy_cost_pct
        self.sell_cost_pct = sell_cost_pct
        self.reward_scaling = reward_scaling
        self.state_space = state_space
        self.tech_indicator_list = tech_indicator_list
        self.action_space = Box(low=-1, high=1, shape=(action_space,))
        self.observation_space = Box(low=-np.inf, high=np.inf, shape=(state_space,))
        self.terminal = False
        self.make_plots = make_plots
        self.print_verbosity = print_verbosity
        self.turbulence_threshold = turbulence_threshold
        self.risk_indicator_col = risk_indicator_col
        self.initial = initial
        self.previous_state = previous_state
        self.model_name, self.mode, self.iteration = model_name, mode, iteration
        self._kw = dict(
            stock_dim=stock_dim, hmax=hmax, initial_amount=initial_amount, num_stock_shares=num_stock_shares,
            buy_cost_pct=buy_cost_pct, sell_cost_pct=sell_cost_pct, reward_scaling=reward_scaling, state_space=state_space,
            action_space=action_space, tech_indicator_list=tech_indicator_list, turbulence_threshold=turbulence_threshold,
            risk_indicator_col=risk_indicator_col, day=day, initial=initial, previous_state=previous_state,
        )
        self._device = device
        self.engine = BatchedStockTradingEnv(df, n_envs=1, device=device, track_asset=True, **self._kw)
        self._tables = self.engine.tables
        self._dates = df.date.to_numpy().reshape(self._tables.n_days, stock_dim)[:, 0] if "date" in df.columns else None
        self.state = self._pull_state()
        self.reward = 0
        self.turbulence = 0
        self.cost = 0
        self.trades = 0
        self.episode = 0
        # the constructor's first entry is initial_amount + shares * prices even when resuming (:85-91)
        self.asset_memory = [self.initial_amount + np.sum(np.array(self.num_stock_shares) * np.array(self.state[1 : 1 + self.stock_dim]))]
        self.rewards_memory = []
        self.actions_memory = []
        self.state_memory = []
        self.date_memory = [self._get_date()]
        self._seed()

    # ------------------------------------------------------------------------------------------
    @property
    def data(self):
        return self.df.loc[self.day, :]

    def _pull_state(self):
        """Device state -> the reference's state list [cash] + close + holdings + tech (fp64 values)."""
        e = self.engine
        self._cash = float(e.cash[0].item())
        self._hold = e.hold[:, 0].cpu().numpy().astype(np.int64)
        sday = int(e.sday[0].item())
        self._sd = -sday - 1 if sday < 0 else sday
        t = self._tables
        return ([self._cash] + t.host_close[self._sd].tolist() + [int(h) for h in self._hold]
                + t.host_tech[:, self._sd, :].reshape(-1).tolist())

    def _priced_holdings(self):
        D = self.stock_dim
        return np.array(self.state[1 : D + 1]) * np.array(self.state[D + 1 : 2 * D + 1])

    def _reset_asset(self):
        # first asset_memory entry after reset() (:364-378): bookkeeping of the episode log, not the step path
        D = self.stock_dim
        if self.initial:
            return self.initial_amount + np.sum(np.array(self.num_stock_shares) * np.array(self.state[1 : 1 + D]))
        return self.previous_state[0] + sum(np.array(self.state[1 : D + 1]) * np.array(self.previous_state[D + 1 : 2 * D + 1]))

    def _get_date(self):
        return None if self._dates is None else self._dates[self.day]

    # ------------------------------------------------------------------------------------------
    def step(self, actions):
        import torch

        e = self.engine
        a = np.asarray(actions)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)  # python ints/floats: `actions * hmax` would be float64
        hold_before = self._hold
        begin_total_asset = self.state[0] + sum(self._priced_holdings())  # for rewards_memory only (:311-314)
        obs, reward, done, flags = e.step(torch.as_tensor(a.reshape(1, -1)), auto_reset=False, want_obs=False)
        self.terminal = bool(done[0].item())
        if self.terminal:
            # terminal branch (:221-301): no state change, previous scaled reward again
            self._episode_report()
            return self.state, self.reward, self.terminal, {}
        self.day = int(e.day[0].item())
        self.state = self._pull_state()
        self.cost = float(e.cost[0].item())
        self.trades = int(e.trades[0].item())
        if self.turbulence_threshold is not None:
            self.turbulence = float(self._tables.host_risk[self.day])
        end_total_asset = float(e.asset[0].item())
        self.actions_memory.append(self._hold - hold_before)  # executed shares, signed (:324, :330)
        self.asset_memory.append(end_total_asset)
        self.date_memory.append(self._get_date())
        self.reward = float(reward[0].item())
        self.rewards_memory.append(end_total_asset - begin_total_asset)  # the unscaled reward (:350-351)
        self.state_memory.append(self.state)
        return self.state, self.reward, self.terminal, {}

    def reset(self):
        self.engine.reset()
        self.state = self._pull_state()  # built from the rows of the day still loaded (stale-day quirk Q1)
        self.asset_memory = [self._reset_asset()]
        self.day = 0
        self.turbulence = 0
        self.cost = 0
        self.trades = 0
        self.terminal = False
        self.rewards_memory = []
        self.actions_memory = []
        self.date_memory = [self._get_date()]
        self.episode += 1
        return self.state

    def render(self, mode="human", close=False):
        return self.state

    def _make_plot(self):
        try:
            import matplotlib.pyplot as plt
        except Exception:  # plotting is optional here; the reference imports matplotlib unconditionally
            return
        plt.plot(self.asset_memory, "r")
        plt.savefig(f"results/account_value_trade_{self.episode}.png")
        plt.close()

    def _episode_report(self):
        """What the reference's terminal branch does besides returning (:223-290): the episode summary print
        every ``print_verbosity`` episodes and, when ``model_name`` and ``mode`` are set (the ensemble agent's
        validation / trade envs), the three CSV files the agent reads back (models.py:208-217)."""
        import pandas as pd

        if self.make_plots:
            self._make_plot()
        end_total_asset = self.state[0] + sum(self._priced_holdings())
        tot_reward = end_total_asset - self.asset_memory[0]
        df_total_value = pd.DataFrame({"account_value": self.asset_memory})
        df_total_value["date"] = self.date_memory
        df_total_value["daily_return"] = df_total_value["account_value"].pct_change(1)
        sd = df_total_value["daily_return"].std()
        df_rewards = pd.DataFrame({"account_rewards": self.rewards_memory})
        df_rewards["date"] = self.date_memory[:-1]
        if self.episode % self.print_verbosity == 0:
            print(f"day: {self.day}, episode: {self.episode}")
            print(f"begin_total_asset: {self.asset_memory[0]:0.2f}")
            print(f"end_total_asset: {end_total_asset:0.2f}")
            print(f"total_reward: {tot_reward:0.2f}")
            print(f"total_cost: {self.cost:0.2f}")
            print(f"total_trades: {self.trades}")
            if sd != 0:
                print(f"Sharpe: {(252 ** 0.5) * df_total_value['daily_return'].mean() / sd:0.3f}")
            print("=================================")
        if self.model_name != "" and self.mode != "":
            tag = f"{self.mode}_{self.model_name}_{self.iteration}"
            self.save_action_memory().to_csv(f"results/actions_{tag}.csv")
            df_total_value.to_csv(f"results/account_value_{tag}.csv", index=False)
            df_rewards.to_csv(f"results/account_rewards_{tag}.csv", index=False)
            try:
                import matplotlib.pyplot as plt

                plt.plot(self.asset_memory, "r")
                plt.savefig(f"results/account_value_{tag}.png")
                plt.close()
            except Exception:
                pass

    # ---- logging / adapter surface (:488-552) ----------------------------------------------------
    def save_asset_memory(self):
        import pandas as pd

        return pd.DataFrame({"date": self.date_memory, "account_value": self.asset_memory})

    def save_action_memory(self):
        import pandas as pd

        date_list = self.date_memory[:-1]
        if self.stock_dim > 1:
            df_actions = pd.DataFrame(self.actions_memory)
            df_actions.columns = self.data.tic.values
            df_actions.index = pd.Index(date_list, name="date")
            return df_actions
        return pd.DataFrame({"date": date_list, "actions": self.actions_memory})

    def save_state_memory(self):
        import pandas as pd

        return pd.DataFrame({"date": self.date_memory[:-1], "states": self.state_memory})

    def _seed(self, seed=None):
        self.np_random = np.random.RandomState(seed)
        return [seed]

    def get_sb_env(self):
        """``DummyVecEnv([lambda: self])`` + its first observation, like the reference (:549-552): the VecEnv
        steps THIS object, so ``env_method("save_asset_memory")``, ``render()`` and the memories work as
        ``DRLAgent.DRL_prediction`` expects (finrl/agents/stablebaselines3/models.py:110-130)."""
        e = dummy_vec_env([lambda: self])
        obs = e.reset()
        return e, obs

    def get_vec_env(self, n_envs, tensor_mode=False, record_memory=False):
        """N copies of this env on the GPU behind the SB3 VecEnv protocol (finrl_b200.vec_env);
        ``record_memory=True`` keeps per-env asset / action memories for ``env_method("save_*_memory")``."""
        eng = BatchedStockTradingEnv(tables=self._tables, n_envs=n_envs, device=self._device, track_asset=record_memory,
                                     **{k: v for k, v in self._kw.items() if k not in ("state_space", "action_space")})
        tics = list(self.df.tic.values[: self.stock_dim]) if "tic" in self.df.columns else None
        return BatchedVecEnv(eng, tensor_mode=tensor_mode, record_memory=record_memory, dates=self._dates, tickers=tics)
