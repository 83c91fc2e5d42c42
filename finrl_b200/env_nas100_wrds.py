"""Drop-in for the sibling ``finrl.meta.env_stock_trading.env_nas100_wrds.StockEnvNAS100``.

Its ``step`` is the numpy env's step line for line (env_nas100_wrds.py:110-154 vs env_stocktrading_np.py:
103-147); the differences are the constructor's hard-coded train/eval/trade slicing of minute-level WRDS
arrays (:37-51), an always-random ``reset`` (:93-108) and ``max(amount, 1e4)`` in ``get_state`` (:157).
The arrays must be float32 like the ones its ``load_data`` returns (:186-189); ``cwd`` loading from disk
is the caller's job (pass the arrays).
"""
from __future__ import annotations

import numpy as np

from .env_stocktrading_np import StockTradingEnv as _NpEnv


class StockEnvNAS100(_NpEnv):
    def __init__(self, cwd=None, price_ary=None, tech_ary=None, turbulence_ary=None, gamma=0.999, turbulence_thresh=30,
                 min_stock_rate=0.1, max_stock=1e2, initial_capital=1e6, buy_cost_pct=1e-3, sell_cost_pct=1e-3, data_gap=4,
                 reward_scaling=2**-11, ticker_list=None, tech_indicator_list=None, initial_stocks=None, if_eval=False,
                 if_trade=False, device="cuda"):
        if cwd is not None or price_ary is None:
            raise ValueError("pass price_ary / tech_ary / turbulence_ary (float32, as load_data returns them) and cwd=None")
        if np.asarray(price_ary).dtype != np.float32 or np.asarray(tech_ary).dtype != np.float32:
            raise TypeError("StockEnvNAS100 arrays must be float32 (the reference's promotion rules depend on it)")
        beg_i, mid_i, end_i = 0, 211210, 422420
        i0, i1 = (beg_i, mid_i) if if_eval else (mid_i, end_i)
        sl = slice(422420, 528026, data_gap) if if_trade else slice(i0, i1, data_gap)
        cfg = {"price_array": np.asarray(price_ary)[sl], "tech_array": np.asarray(tech_ary)[sl],
               "turbulence_array": np.asarray(turbulence_ary)[sl], "if_train": True}  # reset always randomises (:93-108)
        super().__init__(cfg, gamma=gamma, turbulence_thresh=turbulence_thresh, min_stock_rate=min_stock_rate,
                         max_stock=max_stock, initial_capital=initial_capital, buy_cost_pct=buy_cost_pct,
                         sell_cost_pct=sell_cost_pct, reward_scaling=reward_scaling, initial_stocks=initial_stocks,
                         device=device, obs_amount_floor=1e4)
        self.env_name = "StockEnvNAS"
        self.target_return = 2.2

    @property
    def stocks_cd(self):
        return self.stocks_cool_down
