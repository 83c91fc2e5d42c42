#!/usr/bin/env python
"""Small invocations of every kernel/path for compute-sanitizer (memcheck / racecheck / initcheck):
ragged tiles (N not a multiple of 32), both action layouts, all obs modes, unaligned days, resets."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from finrl_b200 import (BatchedNpStockTradingEnv, BatchedStockPortfolioEnv, BatchedStockTradingEnv,  # noqa: E402
                        BatchedStockTradingEnvCashpenalty, CashPenaltyTables, PortfolioTables, TradingTables,
                        synthetic as syn)


def main():
    for D in (30, 13, 5):
        N, T, K = 70, 12, 2
        close, tech, turb = syn.make_tables(T, D, K, seed=0)
        env = BatchedStockTradingEnv(tables=TradingTables.from_arrays(close, tech, turb, "cuda"), n_envs=N, hmax=100,
                                     initial_amount=50_000, turbulence_threshold=60, track_asset=True)
        for dt in (np.float32, np.float64):
            a = torch.from_numpy(syn.make_actions((T + 3, N, D), seed=1, dtype=dt)).cuda()
            for s in range(3):
                env.step(a[s], auto_reset=True, accumulate_stats=True)
            env.rollout(a, layout="KND", obs_mode="all")
            env.rollout(a.permute(1, 0, 2).contiguous(), layout="NKD", obs_mode="last")
            env.rollout(a, obs_mode="none")
        env.set_state(day=np.arange(N) % T, sday=np.arange(N) % T)
        env.step(a[0].float())
        env.reset(mask=(np.arange(N) % 2).astype(np.uint8))
        env.observe()

        pa, ta, tu = syn.make_np_arrays(close, tech, turb)
        nenv = BatchedNpStockTradingEnv({"price_array": pa, "tech_array": ta, "turbulence_array": tu, "if_train": True},
                                        n_envs=N, turbulence_thresh=60)
        a = torch.from_numpy(syn.make_actions((T + 3, N, D), seed=2)).cuda()
        for s in range(3):
            nenv.step(a[s], accumulate_stats=True)
        nenv.rollout(a, obs_mode="all")
        nenv.rollout(a.permute(1, 0, 2).contiguous(), layout="NKD")
        nenv.rollout(a.double(), obs_mode="none")
        nenv.reset(mask=(np.arange(N) % 3 == 0).astype(np.uint8))

        c2, t2, _ = syn.make_tables(T + 6, D, K, seed=1)
        cov, first = syn.make_cov_table(c2, 6)
        penv = BatchedStockPortfolioEnv(tables=PortfolioTables.from_arrays(c2[first:], cov, t2[:, first:], "cuda"), n_envs=N)
        a = torch.from_numpy(syn.make_actions((T + 3, N, D), seed=3, low=0, high=1, dtype=np.float64)).cuda()
        for s in range(3):
            penv.step(a[s], auto_reset=True, accumulate_stats=True)
        penv.rollout(a, obs_mode="all")
        penv.rollout(a.float().permute(1, 0, 2).contiguous(), layout="NKD", obs_mode="last")
        penv.observe()

    for D in (100, 128, 7):
        N, T = 37, 10
        close, _, turb = syn.make_tables(T, D, 0, seed=4)
        o, h, l, v = syn.make_ohlv(close, 4)
        cenv = BatchedStockTradingEnvCashpenalty(
            tables=CashPenaltyTables.from_arrays(close, np.stack([o, close, h, l, v], axis=2), turb, "cuda"), n_envs=N,
            random_start=False, turbulence_threshold=60, hmax=30000, initial_amount=1e5)
        a = torch.from_numpy(syn.make_actions((T + 3, N, D), seed=5)).cuda()
        for s in range(3):
            cenv.step(a[s], auto_reset=True, accumulate_stats=True)
        cenv.rollout(a, obs_mode="all")
        cenv.rollout(a.double().permute(1, 0, 2).contiguous(), layout="NKD", obs_mode="last")
        cenv.reset(mask=(np.arange(N) % 2).astype(np.uint8), start_points=np.arange(N) % 3)
        cenv.observe()
    torch.cuda.synchronize()
    print("sanitize cases ok")


if __name__ == "__main__":
    main()
