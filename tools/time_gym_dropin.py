#!/usr/bin/env python
"""GPU box: step rate of the gym-protocol drop-ins at the reference's own scale (BASELINE config 1: ONE env,
DOW-30, 2500 days, numpy in / numpy out) and of the SB3-style VecEnv in numpy mode at a few batch sizes.  The
reference's gym loop does ~140-160 env-steps/s on one core (profiles/r01_reference_cpu_timing.txt)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from finrl_b200 import synthetic as syn  # noqa: E402


def rate(fn, n, warm=20):
    for i in range(warm):
        fn(i)
    t0 = time.perf_counter()
    for i in range(n):
        fn(warm + i)
    return n / (time.perf_counter() - t0)


def main():
    from finrl_b200.env_stocktrading import StockTradingEnv

    T, D, K = 2500, 30, 8
    close, tech, turb = syn.make_tables(T, D, K, seed=0)
    df = syn.make_frame(close, tech, turb)
    env = StockTradingEnv(df=df, stock_dim=D, hmax=100, initial_amount=1_000_000, num_stock_shares=[0] * D,
                          buy_cost_pct=0.001, sell_cost_pct=0.001, reward_scaling=1e-4, state_space=1 + 2 * D + K * D,
                          action_space=D, tech_indicator_list=syn.INDICATORS[:K], turbulence_threshold=99,
                          print_verbosity=10**9)
    acts = syn.make_actions((2400, D), seed=1)
    env.reset()
    r = rate(lambda i: env.step(acts[i]), 2000)
    print(f"gym drop-in StockTradingEnv.step (1 env, python list state out): {r:10.1f} env-steps/s")
    for n in (1, 8, 64, 1024, 65536):
        vec = env.get_vec_env(n) if hasattr(env, "get_vec_env") else None
        if vec is None:
            break
        vec.reset()
        a = syn.make_actions((8, n, D), seed=2)
        steps = 400 if n <= 1024 else 60
        r = rate(lambda i: vec.step(a[i % 8]), steps, warm=5)
        print(f"BatchedVecEnv numpy mode, {n:6d} envs: {r:9.1f} steps/s = {r * n:14.1f} env-steps/s")
        if n >= 1024:
            vec.copy_outputs = False
            r = rate(lambda i: vec.step(a[i % 8]), steps, warm=5)
            print(f"    copy_outputs=False (views of the pinned buffers): {r:9.1f} steps/s = {r * n:14.1f} env-steps/s")


if __name__ == "__main__":
    main()
