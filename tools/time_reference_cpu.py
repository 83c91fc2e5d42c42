#!/usr/bin/env python
"""Build container only: time the UNMODIFIED reference envs (pure Python) on this host's CPU, next to the
C oracle port on the same shapes.  The GPU box has no /root/reference, so bench.py's reference arm is the
port; this script documents how far the port is from the real thing (SURVEY.md §6 / §8d method: one core,
time.perf_counter around the loop after a warm-up, random pre-generated actions)."""
import contextlib
import io
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from finrl_b200 import synthetic as syn  # noqa: E402
from oracle import oracle as ora  # noqa: E402
from oracle import ref_loader  # noqa: E402


def timeit(fn, n, warm):
    for i in range(warm):
        fn(i)
    t0 = time.perf_counter()
    for i in range(n):
        fn(warm + i)
    return n / (time.perf_counter() - t0)


def main():
    print(f"host: {os.cpu_count()} logical CPUs, numpy {np.__version__}")
    rates = {"where": f"build container, {os.cpu_count()} logical CPUs, numpy {np.__version__}", "unit": "env-steps/s"}
    T, D, K = 2500, 30, 8
    close, tech, turb = syn.make_tables(T, D, K, seed=0)
    acts = syn.make_actions((4000, D), seed=1)
    quiet = contextlib.redirect_stdout(io.StringIO())
    # A1
    mod = ref_loader.load("env_stocktrading")
    env = mod.StockTradingEnv(df=syn.make_frame(close, tech, turb), stock_dim=D, hmax=100, initial_amount=1_000_000,
                              num_stock_shares=[0] * D, buy_cost_pct=0.001, sell_cost_pct=0.001, reward_scaling=1e-4,
                              state_space=1 + 2 * D + K * D, action_space=D, tech_indicator_list=syn.INDICATORS[:K],
                              turbulence_threshold=99, print_verbosity=10**9)
    with quiet:
        r = timeit(lambda i: env.step(acts[i].copy()), 300, 20)
    print(f"reference StockTradingEnv.step        D=30 K=8 T=2500 : {r:10.1f} env-steps/s (1 core)")
    rates["StockTradingEnv_single_env"] = r
    # config 1, second half: SB3's DummyVecEnv over n = 8 reference envs (a sequential loop in ONE process;
    # env_stocktrading.py:549-552 builds exactly this with n = 1) ...
    frame = syn.make_frame(close, tech, turb)
    mk = lambda: mod.StockTradingEnv(df=frame, stock_dim=D, hmax=100, initial_amount=1_000_000,  # noqa: E731
                                     num_stock_shares=[0] * D, buy_cost_pct=0.001, sell_cost_pct=0.001, reward_scaling=1e-4,
                                     state_space=1 + 2 * D + K * D, action_space=D, tech_indicator_list=syn.INDICATORS[:K],
                                     turbulence_threshold=99, print_verbosity=10**9)
    vec = sys.modules["stable_baselines3.common.vec_env"].DummyVecEnv([mk for _ in range(8)])
    vacts = syn.make_actions((200, 8, D), seed=5)
    with quiet:
        vec.reset()
        r8 = 8 * timeit(lambda i: vec.step(vacts[i]), 60, 5)
    print(f"reference DummyVecEnv x 8 (one process, sequential)    : {r8:10.1f} env-steps/s")
    rates["StockTradingEnv_DummyVecEnv_x8"] = r8
    # ... and the same 8 envs as 8 forked processes stepping at once (what SubprocVecEnv buys on 8 cores)
    import multiprocessing as mp

    def worker(q):
        e = mk()
        a = syn.make_actions((400, D), seed=6)
        with contextlib.redirect_stdout(io.StringIO()):
            q.put(timeit(lambda i: e.step(a[i].copy()), 200, 20))

    ctx = mp.get_context("fork")
    q = ctx.Queue()
    ps = [ctx.Process(target=worker, args=(q,)) for _ in range(8)]
    for pr in ps:
        pr.start()
    rp = sum(q.get() for _ in ps)
    for pr in ps:
        pr.join()
    print(f"reference x 8 forked processes ({os.cpu_count()} logical CPUs)          : {rp:10.1f} env-steps/s aggregate")
    rates["StockTradingEnv_8_processes"] = rp
    o = ora.TradingOracle(close, tech, turb, 1, turbulence_threshold=99)
    print(f"  C port, same single env                            : {timeit(lambda i: o.step(acts[i % 4000][None]), 20000, 100):10.1f}")
    # A2
    mod = ref_loader.load("env_stocktrading_np")
    pa, ta, tu = syn.make_np_arrays(close, tech, turb)
    env = mod.StockTradingEnv({"price_array": pa, "tech_array": ta, "turbulence_array": tu, "if_train": False})
    env.reset()
    r = timeit(lambda i: env.step(acts[i]), 2000, 50)
    print(f"reference numpy StockTradingEnv.step  D=30 K=8 T=2500 : {r:10.1f} env-steps/s (1 core)")
    rates["env_stocktrading_np_single_env"] = r
    o = ora.NpTradingOracle(pa, ta, tu, 1)
    print(f"  C port, same single env                            : {timeit(lambda i: o.step(acts[i % 2000][None]), 2000, 10):10.1f}")
    # A3
    mod = ref_loader.load("env_portfolio")
    c2, t2, _ = syn.make_tables(252 + 350, D, 4, seed=0)
    cov, first = syn.make_cov_table(c2, 252)
    df = syn.make_frame(c2[first:], t2[:, first:], np.zeros(350))
    df["cov_list"] = [cov[t] for t in range(350) for _ in range(D)]
    env = mod.StockPortfolioEnv(df=df, stock_dim=D, hmax=100, initial_amount=1_000_000, transaction_cost_pct=0.001,
                                reward_scaling=1e-4, state_space=D, action_space=D, tech_indicator_list=syn.INDICATORS[:4])
    env.reset()
    pacts = syn.make_actions((400, D), seed=2, low=0, high=1, dtype=np.float64)
    with quiet:
        r = timeit(lambda i: env.step(pacts[i]), 300, 20)
    print(f"reference StockPortfolioEnv.step      D=30 K=4        : {r:10.1f} env-steps/s (1 core)")
    rates["StockPortfolioEnv_single_env"] = r
    # A4 (short slice: the constructor's cache is O(T*D) pandas filters)
    mod = ref_loader.load("env_stocktrading_cashpenalty")
    Tc, Dc = 60, 30
    cc, _, tc = syn.make_tables(Tc, Dc, 0, seed=3)
    o_, h_, l_, v_ = syn.make_ohlv(cc, 3)
    dfc = syn.make_frame(cc, np.zeros((0, Tc, Dc)), tc, tech_names=[], extra_cols={"open": o_, "high": h_, "low": l_, "volume": v_}).reset_index(drop=True)
    with quiet:
        t0 = time.perf_counter()
        env = mod.StockTradingEnvCashpenalty(df=dfc, random_start=False, print_verbosity=10**9, hmax=5000)
        ctor = time.perf_counter() - t0
        env.reset()
        cacts = syn.make_actions((60, Dc), seed=4)
        r = timeit(lambda i: env.step(cacts[i]), 40, 5)
    print(f"reference StockTradingEnvCashpenalty  D=30 T=60       : {r:10.1f} env-steps/s (1 core); constructor {ctor:.1f} s")
    rates["Cashpenalty_single_env"] = r
    import json

    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "python_reference_rates.json")
    json.dump(rates, open(out, "w"), indent=1)
    print("wrote", out)


if __name__ == "__main__":
    main()
