import sys; sys.path.insert(0, '/root/repo')
import torch, numpy as np
from finrl_b200 import BatchedStockTradingEnv, TradingTables, _cabi, synthetic as syn
T, D, K = 400, 100, 2
close, tech, turb = syn.make_tables(T, D, K, seed=0)
tb = TradingTables.from_arrays(close, tech, turb, "cuda")
kw = dict(hmax=100, initial_amount=1_000_000, buy_cost_pct=0.001, sell_cost_pct=0.001, reward_scaling=1e-4, turbulence_threshold=99)
for N in (1024, 2048, 4096, 8192, 16384, 32768, 65536):
    res = []
    for mode, wmin in (("small", 2**31 - 1), ("wide", 0)):
        _cabi.set_option("trading_wide_min_envs", wmin)
        env = BatchedStockTradingEnv(tables=tb, n_envs=N, **kw)
        a = [torch.rand((N, D), device="cuda") * 2 - 1 for _ in range(4)]
        for i in range(5): env.step(a[i % 4])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(50): env.step(a[i % 4], want_done=False)
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 50)
    print(N, "small %.4f ms  wide %.4f ms" % tuple(res))
