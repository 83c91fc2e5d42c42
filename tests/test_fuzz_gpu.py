"""GPU: randomised configurations (shapes, costs, thresholds, dtypes, ragged N) of the two bit-exact
envs against the CPU oracle — every state array after every step."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.mark.parametrize("kernel", ["tile", "small"])
@pytest.mark.parametrize("seed", range(14))
def test_trading_fuzz(seed, kernel):
    from finrl_b200 import BatchedStockTradingEnv, TradingTables, _cabi, synthetic as syn
    from oracle import oracle as ora

    _cabi.set_option("trading_small_max", 0 if kernel == "tile" else 2**31 - 1)
    _cabi.set_option("trading_wide_min_envs", 0 if kernel == "tile" else 2**31 - 1)  # D > 32: wide vs 8 lanes

    rng = np.random.default_rng(1000 + seed)
    D = int(rng.integers(1, 33)) if seed < 10 else int(rng.integers(33, 129))  # > 32: thread-per-env wide kernel ("tile") or 8 lanes per env ("small")
    K = int(rng.integers(0, 5))
    T = int(rng.integers(4, 36))
    N = int(rng.choice([1, 31, 33, 100]))
    dtype = [np.float32, np.float64][seed % 2]
    kw = dict(
        hmax=int(rng.choice([1, 3, 10, 100, 1000])), initial_amount=float(rng.choice([1e3, 1e5, 1e6])),
        buy_cost_pct=float(rng.choice([0.0, 0.001, 0.01])), sell_cost_pct=float(rng.choice([0.0, 0.001, 0.02])),
        reward_scaling=float(rng.choice([1e-4, 1.0])), turbulence_threshold=[None, 40, 99][seed % 3],
        num_stock_shares=[int(v) for v in rng.integers(0, 4, D)],
    )
    close, tech, turb = syn.make_tables(T, D, K, seed=seed)
    if K > 0:
        tech[0][rng.random((T, D)) < 0.05] = 1.0  # sprinkle "disable" flags
    env = BatchedStockTradingEnv(tables=TradingTables.from_arrays(close, tech, turb, "cuda"), n_envs=N, **kw)
    o = ora.TradingOracle(close, tech, turb, N, **kw)
    acts = (syn.make_actions((2 * T + 3, N, D), seed=seed, dtype=np.float64) * float(rng.choice([1.0, 1.7]))).astype(dtype)
    assert np.array_equal(env.observe().cpu().numpy(), o.obs())
    for s in range(acts.shape[0]):
        auto = bool((s // 7) % 2 == 0)
        obs, reward, done, flags = env.step(torch.from_numpy(acts[s]).cuda(), auto_reset=auto)
        oobs, orew, ofl = o.step(acts[s], auto_reset=auto)
        ctx = f"seed {seed} step {s} D={D} K={K} T={T} N={N} {kw}"
        assert np.array_equal(flags.cpu().numpy(), ofl), ctx
        assert np.array_equal(reward.cpu().numpy(), orew), ctx
        assert np.array_equal(obs.cpu().numpy(), oobs), ctx
        st = env.get_state()
        for name, ref in (("cash", o.cash), ("hold", o.hold), ("day", o.day), ("sday", o.sday), ("cost", o.cost),
                          ("trades", o.trades), ("episode", o.episode)):
            assert np.array_equal(st[name].cpu().numpy(), ref), ctx + " " + name
        if not auto and s % 7 == 6:
            assert np.array_equal(env.reset().cpu().numpy(), o.reset()), ctx
    _cabi.set_option("trading_small_max", 8192)
    _cabi.set_option("trading_wide_min_envs", 3072)


@pytest.mark.parametrize("seed", range(16))
def test_np_fuzz(seed):
    """Seeds 0-9: D in 1..32 (register kernel; odd seeds forced through the streaming kernel as well);
    seeds 10-15: D in 33..128 (streaming kernel, np_wide.cu)."""
    from finrl_b200 import BatchedNpStockTradingEnv, _cabi, synthetic as syn
    from oracle import oracle as ora

    rng = np.random.default_rng(2000 + seed)
    D = int(rng.integers(1, 33)) if seed < 10 else int(rng.integers(33, 129))
    if seed in (3, 7, 11, 13, 15):
        D = (D + 3) // 4 * 4  # multiples of four take the streaming kernel's bulk-staged variant
    _cabi.set_option("np_wide_min_d", 1 if seed % 2 else 33)
    K = int(rng.integers(0, 4))
    T = int(rng.integers(4, 40))
    N = int(rng.choice([1, 31, 65]))
    kw = dict(gamma=float(rng.choice([0.99, 0.9])), turbulence_thresh=float(rng.choice([30, 99])),
              min_stock_rate=float(rng.choice([0.0, 0.1, 0.3])), max_stock=float(rng.choice([1e2, 37.0, 5.0])),
              initial_capital=float(rng.choice([1e6, 3e4, 2e3])), buy_cost_pct=float(rng.choice([1e-3, 0.0])),
              sell_cost_pct=float(rng.choice([1e-3, 5e-3])), reward_scaling=float(rng.choice([2**-11, 1e-3])),
              initial_stocks=rng.integers(0, 5, D).astype(np.float32))
    close, tech, turb = syn.make_tables(T, D, max(K, 1), seed=seed)
    pa, ta, tu = syn.make_np_arrays(close, tech[:K] if K else tech[:0], turb)
    cfg = {"price_array": pa, "tech_array": ta.reshape(T, D * K), "turbulence_array": tu, "if_train": False}
    env = BatchedNpStockTradingEnv(cfg, n_envs=N, **kw)
    o = ora.NpTradingOracle(pa, ta.reshape(T, D * K), tu, N, **kw)
    acts = syn.make_actions((2 * T + 1, N, D), seed=seed)
    if seed % 2:
        acts[: T // 2] *= 0.06
    rs = np.random.RandomState(seed)

    def reset_both():
        s0 = (kw["initial_stocks"] + rs.randint(0, 64, size=(N, D))).astype(np.float32) if seed % 3 else None
        f = rs.uniform(0.95, 1.05, size=N) if seed % 3 else None
        assert np.array_equal(env.reset(stocks0=s0, factor=f).cpu().numpy(), o.reset(stocks0=s0, factor=f))

    reset_both()
    for s in range(acts.shape[0]):
        obs, reward, done, flags = env.step(torch.from_numpy(acts[s]).cuda())
        oobs, orew, ork, ofl = o.step(acts[s])
        ctx = f"seed {seed} step {s} D={D} K={K} T={T} N={N}"
        assert np.array_equal(flags.cpu().numpy(), ofl | (ork << 4)), ctx
        assert np.array_equal(reward.cpu().numpy(), orew), ctx
        assert np.array_equal(obs.cpu().numpy(), oobs), ctx
        st = env.get_state()
        for name, ref in (("amount", o.amount), ("amount_kind", o.amount_kind), ("stocks", o.stocks), ("cool", o.cool),
                          ("total", o.total), ("total_kind", o.total_kind), ("gamma_reward", o.gamma_reward),
                          ("gr_kind", o.gr_kind), ("day", o.day)):
            assert np.array_equal(st[name].cpu().numpy(), ref), ctx + " " + name
        if done.all():
            reset_both()
    _cabi.set_option("np_wide_min_d", 33)
