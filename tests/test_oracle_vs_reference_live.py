"""CPU, build container only: fuzz the oracle against the UNMODIFIED reference executed live
(skipped where /root/reference does not exist, e.g. on the GPU box — the committed goldens cover
that side).  Random shapes / costs / thresholds / dtypes, every step compared bit-for-bit."""
import contextlib
import io

import numpy as np
import pytest

from oracle import oracle as ora
from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")


@pytest.fixture(autouse=True, scope="module")
def _numpy_argsort_is_the_pinned_one():
    """The bit-exact claim is pinned to numpy's AVX-512 argsort (a bitonic network whose tie order the oracle
    and the kernels restate; goldens: numpy 2.3.x on an AVX-512 host).  On another ISA / numpy release the
    REFERENCE ITSELF orders tied actions differently, so a mismatch here would say nothing about the oracle."""
    try:
        from numpy._core._multiarray_umath import __cpu_features__ as feats
    except Exception:  # pragma: no cover
        feats = {}
    if not feats.get("AVX512_SKX", False):
        pytest.skip("host CPU has no AVX-512: numpy's argsort tie order differs from the pinned goldens")
    if tuple(int(x) for x in np.__version__.split(".")[:2]) < (2, 0):
        pytest.skip(f"numpy {np.__version__}: NEP-50 promotion and the SIMD argsort need numpy >= 2.0")


@pytest.mark.parametrize("seed", range(10))
def test_trading_oracle_vs_live_reference(seed):
    from finrl_b200 import synthetic as syn

    mod = ref_loader.load("env_stocktrading")
    rng = np.random.default_rng(3000 + seed)
    D = int(rng.integers(2, 33))
    K = int(rng.integers(1, 4))
    T = int(rng.integers(5, 30))
    dtype = [np.float32, np.float64][seed % 2]
    hmax = int(rng.choice([1, 3, 10, 100, 1000]))
    init = int(rng.choice([1_000, 100_000, 1_000_000]))
    bc, sc = float(rng.choice([0.0, 0.001, 0.01])), float(rng.choice([0.0, 0.001, 0.02]))
    thr = [None, 40, 99][seed % 3]
    shares = [int(v) for v in rng.integers(0, 4, D)]
    close, tech, turb = syn.make_tables(T, D, K, seed=seed)
    tech[0][rng.random((T, D)) < 0.05] = 1.0
    df = syn.make_frame(close, tech, turb)
    env = mod.StockTradingEnv(df=df, stock_dim=D, hmax=hmax, initial_amount=init, num_stock_shares=list(shares),
                              buy_cost_pct=bc, sell_cost_pct=sc, reward_scaling=1e-4, state_space=1 + 2 * D + K * D,
                              action_space=D, tech_indicator_list=syn.INDICATORS[:K], turbulence_threshold=thr,
                              print_verbosity=10**9)
    o = ora.TradingOracle(close, tech, turb, 1, hmax=hmax, initial_amount=init, buy_cost_pct=bc, sell_cost_pct=sc,
                          reward_scaling=1e-4, turbulence_threshold=thr, num_stock_shares=shares)
    acts = (syn.make_actions((2 * T + 3, D), seed=seed, dtype=np.float64) * float(rng.choice([1.0, 1.7]))).astype(dtype)
    with contextlib.redirect_stdout(io.StringIO()):
        for s in range(acts.shape[0]):
            state, reward, done, _ = env.step(acts[s].copy())
            if done:
                state = env.reset()
            obs, orew, ofl = o.step(acts[s][None, :], auto_reset=True)
            ctx = f"seed {seed} step {s}"
            assert bool(ofl[0] & 1) == bool(done) and orew[0] == reward, ctx
            assert o.cash[0] == state[0] and np.array_equal(o.hold[0], np.asarray(state[1 + D : 1 + 2 * D], dtype=np.float64)), ctx
            assert np.array_equal(obs[0], np.asarray(state, dtype=np.float64).astype(np.float32)), ctx
            if not done:
                assert o.trades[0] == env.trades and o.cost[0] == env.cost, ctx


@pytest.mark.parametrize("seed", range(8))
def test_np_oracle_vs_live_reference(seed):
    from finrl_b200 import synthetic as syn

    mod = ref_loader.load("env_stocktrading_np")
    rng = np.random.default_rng(4000 + seed)
    D, K, T = int(rng.integers(1, 33)), int(rng.integers(1, 4)), int(rng.integers(5, 40))
    kw = dict(gamma=float(rng.choice([0.99, 0.9])), turbulence_thresh=float(rng.choice([30, 99])),
              min_stock_rate=float(rng.choice([0.0, 0.1, 0.3])), max_stock=float(rng.choice([1e2, 37.0, 5.0])),
              initial_capital=float(rng.choice([1e6, 3e4, 2e3])), buy_cost_pct=float(rng.choice([1e-3, 0.0])),
              sell_cost_pct=float(rng.choice([1e-3, 5e-3])), reward_scaling=float(rng.choice([2**-11, 1e-3])),
              initial_stocks=rng.integers(0, 5, D).astype(np.float32))
    close, tech, turb = syn.make_tables(T, D, K, seed=seed)
    pa, ta, tu = syn.make_np_arrays(close, tech, turb)
    if_train = bool(seed % 2)
    env = mod.StockTradingEnv({"price_array": pa, "tech_array": ta, "turbulence_array": tu, "if_train": if_train}, **kw)
    o = ora.NpTradingOracle(pa, ta, tu, 1, **kw)
    np.random.seed(seed)
    rs = np.random.RandomState(seed)

    def reset_both():
        obs = env.reset()
        s0 = f = None
        if if_train:
            s0 = (kw["initial_stocks"] + rs.randint(0, 64, size=D)).astype(np.float32)[None, :]
            f = np.array([rs.uniform(0.95, 1.05)])
        assert np.array_equal(o.reset(stocks0=s0, factor=f)[0], obs)

    reset_both()
    kinds = {float: 0, np.float32: 1, np.float64: 2}
    acts = syn.make_actions((2 * T + 1, D), seed=seed)
    if seed % 3 == 0:
        acts[: T // 2] *= 0.06
    for s in range(acts.shape[0]):
        state, reward, done, _ = env.step(acts[s])
        obs, orew, ork, ofl = o.step(acts[s][None, :])
        ctx = f"seed {seed} step {s}"
        assert bool(ofl[0] & 1) == bool(done) and (orew[0], ork[0]) == (reward, kinds[type(reward)]), ctx
        assert (o.amount[0], o.amount_kind[0]) == (env.amount, kinds[type(env.amount)]), ctx
        assert np.array_equal(o.stocks[0], env.stocks) and np.array_equal(o.cool[0], env.stocks_cool_down), ctx
        assert np.array_equal(obs[0], state), ctx
        if done:
            assert o.episode_return[0] == env.episode_return, ctx
            reset_both()
