"""GPU: the reference-facing gym classes (same names / constructor signatures as the reference) and
the batched SB3-style VecEnv, driven exactly like the reference's own loops and checked against the
goldens the unmodified reference produced."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from test_oracle_golden import NpResetDraws, cashpen_args_from_golden, np_kwargs_from_golden

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def test_stocktrading_gym_class_matches_reference():
    from finrl_b200 import synthetic as syn
    from finrl_b200.env_stocktrading import StockTradingEnv

    g = np.load(os.path.join(GOLDEN, "trading_d30_starved.npz"))
    hmax, init, bc, sc, rs, use_t, thr = g["cfg"]
    T, D = g["close"].shape
    K = g["tech"].shape[0]
    df = syn.make_frame(g["close"], g["tech"], g["risk"])
    env = StockTradingEnv(
        df=df, stock_dim=D, hmax=int(hmax), initial_amount=int(init), num_stock_shares=[0] * D, buy_cost_pct=bc,
        sell_cost_pct=sc, reward_scaling=rs, state_space=1 + 2 * D + K * D, action_space=D,
        tech_indicator_list=syn.INDICATORS[:K], turbulence_threshold=thr if use_t > 0 else None, print_verbosity=10**9,
    )
    assert env.action_space.shape == (D,) and env.observation_space.shape == (1 + 2 * D + K * D,)
    assert np.array_equal(np.asarray(env.state, dtype=np.float64), g["obs0"])  # the fp64 state LIST, not a f32 image
    acts = g["actions"]
    for s in range(acts.shape[0]):
        state, reward, done, info = env.step(acts[s].copy())
        ctx = f"step {s}"
        assert isinstance(state, list) and info == {} and done == bool(g["done"][s]), ctx
        assert reward == g["reward"][s], ctx
        if done:
            state = env.reset()
        else:
            assert env.trades == g["trades"][s] and env.cost == g["cost"][s], ctx
            assert len(env.asset_memory) == len(env.actions_memory) + 1 == len(env.date_memory)
        assert state[0] == g["cash"][s], ctx
        assert np.array_equal(np.asarray(state[1 + D : 1 + 2 * D]), g["hold"][s]), ctx
        assert np.array_equal(np.asarray(state, dtype=np.float64).astype(np.float32), g["obs"][s]), ctx
        assert env.day == g["day"][s], ctx
        assert env.asset_memory[0] == g["begin_asset"][s], ctx
    df_assets, df_actions = env.save_asset_memory(), env.save_action_memory()
    assert list(df_assets.columns) == ["date", "account_value"] and df_actions.shape[1] == D


def test_np_gym_class_matches_reference():
    from finrl_b200.env_stocktrading_np import StockTradingEnv

    g = np.load(os.path.join(GOLDEN, "np_d30_train.npz"))
    kw = np_kwargs_from_golden(g)
    cfg = {"price_array": g["price_array"], "tech_array": g["tech_array"], "turbulence_array": g["turbulence_array"],
           "if_train": True}
    np.random.seed(int(g["rng_seed"]))  # the reference draws its reset randomness from numpy's global RNG
    env = StockTradingEnv(cfg, **kw)  # the engine constructor resets once, like ElegantRL's first env.reset() would
    np.random.seed(int(g["rng_seed"]))
    state = env.reset()
    assert state.dtype == np.float32 and np.array_equal(state, g["obs0"])
    assert type(env.amount) is np.float32 and env.amount == g["init_amount"]
    assert (env.state_dim, env.action_dim, env.max_step, env.if_discrete) == (g["obs0"].shape[0], 30, g["price_array"].shape[0] - 1, False)
    kinds = {0: float, 1: np.float32, 2: np.float64}
    acts = g["actions"]
    for s in range(acts.shape[0]):
        state, reward, done, info = env.step(acts[s])
        ctx = f"step {s}"
        assert done == bool(g["done"][s]) and info == {}, ctx
        assert reward == g["reward"][s] and type(reward) is kinds[int(g["reward_kind"][s])], ctx
        assert env.amount == g["amount"][s] and type(env.amount) is kinds[int(g["amount_kind"][s])], ctx
        assert np.array_equal(env.stocks, g["stocks"][s]) and np.array_equal(state, g["obs"][s]), ctx
        assert env.day == g["day"][s], ctx
        if done:
            assert env.episode_return == g["episode_return"][s], ctx
            env.reset()


def test_portfolio_gym_class_matches_reference():
    from finrl_b200 import synthetic as syn
    from finrl_b200.env_portfolio import StockPortfolioEnv

    g = np.load(os.path.join(GOLDEN, "portfolio_d30_f64.npz"))
    T, D = g["close"].shape
    K = g["tech"].shape[0]
    df = syn.make_frame(g["close"], g["tech"], np.zeros(T))
    df["cov_list"] = [g["cov"][t] for t in range(T) for _ in range(D)]
    env = StockPortfolioEnv(df=df, stock_dim=D, hmax=100, initial_amount=1_000_000, transaction_cost_pct=0.001,
                            reward_scaling=1e-4, state_space=D, action_space=D, tech_indicator_list=syn.INDICATORS[:K])
    assert env.observation_space.shape == (D + K, D)
    assert np.array_equal(env.reset(), g["obs0"])
    import contextlib, io

    acts = g["actions"]
    with contextlib.redirect_stdout(io.StringIO()):
        for s in range(acts.shape[0]):
            state, reward, done, info = env.step(acts[s])
            ctx = f"step {s}"
            assert done == bool(g["done"][s]), ctx
            assert abs(reward - g["reward"][s]) <= 1e-9 * abs(g["reward"][s]), ctx
            if done:
                state = env.reset()
            assert np.array_equal(state, g["obs"][s]) and env.day == g["day"][s], ctx
            assert abs(env.portfolio_value - g["pv"][s]) <= 1e-9 * g["pv"][s], ctx
            if not g["done"][s]:
                np.testing.assert_allclose(env.actions_memory[-1], g["weights"][s], rtol=1e-9)
                assert abs(env.portfolio_return_memory[-1] - g["pret"][s]) <= 1e-9 * max(abs(g["pret"][s]), 1e-3), ctx


def test_cashpenalty_gym_class_matches_reference():
    from finrl_b200 import synthetic as syn
    from finrl_b200.env_stocktrading_cashpenalty import StockTradingEnvCashpenalty

    g = np.load(os.path.join(GOLDEN, "cashpen_d10_turb_patient.npz"))
    close, info, turb, kw = cashpen_args_from_golden(g)
    T, D = close.shape
    df = syn.make_frame(close, np.zeros((0, T, D)), turb, tech_names=[],
                        extra_cols={c: g[c] for c in ("open", "high", "low", "volume")}).reset_index(drop=True)
    env = StockTradingEnvCashpenalty(df=df, random_start=False, print_verbosity=10**9,
                                     daily_information_cols=[str(c) for c in g["cols"]], **kw)
    np.testing.assert_allclose(env.reset(), g["obs0"], rtol=1e-7)
    acts = g["actions"]
    for s in range(acts.shape[0]):
        state, reward, done, info_ = env.step(acts[s])
        ctx = f"step {s}"
        assert done == bool(g["done"][s]), ctx
        np.testing.assert_allclose(reward, g["reward"][s], rtol=1e-9, atol=1e-15, err_msg=ctx)
        if done:
            state = env.reset()
        np.testing.assert_allclose(state[: 1 + D], g["obs"][s][: 1 + D], rtol=1e-9, atol=1e-7, err_msg=ctx)
        np.testing.assert_allclose(state[1 + D :], g["obs"][s][1 + D :], rtol=2e-7, err_msg=ctx)  # info comes from the f32 table
        assert env.date_index == g["date_index"][s], ctx


def test_batched_vec_env_follows_dummy_vec_env_protocol():
    """N envs behind the SB3 VecEnv protocol: float32 obs, auto-reset on done with the terminal
    observation in infos, numpy in / numpy out — env i must equal a reference-style loop of env i."""
    from finrl_b200 import BatchedStockTradingEnv, TradingTables, synthetic as syn
    from finrl_b200.vec_env import BatchedVecEnv
    from oracle import oracle as ora

    N, T, D, K = 6, 25, 30, 8
    close, tech, turb = syn.make_tables(T, D, K, seed=1)
    kw = dict(hmax=100, initial_amount=150_000, buy_cost_pct=0.001, sell_cost_pct=0.001, reward_scaling=1e-4,
              turbulence_threshold=90)
    vec = BatchedVecEnv(BatchedStockTradingEnv(tables=TradingTables.from_arrays(close, tech, turb, "cuda"), n_envs=N, **kw))
    o = ora.TradingOracle(close, tech, turb, N, **kw)
    obs = vec.reset()
    assert obs.dtype == np.float32 and obs.shape == (N, 301) and vec.num_envs == N
    assert np.array_equal(obs, o.reset())
    acts = syn.make_actions((2 * T + 3, N, D), seed=2)
    for s in range(acts.shape[0]):
        prev = o.obs()
        obs, rews, dones, infos = vec.step(acts[s])
        oobs, orew, ofl = o.step(acts[s], auto_reset=True)
        assert np.array_equal(obs, oobs) and np.array_equal(dones, (ofl & 1).astype(bool))
        assert np.array_equal(rews, orew.astype(np.float32)) and len(infos) == N
        for i in range(N):
            if dones[i]:
                assert np.array_equal(infos[i]["terminal_observation"], prev[i])
            else:
                assert infos[i] == {}
    # tensor mode: device tensors, kernel-side auto-reset
    tvec = BatchedVecEnv(BatchedStockTradingEnv(tables=vec.engine.tables, n_envs=N, **kw), tensor_mode=True)
    t_obs = tvec.reset()
    assert t_obs.is_cuda
    o2 = ora.TradingOracle(close, tech, turb, N, **kw)
    for s in range(T + 2):
        t_obs, t_rew, t_done, _ = tvec.step(torch.from_numpy(acts[s]).cuda())
        oobs, orew, ofl = o2.step(acts[s], auto_reset=True)
        assert np.array_equal(t_obs.cpu().numpy(), oobs) and np.array_equal(t_rew.cpu().numpy(), orew)


def test_np_vectorized_elegantrl_convention():
    from finrl_b200 import synthetic as syn
    from finrl_b200.env_stocktrading_np import StockTradingEnv

    close, tech, turb = syn.make_tables(30, 30, 8, seed=5)
    pa, ta, tu = syn.make_np_arrays(close, tech, turb)
    env = StockTradingEnv({"price_array": pa, "tech_array": ta, "turbulence_array": tu, "if_train": False})
    vec = env.vectorized(64)
    assert (vec.env_num, vec.state_dim, vec.action_dim, vec.max_step, vec.if_discrete) == (64, 333, 30, 29, False)
    s = vec.reset()
    assert s.is_cuda and tuple(s.shape) == (64, 333)
    for _ in range(31):
        s, r, d, _ = vec.step(torch.rand((64, 30), device="cuda") * 2 - 1)
    assert tuple(s.shape) == (64, 333) and tuple(r.shape) == (64,) and d.dtype == torch.bool


def test_nas100_sibling_class_matches_reference():
    from finrl_b200.env_nas100_wrds import StockEnvNAS100

    g = np.load(os.path.join(GOLDEN, "np_nas100_d20.npz"))
    kw = np_kwargs_from_golden(g)
    kw.pop("obs_amount_floor")
    np.random.seed(int(g["rng_seed"]))
    env = StockEnvNAS100(cwd=None, price_ary=g["price_array"], tech_ary=g["tech_array"], turbulence_ary=g["turbulence_array"],
                         data_gap=1, if_eval=True, **kw)
    np.random.seed(int(g["rng_seed"]))
    assert np.array_equal(env.reset(), g["obs0"]) and env.env_name == "StockEnvNAS"
    floor_seen = 0
    for s in range(g["actions"].shape[0]):
        state, reward, done, _ = env.step(g["actions"][s])
        assert done == bool(g["done"][s]) and reward == g["reward"][s], f"step {s}"
        assert np.array_equal(state, g["obs"][s]) and np.array_equal(env.stocks_cd, g["cool"][s]), f"step {s}"
        floor_seen += int(state[0] == np.float32(1e4 * 2**-12) and env.amount < 1e4)
        if done:
            env.reset()
    assert floor_seen > 0  # the max(amount, 1e4) branch was exercised
    with pytest.raises(TypeError):
        StockEnvNAS100(cwd=None, price_ary=g["price_array"].astype(np.float64), tech_ary=g["tech_array"],
                       turbulence_ary=g["turbulence_array"])


def test_stoploss_gym_class_and_vec_env_match_reference():
    """The StockTradingEnvStopLoss drop-in on a frame vs the golden of the unmodified reference, and its
    get_multiproc_env VecEnv (n GPU envs instead of n forked processes) vs the oracle."""
    from finrl_b200 import synthetic as syn
    from finrl_b200.env_stocktrading_stoploss import StockTradingEnvStopLoss
    from oracle import oracle as ora
    from test_oracle_golden import stoploss_args_from_golden

    g = np.load(os.path.join(GOLDEN, "stoploss_d10_turb_patient.npz"))
    close, info, turb, kw = stoploss_args_from_golden(g)
    T, D = close.shape
    df = syn.make_frame(close, np.zeros((0, T, D)), turb, tech_names=[],
                        extra_cols={c: g[c] for c in ("open", "high", "low", "volume")}).reset_index(drop=True)
    env = StockTradingEnvStopLoss(df=df, random_start=False, print_verbosity=10**9,
                                  daily_information_cols=[str(c) for c in g["cols"]], **kw)
    np.testing.assert_allclose(env.reset(), g["obs0"], rtol=1e-7)
    acts = g["actions"]
    for s in range(acts.shape[0]):
        state, reward, done, info_ = env.step(acts[s])
        ctx = f"step {s}"
        assert done == bool(g["done"][s]), ctx
        np.testing.assert_allclose(reward, g["reward"][s], rtol=1e-9, atol=1e-15, err_msg=ctx)
        if done:
            state = env.reset()
        np.testing.assert_allclose(state[: 1 + D], g["obs"][s][: 1 + D], rtol=1e-9, atol=1e-7, err_msg=ctx)
        assert env.date_index == g["date_index"][s], ctx
    # VecEnv: 5 envs, numpy in/out, auto-reset
    vec, obs = env.get_multiproc_env(n=5)
    o = ora.StopLossOracle(close, info, turb, 5, **kw)
    np.testing.assert_allclose(obs, o.obs().astype(np.float32), rtol=2e-7, atol=1e-6)
    a = syn.make_actions((T + 5, 5, D), seed=77)
    for s in range(a.shape[0]):
        obs, rews, dones, infos = vec.step(a[s])
        orew, ofl = o.step(a[s], auto_reset=True)
        assert np.array_equal(dones, (ofl & 1).astype(bool)), s
        np.testing.assert_allclose(rews, orew.astype(np.float32), rtol=1e-6, atol=1e-12)
        np.testing.assert_allclose(obs, o.obs().astype(np.float32), rtol=2e-7, atol=1e-6)
