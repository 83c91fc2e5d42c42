"""GPU: the reference's SB3 adapter loops (tests/adapter_loops.py restates DRLAgent.DRL_prediction and the
ensemble agent's trade-window loop, models.py:110-130 / :278-325) driven over the finrl_b200 drop-ins, against
goldens those same loops produced over the UNMODIFIED reference classes (tests/golden/make_golden.py adapter)."""
import contextlib
import io
import os

import numpy as np
import pytest

import adapter_loops
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


def _trading_kwargs(D, K, hmax, initial_amount, cost, threshold, risk_col):
    from finrl_b200 import synthetic as syn

    return dict(stock_dim=D, hmax=int(hmax), initial_amount=int(initial_amount), num_stock_shares=[0] * D, buy_cost_pct=cost,
                sell_cost_pct=cost, reward_scaling=1e-4, state_space=1 + 2 * D + K * D, action_space=D,
                tech_indicator_list=syn.INDICATORS[:K], turbulence_threshold=threshold, risk_indicator_col=risk_col,
                print_verbosity=10**9)


def test_drl_prediction_over_stocktrading_env():
    from finrl_b200 import synthetic as syn
    from finrl_b200.env_stocktrading import StockTradingEnv

    g = np.load(os.path.join(GOLDEN, "adapter_trading_prediction.npz"))
    hmax, init, cost, thr = g["cfg"]
    T, D = g["close"].shape
    K = g["tech"].shape[0]
    df = syn.make_frame(g["close"], g["tech"], g["risk"], risk_col="vix")
    env = StockTradingEnv(df=df, **_trading_kwargs(D, K, hmax, init, cost, thr, "vix"))
    model = adapter_loops.ReplayModel(g["actions"])
    with _quiet():
        df_account, df_actions = adapter_loops.drl_prediction(model, env)
    assert model.calls == T  # T-1 transitions + the terminal call, then "hit end"
    assert list(df_account.columns) == ["date", "account_value"]
    assert np.array_equal(np.array(df_account["date"].tolist(), dtype="U16"), g["account_date"])
    assert np.array_equal(df_account["account_value"].to_numpy(np.float64), g["account_value"])  # bit-exact fp64
    assert df_actions.index.name == "date"
    assert np.array_equal(np.array(df_actions.index.tolist(), dtype="U16"), g["action_date"])
    assert np.array_equal(np.array(list(df_actions.columns), dtype="U16"), g["action_cols"])
    assert np.array_equal(df_actions.to_numpy(np.int64), g["executed"])


def test_ensemble_two_window_hand_off(tmp_path, monkeypatch):
    """``StockTradingEnv(initial=False, previous_state=last_state)`` + ``last_state = trade_env.render()``,
    and the CSV files the terminal branch leaves for the ensemble agent."""
    import pandas as pd

    from finrl_b200 import synthetic as syn
    from finrl_b200.env_stocktrading import StockTradingEnv
    from finrl_b200.vec_env import dummy_vec_env

    g = np.load(os.path.join(GOLDEN, "adapter_ensemble_two_windows.npz"))
    hmax, init, cost, thr, W = g["cfg"]
    W = int(W)
    D, K = g["close"].shape[1], g["tech"].shape[0]
    kw = _trading_kwargs(D, K, hmax, init, cost, thr, "turbulence")
    (tmp_path / "results").mkdir()
    monkeypatch.chdir(tmp_path)
    last_state = []
    for w in range(2):
        sl = slice(w * W, (w + 1) * W)
        df = syn.make_frame(g["close"][sl], g["tech"][:, sl], g["risk"][sl])
        trace = {}
        with _quiet():
            last_state = adapter_loops.ensemble_trade_window(
                dummy_vec_env, StockTradingEnv, df, adapter_loops.ReplayModel(g["actions"][sl]), last_state, w == 0, kw,
                "ens", 100 + w, trace)
        ctx = f"window {w}"
        assert np.array_equal(trace["obs0"], g[f"w{w}_obs0"]), ctx
        assert np.array_equal(np.asarray(trace["obs"], dtype=np.float32), g[f"w{w}_obs"]), ctx
        assert np.array_equal(np.asarray(trace["rewards"], dtype=np.float32), g[f"w{w}_rewards"]), ctx
        assert np.array_equal(np.asarray(trace["dones"], dtype=np.uint8), g[f"w{w}_dones"]), ctx
        assert isinstance(last_state, list)
        assert np.array_equal(np.asarray(last_state, dtype=np.float64), g[f"w{w}_last_state"]), ctx  # fp64 state list
        av = pd.read_csv(f"results/account_value_trade_ens_{100 + w}.csv")
        rw = pd.read_csv(f"results/account_rewards_trade_ens_{100 + w}.csv")
        ac = pd.read_csv(f"results/actions_trade_ens_{100 + w}.csv")
        assert np.array_equal(av["account_value"].to_numpy(np.float64), g[f"w{w}_csv_account_value"]), ctx
        assert np.array_equal(av["daily_return"].to_numpy(np.float64), g[f"w{w}_csv_daily_return"], equal_nan=True), ctx
        assert np.array_equal(rw["account_rewards"].to_numpy(np.float64), g[f"w{w}_csv_rewards"]), ctx
        assert np.array_equal(ac.iloc[:, 1:].to_numpy(np.int64), g[f"w{w}_csv_actions"]), ctx
        assert np.array_equal(np.array(av["date"].tolist(), dtype="U16"), g[f"w{w}_csv_dates"]), ctx
    # window 2 really resumed: its first account value is window 1's portfolio re-priced, not initial_amount
    assert g["w1_csv_account_value"][0] != init


def test_drl_prediction_over_portfolio_env(tmp_path, monkeypatch):
    from finrl_b200 import synthetic as syn
    from finrl_b200.env_portfolio import StockPortfolioEnv

    g = np.load(os.path.join(GOLDEN, "adapter_portfolio_prediction.npz"))
    T, D = g["close"].shape
    K = g["tech"].shape[0]
    df = syn.make_frame(g["close"], g["tech"], np.zeros(T))
    df["cov_list"] = [g["cov"][t] for t in range(T) for _ in range(D)]
    env = StockPortfolioEnv(df=df, stock_dim=D, hmax=100, initial_amount=1_000_000, transaction_cost_pct=0.001,
                            reward_scaling=1e-4, state_space=D, action_space=D, tech_indicator_list=syn.INDICATORS[:K])
    (tmp_path / "results").mkdir()
    monkeypatch.chdir(tmp_path)
    with _quiet():
        df_ret, df_w = adapter_loops.drl_prediction(adapter_loops.ReplayModel(g["actions"]), env)
    assert list(df_ret.columns) == ["date", "daily_return"]
    assert np.array_equal(np.array(df_ret["date"].tolist(), dtype="U16"), g["ret_date"])
    # f32 actions: np.exp(float32) is reproduced to 1 ulp, not bit for bit -> 2e-6 on the weights
    np.testing.assert_allclose(df_ret["daily_return"].to_numpy(np.float64), g["daily_return"], rtol=2e-6, atol=1e-9)
    assert np.array_equal(np.array(df_w.index.tolist(), dtype="U16"), g["weight_date"])
    assert np.array_equal(np.array(list(df_w.columns), dtype="U16"), g["weight_cols"])
    np.testing.assert_allclose(df_w.to_numpy(np.float64), g["weights"], rtol=2e-6)


def test_drl_prediction_over_cashpenalty_env():
    """get_sb_env deep-copies the env (env_stocktrading_cashpenalty.py:374-380): the VecEnv's copy keeps the
    memories, the original object stays untouched."""
    from finrl_b200 import synthetic as syn
    from finrl_b200.env_stocktrading_cashpenalty import StockTradingEnvCashpenalty

    g = np.load(os.path.join(GOLDEN, "adapter_cashpen_prediction.npz"))
    hmax, thr = g["cfg"]
    close = g["close"]
    T, D = close.shape
    df = syn.make_frame(close, np.zeros((0, T, D)), g["turbulence"], tech_names=[],
                        extra_cols={c: g[c] for c in ("open", "high", "low", "volume")}).reset_index(drop=True)
    with _quiet():
        env = StockTradingEnvCashpenalty(df=df, hmax=hmax, turbulence_threshold=None if thr < 0 else thr,
                                         print_verbosity=10**9, random_start=False, cache_indicator_data=True)
        df_account, df_actions = adapter_loops.drl_prediction(adapter_loops.ReplayModel(g["actions"]), env)
    assert list(df_account.columns) == [str(c) for c in g["account_cols"]]
    np.testing.assert_allclose(df_account[["cash", "asset_value", "total_assets", "reward"]].to_numpy(np.float64),
                               g["account"], rtol=1e-9, atol=1e-9)
    assert np.array_equal(np.array(df_account["date"].tolist(), dtype="U16"), g["account_date"])
    assert np.array_equal(np.array(df_actions["date"].tolist(), dtype="U16"), g["action_date"])
    assert np.array_equal(np.stack(df_actions["actions"].tolist()), g["logged_actions"])
    np.testing.assert_allclose(np.stack(df_actions["transactions"].tolist()), g["transactions"], rtol=1e-9, atol=1e-9)
    assert env.current_step == 0 and env.account_information["cash"] == []  # the original was never stepped


def test_batched_vec_env_per_env_methods_and_attrs():
    """SB3 contract for N > 1: env_method / get_attr return ONE entry per env — every env's own memories
    (record_memory=True) and its own state values, equal to a 1-env gym object fed that env's actions."""
    from finrl_b200 import synthetic as syn
    from finrl_b200.env_stocktrading import StockTradingEnv

    N, T, D, K = 4, 14, 30, 8
    close, tech, turb = syn.make_tables(T, D, K, seed=81)
    df = syn.make_frame(close, tech, turb)
    kw = _trading_kwargs(D, K, 100, 120_000, 0.001, 90, "turbulence")
    proto = StockTradingEnv(df=df, **kw)
    vec = proto.get_vec_env(N, record_memory=True)
    acts = syn.make_actions((T + 6, N, D), seed=82)
    vec.reset()
    singles = [StockTradingEnv(df=df, **kw) for _ in range(N)]
    for e in singles:
        e.reset()
    for s in range(acts.shape[0]):
        obs, rews, dones, infos = vec.step(acts[s])
        with _quiet():
            for i, e in enumerate(singles):
                st, r, d, _ = e.step(acts[s, i])
                assert d == dones[i]
                if d:
                    e.reset()
        if s in (5, T - 2, T + 4):  # mid-episode, last transition, and after the auto-reset
            accounts = vec.env_method("save_asset_memory")
            actions = vec.env_method("save_action_memory")
            assert len(accounts) == N and len(actions) == N
            for i, e in enumerate(singles):
                ref_a, ref_x = e.save_asset_memory(), e.save_action_memory()
                assert np.array_equal(accounts[i]["account_value"].to_numpy(), ref_a["account_value"].to_numpy()), (s, i)
                assert accounts[i]["date"].tolist() == ref_a["date"].tolist()
                assert np.array_equal(actions[i].to_numpy(np.int64), ref_x.to_numpy(np.int64)), (s, i)
                assert list(actions[i].columns) == list(ref_x.columns) and actions[i].index.tolist() == ref_x.index.tolist()
            assert not np.array_equal(accounts[0]["account_value"].to_numpy()[1:], accounts[1]["account_value"].to_numpy()[1:])
    cash = vec.get_attr("cash")
    assert len(cash) == N and [float(c) for c in cash] == [e.state[0] for e in singles]
    hold = vec.get_attr("hold", indices=[1, 3])
    assert len(hold) == 2 and np.array_equal(hold[0], np.asarray(singles[1].state[1 + D : 1 + 2 * D]))
    assert vec.get_attr("hmax") == [100] * N and vec.env_is_wrapped(object) == [False] * N
    one = vec.env_method("save_asset_memory", indices=[2])
    assert len(one) == 1
    with pytest.raises(AttributeError):
        proto.get_vec_env(2).env_method("save_asset_memory")
