"""Generate golden vectors by EXECUTING THE UNMODIFIED REFERENCE (superyuri/FinRL).

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference env files are loaded as they lie (oracle/ref_loader.py installs inert stubs for
gym / matplotlib / stable_baselines3, which are not installed and do not touch arithmetic), driven
with seeded synthetic tables and actions, and every step's outputs are written to
``tests/golden/*.npz``.  The fixtures travel to the GPU box; /root/reference does not.

numpy {np} / pandas {pd} versions are recorded inside each fixture (argsort tie order and NEP-50
promotion are numpy-version sensitive).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from finrl_b200 import synthetic as syn  # noqa: E402
from oracle import ref_loader  # noqa: E402

KIND = {float: 0, np.float32: 1, np.float64: 2, int: 0}


def _meta():
    return {"numpy_version": np.__version__, "pandas_version": pd.__version__}


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


# ------------------------------------------------------------------------------------------
# A1  StockTradingEnv  (env_stocktrading.py)
# ------------------------------------------------------------------------------------------
def gen_trading(name, T, D, K, n_steps, seed, threshold, act_dtype, hmax=100, initial_amount=1_000_000,
                num_stock_shares=None, plant_disable=(), cost=0.001, reward_scaling=1e-4, act_scale=1.0):
    mod = ref_loader.load("env_stocktrading")
    close, tech, turb = syn.make_tables(T, D, K, seed=seed)
    for (t, i) in plant_disable:
        tech[0, t, i] = 1.0
    df = syn.make_frame(close, tech, turb)
    shares = list(num_stock_shares) if num_stock_shares is not None else [0] * D
    env = mod.StockTradingEnv(
        df=df, stock_dim=D, hmax=hmax, initial_amount=initial_amount, num_stock_shares=list(shares),
        buy_cost_pct=cost, sell_cost_pct=cost, reward_scaling=reward_scaling, state_space=1 + 2 * D + K * D,
        action_space=D, tech_indicator_list=syn.INDICATORS[:K], turbulence_threshold=threshold,
        print_verbosity=10**9,
    )
    actions = (syn.make_actions((n_steps, D), seed=seed + 1, dtype=np.float64) * act_scale).astype(act_dtype)
    O = 1 + 2 * D + K * D
    out = {
        "obs0": np.asarray(env.state, dtype=np.float64),
        "cash": np.zeros(n_steps), "hold": np.zeros((n_steps, D), dtype=np.int64), "reward": np.zeros(n_steps),
        "done": np.zeros(n_steps, dtype=np.uint8), "liq": np.zeros(n_steps, dtype=np.uint8),
        "trades": np.zeros(n_steps, dtype=np.int64), "cost": np.zeros(n_steps), "day": np.zeros(n_steps, dtype=np.int64),
        "obs": np.zeros((n_steps, O), dtype=np.float32), "term_obs": np.zeros((n_steps, O), dtype=np.float32),
        "begin_asset": np.zeros(n_steps),
    }
    with _quiet():
        for s in range(n_steps):
            liq = threshold is not None and env.turbulence >= threshold and env.day < T - 1
            state, reward, done, _ = env.step(actions[s].copy())
            out["liq"][s] = liq
            out["reward"][s] = reward
            out["done"][s] = done
            # values BEFORE the auto-reset (the env's own post-step state)
            out["trades"][s] = env.trades
            out["cost"][s] = env.cost
            out["term_obs"][s] = np.asarray(state, dtype=np.float64).astype(np.float32)
            if done:  # what DummyVecEnv.step_wait does
                state = env.reset()
            out["cash"][s] = state[0]
            out["hold"][s] = np.asarray(state[1 + D : 1 + 2 * D], dtype=np.float64).astype(np.int64)
            out["day"][s] = env.day
            out["obs"][s] = np.asarray(state, dtype=np.float64).astype(np.float32)
            out["begin_asset"][s] = env.asset_memory[0]
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), close=close, tech=tech, risk=turb, actions=actions,
        cfg=np.array([hmax, initial_amount, cost, cost, reward_scaling,
                      -1.0 if threshold is None else 1.0, 0.0 if threshold is None else threshold]),
        num_stock_shares=np.asarray(shares, dtype=np.int64), **out, **_meta(),
    )
    print(f"{name}: {n_steps} steps, dones={int(out['done'].sum())}, liq={int(out['liq'].sum())}, "
          f"min cash={out['cash'].min():.3f}, trades[-1]={out['trades'][-1]}")


# ------------------------------------------------------------------------------------------
# A2  numpy / ElegantRL StockTradingEnv  (env_stocktrading_np.py)
# ------------------------------------------------------------------------------------------
def _kind(x):
    return {float: 0, int: 0, np.float32: 1, np.float64: 2}[type(x)]


def gen_np(name, T, D, K, n_steps, seed, if_train, thresh=99, quiet_steps=0, nas100=False, **kw):
    close, tech, turb = syn.make_tables(T, D, K, seed=seed)
    price_array, tech_array, turb_array = syn.make_np_arrays(close, tech, turb)
    if nas100:
        # StockEnvNAS100 (env_nas100_wrds.py): the same step, get_state shows max(amount, 1e4); its arrays come
        # from load_data() as float32 and the ctor slices [0:211210:data_gap] in eval mode
        mod = ref_loader.load("env_nas100_wrds")
        price_array, tech_array = price_array.astype(np.float32), tech_array.astype(np.float32)
        env = mod.StockEnvNAS100(cwd=None, price_ary=price_array, tech_ary=tech_array, turbulence_ary=turb_array,
                                 turbulence_thresh=thresh, data_gap=1, if_eval=True, **kw)
        env.stocks_cool_down = None
    else:
        mod = ref_loader.load("env_stocktrading_np")
        cfg = {"price_array": price_array, "tech_array": tech_array, "turbulence_array": turb_array, "if_train": if_train}
        env = mod.StockTradingEnv(cfg, turbulence_thresh=thresh, **kw)
    np.random.seed(seed + 100)
    obs0 = env.reset()
    init = {"init_amount": np.float64(env.amount), "init_amount_kind": _kind(env.amount),
            "init_stocks": env.stocks.copy(), "init_total": np.float64(env.total_asset),
            "init_total_kind": _kind(env.total_asset)}
    actions = syn.make_actions((n_steps, D), seed=seed + 1)
    if quiet_steps:  # dead-band / capped-only prefix: amount stays a Python float or np.float32 for a while
        actions[:quiet_steps] *= 0.09
        actions[quiet_steps : 2 * quiet_steps] = -np.abs(actions[quiet_steps : 2 * quiet_steps])
    O = env.state_dim
    out = {k: np.zeros(n_steps) for k in ("amount", "total", "gamma_reward", "reward", "episode_return")}
    out.update({k: np.zeros(n_steps, dtype=np.uint8) for k in ("amount_kind", "total_kind", "gr_kind", "reward_kind", "done", "liq")})
    out["stocks"] = np.zeros((n_steps, D), dtype=np.float32)
    out["cool"] = np.zeros((n_steps, D), dtype=np.float32)
    out["obs"] = np.zeros((n_steps, O), dtype=np.float32)
    out["day"] = np.zeros(n_steps, dtype=np.int64)
    resets = []
    for s in range(n_steps):
        state, reward, done, _ = env.step(actions[s])
        out["liq"][s] = env.turbulence_bool[env.day] != 0
        out["amount"][s] = env.amount
        out["amount_kind"][s] = _kind(env.amount)
        out["total"][s] = env.total_asset
        out["total_kind"][s] = _kind(env.total_asset)
        out["gamma_reward"][s] = env.gamma_reward
        out["gr_kind"][s] = _kind(env.gamma_reward)
        out["reward"][s] = reward
        out["reward_kind"][s] = _kind(reward)
        out["done"][s] = done
        out["episode_return"][s] = env.episode_return
        out["stocks"][s] = env.stocks
        out["cool"][s] = env.stocks_cd if nas100 else env.stocks_cool_down
        out["obs"][s] = state
        out["day"][s] = env.day
        assert state.dtype == np.float32
        if done:
            env.reset()
            resets.append((np.float64(env.amount), _kind(env.amount), env.stocks.copy()))
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), price_array=price_array, tech_array=tech_array, turbulence_array=turb_array,
        actions=actions, obs0=obs0, if_train=np.array(int(if_train)), thresh=np.array(float(thresh)),
        rng_seed=np.array(seed + 100), nas100=np.array(int(nas100)),
        reset_amount=np.array([r[0] for r in resets]), reset_amount_kind=np.array([r[1] for r in resets], dtype=np.uint8),
        reset_stocks=np.array([r[2] for r in resets], dtype=np.float32).reshape(len(resets), D),
        kw_keys=np.array([k for k in kw if k != "initial_stocks"], dtype="U32"),
        kw_vals=np.array([float(v) for k, v in kw.items() if k != "initial_stocks"]),
        initial_stocks=np.asarray(kw.get("initial_stocks", np.zeros(D)), dtype=np.float32),
        **init, **out, **_meta(),
    )
    print(f"{name}: {n_steps} steps, dones={int(out['done'].sum())}, liq={int(out['liq'].sum())}, "
          f"min stocks={out['stocks'].min()}, kinds amount={sorted(set(out['amount_kind']))} reward={sorted(set(out['reward_kind']))}")


# ------------------------------------------------------------------------------------------
# A3  StockPortfolioEnv  (env_portfolio.py)
# ------------------------------------------------------------------------------------------
def gen_portfolio(name, T, D, K, n_steps, seed, act_dtype, lookback=252):
    mod = ref_loader.load("env_portfolio")
    close, tech, turb = syn.make_tables(T, D, K, seed=seed)
    cov, first = syn.make_cov_table(close, lookback)
    Te = T - first
    close_e, tech_e, turb_e = close[first:], tech[:, first:], turb[first:]
    df = syn.make_frame(close_e, tech_e, turb_e)
    df["cov_list"] = [cov[t] for t in range(Te) for _ in range(D)]
    env = mod.StockPortfolioEnv(
        df=df, stock_dim=D, hmax=100, initial_amount=1_000_000, transaction_cost_pct=0.001, reward_scaling=1e-4,
        state_space=D, action_space=D, tech_indicator_list=syn.INDICATORS[:K],
    )
    obs0 = np.asarray(env.reset(), dtype=np.float64)
    actions = syn.make_actions((n_steps, D), seed=seed + 1, low=0.0, high=1.0, dtype=np.float64).astype(act_dtype)
    out = {"pv": np.zeros(n_steps), "reward": np.zeros(n_steps), "done": np.zeros(n_steps, dtype=np.uint8),
           "day": np.zeros(n_steps, dtype=np.int64), "weights": np.zeros((n_steps, D)),
           "obs": np.zeros((n_steps, D + K, D)), "pret": np.zeros(n_steps)}
    with _quiet():
        for s in range(n_steps):
            state, reward, done, _ = env.step(actions[s])
            out["reward"][s] = reward
            out["done"][s] = done
            out["pret"][s] = env.portfolio_return_memory[-1]
            out["weights"][s] = np.asarray(env.actions_memory[-1], dtype=np.float64)
            if done:
                state = env.reset()
            out["pv"][s] = env.portfolio_value
            out["day"][s] = env.day
            out["obs"][s] = state
    np.savez_compressed(os.path.join(HERE, name + ".npz"), close=close_e, tech=tech_e, cov=cov, actions=actions,
                        obs0=obs0, initial_amount=np.array(1_000_000.0), **out, **_meta())
    print(f"{name}: {n_steps} steps, dones={int(out['done'].sum())}, pv[-1]={out['pv'][-1]:.6f}")


# ------------------------------------------------------------------------------------------
# A4  StockTradingEnvCashpenalty  (env_stocktrading_cashpenalty.py)
# ------------------------------------------------------------------------------------------
def gen_cashpenalty(name, T, D, n_steps, seed, act_dtype, threshold=None, patient=False, discrete=False, hmax=10,
                    initial_amount=1e6, shares_increment=1, cols=("open", "close", "high", "low", "volume"),
                    cost=3e-3, penalty=0.1, stoploss=None):
    """``stoploss=(stoploss_penalty, profit_loss_ratio)`` generates from the sibling StockTradingEnvStopLoss."""
    mod = ref_loader.load("env_stocktrading_stoploss" if stoploss else "env_stocktrading_cashpenalty")
    close, _tech, turb = syn.make_tables(T, D, 0, seed=seed)
    o, h, l, v = syn.make_ohlv(close, seed)
    df = syn.make_frame(close, np.zeros((0, T, D)), turb, tech_names=[], extra_cols={"open": o, "high": h, "low": l, "volume": v})
    df = df.reset_index(drop=True)
    with _quiet():
        kw = dict(df=df, buy_cost_pct=cost, sell_cost_pct=cost, hmax=hmax, discrete_actions=discrete,
                  shares_increment=shares_increment, turbulence_threshold=threshold, print_verbosity=10**9,
                  initial_amount=initial_amount, daily_information_cols=list(cols), cache_indicator_data=True,
                  cash_penalty_proportion=penalty, random_start=False, patient=patient)
        if stoploss:
            env = mod.StockTradingEnvStopLoss(stoploss_penalty=stoploss[0], profit_loss_ratio=stoploss[1], **kw)
        else:
            env = mod.StockTradingEnvCashpenalty(**kw)
        obs0 = np.asarray(env.reset(), dtype=np.float64)
    actions = syn.make_actions((n_steps, D), seed=seed + 1, dtype=np.float64).astype(act_dtype)
    O = env.state_space
    out = {"obs": np.zeros((n_steps, O)), "reward": np.zeros(n_steps), "done": np.zeros(n_steps, dtype=np.uint8),
           "date_index": np.zeros(n_steps, dtype=np.int64), "liq": np.zeros(n_steps, dtype=np.uint8),
           "term_obs": np.zeros((n_steps, O))}
    with _quiet():
        for s in range(n_steps):
            liq = threshold is not None and env.turbulence >= threshold and env.date_index < T - 1
            state, reward, done, _ = env.step(actions[s])
            out["liq"][s] = liq
            out["reward"][s] = reward
            out["done"][s] = done
            out["term_obs"][s] = np.asarray(state, dtype=np.float64)
            if done:
                state = env.reset()
            out["obs"][s] = np.asarray(state, dtype=np.float64)
            out["date_index"][s] = env.date_index
    tables = {"close": close, "open": o, "high": h, "low": l, "volume": v, "turbulence": turb}
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), actions=actions, obs0=obs0, cols=np.array(list(cols), dtype="U16"),
        cfg=np.array([cost, cost, hmax if np.isscalar(hmax) else 0.0, float(discrete), shares_increment,
                      -1.0 if threshold is None else 1.0, 0.0 if threshold is None else threshold, initial_amount, penalty,
                      float(patient)]),
        hmax_vec=np.zeros(0) if np.isscalar(hmax) else np.asarray(hmax),  # per-asset hmax array (dtype kept)
        stoploss=np.array(stoploss if stoploss else [0.0, 0.0]),
        **tables, **out, **_meta(),
    )
    print(f"{name}: {n_steps} steps, dones={int(out['done'].sum())}, liq={int(out['liq'].sum())}, "
          f"min cash={out['obs'][:, 0].min():.3f}")


# ------------------------------------------------------------------------------------------
# sibling  CryptoEnv  (env_cryptocurrency_trading/env_multiple_crypto.py)
# ------------------------------------------------------------------------------------------
def gen_crypto(name, T, D, K, n_steps, seed, act_dtype, lookback=1, initial_capital=1e6, scales=None):
    mod = ref_loader.load("env_multiple_crypto")
    close, tech, turb = syn.make_tables(T, D, K, seed=seed)
    if scales is not None:  # coins span orders of magnitude, which is what the action normaliser is about
        close = close * np.asarray(scales)[None, :] / 100.0
    price_array, tech_array, _ = syn.make_np_arrays(close, tech, turb)
    env = mod.CryptoEnv({"price_array": price_array, "tech_array": tech_array}, lookback=lookback,
                        initial_capital=initial_capital)
    obs0 = env.reset()
    actions = syn.make_actions((n_steps, D), seed=seed + 1, dtype=np.float64).astype(act_dtype)
    O = obs0.shape[0]
    out = {k: np.zeros(n_steps) for k in ("cash", "total", "gamma_return", "reward", "episode_return")}
    out["done"] = np.zeros(n_steps, dtype=np.uint8)
    out["stocks"] = np.zeros((n_steps, D), dtype=np.float32)
    out["obs"] = np.zeros((n_steps, O), dtype=np.float32)
    out["time"] = np.zeros(n_steps, dtype=np.int64)
    for s in range(n_steps):
        state, reward, done, info = env.step(actions[s].copy())
        assert info is None and state.dtype == np.float32
        out["cash"][s], out["total"][s], out["gamma_return"][s] = env.cash, env.total_asset, env.gamma_return
        out["reward"][s], out["done"][s], out["episode_return"][s] = reward, done, env.episode_return
        out["stocks"][s], out["obs"][s], out["time"][s] = env.stocks, state, env.time
        if done:
            env.reset()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), price_array=price_array, tech_array=tech_array, actions=actions,
                        obs0=obs0, lookback=np.array(lookback), initial_capital=np.array(float(initial_capital)),
                        action_norm_vector=np.asarray(env.action_norm_vector), **out, **_meta())
    print(f"{name}: {n_steps} steps, dones={int(out['done'].sum())}, min cash={out['cash'].min():.4f}, "
          f"max stocks={out['stocks'].max():.3f}")


# ------------------------------------------------------------------------------------------
# adapter flows  (finrl/agents/stablebaselines3/models.py: DRL_prediction :110-130, ensemble :278-325)
# ------------------------------------------------------------------------------------------
sys.path.insert(0, os.path.join(ROOT, "tests"))
import adapter_loops  # noqa: E402


def _trading_env_kwargs(D, K, hmax, initial_amount, cost, threshold, risk_col):
    return dict(stock_dim=D, hmax=hmax, initial_amount=initial_amount, num_stock_shares=[0] * D, buy_cost_pct=cost,
                sell_cost_pct=cost, reward_scaling=1e-4, state_space=1 + 2 * D + K * D, action_space=D,
                tech_indicator_list=syn.INDICATORS[:K], turbulence_threshold=threshold, risk_indicator_col=risk_col,
                print_verbosity=10**9)


def gen_adapter_trading_prediction(name, T, D, K, seed, threshold=70, hmax=100, initial_amount=200_000, cost=0.001):
    """DRLAgent.DRL_prediction over the reference StockTradingEnv with a replayed action table."""
    mod = ref_loader.load("env_stocktrading")
    close, tech, turb = syn.make_tables(T, D, K, seed=seed)
    df = syn.make_frame(close, tech, turb, risk_col="vix")
    env = mod.StockTradingEnv(df=df, **_trading_env_kwargs(D, K, hmax, initial_amount, cost, threshold, "vix"))
    actions = syn.make_actions((T, D), seed=seed + 1)
    with _quiet():
        df_account, df_actions = adapter_loops.drl_prediction(adapter_loops.ReplayModel(actions), env)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), close=close, tech=tech, risk=turb, actions=actions,
        cfg=np.array([hmax, initial_amount, cost, threshold]),
        account_date=np.array(df_account["date"].tolist(), dtype="U16"), account_value=df_account["account_value"].to_numpy(np.float64),
        action_date=np.array(df_actions.index.tolist(), dtype="U16"), action_cols=np.array(list(df_actions.columns), dtype="U16"),
        executed=df_actions.to_numpy(np.int64), **_meta())
    print(f"{name}: account rows={len(df_account)}, action rows={df_actions.shape}, final={df_account['account_value'].iloc[-1]:.4f}")


def gen_adapter_ensemble(name, W, D, K, seed, threshold=80, hmax=100, initial_amount=150_000, cost=0.001):
    """Two consecutive trade windows of the ensemble agent: window 2 resumes from window 1's rendered state
    (initial=False, previous_state=last_state).  The terminal branch writes its three CSVs into results/."""
    import tempfile

    mod = ref_loader.load("env_stocktrading")
    close, tech, turb = syn.make_tables(2 * W, D, K, seed=seed)
    kw = _trading_env_kwargs(D, K, hmax, initial_amount, cost, threshold, "turbulence")
    actions = syn.make_actions((2 * W, D), seed=seed + 1)
    out = {}
    last_state, cwd = [], os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "results"))
        os.chdir(tmp)
        try:
            for w in range(2):
                sl = slice(w * W, (w + 1) * W)
                df = syn.make_frame(close[sl], tech[:, sl], turb[sl])  # data_split re-indexes each window from 0
                trace = {}
                with _quiet():
                    last_state = adapter_loops.ensemble_trade_window(
                        sys.modules["stable_baselines3.common.vec_env"].DummyVecEnv, mod.StockTradingEnv, df,
                        adapter_loops.ReplayModel(actions[sl]), last_state, w == 0, kw, "ens", 100 + w, trace)
                out[f"w{w}_obs0"] = trace["obs0"]
                out[f"w{w}_obs"] = np.asarray(trace["obs"], dtype=np.float32)
                out[f"w{w}_rewards"] = np.asarray(trace["rewards"], dtype=np.float32)
                out[f"w{w}_dones"] = np.asarray(trace["dones"], dtype=np.uint8)
                out[f"w{w}_last_state"] = np.asarray(last_state, dtype=np.float64)
                av = pd.read_csv(f"results/account_value_trade_ens_{100 + w}.csv")
                rw = pd.read_csv(f"results/account_rewards_trade_ens_{100 + w}.csv")
                ac = pd.read_csv(f"results/actions_trade_ens_{100 + w}.csv")
                out[f"w{w}_csv_account_value"] = av["account_value"].to_numpy(np.float64)
                out[f"w{w}_csv_daily_return"] = av["daily_return"].to_numpy(np.float64)
                out[f"w{w}_csv_rewards"] = rw["account_rewards"].to_numpy(np.float64)
                out[f"w{w}_csv_actions"] = ac.iloc[:, 1:].to_numpy(np.int64)
                out[f"w{w}_csv_dates"] = np.array(av["date"].tolist(), dtype="U16")
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), close=close, tech=tech, risk=turb, actions=actions,
                        cfg=np.array([hmax, initial_amount, cost, threshold, W]), **out, **_meta())
    print(f"{name}: 2 windows x {W} days; w0 end cash={out['w0_last_state'][0]:.3f}, w1 first account value="
          f"{out['w1_csv_account_value'][0]:.3f}, w1 last={out['w1_csv_account_value'][-1]:.3f}")


def gen_adapter_portfolio_prediction(name, T, D, K, seed, lookback=60):
    mod = ref_loader.load("env_portfolio")
    close, tech, turb = syn.make_tables(T, D, K, seed=seed)
    cov, first = syn.make_cov_table(close, lookback)
    Te = T - first
    df = syn.make_frame(close[first:], tech[:, first:], turb[first:])
    df["cov_list"] = [cov[t] for t in range(Te) for _ in range(D)]
    env = mod.StockPortfolioEnv(df=df, stock_dim=D, hmax=100, initial_amount=1_000_000, transaction_cost_pct=0.001,
                                reward_scaling=1e-4, state_space=D, action_space=D, tech_indicator_list=syn.INDICATORS[:K])
    actions = syn.make_actions((Te, D), seed=seed + 1, low=0.0, high=1.0)
    os.makedirs("results", exist_ok=True)
    with _quiet():
        df_ret, df_w = adapter_loops.drl_prediction(adapter_loops.ReplayModel(actions), env)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), close=close[first:], tech=tech[:, first:], cov=cov, actions=actions,
        ret_date=np.array(df_ret["date"].tolist(), dtype="U16"), daily_return=df_ret["daily_return"].to_numpy(np.float64),
        weight_date=np.array(df_w.index.tolist(), dtype="U16"), weight_cols=np.array(list(df_w.columns), dtype="U16"),
        weights=df_w.to_numpy(np.float64), **_meta())
    print(f"{name}: return rows={len(df_ret)}, weights={df_w.shape}")


def gen_adapter_cashpenalty_prediction(name, T, D, seed, hmax=5000, threshold=None):
    mod = ref_loader.load("env_stocktrading_cashpenalty")
    close, _tech, turb = syn.make_tables(T, D, 0, seed=seed)
    o, h, l, v = syn.make_ohlv(close, seed)
    df = syn.make_frame(close, np.zeros((0, T, D)), turb, tech_names=[], extra_cols={"open": o, "high": h, "low": l, "volume": v})
    df = df.reset_index(drop=True)
    actions = syn.make_actions((T, D), seed=seed + 1)
    with _quiet():
        env = mod.StockTradingEnvCashpenalty(df=df, hmax=hmax, turbulence_threshold=threshold, print_verbosity=10**9,
                                             random_start=False, cache_indicator_data=True)
        df_account, df_actions = adapter_loops.drl_prediction(adapter_loops.ReplayModel(actions), env)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), close=close, open=o, high=h, low=l, volume=v, turbulence=turb, actions=actions,
        cfg=np.array([hmax, -1.0 if threshold is None else threshold]),
        account_cols=np.array(list(df_account.columns), dtype="U16"),
        account=df_account[["cash", "asset_value", "total_assets", "reward"]].to_numpy(np.float64),
        account_date=np.array(df_account["date"].tolist(), dtype="U16"),
        action_date=np.array(df_actions["date"].tolist(), dtype="U16"),
        logged_actions=np.stack(df_actions["actions"].tolist()), transactions=np.stack(df_actions["transactions"].tolist()),
        **_meta())
    print(f"{name}: account rows={len(df_account)}, action rows={len(df_actions)}")


# ------------------------------------------------------------------------------------------
# table precompute: FeatureEngineer.calculate_turbulence (finrl/meta/preprocessor/preprocessors.py:215-267)
# ------------------------------------------------------------------------------------------
def gen_turbulence(name, T, D, seed, duplicate=None):
    """The reference function itself on a synthetic close frame.  ``duplicate=(a, b)`` makes ticker b an exact
    copy of ticker a: a rank-deficient covariance, so np.linalg.pinv's rcond cut-off is exercised."""
    mod = ref_loader.load("preprocessors")
    close, _, _ = syn.make_tables(T, D, 0, seed=seed)
    if duplicate is not None:
        close[:, duplicate[1]] = close[:, duplicate[0]]
    df = syn.make_frame(close, np.zeros((0, T, D)), np.zeros(T), tech_names=[])
    fe = mod.FeatureEngineer(use_technical_indicator=False, use_turbulence=True)
    out = fe.calculate_turbulence(df)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), close=close, turbulence=out["turbulence"].to_numpy(np.float64),
                        date=np.array(out["date"].tolist(), dtype="U16"), **_meta())
    t = out["turbulence"].to_numpy()
    print(f"{name}: T={T} D={D}, nonzero={int((t > 0).sum())}, max={t.max():.4f}")


def main():
    assert ref_loader.available(), "needs /root/reference"
    which = set(sys.argv[1:])

    def want(k):
        return not which or k in which

    if want("trading"):
        # DOW-30 shape, f32 actions, turbulence liquidation, planted disable flags, 3 episodes + stale resets
        gen_trading("trading_d30_f32", T=40, D=30, K=8, n_steps=125, seed=0, threshold=70, act_dtype=np.float32,
                    plant_disable=[(3, 4), (3, 17), (10, 0), (39, 2), (39, 29)])
        # same shape but cash-starved (the sequential cash-constrained buys decide holdings)
        gen_trading("trading_d30_starved", T=50, D=30, K=8, n_steps=110, seed=2, threshold=90, act_dtype=np.float32,
                    initial_amount=150_000, plant_disable=[(5, 1), (20, 11)])
        # D=5 (8 sort slots), f64 actions, no turbulence, non-zero initial holdings, tight cash
        gen_trading("trading_d5_f64", T=25, D=5, K=2, n_steps=60, seed=3, threshold=None, act_dtype=np.float64,
                    initial_amount=20_000, num_stock_shares=[5, 0, 7, 1, 0], hmax=50)
        # D=13 (16 slots), f32, threshold 99, hmax 10 -> many tied actions
        gen_trading("trading_d13_ties", T=30, D=13, K=3, n_steps=70, seed=5, threshold=99, act_dtype=np.float32, hmax=10,
                    initial_amount=50_000)
        # D=30, hmax=3: almost everything ties -> stresses the argsort tie rule; cash-starved
        gen_trading("trading_d30_ties", T=60, D=30, K=1, n_steps=70, seed=7, threshold=None, act_dtype=np.float32, hmax=3,
                    initial_amount=3_000)
        # single-stock branch of the reference (len(df.tic.unique()) == 1, :415-422, :441-450, :469-476)
        gen_trading("trading_d1_single", T=30, D=1, K=2, n_steps=70, seed=10, threshold=60, act_dtype=np.float32, hmax=50,
                    initial_amount=3_000)
        # wide universes (64- and 128-slot argsort networks): many ties, cash-starved
        gen_trading("trading_d60_wide", T=25, D=60, K=2, n_steps=60, seed=61, threshold=90, act_dtype=np.float32, hmax=10,
                    initial_amount=40_000, plant_disable=[(4, 33), (4, 59), (12, 0)])
        gen_trading("trading_d100_nasdaq", T=20, D=100, K=1, n_steps=45, seed=62, threshold=None, act_dtype=np.float32,
                    hmax=5, initial_amount=60_000, plant_disable=[(3, 64), (3, 99)])
        # actions outside [-1,1] (the reference does not clip) and a high cost
        gen_trading("trading_d8_wide", T=30, D=8, K=2, n_steps=40, seed=9, threshold=50, act_dtype=np.float64, hmax=100,
                    initial_amount=100_000, cost=0.01, act_scale=3.0)
    if want("np"):
        gen_np("np_d30_eval", T=60, D=30, K=8, n_steps=130, seed=11, if_train=False)
        gen_np("np_d30_train", T=40, D=30, K=8, n_steps=90, seed=12, if_train=True, thresh=60)
        gen_np("np_d7_small", T=50, D=7, K=2, n_steps=110, seed=13, if_train=False, initial_capital=2e4, max_stock=50.0,
               thresh=80)
        gen_np("np_d30_kinds", T=40, D=30, K=8, n_steps=90, seed=14, if_train=True, thresh=99, quiet_steps=6)
        gen_np("np_nas100_d20", T=45, D=20, K=4, n_steps=100, seed=16, if_train=True, thresh=60, nas100=True,
               initial_capital=6e4, gamma=0.999)
        # D > 32: the streaming kernel (np_wide.cu); StockEnvNAS100 at its natural size
        gen_np("np_nas100_d100", T=40, D=100, K=2, n_steps=90, seed=17, if_train=True, thresh=60, nas100=True,
               initial_capital=5e5, gamma=0.999)
        gen_np("np_d50_train", T=36, D=50, K=3, n_steps=80, seed=18, if_train=True, thresh=70)
        gen_np("np_d128_eval", T=30, D=128, K=1, n_steps=65, seed=19, if_train=False, thresh=80, initial_capital=3e5)
        gen_np("np_d12_kinds_eval", T=40, D=12, K=3, n_steps=90, seed=15, if_train=False, thresh=70, quiet_steps=5,
               initial_stocks=np.arange(12, dtype=np.float32))
    if want("portfolio"):
        gen_portfolio("portfolio_d30_f64", T=252 + 24, D=30, K=4, n_steps=50, seed=21, act_dtype=np.float64)
        gen_portfolio("portfolio_d6_f32", T=40 + 12, D=6, K=2, n_steps=30, seed=22, act_dtype=np.float32, lookback=40)
        # D > 32: the two-sweep kernel
        gen_portfolio("portfolio_d50_f64", T=60 + 16, D=50, K=2, n_steps=36, seed=23, act_dtype=np.float64, lookback=60)
        gen_portfolio("portfolio_d72_f32", T=80 + 8, D=72, K=1, n_steps=18, seed=24, act_dtype=np.float32, lookback=80)
    if want("stoploss"):
        gen_cashpenalty("stoploss_d7_hmaxvec", T=30, D=7, n_steps=50, seed=55, act_dtype=np.float32,
                        hmax=np.linspace(3000.0, 9000.0, 7), stoploss=(0.92, 2))
        gen_cashpenalty("stoploss_d10", T=40, D=10, n_steps=85, seed=51, act_dtype=np.float32, hmax=8000, stoploss=(0.9, 2))
        gen_cashpenalty("stoploss_d10_turb_patient", T=40, D=10, n_steps=85, seed=52, act_dtype=np.float64, threshold=70,
                        patient=True, hmax=30000, initial_amount=1e5, stoploss=(0.95, 3))
        gen_cashpenalty("stoploss_d8_shortage", T=30, D=8, n_steps=45, seed=53, act_dtype=np.float32, hmax=25000,
                        initial_amount=1e5, stoploss=(0.9, 2))
        gen_cashpenalty("stoploss_d6_discrete", T=30, D=6, n_steps=62, seed=54, act_dtype=np.float32, hmax=4000,
                        discrete=True, shares_increment=2, threshold=90, stoploss=(0.97, 1.5))
    if want("crypto"):
        gen_crypto("crypto_d5_f32", T=45, D=5, K=3, n_steps=100, seed=41, act_dtype=np.float32, lookback=1,
                   initial_capital=2e5, scales=[300.0, 1.5, 0.02, 45.0, 7000.0])
        gen_crypto("crypto_d8_lb3_f64", T=40, D=8, K=2, n_steps=80, seed=42, act_dtype=np.float64, lookback=3,
                   initial_capital=1e6, scales=[30000.0, 2000.0, 1.0, 0.5, 150.0, 20.0, 6.0, 0.08])
    if want("turbulence"):
        gen_turbulence("turbulence_d30", T=300, D=30, seed=91)
        gen_turbulence("turbulence_d7_dup", T=290, D=7, seed=92, duplicate=(2, 5))
        gen_turbulence("turbulence_d100", T=272, D=100, seed=93)
    if want("adapter"):
        gen_adapter_trading_prediction("adapter_trading_prediction", T=36, D=30, K=8, seed=71)
        gen_adapter_ensemble("adapter_ensemble_two_windows", W=21, D=30, K=8, seed=72)
        gen_adapter_portfolio_prediction("adapter_portfolio_prediction", T=60 + 20, D=12, K=3, seed=73)
        gen_adapter_cashpenalty_prediction("adapter_cashpen_prediction", T=26, D=8, seed=74, threshold=85)
    if want("cashpenalty"):
        # the BASELINE shape (config 5): 100 assets -> np.dot takes the blocked BLAS path
        gen_cashpenalty("cashpen_d100_nasdaq", T=40, D=100, n_steps=85, seed=37, act_dtype=np.float32, hmax=2000,
                        threshold=95)
        gen_cashpenalty("cashpen_d7_hmaxvec", T=30, D=7, n_steps=50, seed=35, act_dtype=np.float32,
                        hmax=np.linspace(2000.0, 8000.0, 7))
        gen_cashpenalty("cashpen_d5_hmaxvec_f32", T=24, D=5, n_steps=40, seed=36, act_dtype=np.float32,
                        hmax=np.linspace(2000.0, 8000.0, 5).astype(np.float32), threshold=90)
        gen_cashpenalty("cashpen_d10", T=30, D=10, n_steps=65, seed=31, act_dtype=np.float32, hmax=5000)
        gen_cashpenalty("cashpen_d10_turb_patient", T=30, D=10, n_steps=65, seed=32, act_dtype=np.float64, threshold=70,
                        patient=True, hmax=20000, initial_amount=1e5)
        gen_cashpenalty("cashpen_d10_shortage", T=30, D=10, n_steps=40, seed=33, act_dtype=np.float32, hmax=20000,
                        initial_amount=1e5)
        gen_cashpenalty("cashpen_d6_discrete", T=24, D=6, n_steps=50, seed=34, act_dtype=np.float32, hmax=3000,
                        discrete=True, shares_increment=3, threshold=90)


if __name__ == "__main__":
    main()
