"""GPU parity: the CUDA StockTradingEnvStopLoss path vs reference goldens and the CPU oracle.
The reward's dot products go through BLAS in the reference, so fp64 values are compared at 1e-9
relative (north_star); done / liquidation / shortage flags and the date index are exact."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN
from test_oracle_golden import stoploss_args_from_golden

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

SL = sorted(glob.glob(os.path.join(GOLDEN, "stoploss_*.npz")))
RT = 1e-9


def _close(a, b, ctx="", atol=1e-7):
    np.testing.assert_allclose(a, b, rtol=RT, atol=atol, err_msg=ctx)


@pytest.mark.parametrize("path", SL, ids=[os.path.basename(p)[:-4] for p in SL])
def test_golden_single_env(path):
    from finrl_b200 import BatchedStockTradingEnvStopLoss, CashPenaltyTables

    g = np.load(path)
    close, info, turb, kw = stoploss_args_from_golden(g)
    env = BatchedStockTradingEnvStopLoss(tables=CashPenaltyTables.from_arrays(close, info, turb, "cuda"), n_envs=1,
                                         random_start=False, **kw)
    D = close.shape[1]
    assert np.array_equal(env.reset().cpu().numpy()[0], g["obs0"].astype(np.float32))
    acts = g["actions"]
    for s in range(acts.shape[0]):
        obs, reward, done, flags = env.step(torch.from_numpy(acts[s][None, :]).cuda(), auto_reset=True)
        ctx = f"step {s}"
        assert bool(done[0]) == bool(g["done"][s]), ctx
        assert bool(int(flags[0]) & 2) == bool(g["liq"][s]), ctx
        assert env.date_index[0].item() == g["date_index"][s], ctx
        _close(reward[0].item(), g["reward"][s], ctx, atol=1e-15)
        _close(env.cash[0].item(), g["obs"][s][0], ctx)
        _close(env.holdings[0].cpu().numpy(), g["obs"][s][1 : 1 + D], ctx)
        np.testing.assert_allclose(obs[0].cpu().numpy(), g["obs"][s].astype(np.float32), rtol=2e-7, atol=1e-6, err_msg=ctx)


def _make(N, T=40, D=30, seed=0, **kw):
    from finrl_b200 import BatchedStockTradingEnvStopLoss, CashPenaltyTables, synthetic as syn
    from oracle import oracle as ora

    close, _, turb = syn.make_tables(T, D, 0, seed=seed)
    o_, h_, l_, v_ = syn.make_ohlv(close, seed)
    info = np.stack([o_, close, h_, l_, v_], axis=2)
    args = dict(hmax=5000, initial_amount=1e6)
    args.update(kw)
    env = BatchedStockTradingEnvStopLoss(tables=CashPenaltyTables.from_arrays(close, info, turb, "cuda"), n_envs=N,
                                         random_start=False, **args)
    return env, ora.StopLossOracle(close, info, turb, N, **args)


def _compare_state(env, o, ctx):
    _close(env.cash.cpu().numpy(), o.cash, ctx)
    for name in ("hold", "prev_hold", "avg_buy", "n_buys", "cdiff", "pdiff"):
        _close(getattr(env, name).t().cpu().numpy(), getattr(o, name), f"{ctx} {name}")
    _close(env.last_total.cpu().numpy(), o.last_total, ctx)
    _close(env.last_cash.cpu().numpy(), o.last_cash, ctx)


@pytest.mark.parametrize("N,D,dtype,kw", [
    (1, 30, np.float32, {}),
    (777, 100, np.float32, dict(turbulence_threshold=70)),
    (512, 30, np.float64, dict(patient=True, hmax=60000, initial_amount=2e5, turbulence_threshold=90)),
    (512, 30, np.float32, dict(hmax=40000, initial_amount=1e5)),            # CASH SHORTAGE terminations
    (300, 128, np.float32, dict(discrete_actions=True, shares_increment=2, stoploss_penalty=0.98)),
    (300, 7, np.float32, dict(discrete_actions=True, shares_increment=1, turbulence_threshold=80, profit_loss_ratio=0.5)),
    (400, 30, np.float32, dict(stoploss_penalty=0.999, profit_loss_ratio=1, hmax=1)),  # stop-loss fires often
])
def test_step_vs_oracle(N, D, dtype, kw):
    from finrl_b200 import synthetic as syn

    T = 40
    env, o = _make(N, T=T, D=D, **kw)
    acts = syn.make_actions((T + 25, N, D), seed=7, dtype=dtype)
    seen = 0
    for s in range(acts.shape[0]):
        auto = s % 3 != 0  # exercise both the terminal-without-reset and the auto-reset paths
        obs, reward, done, flags = env.step(torch.from_numpy(acts[s]).cuda(), auto_reset=auto)
        orew, ofl = o.step(acts[s], auto_reset=auto)
        ctx = f"step {s}"
        assert np.array_equal(flags.cpu().numpy(), ofl), ctx
        assert np.array_equal(env.date_index.cpu().numpy(), o.date_index), ctx
        _close(reward.cpu().numpy(), orew, ctx, atol=1e-15)
        _compare_state(env, o, ctx)
        np.testing.assert_allclose(obs.cpu().numpy(), o.obs().astype(np.float32), rtol=2e-7, atol=1e-6, err_msg=ctx)
        seen |= int(np.bitwise_or.reduce(ofl))
        if not auto:  # reset the finished envs by hand, like a gym caller would
            m = (ofl & 1).astype(bool)
            if m.any():
                env.reset(mask=torch.from_numpy(m).cuda())
                o.reset(mask=m)
    assert seen & 1
    if kw.get("turbulence_threshold") is not None:
        assert seen & 2
    if kw.get("hmax", 0) >= 40000:
        assert seen & 4


def test_stoploss_actually_fires():
    """With a tight stop-loss the forced liquidation path must be hit: some env holds an asset at step s and
    holds none of it at s+1 although its action was a buy.  The override is only armed while cash is at least
    stoploss_penalty x initial_amount (:357), hence the tiny hmax."""
    from finrl_b200 import synthetic as syn

    N, D, T = 256, 10, 40
    env, o = _make(N, T=T, D=D, stoploss_penalty=0.999, hmax=1)
    acts = np.abs(syn.make_actions((T - 1, N, D), seed=3))  # buys only
    fired = 0
    for s in range(T - 1):
        before = env.holdings.clone()
        env.step(torch.from_numpy(acts[s]).cuda())
        o.step(acts[s])
        after = env.holdings
        fired += int(((before > 0) & (after == 0)).sum().item())
    assert fired > 0
    _compare_state(env, o, "end")


def test_rollout_matches_stepping():
    from finrl_b200 import synthetic as syn

    N, K, T, D = 1024, 24, 60, 30
    env, o = _make(N, T=T, D=D, turbulence_threshold=85)
    for r in range(3):
        acts = syn.make_actions((K, N, D), seed=40 + r)
        obs, rewards, flags = env.rollout(torch.from_numpy(acts).cuda(), obs_mode="last", auto_reset=True)
        orew = np.empty((K, N))
        ofl = np.empty((K, N), dtype=np.uint8)
        for k in range(K):
            orew[k], ofl[k] = o.step(acts[k], auto_reset=True)
        assert np.array_equal(flags.cpu().numpy(), ofl)
        _close(rewards.cpu().numpy(), orew, atol=1e-15)
        _compare_state(env, o, f"rollout {r}")
        np.testing.assert_allclose(obs.cpu().numpy(), o.obs().astype(np.float32), rtol=2e-7, atol=1e-6)
    st = env.read_stats()
    assert st["env_steps"] == 3 * K * N and st["done_count"] == float(N * (3 * K // T))


def test_gym_dropin_matches_golden_frame():
    """The gym-protocol class on a frame: same numbers as the batched engine fed the same tables."""
    import pandas as pd

    from finrl_b200 import synthetic as syn
    from finrl_b200.env_stocktrading_stoploss import StockTradingEnvStopLoss
    from oracle import oracle as ora

    T, D = 25, 4
    close, _, turb = syn.make_tables(T, D, 0, seed=5)
    o_, h_, l_, v_ = syn.make_ohlv(close, 5)
    dates = pd.date_range("2020-01-01", periods=T).strftime("%Y-%m-%d")
    rows = []
    for t in range(T):
        for d in range(D):
            rows.append(dict(date=dates[t], tic=f"T{d}", open=o_[t, d], close=close[t, d], high=h_[t, d], low=l_[t, d],
                             volume=v_[t, d], turbulence=turb[t]))
    df = pd.DataFrame(rows)
    env = StockTradingEnvStopLoss(df, hmax=5000, random_start=False, turbulence_threshold=90)
    info = np.stack([o_, close, h_, l_, v_], axis=2)
    o = ora.StopLossOracle(close, info, turb, 1, hmax=5000, turbulence_threshold=90)
    acts = syn.make_actions((T - 1, 1, D), seed=11, dtype=np.float64)
    for s in range(T - 1):
        state, reward, done, _ = env.step(acts[s, 0])
        orew, ofl = o.step(acts[s])
        assert done == bool(ofl[0] & 1)
        _close(reward, orew[0], f"step {s}", atol=1e-15)
        if not done:
            np.testing.assert_allclose(state, o.obs()[0], rtol=2e-7, atol=1e-6)
