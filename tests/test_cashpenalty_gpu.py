"""GPU parity: the CUDA StockTradingEnvCashpenalty path vs reference goldens and the CPU oracle.
np.dot's summation order is BLAS-specific, so fp64 values are compared at 1e-9 relative
(north_star); done / liquidation / shortage flags and the date index are exact."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN
from test_oracle_golden import cashpen_args_from_golden

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

CP = sorted(glob.glob(os.path.join(GOLDEN, "cashpen_*.npz")))
RT = 1e-9


def _close(a, b, ctx="", atol=1e-7):
    np.testing.assert_allclose(a, b, rtol=RT, atol=atol, err_msg=ctx)


@pytest.mark.parametrize("path", CP, ids=[os.path.basename(p)[:-4] for p in CP])
def test_golden_single_env(path):
    from finrl_b200 import BatchedStockTradingEnvCashpenalty, CashPenaltyTables

    g = np.load(path)
    close, info, turb, kw = cashpen_args_from_golden(g)
    env = BatchedStockTradingEnvCashpenalty(tables=CashPenaltyTables.from_arrays(close, info, turb, "cuda"), n_envs=1,
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


def _make(N, T=40, D=100, seed=0, **kw):
    from finrl_b200 import BatchedStockTradingEnvCashpenalty, CashPenaltyTables, synthetic as syn
    from oracle import oracle as ora

    close, _, turb = syn.make_tables(T, D, 0, seed=seed)
    o_, h_, l_, v_ = syn.make_ohlv(close, seed)
    info = np.stack([o_, close, h_, l_, v_], axis=2)
    args = dict(hmax=5000, initial_amount=1e6)
    args.update(kw)
    env = BatchedStockTradingEnvCashpenalty(tables=CashPenaltyTables.from_arrays(close, info, turb, "cuda"), n_envs=N,
                                            random_start=False, **args)
    return env, ora.CashPenaltyOracle(close, info, turb, N, **args)


@pytest.mark.parametrize("N,D,dtype,kw", [
    (1, 100, np.float32, {}),
    (777, 100, np.float32, dict(turbulence_threshold=70)),
    (512, 100, np.float64, dict(patient=True, hmax=60000, initial_amount=2e5, turbulence_threshold=90)),
    (512, 30, np.float32, dict(hmax=40000, initial_amount=1e5)),            # CASH SHORTAGE terminations
    (300, 128, np.float32, dict(discrete_actions=True, shares_increment=2)),
    (300, 7, np.float32, dict(discrete_actions=True, shares_increment=1, turbulence_threshold=80)),
])
def test_step_vs_oracle(N, D, dtype, kw):
    from finrl_b200 import synthetic as syn

    T = 40
    env, o = _make(N, T=T, D=D, **kw)
    acts = syn.make_actions((T + 25, N, D), seed=7, dtype=dtype)
    seen = 0
    for s in range(acts.shape[0]):
        obs, reward, done, flags = env.step(torch.from_numpy(acts[s]).cuda(), auto_reset=True)
        orew, ofl = o.step(acts[s], auto_reset=True)
        ctx = f"step {s}"
        assert np.array_equal(flags.cpu().numpy(), ofl), ctx
        assert np.array_equal(env.date_index.cpu().numpy(), o.date_index), ctx
        _close(reward.cpu().numpy(), orew, ctx, atol=1e-15)
        _close(env.cash.cpu().numpy(), o.cash, ctx)
        _close(env.holdings.cpu().numpy(), o.hold, ctx)
        np.testing.assert_allclose(obs.cpu().numpy(), o.obs().astype(np.float32), rtol=2e-7, atol=1e-6, err_msg=ctx)
        seen |= int(np.bitwise_or.reduce(ofl))
    assert seen & 1
    if kw.get("turbulence_threshold") is not None:
        assert seen & 2
    if kw.get("hmax", 0) >= 40000:
        assert seen & 4


def test_rollout_config5_shape():
    """Config 5 shape: D=100 assets, 5 info columns (O=601), fused rollouts vs the oracle."""
    from finrl_b200 import synthetic as syn

    N, K, T, D = 2048, 24, 60, 100
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
        _close(env.cash.cpu().numpy(), o.cash)
        _close(env.holdings.cpu().numpy(), o.hold)
        np.testing.assert_allclose(obs.cpu().numpy(), o.obs().astype(np.float32), rtol=2e-7, atol=1e-6)
    st = env.read_stats()
    assert st["env_steps"] == 3 * K * N and st["done_count"] == float(N * (3 * K // T))


def test_full_size_properties_config5():
    """BASELINE config 5 at its full per-GPU size (262144 envs, D = 100, O = 601): identical actions -> every env equals
    the 1-env oracle; with distinct actions the observation rows are consistent with the state arrays, a sample of envs
    from both ends of the batch (incl. the last tile) equals the oracle, and the statistics add up."""
    from finrl_b200 import synthetic as syn

    N, T, K, D = 262144, 40, 12, 100
    env, _ = _make(N, T=T, D=D, seed=3, turbulence_threshold=85)
    o1 = _make(1, T=T, D=D, seed=3, turbulence_threshold=85)[1]
    acts = syn.make_actions((K, 1, D), seed=21)
    obs, rewards, flags = env.rollout(torch.from_numpy(acts).cuda().expand(K, N, D).contiguous(), obs_mode="last",
                                      auto_reset=True)
    for k in range(K):
        orew, ofl = o1.step(acts[k], auto_reset=True)
        assert bool((flags[k] == int(ofl[0])).all()), k
        assert bool((rewards[k] == rewards[k][0]).all()), k       # every env took the same path, bit for bit
        _close(rewards[k][0].item(), orew[0], f"step {k}", atol=1e-15)
    assert bool((env.cash == env.cash[0]).all()) and bool((env.holdings == env.holdings[0:1]).all())
    _close(env.cash[0].item(), o1.cash[0])
    _close(env.holdings[0].cpu().numpy(), o1.hold[0])
    assert bool((obs == obs[0:1]).all())
    np.testing.assert_allclose(obs[0].cpu().numpy(), o1.obs()[0].astype(np.float32), rtol=2e-7, atol=1e-6)
    # distinct actions per env: rows vs state, a sample vs the oracle
    env.reset()
    M = 70
    sample = np.r_[0:M, N - M:N]
    so = _make(len(sample), T=T, D=D, seed=3, turbulence_threshold=85)[1]
    env.read_stats(reset=True)
    steps = 5
    rsum = 0.0
    for s in range(steps):
        a = torch.from_numpy(syn.make_actions((N, D), seed=60 + s)).cuda()
        obs, reward, done, fl = env.step(a, auto_reset=True, accumulate_stats=True)
        orew, ofl = so.step(a[sample].cpu().numpy(), auto_reset=True)
        assert np.array_equal(fl[sample].cpu().numpy(), ofl), s
        _close(reward[sample].cpu().numpy(), orew, f"step {s}", atol=1e-15)
        assert bool((obs[:, 0] == env.cash.float()).all())
        assert bool((obs[:, 1:1 + D] == env.holdings.float()).all())
        rsum += float(reward.sum())
    _close(env.cash[sample].cpu().numpy(), so.cash)
    _close(env.holdings[sample].cpu().numpy(), so.hold)
    st = env.read_stats()
    assert st["env_steps"] == steps * N
    np.testing.assert_allclose(st["reward_sum"], rsum, rtol=1e-9)


def test_reference_invariants_zero_step_and_patient():
    """The two data-independent invariants of the reference's own tests, on synthetic frames
    (/root/reference/tests/environments/test_cash_penalty.py:29-52 and :55-75)."""
    from finrl_b200 import BatchedStockTradingEnvCashpenalty, CashPenaltyTables, synthetic as syn

    T, D = 20, 2
    close, _, turb = syn.make_tables(T, D, 0, seed=1)
    o_, h_, l_, v_ = syn.make_ohlv(close, 1)
    info = np.stack([o_, close, h_, l_, v_], axis=2)
    tables = CashPenaltyTables.from_arrays(close, info, turb, "cuda")
    # test_zero_step: zero actions -> cash stays initial, holdings 0, asset value 0, step counter advances
    init = 1e6
    env = BatchedStockTradingEnvCashpenalty(tables=tables, n_envs=4, initial_amount=init, random_start=False)
    for i in range(T - 1):
        obs, reward, done, flags = env.step(torch.zeros((4, D), device="cuda"))
        assert bool((env.cash == init).all()) and bool((env.holdings == 0).all())
        assert bool((env.last_total - env.last_cash == 0).all())
        assert bool((env.date_index - env.starting_point == i + 1).all())
    # test_patient: cash == one close of asset 0, hmax = 100 x that close, buy everything, patient -> no done, no holdings
    aapl_first_close = float(close[0, 0])
    env = BatchedStockTradingEnvCashpenalty(tables=tables, n_envs=3, initial_amount=aapl_first_close,
                                            hmax=aapl_first_close * 100, cash_penalty_proportion=0, patient=True,
                                            random_start=False)
    for _ in range(T - 2):
        obs, reward, done, flags = env.step(torch.ones((3, D), device="cuda"))
        assert not bool(done.any())
        assert bool((env.holdings == 0).all())


@pytest.mark.parametrize("kind", ["cashpenalty", "stoploss"])
def test_auto_reset_redraws_the_random_start(kind):
    """random_start=True (the reference's default) must hold for the in-kernel auto-reset too:
    starting_point = random.choice(range(int(T * 0.5))) per env and episode, not 0."""
    from finrl_b200 import BatchedStockTradingEnvCashpenalty, BatchedStockTradingEnvStopLoss, CashPenaltyTables, synthetic as syn

    N, T, D = 4000, 40, 6
    close, _, turb = syn.make_tables(T, D, 0, seed=3)
    o, h, l, v = syn.make_ohlv(close, 3)
    tables = CashPenaltyTables.from_arrays(close, np.stack([o, close, h, l, v], axis=2), turb, "cuda")
    cls = BatchedStockTradingEnvCashpenalty if kind == "cashpenalty" else BatchedStockTradingEnvStopLoss

    def run(seed):
        env = cls(tables=tables, n_envs=N, random_start=True, hmax=100)
        env.seed(seed)
        env.reset(start_points=np.full(N, T - 3))  # everyone two steps from the end
        zero = torch.zeros((N, D), device="cuda")
        for _ in range(3):
            obs, rew, done, fl = env.step(zero, auto_reset=True)
        assert bool(done.all())
        return env.starting_point.cpu().numpy().copy(), env.date_index.cpu().numpy().copy(), obs.cpu().numpy()

    sp, di, obs = run(5)
    assert np.array_equal(sp, di) and sp.min() >= 0 and sp.max() < T // 2
    assert len(np.unique(sp)) == T // 2 and abs(sp.mean() - (T // 2 - 1) / 2) < 0.5  # uniform over range(int(T * 0.5))
    tmpl = tables.obs_tmpl.cpu().numpy()
    assert np.array_equal(obs[:, 1 + D :], tmpl[sp][:, 1 + D :]) and (obs[:, 0] == 1e6).all()  # the reset observation
    assert np.array_equal(run(5)[0], sp) and not np.array_equal(run(6)[0], sp)
