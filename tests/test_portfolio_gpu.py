"""GPU parity: the CUDA StockPortfolioEnv path vs reference goldens and the CPU oracle.
Tolerance (north_star): 1e-9 relative on fp64 values with float64 actions; np.exp float32 differs
from CUDA expf by up to ~2e-7, so float32 actions are checked at 2e-6.  Day/done flags and the
table-derived observation are exact."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

PF = sorted(glob.glob(os.path.join(GOLDEN, "portfolio_*.npz")))


@pytest.mark.parametrize("path", PF, ids=[os.path.basename(p)[:-4] for p in PF])
def test_golden_single_env(path):
    from finrl_b200 import BatchedStockPortfolioEnv, PortfolioTables

    g = np.load(path)
    acts = g["actions"]
    tol = 1e-9 if acts.dtype == np.float64 else 2e-6
    env = BatchedStockPortfolioEnv(tables=PortfolioTables.from_arrays(g["close"], g["cov"], g["tech"], "cuda"), n_envs=1,
                                   initial_amount=float(g["initial_amount"]))
    obs = env.reset()
    assert np.array_equal(obs[0].cpu().numpy(), g["obs0"].astype(np.float32))
    for s in range(acts.shape[0]):
        obs, reward, done, flags = env.step(torch.from_numpy(acts[s][None, :]).cuda(), auto_reset=True)
        ctx = f"step {s}"
        assert bool(done[0]) == bool(g["done"][s]), ctx
        assert env.day[0].item() == g["day"][s], ctx
        assert abs(reward[0].item() - g["reward"][s]) <= tol * abs(g["reward"][s]), ctx
        assert abs(env.portfolio_value[0].item() - g["pv"][s]) <= tol * abs(g["pv"][s]), ctx
        assert np.array_equal(obs[0].cpu().numpy(), g["obs"][s].astype(np.float32)), ctx
        assert np.array_equal(env.observe_view()[0].cpu().numpy(), g["obs"][s].astype(np.float32)), ctx


def _make(N, T=40, D=30, K=4, seed=0):
    from finrl_b200 import BatchedStockPortfolioEnv, PortfolioTables, synthetic as syn
    from oracle import oracle as ora

    close, tech, _ = syn.make_tables(T + 20, D, K, seed=seed)
    cov, first = syn.make_cov_table(close, 20)
    close, tech = close[first:], tech[:, first:]
    env = BatchedStockPortfolioEnv(tables=PortfolioTables.from_arrays(close, cov, tech, "cuda"), n_envs=N)
    return env, ora.PortfolioOracle(close, cov, tech, N)


@pytest.mark.parametrize("N,D,dtype", [(1, 30, np.float64), (4096 + 3, 30, np.float64), (500, 6, np.float32), (500, 13, np.float64),
                                           (700, 100, np.float64), (333, 128, np.float32), (64, 33, np.float64)])
def test_step_vs_oracle(N, D, dtype):
    from finrl_b200 import synthetic as syn

    T = 40
    tol = 1e-9 if dtype == np.float64 else 2e-6
    env, o = _make(N, T=T, D=D)
    env.reset(want_obs=False)
    acts = syn.make_actions((2 * T + 5, N, D), seed=3, low=0.0, high=1.0, dtype=dtype)
    for s in range(acts.shape[0]):
        obs, reward, done, flags = env.step(torch.from_numpy(acts[s]).cuda(), auto_reset=True, want_obs=(s % 7 == 0))
        oreward, oflags, _, _ = o.step(acts[s], auto_reset=True)
        assert np.array_equal(flags.cpu().numpy(), oflags)
        assert np.array_equal(env.day.cpu().numpy(), o.day)
        np.testing.assert_allclose(reward.cpu().numpy(), oreward, rtol=tol, atol=0)
        np.testing.assert_allclose(env.portfolio_value.cpu().numpy(), o.pv, rtol=tol, atol=0)
        if obs is not None:
            assert np.array_equal(obs.cpu().numpy(), o.obs().astype(np.float32))


def test_rollout_config4():
    """Config 4: 262,144 envs, fused rollout (no obs traffic) then a materialised observation;
    first 2048 envs against the oracle, the rest through the property that envs fed identical
    actions stay identical."""
    from finrl_b200 import synthetic as syn

    N, K, T, D = 262144, 48, 40, 30
    env, _ = _make(N, T=T)
    o = _make(2048, T=T)[1]
    env.reset(want_obs=False)
    acts = syn.make_actions((K, 2048, D), seed=5, low=0.0, high=1.0, dtype=np.float64)
    a_dev = torch.from_numpy(acts).cuda().repeat(1, N // 2048, 1)  # env n uses actions of n % 2048
    obs, rewards, flags = env.rollout(a_dev, obs_mode="last", auto_reset=True)
    orew = np.empty((K, 2048))
    for k in range(K):
        orew[k], ofl, _, _ = o.step(acts[k], auto_reset=True)
        assert bool((flags[k] == int(ofl[0])).all())
    np.testing.assert_allclose(rewards[:, :2048].cpu().numpy(), orew, rtol=1e-9, atol=0)
    assert bool((rewards.view(K, N // 2048, 2048) == rewards[:, None, :2048]).all())
    assert np.array_equal(obs[:2048].cpu().numpy(), o.obs().astype(np.float32).reshape(2048, -1))
    assert bool((obs.view(N, -1) == env.tables.obs_table[env.day.long()]).all())
    st = env.read_stats()
    assert st["env_steps"] == K * N and st["done_count"] == N * (K // T)
