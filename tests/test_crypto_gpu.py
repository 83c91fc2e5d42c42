"""GPU parity of the sibling CryptoEnv: goldens from the unmodified reference + the CPU oracle, bit-exact."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

CRYPTO = sorted(glob.glob(os.path.join(GOLDEN, "crypto_*.npz")))


@pytest.mark.parametrize("path", CRYPTO, ids=[os.path.basename(p)[:-4] for p in CRYPTO])
def test_golden_gym_class(path):
    from finrl_b200.crypto import CryptoEnv

    g = np.load(path)
    env = CryptoEnv({"price_array": g["price_array"], "tech_array": g["tech_array"]}, lookback=int(g["lookback"]),
                    initial_capital=float(g["initial_capital"]))
    assert np.array_equal(env.action_norm_vector, g["action_norm_vector"])
    assert np.array_equal(env.reset(), g["obs0"])
    acts = g["actions"]
    for s in range(acts.shape[0]):
        state, reward, done, info = env.step(acts[s].copy())
        ctx = f"step {s}"
        assert info is None and done == bool(g["done"][s]) and env.time == g["time"][s], ctx
        assert env.cash == g["cash"][s] and np.array_equal(env.stocks, g["stocks"][s]), ctx
        assert env.total_asset == g["total"][s] and env.gamma_return == g["gamma_return"][s] and reward == g["reward"][s], ctx
        assert state.dtype == np.float32 and np.array_equal(state, g["obs"][s]), ctx
        if done:
            assert env.episode_return == g["episode_return"][s], ctx
            env.reset()


@pytest.mark.parametrize("N,D,lookback,dtype", [(1, 5, 1, np.float32), (1000, 32, 2, np.float32), (333, 11, 3, np.float64)])
def test_step_and_rollout_vs_oracle(N, D, lookback, dtype):
    from finrl_b200 import synthetic as syn
    from finrl_b200.crypto import BatchedCryptoEnv
    from oracle import oracle as ora

    T, K = 40, 2
    close, tech, turb = syn.make_tables(T, D, K, seed=D)
    close = close * (10.0 ** np.random.default_rng(D).integers(-2, 4, D))[None, :]
    pa, ta, _ = syn.make_np_arrays(close, tech, turb)
    env = BatchedCryptoEnv({"price_array": pa, "tech_array": ta}, lookback=lookback, initial_capital=5e5, n_envs=N)
    o = ora.CryptoOracle(pa, ta, N, lookback=lookback, initial_capital=5e5)
    acts = syn.make_actions((2 * T, N, D), seed=3, dtype=dtype)
    for s in range(T + 5):
        obs, reward, done, flags = env.step(torch.from_numpy(acts[s]).cuda())
        oobs, orew, ofl = o.step(acts[s])
        ctx = f"step {s}"
        assert np.array_equal(flags.cpu().numpy(), ofl) and np.array_equal(reward.cpu().numpy(), orew), ctx
        assert np.array_equal(obs.cpu().numpy(), oobs), ctx
        assert np.array_equal(env.cash.cpu().numpy(), o.cash) and np.array_equal(env.stocks.t().cpu().numpy(), o.stocks), ctx
        assert np.array_equal(env.total_asset.cpu().numpy(), o.total) and np.array_equal(env.gamma_return.cpu().numpy(), o.gamma_return), ctx
        if done.all():
            assert np.array_equal(env.reset().cpu().numpy(), o.reset())
    # fused rollout with kernel-side auto reset
    K2 = T - 3
    obs, rewards, flags = env.rollout(torch.from_numpy(acts[:K2]).cuda(), obs_mode="last", auto_reset=True)
    for k in range(K2):
        oobs, orew, ofl = o.step(acts[k])
        assert np.array_equal(rewards[k].cpu().numpy(), orew) and np.array_equal(flags[k].cpu().numpy(), ofl)
        if ofl[0] & 1:
            oobs = o.reset()
    assert np.array_equal(obs.cpu().numpy(), oobs)
