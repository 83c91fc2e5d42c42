"""GPU parity: the CUDA numpy/ElegantRL StockTradingEnv path (through the C-ABI) vs golden vectors
made by the unmodified reference and vs the CPU oracle.  Bit-exact, including the data-dependent
numpy kinds (Python float / np.float32 / np.float64) of amount, total_asset, gamma_reward, reward."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN
from test_oracle_golden import NpResetDraws, np_kwargs_from_golden

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

NP = sorted(glob.glob(os.path.join(GOLDEN, "np_*.npz")))


@pytest.fixture(autouse=True, params=["regs", "wide", "wide_generic"])
def np_kernel(request):
    """Every test of this file runs three times: with the register kernel for D <= 32 (nptrading.cu), with the
    streaming kernel (np_wide.cu, the one D > 32 always uses) forced for every stock count, and with the streaming
    kernel's bulk-staged variant (float32 actions, D a multiple of four) switched off."""
    from finrl_b200 import _cabi

    _cabi.set_option("np_wide_min_d", 33 if request.param == "regs" else 1)
    _cabi.set_option("np_wide_bulk", 0 if request.param == "wide_generic" else 1)
    yield request.param
    _cabi.set_option("np_wide_min_d", 33)
    _cabi.set_option("np_wide_bulk", 1)


@pytest.mark.parametrize("path", NP, ids=[os.path.basename(p)[:-4] for p in NP])
def test_golden_single_env(path):
    from finrl_b200 import BatchedNpStockTradingEnv

    g = np.load(path)
    cfg = {"price_array": g["price_array"], "tech_array": g["tech_array"], "turbulence_array": g["turbulence_array"],
           "if_train": False}
    env = BatchedNpStockTradingEnv(cfg, n_envs=1, **np_kwargs_from_golden(g))
    draws = NpResetDraws(g)
    s0, f = draws.draw()
    obs = env.reset(stocks0=s0, factor=f)
    assert np.array_equal(obs.cpu().numpy()[0], g["obs0"])
    st = env.get_state()
    assert st["amount"][0].item() == g["init_amount"] and st["amount_kind"][0].item() == g["init_amount_kind"]
    assert st["total"][0].item() == g["init_total"] and st["total_kind"][0].item() == g["init_total_kind"]
    acts = g["actions"]
    nreset = 0
    for s in range(acts.shape[0]):
        obs, reward, done, flags = env.step(torch.from_numpy(acts[s][None, :]).cuda())
        st = env.get_state()
        ctx = f"step {s}"
        fl = int(flags[0])
        assert bool(fl & 1) == bool(g["done"][s]), ctx
        assert bool(fl & 2) == bool(g["liq"][s]), ctx
        assert np.array_equal(st["stocks"][0].cpu().numpy(), g["stocks"][s]), ctx
        assert np.array_equal(st["cool"][0].cpu().numpy(), g["cool"][s]), ctx
        assert (st["amount"][0].item(), st["amount_kind"][0].item()) == (g["amount"][s], g["amount_kind"][s]), ctx
        assert (st["total"][0].item(), st["total_kind"][0].item()) == (g["total"][s], g["total_kind"][s]), ctx
        assert (st["gamma_reward"][0].item(), st["gr_kind"][0].item()) == (g["gamma_reward"][s], g["gr_kind"][s]), ctx
        assert (reward[0].item(), fl >> 4) == (g["reward"][s], g["reward_kind"][s]), ctx
        assert st["day"][0].item() == g["day"][s], ctx
        assert np.array_equal(obs[0].cpu().numpy(), g["obs"][s]), ctx
        if g["done"][s]:
            assert st["episode_return"][0].item() == g["episode_return"][s], ctx
            s0, f = draws.draw()
            env.reset(stocks0=s0, factor=f)
            st = env.get_state()
            assert st["amount"][0].item() == g["reset_amount"][nreset]
            assert st["amount_kind"][0].item() == g["reset_amount_kind"][nreset]
            assert np.array_equal(st["stocks"][0].cpu().numpy(), g["reset_stocks"][nreset])
            nreset += 1


def _make(N, T=50, D=30, K=8, seed=0, thresh=80, **kw):
    from finrl_b200 import BatchedNpStockTradingEnv, synthetic as syn
    from oracle import oracle as ora

    close, tech, turb = syn.make_tables(T, D, K, seed=seed)
    pa, ta, tu = syn.make_np_arrays(close, tech, turb)
    args = dict(turbulence_thresh=thresh, initial_capital=3e5)
    args.update(kw)
    env = BatchedNpStockTradingEnv({"price_array": pa, "tech_array": ta, "turbulence_array": tu, "if_train": False},
                                   n_envs=N, **args)
    o = ora.NpTradingOracle(pa, ta, tu, N, **args)
    return env, o


def _compare(env, o, ctx=""):
    st = env.get_state()
    assert np.array_equal(st["amount"].cpu().numpy(), o.amount), ctx
    assert np.array_equal(st["amount_kind"].cpu().numpy(), o.amount_kind), ctx
    assert np.array_equal(st["stocks"].cpu().numpy(), o.stocks), ctx
    assert np.array_equal(st["cool"].cpu().numpy(), o.cool), ctx
    assert np.array_equal(st["day"].cpu().numpy(), o.day), ctx
    assert np.array_equal(st["total"].cpu().numpy(), o.total), ctx
    assert np.array_equal(st["total_kind"].cpu().numpy(), o.total_kind), ctx
    assert np.array_equal(st["gamma_reward"].cpu().numpy(), o.gamma_reward), ctx
    assert np.array_equal(st["gr_kind"].cpu().numpy(), o.gr_kind), ctx


@pytest.mark.parametrize("N,D", [(1, 30), (77, 30), (2048 + 5, 30), (300, 7), (300, 16), (300, 28), (100, 100), (65, 64)])
def test_step_vs_oracle(N, D):
    """Random train-style initial positions, distinct actions, several episodes (manual resets)."""
    from finrl_b200 import synthetic as syn

    T = 50
    env, o = _make(N, T=T, D=D, K=3)
    rng = np.random.RandomState(3)
    acts = syn.make_actions((2 * T + 7, N, D), seed=6)
    acts[:4] *= 0.08  # dead band: amount keeps its Python-float / f32 kind for a while

    def reset_both():
        s0 = rng.randint(0, 64, size=(N, D)).astype(np.float32)
        f = rng.uniform(0.95, 1.05, size=N)
        obs = env.reset(stocks0=s0, factor=f)
        assert np.array_equal(obs.cpu().numpy(), o.reset(stocks0=s0, factor=f))

    reset_both()
    for s in range(acts.shape[0]):
        obs, reward, done, flags = env.step(torch.from_numpy(acts[s]).cuda())
        oobs, orew, ork, ofl = o.step(acts[s])
        ctx = f"step {s}"
        assert np.array_equal(flags.cpu().numpy(), ofl | (ork << 4)), ctx
        assert np.array_equal(reward.cpu().numpy(), orew), ctx
        assert np.array_equal(obs.cpu().numpy(), oobs), ctx
        _compare(env, o, ctx)
        if done.all():
            assert np.array_equal(env.episode_return.cpu().numpy(), o.episode_return)
            reset_both()


def test_ragged_days_and_mixed_kinds_inside_a_tile():
    """Masked resets in mid-episode leave the 32 envs of a tile on different days and with different numpy kinds
    (freshly reset envs carry a Python-float amount, the others float64): the observation writer's mixed-day path and
    the general (not all-float64) instantiation of the step body, at DOW-30 size and at a generic one."""
    from finrl_b200 import synthetic as syn

    for D in (30, 23, 24):
        N, T = 100, 30
        env, o = _make(N, T=T, D=D, K=8)
        acts = syn.make_actions((3 * T, N, D), seed=40 + D)
        rng = np.random.RandomState(D)
        assert np.array_equal(env.reset().cpu().numpy(), o.reset())
        for s in range(acts.shape[0]):
            obs, reward, done, flags = env.step(torch.from_numpy(acts[s]).cuda())
            oobs, orew, ork, ofl = o.step(acts[s])
            ctx = f"D={D} step {s}"
            assert np.array_equal(flags.cpu().numpy(), ofl | (ork << 4)), ctx
            assert np.array_equal(reward.cpu().numpy(), orew), ctx
            assert np.array_equal(obs.cpu().numpy(), oobs), ctx
            mask = (rng.rand(N) < 0.15) | ((ofl & 1) != 0)  # every done env, plus a random 15 %
            if mask.any():
                m = mask.astype(np.uint8)
                assert np.array_equal(env.reset(mask=torch.from_numpy(m).cuda()).cpu().numpy(), o.reset(mask=m)), ctx
            _compare(env, o, ctx)
        assert len(np.unique(o.day)) > 3


@pytest.mark.parametrize("D", [30, 64, 100])
@pytest.mark.parametrize("layout", ["KND", "NKD"])
def test_rollout_auto_reset_vs_oracle(layout, D):
    """Fused K-step rollouts with the deterministic auto-reset after each done step (D = 64 / 100: the streaming
    kernel, bulk-staged for the KND layout — with the padded and the natural row pitch — and generic for NKD)."""
    from finrl_b200 import synthetic as syn

    N, K, T = (1024, 40, 60) if D == 30 else (200, 25, 30)
    env, o = _make(N, T=T, D=D, K=8 if D == 30 else 2)
    for r in range(4):
        acts = syn.make_actions((K, N, D), seed=30 + r)
        a_dev = torch.from_numpy(acts).cuda()
        if layout == "NKD":
            a_dev = a_dev.permute(1, 0, 2).contiguous()
        obs, rewards, flags = env.rollout(a_dev, layout=layout, obs_mode="last", auto_reset=True)
        orew = np.empty((K, N))
        ofl = np.empty((K, N), dtype=np.uint8)
        for k in range(K):
            oobs, orew[k], ork, f = o.step(acts[k])
            ofl[k] = f | (ork << 4)
            if f[0] & 1:
                oobs = o.reset()
        assert np.array_equal(flags.cpu().numpy(), ofl)
        assert np.array_equal(rewards.cpu().numpy(), orew)
        assert np.array_equal(obs.cpu().numpy(), oobs)
        _compare(env, o, f"rollout {r}")
    st = env.read_stats()
    assert st["env_steps"] == 4 * K * N and st["done_count"] == N * (4 * K // (T - 1))


@pytest.mark.parametrize("D", [30, 100, 64])
def test_output_buffers_have_no_out_of_bounds_writes(D):
    """compute-sanitizer is closed on this GPU pool: ragged last tile, outputs embedded in sentinel-filled allocations,
    every obs mode — the register kernel, and the streaming kernel with its bulk-staged variant (natural and padded row
    pitch)."""
    from finrl_b200 import synthetic as syn

    N, K, T = 2048 + 13, 4, 30
    env, _ = _make(N, T=T, D=D, K=8 if D == 30 else 2)
    O = env.state_dim
    acts = torch.from_numpy(syn.make_actions((K, N, D), seed=3)).cuda()
    pad = 64
    for mode, oshape in (("all", (K, N, O)), ("last", (N, O))):
        big_obs = torch.full((int(np.prod(oshape)) + 2 * pad,), -7.25, dtype=torch.float32, device="cuda")
        big_rew = torch.full((K * N + 2 * pad,), -7.25, dtype=torch.float64, device="cuda")
        big_fl = torch.full((K * N + 2 * pad,), 199, dtype=torch.uint8, device="cuda")
        obs = big_obs[pad:-pad].view(*oshape)
        env.rollout(acts, obs_mode=mode, rewards=big_rew[pad:-pad].view(K, N), flags=big_fl[pad:-pad].view(K, N), obs=obs)
        for big, val in ((big_obs, -7.25), (big_rew, -7.25), (big_fl, 199)):
            assert bool((big[:pad] == val).all()) and bool((big[-pad:] == val).all())
        assert not bool((obs == -7.25).any()) and not bool((big_fl[pad:-pad] == 199).any())


def test_full_size_properties():
    """1M envs (config 3): identical actions -> every env equals the 1-env oracle; obs consistent."""
    from finrl_b200 import synthetic as syn

    N, T, K = 1 << 20, 40, 16
    env, _ = _make(N, T=T, seed=4)
    o1 = _make(1, T=T, seed=4)[1]
    acts = syn.make_actions((K, 1, 30), seed=8)
    obs, rewards, flags = env.rollout(torch.from_numpy(acts).cuda().expand(K, N, 30).contiguous(), obs_mode="last",
                                      auto_reset=False)
    for k in range(K):
        oobs, orew, ork, ofl = o1.step(acts[k])
        assert bool((rewards[k] == float(orew[0])).all())
        assert bool((flags[k] == int(ofl[0] | (ork[0] << 4))).all())
    assert bool((env.amount == float(o1.amount[0])).all())
    assert bool((env.stocks == torch.from_numpy(o1.stocks[0]).cuda()[:, None]).all())
    assert bool((obs == torch.from_numpy(oobs[0]).cuda()[None, :]).all())
    acts2 = torch.from_numpy(syn.make_actions((N, 30), seed=9)).cuda()
    obs, reward, done, fl = env.step(acts2)
    assert bool((obs[:, 0] == (env.amount.float() * 2.0**-12)).all())
    assert bool((obs[:, 33:63] == env.stocks.t() * 2.0**-6).all())
    assert bool((obs[:, 63:93] == env.cool.t()).all())


def test_auto_reset_redraws_the_if_train_position():
    """with if_train the in-kernel auto-reset must take the reference's RANDOM branch (:85-92), not
    restart every episode from initial_capital / initial_stocks.  Draws are counter-based (seed, launch, env,
    step): reproducible under a seed, different per env and per episode, and in the reference's ranges."""
    from finrl_b200 import BatchedNpStockTradingEnv, synthetic as syn

    for D in (30, 50):  # register kernel and streaming kernel
        N, T, K = 3000, 6, 2
        pa, ta, tu = syn.make_np_arrays(*syn.make_tables(T, D, K, seed=2))
        cfg = {"price_array": pa, "tech_array": ta, "turbulence_array": tu, "if_train": True}
        init = np.arange(D, dtype=np.float32) % 3

        def run(seed):
            env = BatchedNpStockTradingEnv(cfg, n_envs=N, initial_stocks=init, turbulence_thresh=1e9)
            env.seed(seed)
            zero = torch.zeros((N, D), device="cuda")  # inside the dead-band: nothing trades, positions stay as drawn
            snaps = []
            for s in range(2 * (T - 1)):
                obs, rew, done, fl = env.step(zero, auto_reset=True)
                if bool(done[0]):
                    assert bool(done.all())
                    snaps.append((env.stocks.t().cpu().numpy().copy(), env.amount.cpu().numpy().copy(),
                                  (env.kinds & 3).cpu().numpy().copy(), env.day.cpu().numpy().copy(), obs.cpu().numpy().copy()))
            return snaps

        a, b, c = run(11), run(11), run(12)
        assert len(a) == 2
        price0 = pa[0].astype(np.float32)
        for ep, (stocks, amount, kind, day, obs) in enumerate(a):
            extra = stocks - init[None, :]
            assert extra.min() >= 0 and extra.max() <= 63 and np.array_equal(extra, np.round(extra))
            assert len(np.unique(extra[:, 0])) > 40 and abs(extra.mean() - 31.5) < 0.5  # uniform over 0..63
            assert (day == 0).all() and (kind == 1).all()  # amount is np.float32 after an if_train reset
            asset = (stocks * price0[None, :]).astype(np.float32).sum(axis=1, dtype=np.float64)
            factor = (amount + asset) / 1e6
            assert factor.min() >= 0.9499 and factor.max() <= 1.0501 and 0.99 < factor.mean() < 1.01 and factor.std() > 0.02
            assert np.array_equal(obs[:, 3 + D : 3 + 2 * D], (stocks * np.float32(2**-6)).astype(np.float32))  # obs shows the reset state
            assert np.array_equal(b[ep][0], stocks) and np.array_equal(b[ep][1], amount)  # same seed: same draws
            assert not np.array_equal(c[ep][0], stocks)                                    # other seed: other draws
        assert not np.array_equal(a[0][0], a[1][0])  # a new episode redraws
        assert not np.array_equal(a[0][0][0], a[0][0][1])  # envs differ
        # a fused rollout resets in-kernel as well (step index enters the counter)
        env = BatchedNpStockTradingEnv(cfg, n_envs=N, initial_stocks=init, turbulence_thresh=1e9)
        env.rollout(torch.zeros((2 * T, N, D), device="cuda"), obs_mode="none", auto_reset=True)
        extra = env.stocks.t().cpu().numpy() - init[None, :]
        assert 0 <= extra.min() and extra.max() <= 63 and len(np.unique(extra)) == 64
