"""CPU: pin the oracle (oracle/oracle.c) against golden vectors made by the UNMODIFIED reference
(tests/golden/make_golden.py) and against numpy itself for the numpy-defined building blocks."""
import glob
import os

import numpy as np
import pytest

from oracle import oracle as ora

from conftest import GOLDEN


# ---------------------------------------------------------------------------- building blocks
def test_floor_divide_matches_numpy():
    rng = np.random.default_rng(0)
    a = np.concatenate([rng.uniform(-1e6, 1e6, 20000), rng.uniform(0, 50, 5000), [1.0, 0.0, -0.0, 5.0, -5.0, 1e-300]])
    b = np.concatenate([rng.uniform(0.5, 500, 20000), rng.uniform(0.01, 3, 5000), [0.1, 3.0, 3.0, 5.0, 5.0, 7.0]])
    # exact multiples (where floor(a/b) != numpy's divmod result most often)
    k = rng.integers(-500, 500, 5000).astype(np.float64)
    bb = rng.uniform(0.01, 300, 5000)
    a = np.concatenate([a, k * bb, k * 0.1, k * 100.1])
    b = np.concatenate([b, bb, np.full(5000, 0.1), np.full(5000, 100.1)])
    want = np.floor_divide(a, b)
    got = np.array([ora.floor_divide_f64(x, y) for x, y in zip(a, b)])
    assert np.array_equal(want, got)
    assert np.array_equal(np.signbit(want), np.signbit(got))
    a32, b32 = a.astype(np.float32), b.astype(np.float32)
    want32 = np.floor_divide(a32, b32)
    got32 = np.array([ora.floor_divide_f32(x, y) for x, y in zip(a32, b32)], dtype=np.float32)
    assert np.array_equal(want32, got32)
    assert ora.floor_divide_f64(1.0, 0.1) == 9.0


@pytest.mark.parametrize("n", [1, 2, 5, 8, 13, 16, 30, 32, 60, 64, 100, 128, 200, 256])
def test_argsort_rule_matches_numpy(n):
    """np.argsort's default tie order == the bitonic network rule.  Checked against the
    numpy of THIS host; skipped (not failed) if the host's numpy dispatches another sort."""
    rng = np.random.default_rng(n)
    # self-check the host first: if numpy here disagrees on a canonical vector, the host lacks the
    # AVX2/AVX-512 path the goldens were made with and the comparison is meaningless.
    bad = 0
    for trial in range(300):
        hi = [2, 4, 11, 101][trial % 4]
        keys = rng.integers(-hi, hi + 1, size=n).astype(np.int64)
        if not np.array_equal(np.argsort(keys), ora.argsort_i64(keys)):
            bad += 1
    if bad and n <= 16:
        pytest.skip(f"host numpy argsort does not use the SIMD network for n={n} ({bad}/300 differ)")
    assert bad == 0


@pytest.mark.parametrize("n", [0, 1, 5, 7, 8, 9, 30, 100, 127, 128, 129, 200, 1000])
def test_pairwise_sum_matches_numpy(n):
    rng = np.random.default_rng(n + 1)
    for _ in range(50):
        x = (rng.normal(0, 1000, n) * rng.uniform(0, 64, n)).astype(np.float32)
        assert ora.pairwise_sum_f32(x) == x.sum()
        y = x.astype(np.float64) * 1.000001
        assert ora.pairwise_sum_f64(y) == y.sum()


# ---------------------------------------------------------------------------- A1
TRADING = sorted(glob.glob(os.path.join(GOLDEN, "trading_*.npz")))


def trading_oracle_from_golden(g, n_envs=1):
    hmax, init, bc, sc, rs, use_t, thr = g["cfg"]
    return ora.TradingOracle(
        g["close"], g["tech"], g["risk"], n_envs, hmax=hmax, initial_amount=init, buy_cost_pct=bc, sell_cost_pct=sc,
        reward_scaling=rs, turbulence_threshold=(thr if use_t > 0 else None), num_stock_shares=g["num_stock_shares"],
    )


@pytest.mark.parametrize("path", TRADING, ids=[os.path.basename(p)[:-4] for p in TRADING])
def test_trading_oracle_vs_reference(path):
    g = np.load(path)
    o = trading_oracle_from_golden(g)
    assert np.array_equal(o.obs()[0], g["obs0"].astype(np.float32))
    acts = g["actions"]
    for s in range(acts.shape[0]):
        obs, reward, flags = o.step(acts[s][None, :], auto_reset=True)
        ctx = f"step {s}"
        assert bool(flags[0] & ora.FLAG_DONE) == bool(g["done"][s]), ctx
        assert bool(flags[0] & ora.FLAG_LIQUIDATE) == bool(g["liq"][s]), ctx
        assert reward[0] == g["reward"][s], ctx  # bit-exact f64
        assert o.cash[0] == g["cash"][s], ctx
        assert np.array_equal(o.hold[0], g["hold"][s]), ctx
        assert o.day[0] == g["day"][s], ctx
        assert np.array_equal(obs[0], g["obs"][s]), ctx
        if not g["done"][s]:
            assert o.trades[0] == g["trades"][s], ctx
            assert o.cost[0] == g["cost"][s], ctx


def test_trading_oracle_no_auto_reset_terminal_is_noop():
    g = np.load(os.path.join(GOLDEN, "trading_d5_f64.npz"))
    o = trading_oracle_from_golden(g)
    T = g["close"].shape[0]
    acts = g["actions"]
    for s in range(T - 1):
        o.step(acts[s][None, :])
    cash, hold, rew = o.cash.copy(), o.hold.copy(), o.reward.copy()
    for s in range(3):  # terminal step: state unchanged, previous reward again (Q3)
        obs, reward, flags = o.step(acts[T - 1 + s][None, :])
        assert flags[0] & ora.FLAG_DONE
        assert reward[0] == rew[0] and o.cash[0] == cash[0] and np.array_equal(o.hold, hold)
        assert np.array_equal(obs[0], g["term_obs"][T - 1])


# ---------------------------------------------------------------------------- A2
NP = sorted(glob.glob(os.path.join(GOLDEN, "np_*.npz")))


def np_kwargs_from_golden(g):
    kw = {str(k): float(v) for k, v in zip(g["kw_keys"], g["kw_vals"])}
    kw["turbulence_thresh"] = float(g["thresh"])
    kw["initial_stocks"] = g["initial_stocks"]
    if "nas100" in g.files and int(g["nas100"]):
        kw["obs_amount_floor"] = 1e4  # StockEnvNAS100.get_state (env_nas100_wrds.py:157)
    return kw


class NpResetDraws:
    """Replays the reference's global-RNG draws of reset() (env_stocktrading_np.py:85-92): per reset
    rd.randint(0, 64, D) then rd.uniform(0.95, 1.05)."""

    def __init__(self, g):
        self.rs = np.random.RandomState(int(g["rng_seed"]))
        self.init = g["initial_stocks"]
        self.train = bool(g["if_train"])

    def draw(self):
        if not self.train:
            return None, None
        stocks0 = (self.init + self.rs.randint(0, 64, size=self.init.shape)).astype(np.float32)
        return stocks0[None, :], np.array([self.rs.uniform(0.95, 1.05)])


@pytest.mark.parametrize("path", NP, ids=[os.path.basename(p)[:-4] for p in NP])
def test_np_oracle_vs_reference(path):
    g = np.load(path)
    o = ora.NpTradingOracle(g["price_array"], g["tech_array"], g["turbulence_array"], 1, **np_kwargs_from_golden(g))
    draws = NpResetDraws(g)
    s0, f = draws.draw()
    obs = o.reset(stocks0=s0, factor=f)
    assert np.array_equal(obs[0], g["obs0"])
    assert o.amount[0] == g["init_amount"] and o.amount_kind[0] == g["init_amount_kind"]
    assert o.total[0] == g["init_total"] and o.total_kind[0] == g["init_total_kind"]
    acts = g["actions"]
    nreset = 0
    for s in range(acts.shape[0]):
        obs, reward, rk, flags = o.step(acts[s][None, :])
        ctx = f"step {s}"
        assert bool(flags[0] & ora.FLAG_DONE) == bool(g["done"][s]), ctx
        assert bool(flags[0] & ora.FLAG_LIQUIDATE) == bool(g["liq"][s]), ctx
        assert np.array_equal(o.stocks[0], g["stocks"][s]), ctx
        assert np.array_equal(o.cool[0], g["cool"][s]), ctx
        assert (o.amount[0], o.amount_kind[0]) == (g["amount"][s], g["amount_kind"][s]), ctx
        assert (o.total[0], o.total_kind[0]) == (g["total"][s], g["total_kind"][s]), ctx
        assert (o.gamma_reward[0], o.gr_kind[0]) == (g["gamma_reward"][s], g["gr_kind"][s]), ctx
        assert (reward[0], rk[0]) == (g["reward"][s], g["reward_kind"][s]), ctx
        assert o.day[0] == g["day"][s], ctx
        assert np.array_equal(obs[0], g["obs"][s]), ctx
        if g["done"][s]:
            assert o.episode_return[0] == g["episode_return"][s], ctx
            s0, f = draws.draw()
            o.reset(stocks0=s0, factor=f)
            assert o.amount[0] == g["reset_amount"][nreset] and o.amount_kind[0] == g["reset_amount_kind"][nreset]
            assert np.array_equal(o.stocks[0], g["reset_stocks"][nreset])
            nreset += 1


# ---------------------------------------------------------------------------- A3
PF = sorted(glob.glob(os.path.join(GOLDEN, "portfolio_*.npz")))


@pytest.mark.parametrize("path", PF, ids=[os.path.basename(p)[:-4] for p in PF])
def test_portfolio_oracle_vs_reference(path):
    """np.exp is a black box that differs from libm by an ulp on a few percent of inputs (SURVEY.md
    §8c), so values are compared at 1e-12 (f64 actions) / 2e-6 (f32 actions, np.exp float32 is looser);
    day / done and the table-derived observation are exact."""
    g = np.load(path)
    acts = g["actions"]
    tol = 1e-12 if acts.dtype == np.float64 else 2e-6
    o = ora.PortfolioOracle(g["close"], g["cov"], g["tech"], 1, initial_amount=float(g["initial_amount"]))
    assert np.array_equal(o.obs()[0], g["obs0"])
    for s in range(acts.shape[0]):
        reward, flags, w, pr = o.step(acts[s][None, :], auto_reset=True)
        ctx = f"step {s}"
        assert bool(flags[0] & 1) == bool(g["done"][s]), ctx
        assert o.day[0] == g["day"][s], ctx
        assert abs(reward[0] - g["reward"][s]) <= tol * abs(g["reward"][s]), ctx
        assert abs(o.pv[0] - g["pv"][s]) <= tol * abs(g["pv"][s]), ctx
        if not g["done"][s]:
            assert np.allclose(w[0], g["weights"][s], rtol=tol, atol=0), ctx
            assert abs(pr[0] - g["pret"][s]) <= tol * max(abs(g["pret"][s]), 1e-3), ctx
        assert np.array_equal(o.obs()[0], g["obs"][s]), ctx


# ---------------------------------------------------------------------------- A4
CP = sorted(glob.glob(os.path.join(GOLDEN, "cashpen_*.npz")))


def cashpen_args_from_golden(g):
    bc, sc, hmax, disc, inc, use_t, thr, init, pen, patient = g["cfg"]
    cols = [str(c) for c in g["cols"]]
    info = np.stack([g[c] for c in cols], axis=2)  # [T, D, C] asset-major like get_date_vector
    if "hmax_vec" in g.files and g["hmax_vec"].size:
        hmax = g["hmax_vec"]  # per-asset array, dtype as the reference received it
    kw = dict(buy_cost_pct=bc, sell_cost_pct=sc, hmax=hmax, discrete_actions=bool(disc), shares_increment=int(inc),
              turbulence_threshold=(thr if use_t > 0 else None), initial_amount=init, cash_penalty_proportion=pen,
              patient=bool(patient))
    return g["close"], info, g["turbulence"], kw


@pytest.mark.parametrize("path", CP, ids=[os.path.basename(p)[:-4] for p in CP])
def test_cashpenalty_oracle_vs_reference(path):
    """np.dot is BLAS ddot (order unspecified) -> 1e-12 relative on values; flags/date exact."""
    g = np.load(path)
    close, info, turb, kw = cashpen_args_from_golden(g)
    o = ora.CashPenaltyOracle(close, info, turb, 1, **kw)
    np.testing.assert_allclose(o.obs()[0], g["obs0"], rtol=0, atol=0)
    acts = g["actions"]
    for s in range(acts.shape[0]):
        reward, flags = o.step(acts[s][None, :], auto_reset=True)
        ctx = f"step {s}"
        assert bool(flags[0] & ora.FLAG_DONE) == bool(g["done"][s]), ctx
        assert bool(flags[0] & ora.FLAG_LIQUIDATE) == bool(g["liq"][s]), ctx
        assert o.date_index[0] == g["date_index"][s], ctx
        np.testing.assert_allclose(reward[0], g["reward"][s], rtol=1e-12, atol=1e-18, err_msg=ctx)
        np.testing.assert_allclose(o.obs()[0], g["obs"][s], rtol=1e-12, atol=1e-9, err_msg=ctx)


# ---------------------------------------------------------------------------- sibling: CryptoEnv
CRYPTO = sorted(glob.glob(os.path.join(GOLDEN, "crypto_*.npz")))


@pytest.mark.parametrize("path", CRYPTO, ids=[os.path.basename(p)[:-4] for p in CRYPTO])
def test_crypto_oracle_vs_reference(path):
    g = np.load(path)
    o = ora.CryptoOracle(g["price_array"], g["tech_array"], 1, lookback=int(g["lookback"]),
                         initial_capital=float(g["initial_capital"]))
    assert np.array_equal(o.norm, g["action_norm_vector"])
    assert np.array_equal(o.obs()[0], g["obs0"])
    acts = g["actions"]
    for s in range(acts.shape[0]):
        obs, reward, flags = o.step(acts[s][None, :])
        ctx = f"step {s}"
        assert bool(flags[0] & 1) == bool(g["done"][s]) and o.time[0] == g["time"][s], ctx
        assert o.cash[0] == g["cash"][s] and np.array_equal(o.stocks[0], g["stocks"][s]), ctx
        assert o.total[0] == g["total"][s] and o.gamma_return[0] == g["gamma_return"][s] and reward[0] == g["reward"][s], ctx
        assert np.array_equal(obs[0], g["obs"][s]), ctx
        if g["done"][s]:
            assert o.episode_return[0] == g["episode_return"][s], ctx
            o.reset()


# ---------------------------------------------------------------------------- sibling: StockTradingEnvStopLoss
SL = sorted(glob.glob(os.path.join(GOLDEN, "stoploss_*.npz")))


def stoploss_args_from_golden(g):
    close, info, turb, kw = cashpen_args_from_golden(g)
    kw["stoploss_penalty"], kw["profit_loss_ratio"] = float(g["stoploss"][0]), float(g["stoploss"][1])
    return close, info, turb, kw


@pytest.mark.parametrize("path", SL, ids=[os.path.basename(p)[:-4] for p in SL])
def test_stoploss_oracle_vs_reference(path):
    g = np.load(path)
    close, info, turb, kw = stoploss_args_from_golden(g)
    o = ora.StopLossOracle(close, info, turb, 1, **kw)
    np.testing.assert_allclose(o.obs()[0], g["obs0"], rtol=0, atol=0)
    acts = g["actions"]
    for s in range(acts.shape[0]):
        reward, flags = o.step(acts[s][None, :], auto_reset=True)
        ctx = f"step {s}"
        assert bool(flags[0] & ora.FLAG_DONE) == bool(g["done"][s]), ctx
        assert bool(flags[0] & ora.FLAG_LIQUIDATE) == bool(g["liq"][s]), ctx
        assert o.date_index[0] == g["date_index"][s], ctx
        np.testing.assert_allclose(reward[0], g["reward"][s], rtol=1e-11, atol=1e-16, err_msg=ctx)
        np.testing.assert_allclose(o.obs()[0], g["obs"][s], rtol=1e-12, atol=1e-9, err_msg=ctx)
