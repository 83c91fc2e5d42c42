"""CPU-only tests: the C-ABI library loads and exports every declared symbol (no compute calls),
the host-side table builders, the optional-dependency shims, and the 2-rank sharding layer on gloo."""
import os
import re
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_library_exports_every_declared_symbol():
    from finrl_b200 import _cabi, build

    build.build()
    lib = _cabi.lib()  # resolves every name in SIGNATURES or raises
    header = open(os.path.join(ROOT, "include", "finrl_b200.h")).read()
    declared = set(re.findall(r"FRL_API\s+[\w\s\*]+?\b(frl_\w+)\s*\(", header))
    assert declared == set(_cabi.SIGNATURES), declared ^ set(_cabi.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.frl_abi_version() == _cabi.ABI_VERSION
    assert lib.frl_last_error() is not None


def test_abi_rejects_bad_arguments_without_a_gpu():
    """Argument validation happens on the host before any CUDA call: status + message, no crash."""
    import ctypes as C

    from finrl_b200 import _cabi

    lib = _cabi.lib()
    p = _cabi.TradingParams()
    p.n_envs, p.stock_dim, p.n_tech, p.n_days, p.obs_dim, p.env_stride = 4, 40, 1, 10, 121, 4
    rc = lib.frl_trading_step(C.byref(p), None, 0, None, None, None, 0, None, None)
    assert rc == -1 and b"stock_dim" in lib.frl_last_error()
    q = _cabi.CashPenaltyParams()
    q.n_envs, q.stock_dim, q.n_cols, q.n_days, q.obs_dim, q.env_stride = 1, 500, 5, 10, 3001, 1
    assert lib.frl_cashpenalty_observe(C.byref(q), None, None) == -1
    assert lib.frl_trading_step(None, None, 0, None, None, None, 0, None, None) == -1


def test_struct_layouts_match_the_header():
    """ctypes mirrors vs the C structs: compile a tiny C program printing sizeof/offsetof."""
    import ctypes as C
    import subprocess
    import tempfile

    from finrl_b200 import _cabi

    checks = {
        "frl_trading_params": (_cabi.TradingParams, ["n_envs", "hmax", "turbulence_threshold", "close", "cash", "episode", "asset_out", "obs_tmpl4"]),
        "frl_np_params": (_cabi.NpParams, ["gamma", "initial_capital", "obs_amount_floor", "price", "amount", "episode_return", "price_pitch"]),
        "frl_portfolio_params": (_cabi.PortfolioParams, ["initial_amount", "ret", "reward", "ret_out", "weights_out", "ret_pitch"]),
        "frl_crypto_params": (_cabi.CryptoParams, ["lookback", "env_stride", "initial_capital", "gamma", "price", "episode_return"]),
        "frl_stoploss_params": (_cabi.StopLossParams, ["patient", "env_stride", "buy_cost_pct", "stoploss_penalty", "min_profit_penalty", "close", "assets", "fresh", "sum_trades", "hmax_vec", "hmax_vec_f32"]),
        "frl_cashpenalty_params": (_cabi.CashPenaltyParams, ["patient", "env_stride", "buy_cost_pct", "cash_penalty_proportion", "close", "hold_alt", "sum_trades", "hmax_vec", "hmax_vec_f32"]),
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "finrl_b200.h"', "int main(){"]
    for st, (_, fields) in checks.items():
        lines.append(f'printf("%zu\\n", sizeof({st}));')
        for f in fields:
            lines.append(f'printf("%zu\\n", offsetof({st}, {f}));')
    lines.append("return 0;}")
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write("\n".join(lines))
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", os.path.join(d, "t")])
        vals = [int(x) for x in subprocess.check_output([os.path.join(d, "t")]).split()]
    it = iter(vals)
    for st, (cls, fields) in checks.items():
        assert next(it) == C.sizeof(cls), st
        for f in fields:
            assert next(it) == getattr(cls, f).offset, (st, f)


def test_frame_to_arrays_and_templates():
    from finrl_b200 import synthetic as syn
    from finrl_b200.tables import frame_to_arrays

    T, D, K = 12, 5, 3
    close, tech, turb = syn.make_tables(T, D, K, seed=2)
    df = syn.make_frame(close, tech, turb)
    c2, t2, r2 = frame_to_arrays(df, D, syn.INDICATORS[:K], "turbulence")
    assert np.array_equal(c2, close) and np.array_equal(t2, tech) and np.array_equal(r2, turb)
    with pytest.raises(ValueError):
        frame_to_arrays(df.iloc[:-1], D, syn.INDICATORS[:K], "turbulence")
    df_bad = df.copy()
    df_bad.index = np.arange(len(df_bad))
    with pytest.raises(ValueError):
        frame_to_arrays(df_bad, D, syn.INDICATORS[:K], "turbulence")


def test_cashpenalty_frame_ingest_matches_reference_order():
    from finrl_b200 import synthetic as syn
    from finrl_b200.cashpenalty import frame_to_cashpenalty_arrays

    T, D = 6, 4
    close, _, turb = syn.make_tables(T, D, 0, seed=3)
    o, h, l, v = syn.make_ohlv(close, 3)
    df = syn.make_frame(close, np.zeros((0, T, D)), turb, tech_names=[], extra_cols={"open": o, "high": h, "low": l, "volume": v})
    df = df.reset_index(drop=True).sample(frac=1.0, random_state=0)  # row order must not matter beyond asset order
    c2, info, t2, dates, assets = frame_to_cashpenalty_arrays(df, ["open", "close", "high", "low", "volume"])
    order = [syn.tickers(D).index(a) for a in assets]
    assert np.array_equal(c2, close[:, order]) and np.array_equal(t2, turb)
    assert np.array_equal(info[:, :, 0], o[:, order]) and np.array_equal(info[:, :, 4], v[:, order])


def test_no_cpu_fallback_without_cuda():
    """Constructing an env on a box without CUDA must fail loudly, not fall back."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from finrl_b200 import BatchedStockTradingEnv, EngineError, synthetic as syn

    close, tech, turb = syn.make_tables(8, 3, 1, seed=0)
    with pytest.raises((EngineError, RuntimeError, AssertionError)):
        BatchedStockTradingEnv(df=syn.make_frame(close, tech, turb), stock_dim=3, tech_indicator_list=syn.INDICATORS[:1], n_envs=2)
    with pytest.raises(EngineError):
        BatchedStockTradingEnv(df=syn.make_frame(close, tech, turb), stock_dim=3, tech_indicator_list=syn.INDICATORS[:1], device="cpu")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "finrl_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
            assert "liboracle" not in src, fn


def test_spaces_and_lazy_infos():
    from finrl_b200.spaces import Box
    from finrl_b200.vec_env import _LazyInfos

    b = Box(low=-1, high=1, shape=(7,))
    assert tuple(b.shape) == (7,) and b.sample().shape == (7,)
    infos = _LazyInfos(5, {3: np.ones(2)})
    assert len(infos) == 5 and infos[0] == {} and np.array_equal(infos[3]["terminal_observation"], np.ones(2))
    assert [bool(i) for i in infos] == [False, False, False, True, False]


def test_gym_vec_env_is_dummy_vec_env_protocol():
    """GymVecEnv = SB3's DummyVecEnv over env OBJECTS: sequential stepping, float32 buffer, auto-reset with
    terminal_observation, env_method/get_attr/render reaching the objects themselves (one entry per env)."""
    from finrl_b200.spaces import Box
    from finrl_b200.vec_env import GymVecEnv

    class Counter:
        def __init__(self, horizon):
            self.observation_space, self.action_space = Box(-np.inf, np.inf, (3,)), Box(-1, 1, (2,))
            self.horizon, self.t, self.log = horizon, 0, []

        def reset(self):
            self.t, self.log = 0, []
            return [0.0, 0.0, float(self.horizon)]

        def step(self, a):
            self.t += 1
            self.log.append(float(a[0]))
            return [float(self.t), float(a[1]), float(self.horizon)], 0.5 * self.t, self.t >= self.horizon, {}

        def render(self, mode="human"):
            return ["state", self.t]

        def memory(self, scale=1):
            return [scale * x for x in self.log]

    envs = [Counter(3), Counter(5)]
    vec = GymVecEnv([lambda e=e: e for e in envs])
    assert vec.num_envs == 2 and vec.envs[0] is envs[0]
    obs = vec.reset()
    assert obs.dtype == np.float32 and obs.shape == (2, 3)
    for s in range(3):
        obs, rews, dones, infos = vec.step(np.full((2, 2), s + 1.0, dtype=np.float32))
    assert dones.tolist() == [True, False] and rews.dtype == np.float32 and rews.tolist() == [1.5, 1.5]
    assert infos[0]["terminal_observation"] == [3.0, 3.0, 3.0] and infos[1] == {}
    assert obs[0].tolist() == [0.0, 0.0, 3.0] and obs[1].tolist() == [3.0, 3.0, 5.0]  # env 0 was reset at once
    assert vec.env_method("memory", scale=2) == [[], [2.0, 4.0, 6.0]]
    assert vec.env_method(method_name="memory", indices=[1]) == [[1.0, 2.0, 3.0]]
    assert vec.get_attr("horizon") == [3, 5] and vec.get_attr("t", indices=1) == [3]
    vec.set_attr("horizon", 9, indices=[0])
    assert envs[0].horizon == 9 and envs[1].horizon == 5
    assert vec.render() == [["state", 0], ["state", 3]]
    assert GymVecEnv([lambda: envs[1]]).render() == ["state", 3]  # one env: its own render(), as DummyVecEnv does


def test_tables_refuse_nonpositive_prices_of_tradable_stocks():
    from finrl_b200 import synthetic as syn
    from finrl_b200.tables import TradingTables

    close, tech, turb = syn.make_tables(6, 4, 2, seed=1)
    bad = close.copy()
    bad[3, 2] = 0.0
    with pytest.raises(ValueError, match=r"close\[3, 2\]"):
        TradingTables.from_arrays(bad, tech, turb, "cpu")
    tech2 = tech.copy()
    tech2[0, 3, 2] = 1.0  # the reference's disable flag: that stock is never traded that day
    assert TradingTables.from_arrays(bad, tech2, turb, "cpu").n_days == 6
    assert TradingTables.from_arrays(bad, tech, turb, "cpu", allow_nonpositive_close=True).stock_dim == 4


def test_expand_obs_host_matches_numpy():
    """frl_expand_obs_host (host threads, no GPU): factored rows -> dense rows, against plain numpy indexing."""
    from finrl_b200.trading import FactoredObs

    rng = np.random.default_rng(0)
    T, D, K, N = 17, 5, 3, 20_003
    O = 1 + 2 * D + K * D
    tmpl = rng.normal(size=(T, O)).astype(np.float32)
    env_part = rng.normal(size=(N, 1 + D)).astype(np.float32)
    sday = rng.integers(0, T, size=N).astype(np.int32)
    fo = FactoredObs(env_part, sday, tmpl, D)
    want = tmpl[sday].copy()
    want[:, 0] = env_part[:, 0]
    want[:, 1 + D : 1 + 2 * D] = env_part[:, 1:]
    for threads in (1, 3, 0):
        out = np.full((N, O), np.nan, dtype=np.float32)
        assert fo.dense(out=out, n_threads=threads) is out and np.array_equal(out, want)
    assert np.array_equal(fo[123], want[123]) and fo.shape == (N, O) and len(fo) == N
    sday[77] = T  # out of range: reported, not a crash
    from finrl_b200 import EngineError

    with pytest.raises(EngineError):
        fo.dense()


def test_shard_range_partitions_exactly():
    from finrl_b200.dist import shard_range

    for n in (1, 7, 8, 1 << 20, (1 << 20) + 5):
        for w in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def _gloo_worker(rank, world, port, N, K, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from finrl_b200 import synthetic as syn
    from finrl_b200.dist import StatsExchange, allreduce_stats, init_from_env, shard_range, summarize
    from oracle import oracle as ora

    r, w, _ = init_from_env(backend="gloo")
    start, count = shard_range(N, r, w)
    close, tech, turb = syn.make_tables(30, 30, 2, seed=0)  # tables replicated on every rank
    o = ora.TradingOracle(close, tech, turb, count, hmax=100, initial_amount=2e5, turbulence_threshold=90)
    acts = syn.make_actions((K, N, 30), seed=1)[:, start : start + count]  # this rank's env-index slice
    stats = torch.zeros(8, dtype=torch.float64)
    ex = StatsExchange("cpu")  # host-logic leg of the exchange: snapshots + all-reduce (gloo), flushed mid-run
    assert ex.mode == "collective" and ex.world == w
    for k in range(K):
        _, rew, fl = o.step(acts[k], auto_reset=True, want_obs=False)
        part = torch.tensor([rew.sum(), (rew ** 2).sum(), float((fl & 1).sum()), 0, 0, 0, count, 0], dtype=torch.float64)
        stats += part
        ex.sum += part  # what a kernel epilogue does on the device
        if k % 8 == 7:
            ex.flush()
    allreduce_stats(stats)  # the plain collective
    totals = ex.totals()
    np.savez(os.path.join(out_dir, f"rank{r}.npz"), cash=o.cash, hold=o.hold, stats=stats.numpy(), start=start,
             exchanged=np.asarray(totals))
    s = summarize(stats)
    assert s["env_steps"] == N * K
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_equals_single_process(tmp_path):
    """world_size 2 on gloo: env-index sharding with replicated tables reproduces the unsharded run
    env for env, and the all-reduced statistics equal the global sums."""
    import torch.multiprocessing as mp

    from finrl_b200 import synthetic as syn
    from oracle import oracle as ora

    N, K, world = 101, 35, 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(world, port, N, K, str(tmp_path)), nprocs=world, join=True)
    close, tech, turb = syn.make_tables(30, 30, 2, seed=0)
    o = ora.TradingOracle(close, tech, turb, N, hmax=100, initial_amount=2e5, turbulence_threshold=90)
    acts = syn.make_actions((K, N, 30), seed=1)
    tot = np.zeros(8)
    for k in range(K):
        _, rew, fl = o.step(acts[k], auto_reset=True, want_obs=False)
        tot[0] += rew.sum(); tot[1] += (rew ** 2).sum(); tot[2] += (fl & 1).sum(); tot[6] += N
    parts = [np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(world)]
    assert np.array_equal(np.concatenate([p["cash"] for p in parts]), o.cash)
    assert np.array_equal(np.concatenate([p["hold"] for p in parts]), o.hold)
    for p in parts:
        np.testing.assert_allclose(p["stats"], tot, rtol=1e-12)
        assert p["stats"][2] == tot[2] and p["stats"][6] == tot[6]
        np.testing.assert_allclose(p["exchanged"], tot, rtol=1e-12)  # StatsExchange (flushed every 8 steps) agrees
        assert p["exchanged"][2] == tot[2] and p["exchanged"][6] == tot[6]


def test_missing_library_fails_loudly(tmp_path):
    """No silent fallback when the CUDA extension is absent: the first use raises EngineError."""
    import subprocess

    code = ("import sys; sys.path.insert(0, %r)\n"
            "from finrl_b200 import _cabi\n"
            "try:\n    _cabi.lib()\nexcept _cabi.EngineError as e:\n    print('RAISED', 'no CPU fallback' in str(e))\n" % ROOT)
    env = dict(os.environ, FINRL_B200_LIB=str(tmp_path / "missing.so"))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert "RAISED True" in out.stdout, out.stdout + out.stderr


def test_register_network_decomposition_equals_the_full_network():
    """trading_wide.cu's compiled-in NASDAQ-100 path runs np.argsort's 128-slot network as 32-slot networks on the four
    quarters, the block-size-64 level on each half, the 128 flip stage between the halves and its half-cleaners on each
    half (macros generated by csrc/gen_sort_network.py).  Same permutation as the full network, ties included; and the
    committed sort_network.inc is what the generator emits."""
    import contextlib
    import io
    import os
    import random
    import runpy
    import sys

    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "finrl_b200", "csrc")
    sys.path.insert(0, csrc)
    try:
        from gen_sort_network import merge, network
    finally:
        sys.path.remove(csrc)

    def run(pairs, keys, off=0):
        for lo, hi in pairs:
            if keys[off + lo][0] > keys[off + hi][0]:  # strict: ties keep their slots
                keys[off + lo], keys[off + hi] = keys[off + hi], keys[off + lo]

    rnd = random.Random(7)
    for trial in range(60):
        D = rnd.choice([65, 100, 100, 127, 128])
        vals = [(rnd.randint(-4, 4), i) for i in range(D)] + [(10**9, i) for i in range(D, 128)]
        a = list(vals)
        run(network(128), a)
        b = list(vals)
        for g in range(0, 128, 32):
            run(network(32), b, g)
        for g in (0, 64):
            run(merge(64, True), b, g)
        run([(i, 127 - i) for i in range(64)], b)
        for g in (0, 64):
            run(merge(64, False), b, g)
        assert a == b
        assert all(k[0] == 10**9 for k in a[D:])  # pads never move
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        runpy.run_path(os.path.join(csrc, "gen_sort_network.py"), run_name="__main__")
    assert buf.getvalue() == open(os.path.join(csrc, "sort_network.inc")).read()
