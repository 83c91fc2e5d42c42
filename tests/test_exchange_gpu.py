"""GPU: the one-sided statistics exchange fused into the step kernels (frl_stats_block,
finrl_b200.dist.StatsExchange).  One GPU: launches alternate between the block's two accumulators and the first
thread block of every launch moves the previous launch's sums into ``total`` (n_peers = 1: the own block) — for
every env kind, equal to what a plain engine accumulates.  Two
GPUs (skipped on a 1-GPU box): two processes push into each other's peer-mapped blocks over NVLink and both
read the global sums, equal to an NCCL all-reduce of the plain vectors."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _engines():
    """(name, factory(n_envs) -> engine, action maker) for every env kind, small shapes."""
    from finrl_b200 import (BatchedNpStockTradingEnv, BatchedStockPortfolioEnv, BatchedStockTradingEnv,
                            BatchedStockTradingEnvCashpenalty, BatchedStockTradingEnvStopLoss, CashPenaltyTables,
                            PortfolioTables, TradingTables, synthetic as syn)

    close, tech, turb = syn.make_tables(24, 30, 4, seed=3)
    tkw = dict(hmax=100, initial_amount=2e5, turbulence_threshold=80)
    tt = TradingTables.from_arrays(close, tech, turb, "cuda")
    pa, ta, tu = syn.make_np_arrays(close, tech, turb)
    cl2, tech2, _ = syn.make_tables(40 + 20, 30, 4, seed=4)
    cov, first = syn.make_cov_table(cl2, 40)
    pt = PortfolioTables.from_arrays(cl2[first:], cov, tech2[:, first:], "cuda")
    cl3, _, turb3 = syn.make_tables(24, 12, 0, seed=5)
    o, h, l, v = syn.make_ohlv(cl3, 5)
    ct = CashPenaltyTables.from_arrays(cl3, np.stack([o, cl3, h, l, v], axis=2), turb3, "cuda")
    return [
        ("trading", lambda n: BatchedStockTradingEnv(tables=tt, n_envs=n, **tkw), 30, (-1, 1)),
        ("np", lambda n: BatchedNpStockTradingEnv({"price_array": pa, "tech_array": ta, "turbulence_array": tu,
                                                    "if_train": False}, n_envs=n), 30, (-1, 1)),
        ("portfolio", lambda n: BatchedStockPortfolioEnv(tables=pt, n_envs=n), 30, (0, 1)),
        ("cashpenalty", lambda n: BatchedStockTradingEnvCashpenalty(tables=ct, n_envs=n, random_start=False, hmax=5000,
                                                                    turbulence_threshold=80), 12, (-1, 1)),
        ("stoploss", lambda n: BatchedStockTradingEnvStopLoss(tables=ct, n_envs=n, random_start=False, hmax=5000,
                                                              turbulence_threshold=80), 12, (-1, 1)),
    ]


@pytest.mark.parametrize("n_envs", [1000, 20000])  # 8-lanes-per-env and thread-per-env trading kernels
def test_epilogue_exchange_equals_plain_accumulation(n_envs):
    from finrl_b200 import synthetic as syn
    from finrl_b200.dist import StatsExchange

    for name, make, D, (lo, hi) in _engines():
        plain, fused = make(n_envs), make(n_envs)
        ex = StatsExchange("cuda", mode="p2p")
        assert ex.mode == "p2p", getattr(ex, "fallback_reason", None)
        ex.attach(fused)
        K = 30
        acts = torch.from_numpy(syn.make_actions((K, n_envs, D), seed=7, low=lo, high=hi)).cuda()
        for k in range(0, K, 10):  # three fused 10-step rollouts
            plain.rollout(acts[k : k + 10], obs_mode="none", auto_reset=True, accumulate_stats=True)
            fused.rollout(acts[k : k + 10], obs_mode="none", auto_reset=True, accumulate_stats=True)
            torch.cuda.synchronize()
            # the launch that just ran pushed its predecessor: only ONE accumulator holds anything now
            acc = ex._acc.view(2, 8).abs().sum(dim=1).cpu().numpy()
            assert (acc != 0).sum() == 1 and (k == 0) == (ex.total.abs().sum().item() == 0.0), name
        plain.step(acts[0], auto_reset=True, accumulate_stats=True)
        fused.step(acts[0], auto_reset=True, accumulate_stats=True)
        want = plain.stats.cpu().numpy()
        got = np.asarray(ex.totals())
        # same addends, different atomic order: exact for the counters, 1e-12 for the fp sums
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-9, err_msg=name)
        assert got[2] == want[2] and got[5] == want[5] and got[6] == want[6] == (K + 1) * n_envs, name
        assert ex._acc.abs().sum().item() == 0.0, name  # totals() flushed the last launch
        assert list(fused.read_stats().values()) == list(got), name
        assert np.asarray(ex.totals(reset=True)).tolist() == got.tolist() and sum(ex.totals()) == 0.0
        ex.close()


def _two_gpu_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from finrl_b200 import BatchedStockTradingEnv, TradingTables, synthetic as syn
    from finrl_b200.dist import StatsExchange, allreduce_stats, init_from_env, shard_range

    r, w, local = init_from_env()
    dev = torch.device("cuda", local)
    N, K = 50_001, 12
    start, count = shard_range(N, r, w)
    close, tech, turb = syn.make_tables(20, 30, 8, seed=0)
    kw = dict(hmax=100, initial_amount=2e5, turbulence_threshold=80)
    tables = TradingTables.from_arrays(close, tech, turb, dev)
    plain = BatchedStockTradingEnv(tables=tables, n_envs=count, device=dev, **kw)
    fused = BatchedStockTradingEnv(tables=tables, n_envs=count, device=dev, **kw)
    ex = StatsExchange(dev)
    ex.attach(fused)
    acts = torch.from_numpy(syn.make_actions((K, N, 30), seed=1)[:, start : start + count].copy()).to(dev)
    for k in range(K):
        plain.step(acts[k], auto_reset=True, accumulate_stats=True)
        fused.step(acts[k], auto_reset=True, accumulate_stats=True)
        if r == 1 and k == 5:
            torch.cuda.synchronize()  # a rank that falls behind: nobody waits for it inside the loop
    ref = plain.stats.clone()
    allreduce_stats(ref)
    got = ex.totals()
    np.savez(os.path.join(out_dir, f"rank{r}.npz"), ref=ref.cpu().numpy(), got=np.asarray(got), mode=np.array(ex.mode),
             reason=np.array(str(getattr(ex, "fallback_reason", None))))
    ex.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_gpu_one_sided_exchange_equals_nccl(tmp_path):
    import torch.multiprocessing as mp

    port = 29600 + (os.getpid() % 2000)
    mp.spawn(_two_gpu_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    parts = [np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(2)]
    for p in parts:
        assert str(p["mode"]) == "p2p", str(p["reason"])
        np.testing.assert_allclose(p["got"], p["ref"], rtol=1e-12, atol=1e-9)
        assert p["got"][6] == p["ref"][6] == 12 * 50_001 and p["got"][2] == p["ref"][2]
    np.testing.assert_allclose(parts[0]["got"], parts[1]["got"], rtol=1e-12)  # both ranks hold the global sums
