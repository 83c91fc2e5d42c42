#!/usr/bin/env python
"""bench.py — env-steps/s of the FinRL env step path on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload W] [--envs E]

Default workload (the one BASELINE.json's metric is quoted on): StockTradingEnv, DOW-30 shape (D=30
stocks, K=8 indicators, T=2500 days, O=301), E = 1,048,576 envs PER GPU, one ``step()`` of every env per
bench step, float32 actions U(-1,1) distinct per env and step, turbulence threshold 99, auto-reset at
episode ends.  A "step" is one pass of the hot path over the whole batch.  Other workloads (the other
BASELINE configs) are selected with --workload; they are not the driver's bench line.

* value        env-steps/s with inputs resident in HBM (CUDA events, max over ranks, whole job)
* e2e          same metric through the numpy-in / numpy-out VecEnv-style call: pinned HOST actions
               -> H2D -> step -> D2H of obs + reward + flags, every step inside the timed region
* roofline     algorithmic bytes per launch (SURVEY.md §8d figure x envs) / measured kernel time
* cpu_baseline the CPU oracle port (oracle/oracle.c, OpenMP over all host threads) on a bounded sample

--impl reference times that CPU port as the reference arm (the reference itself is pure Python and
cannot travel to the GPU box; its survey-time single-core numbers are in BASELINE.md).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# BASELINE.json's metric, verbatim; `value` is the env-steps/sec part, `roofline` carries the GB/s-vs-peak part
METRIC = "env-steps/sec at 1/2/4/8 B200 (DOW-30, 1M envs); achieved HBM GB/s vs peak"
UNIT = "env-steps/s"
T_DAYS = 2500


# ------------------------------------------------------------------------------------------------
# workloads: how to build the engine, its CPU port, and the algorithmic bytes per env-step
# ------------------------------------------------------------------------------------------------
class Workload:
    name = ""
    D = 30
    act_low, act_high = -1.0, 1.0
    act_dtype = "float32"
    bytes_per_env_step = 0
    kernel = ""
    default_envs = 1 << 20
    rollout_k = 1            # env steps fused per launch

    def describe(self, envs):
        raise NotImplementedError

    def make_env(self, dev, envs):
        raise NotImplementedError

    def make_cpu(self, envs):
        raise NotImplementedError

    def obs_numel(self, env):
        return int(env._obs.shape[1]) if getattr(env, "_obs", None) is not None else 0


class TradingStep(Workload):
    name = "trading_step"
    K_TECH = 8
    bytes_per_env_step = 1621  # "Config 2' single step (DOW-30)"
    kernel = "trading_rollout_kernel<32,30,float,4>"
    KW = dict(hmax=100, initial_amount=1_000_000, buy_cost_pct=0.001, sell_cost_pct=0.001, reward_scaling=1e-4,
              turbulence_threshold=99)

    def describe(self, envs):
        return f"StockTradingEnv DOW-30 single step, {envs} envs/GPU, D=30 K=8 T=2500 O=301, f32 actions, f32 obs"

    def tables(self):
        from finrl_b200 import synthetic as syn

        return syn.make_tables(T_DAYS, self.D, self.K_TECH, seed=0)

    def make_env(self, dev, envs):
        from finrl_b200 import BatchedStockTradingEnv, TradingTables

        return BatchedStockTradingEnv(tables=TradingTables.from_arrays(*self.tables(), dev), n_envs=envs, device=dev, **self.KW)

    def make_cpu(self, envs):
        from oracle import oracle as ora

        o = ora.TradingOracle(*self.tables(), envs, **self.KW)
        return lambda a: o.step(a, auto_reset=True)


class TradingRollout(TradingStep):
    name = "trading_rollout"
    default_envs = 4096
    rollout_k = 64
    bytes_per_env_step = 152.3  # config 2: K=64 fused, obs_mode=last
    kernel = "trading_small_kernel<4,float,2,30> (n_envs <= 8192), else trading_rollout_kernel<32,30,float,4>"

    def describe(self, envs):
        return f"StockTradingEnv DOW-30 fused K=64 rollout (obs after the last step), {envs} envs/GPU, D=30 K=8 T=2500"


class TradingNas100Step(TradingStep):
    name = "trading_nas100_step"
    D = 100
    K_TECH = 2
    default_envs = 1 << 18
    # actions 400 + state 2*(8 cash + 400 hold + 4 day + 4 sday + 8 cost + 4 trades + 8 reward) + obs 4*401 + reward 8 + done 1
    bytes_per_env_step = 2885
    kernel = "trading_wide_kernel<float,2,100> (stock count compiled in)"

    def describe(self, envs):
        return f"StockTradingEnv at NASDAQ-100 size, {envs} envs/GPU, D=100 K=2 T=2500 O=401, f32 actions, f32 obs"


class NpStep(Workload):
    name = "np_step"
    bytes_per_env_step = 2015  # config 3
    kernel = "np_rollout_kernel<32,30,float,4> (all-float64 instantiation of the step body in steady state)"

    def describe(self, envs):
        return f"env_stocktrading_np (ElegantRL) single step, {envs} envs/GPU, D=30 K=8 T=2500 O=333, turbulence_thresh 99"

    def arrays(self):
        from finrl_b200 import synthetic as syn

        return syn.make_np_arrays(*syn.make_tables(T_DAYS, self.D, 8, seed=0))

    def make_env(self, dev, envs):
        from finrl_b200 import BatchedNpStockTradingEnv

        pa, ta, tu = self.arrays()
        return BatchedNpStockTradingEnv({"price_array": pa, "tech_array": ta, "turbulence_array": tu, "if_train": False},
                                        n_envs=envs, device=dev)

    def make_cpu(self, envs):
        from oracle import oracle as ora

        o = ora.NpTradingOracle(*self.arrays(), envs)

        def step(a):
            _, _, _, fl = o.step(a)
            if fl[0] & 1:
                o.reset()

        return step


class NpNas100Step(NpStep):
    name = "np_nas100_step"
    D = 100
    K_TECH = 2
    default_envs = 1 << 18
    # actions 400 + state 2*(8 amount + 1 kind + 400 stocks + 400 cool-down + 4 day + 8 total + 8 gamma_reward + 8 init_total)
    # + obs 4*503 + reward 8 + done 1
    bytes_per_env_step = 4095
    kernel = "np_wide_kernel<float,4,bulk,100> (bulk-staged, stock count compiled in)"

    def describe(self, envs):
        return (f"env_stocktrading_np at NASDAQ-100 size (streaming kernel), {envs} envs/GPU, D=100 K=2 T=2500 O=503, "
                "turbulence_thresh 99")

    def arrays(self):
        from finrl_b200 import synthetic as syn

        return syn.make_np_arrays(*syn.make_tables(T_DAYS, self.D, self.K_TECH, seed=0))


class PortfolioStep(Workload):
    name = "portfolio_step"
    act_low, act_high = 0.0, 1.0
    act_dtype = "float64"
    default_envs = 262144
    bytes_per_env_step = 4353  # config 4 with the (34,30) f32 observation materialised per env
    kernel = "portfolio_rollout_kernel<32,double,2>"

    def describe(self, envs):
        return (f"StockPortfolioEnv softmax allocation, {envs} envs/GPU, D=30 K=4, 252-day covariance state "
                "materialised per env (f32), f64 actions")

    def arrays(self):
        from finrl_b200 import synthetic as syn

        close, tech, _ = syn.make_tables(252 + 600, self.D, 4, seed=0)
        cov, first = syn.make_cov_table(close, 252)
        return close[first:], cov, tech[:, first:]

    def make_env(self, dev, envs):
        from finrl_b200 import BatchedStockPortfolioEnv, PortfolioTables

        env = BatchedStockPortfolioEnv(tables=PortfolioTables.from_arrays(*self.arrays(), dev), n_envs=envs, device=dev)
        env._obs_buf()
        return env

    def make_cpu(self, envs):
        from oracle import oracle as ora

        o = ora.PortfolioOracle(*self.arrays(), envs)

        def step(a):
            o.step(a, auto_reset=True)
            o.obs()

        return step


class CashPenaltyStep(Workload):
    name = "cashpenalty_step"
    D = 100
    default_envs = 1 << 18
    bytes_per_env_step = 4485  # config 5
    kernel = "cashpenalty_rollout_kernel<float,4,false,true> (bulk-staged)"

    def describe(self, envs):
        return f"StockTradingEnvCashpenalty NASDAQ-100 shape, {envs} envs/GPU, D=100 C=5 T=5000 O=601, random_start=False"

    def arrays(self):
        from finrl_b200 import synthetic as syn

        close, _, turb = syn.make_tables(5000, self.D, 0, seed=0)
        o, h, l, v = syn.make_ohlv(close, 0)
        return close, np.stack([o, close, h, l, v], axis=2), turb

    def make_env(self, dev, envs):
        from finrl_b200 import BatchedStockTradingEnvCashpenalty, CashPenaltyTables

        return BatchedStockTradingEnvCashpenalty(tables=CashPenaltyTables.from_arrays(*self.arrays(), dev), n_envs=envs,
                                                 device=dev, random_start=False, turbulence_threshold=99)

    def make_cpu(self, envs):
        from oracle import oracle as ora

        close, info, turb = self.arrays()
        o = ora.CashPenaltyOracle(close, info, turb, envs, turbulence_threshold=99)

        def step(a):
            o.step(a, auto_reset=True)
            o.obs()

        return step


class StopLossStep(CashPenaltyStep):
    name = "stoploss_step"
    # cash-penalty traffic + five more fp64 per-asset arrays read and written (previous holdings, average buy
    # price, buy counts, closing / profit-sell diffs): 2404 obs + 400 actions + 12 x 800 state + ~60 scalars
    bytes_per_env_step = 12464
    kernel = "stoploss_rollout_kernel<float,4>"

    def describe(self, envs):
        return f"StockTradingEnvStopLoss NASDAQ-100 shape, {envs} envs/GPU, D=100 C=5 T=5000 O=601, random_start=False"

    def make_env(self, dev, envs):
        from finrl_b200 import BatchedStockTradingEnvStopLoss, CashPenaltyTables

        return BatchedStockTradingEnvStopLoss(tables=CashPenaltyTables.from_arrays(*self.arrays(), dev), n_envs=envs,
                                              device=dev, random_start=False, turbulence_threshold=99)

    def make_cpu(self, envs):
        from oracle import oracle as ora

        close, info, turb = self.arrays()
        o = ora.StopLossOracle(close, info, turb, envs, turbulence_threshold=99)

        def step(a):
            o.step(a, auto_reset=True)
            o.obs()

        return step


WORKLOADS = {w.name: w for w in (TradingStep, TradingRollout, TradingNas100Step, NpStep, NpNas100Step, PortfolioStep, CashPenaltyStep, StopLossStep)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock + throttle reasons of one GPU during the timed region (pynvml)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def _cpu_actions(wl, envs, n=4):
    from finrl_b200 import synthetic as syn

    return [syn.make_actions((envs, wl.D), seed=100 + i, low=wl.act_low, high=wl.act_high, dtype=np.dtype(wl.act_dtype))
            for i in range(n)]


def cpu_port_rate(wl, budget_s=12.0, envs=65536):  # callers pass the GPU arm's full per-GPU batch
    """Time the CPU oracle port on a bounded sample of the same workload.  Returns (env-steps/s,
    threads, description).  This is the one place bench.py executes oracle/."""
    from oracle import oracle as ora

    try:  # the GPU arm pins the process next to its GPU; the CPU baseline gets every host core back
        os.sched_setaffinity(0, range(os.cpu_count() or 1))
    except OSError:
        pass
    threads = ora.set_threads(os.cpu_count() or 1)  # all host threads, even under torchrun's OMP_NUM_THREADS=1
    step = wl.make_cpu(envs)
    pool = _cpu_actions(wl, envs)
    for i in range(3):
        step(pool[i % 4])
    steps, t0 = 0, time.perf_counter()
    while True:
        step(pool[steps % 4])
        steps += 1
        el = time.perf_counter() - t0
        if el >= budget_s:
            break
    return envs * steps / el, threads, f"{envs} envs x {steps} steps ({el:.1f} s), obs written, OpenMP static over envs"


def workload_config(wl, envs, world):
    """The `config` object both arms print: it names the workload and nothing arm-specific."""
    big = wl.bytes_per_env_step * envs * wl.rollout_k > 2.5e8
    return {
        "workload": wl.describe(envs), "envs_per_gpu": envs, "env_steps_per_launch": envs * wl.rollout_k, "auto_reset": True,
        "l2": ("per-step traffic exceeds the 126 MB L2 and consecutive steps use different action buffers; no explicit flush")
              if big else "working set fits in L2 (latency-bound configuration); the roofline is quoted on the 1M-env workload",
    }


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as ora

    threads = ora.set_threads(os.cpu_count() or 1)  # all host threads, even under torchrun's OMP_NUM_THREADS=1
    envs = args.ref_envs or args.envs  # by default the FULL per-GPU batch of the GPU arm (same config)
    step = wl.make_cpu(envs)
    pool = _cpu_actions(wl, envs)
    for i in range(args.warmup):
        step(pool[i % 4])
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(pool[i % 4])
    el = time.perf_counter() - t0
    v = envs * args.steps * wl.rollout_k / el
    sample = (f"{envs} envs per step" + (" (the full per-GPU batch of the GPU arm)" if envs == args.envs else
              f" (bounded sample of the {args.envs} envs/GPU workload)") + ", obs written")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(wl, args.envs, args.gpus),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "arm": {"cpu_envs_per_step": envs, "host_threads": threads},
        "python_reference": PY_REFERENCE,
        "note": "the reference is pure Python (pandas/numpy) and cannot travel to the GPU box; this arm is its C "
                "restatement (oracle/oracle.c, pinned to the reference by tests/golden) on all host threads, on ONE GPU's "
                "batch. `python_reference` holds the rates of the unmodified Python classes measured in the build container.",
    }))


# the unmodified Python reference, measured in the build container (tools/time_reference_cpu.py ->
# profiles/r02_reference_cpu_timing.txt); it cannot run on the GPU box (no /root/reference there)
PY_REFERENCE = {
    "where": "build container, 1 core per env", "unit": UNIT,
    "StockTradingEnv_single_env": 159.0, "StockTradingEnv_DummyVecEnv_x8": None,
    "env_stocktrading_np_single_env": 5400.0, "StockPortfolioEnv_single_env": 930.0, "Cashpenalty_single_env": 4.0,
}
_pyref = os.path.join(ROOT, "profiles", "python_reference_rates.json")
if os.path.exists(_pyref):
    try:
        PY_REFERENCE.update(json.load(open(_pyref)))
    except Exception:
        pass


def time_device(env, wl, pool, steps, warmup, world, ex, dist, torch, sampler=None):
    """W warm-up launches, then K timed ones bracketed by CUDA events on the launching stream.  The ranks'
    GPU timelines are aligned by a DEVICE-side collective enqueued right before the start event (a host
    barrier leaves them up to a millisecond apart); nothing inside the timed loop waits for a peer."""
    KR = wl.rollout_k
    ev = []

    def one(i):
        if KR == 1:  # env.kernel_events: CUDA events recorded right around the C-ABI launch, on the launching stream
            env.step(pool[i % len(pool)], auto_reset=True, want_obs=True, accumulate_stats=True, want_done=False)
            return None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        env.rollout(pool[i % len(pool)], obs_mode="last", auto_reset=True, accumulate_stats=True)
        e1.record()
        return e0, e1

    for i in range(warmup):
        one(i)
    ex.flush()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    if sampler is not None:
        sampler.start()
    launches0 = env.launches
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tick = torch.zeros(1, device=env.device)
    if world > 1:
        dist.all_reduce(tick)  # completes on all GPUs within microseconds of each other
    if KR == 1:
        env.kernel_events = ev
    t_start.record()
    for i in range(steps):
        e = one(i)
        if e is not None:
            ev.append(e)
        if (i + 1) % 16 == 0:
            ex.flush()  # no-op with the fused peer exchange; side-stream all-reduce in the NCCL fallback
    t_end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    clocks = sampler.stop() if sampler is not None else None
    el = torch.tensor([t_start.elapsed_time(t_end)], dtype=torch.float64, device=env.device)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    env.kernel_events = None
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    return float(el.item()), kernel_ms, env.launches - launches0, clocks


def make_pool(wl, N, dev, torch, rank, n=3):
    tdt = torch.float32 if wl.act_dtype == "float32" else torch.float64
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)  # every rank owns different envs, hence different actions
    shape = (N, wl.D) if wl.rollout_k == 1 else (wl.rollout_k, N, wl.D)
    span = wl.act_high - wl.act_low
    return [(torch.rand(shape, generator=g, device=dev, dtype=tdt) * span + wl.act_low) for _ in range(n)], shape, tdt


def run_ours(args, wl):
    import torch
    import torch.distributed as dist

    from finrl_b200.dist import StatsExchange, init_from_env, shard_range

    all_cpus = os.sched_getaffinity(0)
    rank, world, local = init_from_env()  # binds the GPU and its NUMA-local CPUs, creates the NCCL group
    numa_cpus = len(os.sched_getaffinity(0))
    dev = torch.device("cuda", local)
    start, N = shard_range(args.envs * world, rank, world)  # env-index sharding: this rank owns [start, start+N)
    KR = wl.rollout_k
    env = wl.make_env(dev, N)  # tables replicated per GPU
    ex = StatsExchange(dev)    # the only exchange on the path: 64 B of statistics, pushed by the kernels' epilogue
    ex.attach(env)
    pool, shape, tdt = make_pool(wl, N, dev, torch, rank)
    POOL = len(pool)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    elapsed_ms, kernel_ms, launches, clocks = time_device(env, wl, pool, args.steps, args.warmup, world, ex, dist, torch,
                                                          ClockSampler(local))
    value = N * world * args.steps * KR / (elapsed_ms * 1e-3)
    totals = ex.totals()

    # ---- end to end through the host-buffer call: pinned host actions in, host obs/reward/flags out -------
    e2e = {}
    e2e_steps = max(3, min(args.steps, args.e2e_steps))
    h_act_pool = [torch.empty(shape, dtype=tdt).pin_memory() for _ in range(POOL)]
    for h, d in zip(h_act_pool, pool):
        h.copy_(d)
    use_pipelined = hasattr(env, "step_host") and KR == 1 and not args.no_pipeline

    def time_e2e(fn):
        for i in range(5):
            fn(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            fn(i)
        barrier()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return N * world * e2e_steps * KR / float(t.item())

    if use_pipelined:
        _, h_obs, h_rew, h_flag = env.make_host_buffers("dense", tdt)
        h2d = h_act_pool[0].numel() * h_act_pool[0].element_size()
        d2h_dense = h_obs.numel() * 4 + h_rew.numel() * 8 + h_flag.numel()
        d2h_fact = N * (4 * (1 + wl.D) + 4) + h_rew.numel() * 8 + h_flag.numel()
        # the call a host-resident user makes: dense obs[N,O] in THEIR buffer.  By default the observation crosses PCIe
        # in factored form and host threads rebuild the dense rows slice by slice as the transfers land
        dense = time_e2e(lambda i: env.step_host(h_act_pool[i % POOL], h_obs, h_rew, h_flag, auto_reset=True,
                                                 n_chunks=args.e2e_chunks))
        expanded = env._expand_threads() >= 4
        dma = time_e2e(lambda i: env.step_host(h_act_pool[i % POOL], h_obs, h_rew, h_flag, auto_reset=True,
                                               n_chunks=args.e2e_chunks, host_expand=False))
        e2e = {"value": dense, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h_fact if expanded else d2h_dense, "steps": e2e_steps, "layout": "dense",
               "api": ("BatchedStockTradingEnv.step_host: pinned host actions in, dense obs[N,O] + reward + flags in the caller's "
                       f"host buffers, pipelined in {args.e2e_chunks} env slices over 3 streams; "
                       + (f"the observation crosses PCIe factored and {env._expand_threads()} host threads rebuild the dense rows "
                          "(frl_expand_obs_host_chunks)" if expanded else "dense rows over PCIe")),
               "dense_over_pcie": {"value": dma, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_dense,
                                   "api": "step_host(host_expand=False): the dense rows themselves cross PCIe (round 1's path)"}}
        del h_obs
        _, f_obs, h_rew, h_flag = env.make_host_buffers("factored", tdt)
        fact = time_e2e(lambda i: env.step_host(h_act_pool[i % POOL], f_obs, h_rew, h_flag, auto_reset=True,
                                                n_chunks=args.e2e_chunks))
        d2h_f = f_obs.env_part.numel() * 4 + f_obs.state_day.numel() * 4 + h_rew.numel() * 8 + h_flag.numel()
        e2e["factored"] = {
            "value": fact, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_f,
            "api": "step_host(obs=FactoredObs): env_part[N,1+D] f32 + state_day[N] i32 + reward + flags come back; the "
                   "[T,O] per-day template stays pinned on the host (FactoredObs[n] / .dense() rebuild rows bit-exactly)"}
    else:
        OBS = wl.obs_numel(env)
        h_obs = torch.empty((N, OBS), dtype=torch.float32).pin_memory()
        rshape = (N,) if KR == 1 else (KR, N)
        h_rew = torch.empty(rshape, dtype=torch.float64).pin_memory()
        h_flag = torch.empty(rshape, dtype=torch.uint8).pin_memory()
        d_act = torch.empty(shape, dtype=tdt, device=dev)

        def plain(i):
            d_act.copy_(h_act_pool[i % POOL], non_blocking=True)
            if KR == 1:
                obs, rew, _, fl = env.step(d_act, auto_reset=True, want_obs=True, want_done=False)
            else:
                obs, rew, fl = env.rollout(d_act, obs_mode="last", auto_reset=True, accumulate_stats=False)
            h_obs.copy_(obs.reshape(N, OBS), non_blocking=True)
            h_rew.copy_(rew, non_blocking=True)
            h_flag.copy_(fl, non_blocking=True)
            torch.cuda.synchronize()  # the caller reads numpy arrays after every step

        e2e = {"value": time_e2e(plain), "unit": UNIT, "h2d_bytes_per_step": d_act.numel() * d_act.element_size(),
               "d2h_bytes_per_step": h_obs.numel() * 4 + h_rew.numel() * 8 + h_flag.numel(), "steps": e2e_steps,
               "layout": "dense",
               "api": "Batched*Env.step/rollout with pinned host actions in; obs + reward + flags copied back to host"}

    # ---- the other BASELINE configs, short runs in the same process (every N, so SCALE sees them) ---------
    peak, peak_src = measured_peak()
    workloads = {}
    if args.workload == "trading_step" and not args.no_extra:
        del env, pool, h_act_pool
        torch.cuda.empty_cache()
        for name in EXTRA_WORKLOADS:
            w2 = WORKLOADS[name]()
            n2 = w2.default_envs
            env2 = w2.make_env(dev, n2)
            ex.attach(env2)
            pool2, _, _ = make_pool(w2, n2, dev, torch, rank)
            ms, kms, _, _ = time_device(env2, w2, pool2, args.extra_steps, 5, world, ex, dist, torch)
            ach = w2.bytes_per_env_step * n2 * w2.rollout_k / (kms * 1e-3) / 1e9
            workloads[name] = {
                "value": n2 * world * args.extra_steps * w2.rollout_k / (ms * 1e-3), "unit": UNIT, "envs_per_gpu": n2,
                "steps": args.extra_steps, "kernel_ms": kms, "ms_per_step": ms / args.extra_steps,
                "algorithmic_bytes": w2.bytes_per_env_step * n2 * w2.rollout_k, "achieved_gbs": ach, "frac": ach / peak,
                "kernel": w2.kernel, "workload": w2.describe(n2)}
            del env2, pool2
            torch.cuda.empty_cache()
        ex.totals(reset=True)

    if rank == 0:
        achieved = wl.bytes_per_env_step * N * KR / (kernel_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(wl.name + "_dram_bytes_per_launch")
                traffic_src = "ncu --set full capture (profiles/roofline_traffic.json), not measured in this run"
            except Exception:
                traffic = None
        cfg = workload_config(wl, N, world)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg,
            "arm": {"total_envs": N * world,
                    "parallelism": f"env-index sharding x{world} (finrl_b200.dist.shard_range), tables replicated; statistics "
                                   f"exchange: {ex.mode}" + (" (fused into the step kernel's epilogue: fp64 atomics into every "
                                   "rank's peer-mapped totals, no collective launch)" if ex.mode == "p2p" else
                                   f" (fallback: {getattr(ex, 'fallback_reason', None)})"),
                    "timeline_alignment": "device-side all-reduce right before the start event" if world > 1 else "single GPU"},
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "frac_of_nominal_8000": achieved / 8000.0,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel": wl.kernel,
                "kernel_ms": kernel_ms, "algorithmic_bytes_per_env_step": wl.bytes_per_env_step,
            },
            "e2e": e2e,
            "gpu_launches": launches,
            "host": {"cpus": len(all_cpus), "cpus_bound_to_gpu_numa": numa_cpus},
            "clocks": clocks,
            "stats": dict(zip(("reward_sum", "reward_sqsum", "done_count", "episode_asset_sum", "asset_sum", "liq_count",
                               "env_steps", "trades_sum"), totals)),
        }
        if workloads:
            out["workloads"] = workloads
        if world == 1 and not args.no_cpu:
            v, cores, sample = cpu_port_rate(wl, args.cpu_seconds, args.ref_envs or args.envs)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
            out["python_reference"] = PY_REFERENCE
        print(json.dumps(out))
    ex.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# the other BASELINE configs, then the NASDAQ-100-sized StockTradingEnv / numpy env (not BASELINE configs: the wide kernels)
EXTRA_WORKLOADS = ("trading_rollout", "np_step", "portfolio_step", "cashpenalty_step", "trading_nas100_step", "np_nas100_step")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="trading_step", choices=sorted(WORKLOADS))
    ap.add_argument("--envs", type=int, default=None, help="envs per GPU")
    ap.add_argument("--ref-envs", type=int, default=None, help="envs per step of the CPU arms (default: the GPU arm's envs per GPU)")
    ap.add_argument("--no-extra", action="store_true", help="skip the short runs of the other BASELINE configs")
    ap.add_argument("--extra-steps", type=int, default=20)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--e2e-steps", type=int, default=30)  # ~10 ms each at 1M envs; short runs of 10 scattered 66..110 M on busy hosts
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true", help="e2e through plain step() + copies instead of step_host()")
    ap.add_argument("--e2e-chunks", type=int, default=8)  # measured: 1 -> 42.2, 4 -> 45.2, 8 -> 45.6, 16 -> 45.7 M env-steps/s
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]()
    args.envs = args.envs or wl.default_envs
    args.warmup = max(args.warmup, 3)
    if args.steps is None:
        args.steps = 25 if args.impl == "reference" else 2000
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
