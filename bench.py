#!/usr/bin/env python
"""bench.py — env-steps/s of the StockTradingEnv step path on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--envs E]

Workload (config.workload): StockTradingEnv, DOW-30 shape (D=30 stocks, K=8 indicators, T=2500 days,
O=301), E = 1,048,576 envs PER GPU, one ``step()`` of every env per bench step, float32 actions
U(-1,1) distinct per env and step, turbulence threshold 99, auto-reset at episode ends.  A "step" is
one pass of the hot path over the whole batch.

* value        env-steps/s with inputs resident in HBM (CUDA events, max over ranks, whole job)
* e2e          same metric through the numpy-in / numpy-out VecEnv-style call: pinned HOST actions
               -> H2D -> step -> D2H of obs + reward + done, every step inside the timed region
* roofline     algorithmic bytes per launch (1621 B/env-step) / measured kernel time
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

METRIC = "env-steps/sec"
UNIT = "env-steps/s"
D, K_TECH, T_DAYS = 30, 8, 2500
OBS = 1 + 2 * D + K_TECH * D
BYTES_PER_ENV_STEP = 1621  # SURVEY.md §8(d) "Config 2' single step (DOW-30)"
ENV_KW = dict(hmax=100, initial_amount=1_000_000, buy_cost_pct=0.001, sell_cost_pct=0.001, reward_scaling=1e-4,
              turbulence_threshold=99)


def workload_name(envs):
    return f"StockTradingEnv DOW-30 single step, {envs} envs/GPU, D=30 K=8 T=2500 O=301, f32 actions, f32 obs"


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
            self._stop.wait(0.05)

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


def cpu_port_rate(budget_s=12.0, envs=65536, threads=None):
    """Time the CPU oracle port on a bounded sample of the same workload.  Returns (env-steps/s,
    threads, description).  This is the one place bench.py executes oracle/."""
    from finrl_b200 import synthetic as syn
    from oracle import oracle as ora

    threads = threads or os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(threads))
    close, tech, turb = syn.make_tables(T_DAYS, D, K_TECH, seed=0)
    o = ora.TradingOracle(close, tech, turb, envs, **ENV_KW)
    pool = [syn.make_actions((envs, D), seed=100 + i) for i in range(4)]
    for i in range(3):
        o.step(pool[i % 4], auto_reset=True)
    steps, t0 = 0, time.perf_counter()
    while True:
        o.step(pool[steps % 4], auto_reset=True)
        steps += 1
        el = time.perf_counter() - t0
        if el >= budget_s:
            break
    return envs * steps / el, threads, f"{envs} envs x {steps} steps ({el:.1f} s), obs written, OpenMP static over envs"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # each "step" is one pass of the CPU port over a bounded sample of the workload
    from finrl_b200 import synthetic as syn
    from oracle import oracle as ora

    threads = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(threads))
    envs = args.ref_envs
    close, tech, turb = syn.make_tables(T_DAYS, D, K_TECH, seed=0)
    o = ora.TradingOracle(close, tech, turb, envs, **ENV_KW)
    pool = [syn.make_actions((envs, D), seed=100 + i) for i in range(4)]
    for i in range(args.warmup):
        o.step(pool[i % 4], auto_reset=True)
    t0 = time.perf_counter()
    for i in range(args.steps):
        o.step(pool[i % 4], auto_reset=True)
    el = time.perf_counter() - t0
    v = envs * args.steps / el
    sample = f"{envs} envs per step (bounded sample of the {args.envs} envs/GPU workload), obs written"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.envs), "cpu_sample_envs": envs},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is pure Python (pandas/numpy) and cannot travel; this arm is its C restatement "
                "(oracle/oracle.c) on all host threads. Survey-time reference itself: ~159 env-steps/s on 1 core.",
    }))


def run_ours(args):
    import torch
    import torch.distributed as dist

    from finrl_b200 import BatchedStockTradingEnv, TradingTables, synthetic as syn

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N = args.envs
    close, tech, turb = syn.make_tables(T_DAYS, D, K_TECH, seed=0)
    tables = TradingTables.from_arrays(close, tech, turb, dev)  # replicated per GPU
    env = BatchedStockTradingEnv(tables=tables, n_envs=N, device=dev, **ENV_KW)
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)  # envs are sharded by index: every rank owns different envs/actions
    POOL = 3
    pool = [(torch.rand((N, D), generator=g, device=dev, dtype=torch.float32) * 2.0 - 1.0) for _ in range(POOL)]
    stats_global = torch.zeros_like(env.stats)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(i):
        env.step(pool[i % POOL], auto_reset=True, want_obs=True, accumulate_stats=True, want_done=False)

    def reduce_stats():
        # the only collective on the path: <=64 B episode-return / asset statistics over NVLink
        stats_global.copy_(env.stats)
        if world > 1:
            dist.all_reduce(stats_global)

    for i in range(args.warmup):
        one_step(i)
    reduce_stats()
    barrier()

    sampler = ClockSampler(local)
    sampler.start()
    launches0 = env.launches
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.kernel_events = []  # CUDA events right around every kernel launch, on the launching stream
    barrier()
    t_start.record()
    for i in range(args.steps):
        one_step(i)
        if (i + 1) % 16 == 0 or i == args.steps - 1:
            reduce_stats()
    t_end.record()
    barrier()
    clocks = sampler.stop()
    elapsed_ms = t_start.elapsed_time(t_end)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in env.kernel_events]))
    env.kernel_events = None
    launches = env.launches - launches0
    el = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    elapsed_ms = float(el.item())
    value = N * world * args.steps / (elapsed_ms * 1e-3)

    # ---- end to end through the numpy-facing call: host actions in, host obs/reward/done out ----
    h_act = [torch.empty((N, D), dtype=torch.float32).pin_memory() for _ in range(POOL)]
    for h, d in zip(h_act, pool):
        h.copy_(d)
    h_obs = torch.empty((N, OBS), dtype=torch.float32).pin_memory()
    h_rew = torch.empty(N, dtype=torch.float64).pin_memory()
    h_flag = torch.empty(N, dtype=torch.uint8).pin_memory()
    d_act = torch.empty((N, D), dtype=torch.float32, device=dev)

    def e2e_step(i):
        d_act.copy_(h_act[i % POOL], non_blocking=True)
        obs, rew, done, fl = env.step(d_act, auto_reset=True, want_obs=True)
        h_obs.copy_(obs, non_blocking=True)
        h_rew.copy_(rew, non_blocking=True)
        h_flag.copy_(fl, non_blocking=True)
        torch.cuda.synchronize()  # the caller reads numpy arrays after every step

    e2e_steps = max(3, min(args.steps, args.e2e_steps))
    for i in range(3):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    barrier()
    e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = N * world * e2e_steps / float(e2e_t.item())
    h2d = N * D * 4
    d2h = N * OBS * 4 + N * 8 + N

    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = BYTES_PER_ENV_STEP * N / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("trading_step_dram_bytes_per_launch")
            except Exception:
                traffic = None
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": workload_name(N), "envs_per_gpu": N, "total_envs": N * world,
                "l2": "per-step traffic (~1.7 GB) exceeds the 126 MB L2; no explicit flush",
                "parallelism": f"env-index sharding x{world}, tables replicated, NCCL all-reduce of 64 B stats every 16 steps",
                "auto_reset": True,
            },
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel": "trading_rollout_kernel<32,float,4>",
                "kernel_ms": kernel_ms, "algorithmic_bytes_per_env_step": BYTES_PER_ENV_STEP,
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "BatchedStockTradingEnv.step with pinned host actions in, obs+reward+done copied to host"},
            "gpu_launches": launches,
            "clocks": clocks,
            "stats": dict(zip(("reward_sum", "reward_sqsum", "done_count", "episode_asset_sum", "asset_sum", "liq_count",
                               "env_steps", "trades_sum"), stats_global.tolist())),
        }
        if world == 1 and not args.no_cpu:
            v, cores, sample = cpu_port_rate(args.cpu_seconds)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=1 << 20, help="envs per GPU")
    ap.add_argument("--ref-envs", type=int, default=65536, help="envs per step of the CPU reference arm")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
