"""Multi-GPU layer: env-index sharding + the one collective on the path.

Envs are independent, so N envs are split into contiguous index ranges, one per rank (one process
per GPU, ``torch.distributed`` over NCCL/NVLink); the market tables (5-20 MB) are replicated on every
GPU and no data-path collective exists.  The only exchange is an all-reduce of the 64-byte
statistics vector (episode-return / asset / reward sums) that the step kernels accumulate in their
epilogue — latency bound (~tens of microseconds), so it is issued once per rollout, asynchronously on
the compute stream, never per env.  The reference has no counterpart.
"""
from __future__ import annotations

import math
import os
from typing import Optional, Tuple

from ._cabi import N_STATS, STAT_NAMES


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of [0, n_total): returns (start, count) of ``rank``."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside [0, {world})")
    base, extra = divmod(int(n_total), int(world))
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def bind_cpu_affinity(device_index: int) -> Optional[int]:
    """Pin this process to the CPU cores NVML reports as local to the GPU (same NUMA node / PCIe root), so
    that pinned host buffers allocated afterwards are local to the GPU's link and the host-buffer step does
    not cross sockets.  Returns the number of CPUs in the new affinity set, or None when NVML is unavailable.
    ``FRL_NO_NUMA_BIND=1`` disables it."""
    if os.environ.get("FRL_NO_NUMA_BIND") == "1":
        return None
    try:
        import pynvml
        import torch

        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(device_index)
        try:
            bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


def init_from_env(backend: Optional[str] = None):
    """Read RANK / LOCAL_RANK / WORLD_SIZE (torchrun), bind the GPU, create the process group.
    Returns (rank, world, local_rank)."""
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    use_cuda = torch.cuda.is_available()
    if use_cuda:
        torch.cuda.set_device(local)
        bind_cpu_affinity(local)
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if use_cuda else "gloo")
        kw = {"device_id": torch.device("cuda", local)} if backend == "nccl" else {}
        dist.init_process_group(backend, **kw)
    return rank, world, local


def allreduce_stats(stats, group=None, async_op: bool = False):
    """Sum the statistics vector over all ranks in place (NCCL for CUDA tensors, gloo for CPU)."""
    import torch.distributed as dist

    if stats.numel() != N_STATS:
        raise ValueError(f"stats must have {N_STATS} elements")
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    return dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


def summarize(stats) -> dict:
    """Human-level quantities from a (globally reduced) statistics vector."""
    v = dict(zip(STAT_NAMES, [float(x) for x in stats.tolist()]))
    n = max(v["env_steps"], 1.0)
    mean = v["reward_sum"] / n
    var = max(v["reward_sqsum"] / n - mean * mean, 0.0)
    out = dict(v)
    out["reward_mean"] = mean
    out["reward_std"] = math.sqrt(var)
    out["episode_asset_mean"] = v["episode_asset_sum"] / v["done_count"] if v["done_count"] > 0 else float("nan")
    return out
