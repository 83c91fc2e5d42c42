"""Multi-GPU layer: env-index sharding + the one collective on the path.

Envs are independent, so N envs are split into contiguous index ranges, one per rank (one process
per GPU, ``torch.distributed`` over NCCL/NVLink for the plumbing); the market tables (5-20 MB) are
replicated on every GPU and no data-path collective exists.  The only exchange is the sum of the 64-byte
statistics vector (episode-return / asset / reward sums) that the step kernels accumulate in their
epilogue.  :class:`StatsExchange` does it ONE-SIDED inside those kernels: every rank's 384-byte statistics
block is cudaMalloc'ed, exported over CUDA IPC and peer-mapped by all ranks of the node; launches alternate
between the block's two accumulators, and the first thread block of each step / rollout launch adds the
PREVIOUS launch's sums to every rank's totals with fp64 atomics over NVLink / NVSwitch
(``stats_exchange_previous`` in csrc/common.cuh).  No collective is launched, nothing is fenced or counted,
and no rank waits for a peer inside its step loop, so one slow host cannot stall seven GPUs (a blocking
64-byte NCCL all-reduce on the compute stream cost 14 % at 8 GPUs in round 1).  When the blocks cannot be peer-mapped
(no IPC in the container, no P2P between the devices, ranks on several nodes) it falls back to an NCCL
all-reduce of double-buffered snapshots on a side stream.  The reference has no counterpart.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional, Tuple

from . import _cabi
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


class StatsExchange:
    """Global (all ranks) sums of the engines' statistics vectors without a collective in the step loop.

    ``mode`` is ``"p2p"`` (peer-mapped blocks, the kernels push in their epilogue), ``"collective"``
    (``torch.distributed`` all-reduce of snapshots on a side stream: NCCL fallback, and gloo for CPU tensors in
    the host-logic tests) or ``"local"`` (one rank).  Usage::

        ex = StatsExchange(device)          # collective call: every rank of the group constructs one
        ex.attach(env)                      # env.step/rollout(accumulate_stats=True) now feed the exchange
        ... step loop: ex.flush() every so often (a no-op in p2p mode) ...
        totals = ex.totals()                # synchronises the ranks and returns the 8 global sums
    """

    def __init__(self, device, group=None, mode: Optional[str] = None):
        import torch
        import torch.distributed as dist

        self._torch, self._dist, self.group = torch, dist, group
        self.device = torch.device(device)
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self._own = None          # cudaMalloc'ed block (p2p mode)
        self._opened = []         # peer mappings to close
        want = mode or os.environ.get("FRL_STATS_EXCHANGE") or ("p2p" if self.device.type == "cuda" else "collective")
        if self.world == 1 and want != "p2p":
            want = "local"
        self.mode = want
        if want == "p2p":
            self._setup_p2p()
        if self.mode != "p2p":
            self.block = (_cabi.new_stats_block(torch, self.device) if self.device.type == "cuda"
                          else torch.zeros(_cabi.STATS_BLOCK_DOUBLES, dtype=torch.float64))
        self.sum, self.total = self.block[:N_STATS], self.block[2 * N_STATS : 3 * N_STATS]
        self._acc = self.block[: 2 * N_STATS]  # both accumulators
        if self.mode == "collective":
            self._snap = [torch.zeros_like(self.sum) for _ in range(2)]
            self._turn = 0
            if self.device.type == "cuda":
                self._side = torch.cuda.Stream(self.device)
                self._free = [torch.cuda.Event(), torch.cuda.Event()]  # snapshot i may be overwritten again
                for e in self._free:
                    e.record(torch.cuda.current_stream(self.device))

    # ---- p2p setup: allocate, export, gather handles, map, bind ------------------------------------------
    def _setup_p2p(self):
        torch, dist, lib = self._torch, self._dist, _cabi.lib()
        ok, err, handle = True, "", b""
        try:
            with torch.cuda.device(self.device):
                blk = C.c_void_p()
                _cabi.check(lib.frl_exchange_alloc(C.byref(blk)), "frl_exchange_alloc")
                self._own = blk.value
                buf = C.create_string_buffer(64)
                _cabi.check(lib.frl_exchange_export(C.c_void_p(self._own), buf), "frl_exchange_export")
                handle = buf.raw
        except _cabi.EngineError as e:
            ok, err = False, str(e)
        infos = [(ok, handle, self.device.index)]
        if self.world > 1:
            infos = [None] * self.world
            dist.all_gather_object(infos, (ok, handle, self.device.index), group=self.group)
        ptrs = []
        if all(i[0] for i in infos):
            try:
                with torch.cuda.device(self.device):
                    for r, (_, h, _) in enumerate(infos):
                        if r == self.rank:
                            ptrs.append(self._own)
                            continue
                        q = C.c_void_p()
                        _cabi.check(lib.frl_exchange_open(h, C.byref(q)), f"frl_exchange_open(rank {r})")
                        self._opened.append(q.value)
                        ptrs.append(q.value)
            except _cabi.EngineError as e:
                ok, err = False, str(e)
        else:
            ok = False
        if self.world > 1:  # every rank must take the same path
            flags = [None] * self.world
            dist.all_gather_object(flags, (ok, err), group=self.group)
            bad = [f"rank {r}: {m}" for r, (k, m) in enumerate(flags) if not k]
            ok = not bad
            err = "; ".join(bad)
        if not ok:
            self._release()
            self.mode = "collective" if self.world > 1 else "local"
            self.fallback_reason = err or "a peer could not export its block"
            return
        with torch.cuda.device(self.device):
            arr = (C.c_void_p * len(ptrs))(*ptrs)
            _cabi.check(lib.frl_exchange_bind(C.c_void_p(self._own), arr, len(ptrs), _cabi.current_stream(self.device)),
                        "frl_exchange_bind")
            torch.cuda.current_stream(self.device).synchronize()
            self.block = torch.as_tensor(_cabi._DevicePointerView(self._own, _cabi.STATS_BLOCK_DOUBLES), device=self.device)
        self.fallback_reason = None
        if self.world > 1:
            dist.barrier(group=self.group)  # nobody pushes before every rank's block is bound

    def _release(self):
        lib = _cabi.lib()
        for q in self._opened:
            lib.frl_exchange_close(C.c_void_p(q))
        self._opened = []
        if self._own is not None:
            lib.frl_exchange_free(C.c_void_p(self._own))
            self._own = None

    def close(self):
        """Unmap the peers and free the block (collective: call on every rank, after the last launch)."""
        if self.mode == "p2p":
            self._torch.cuda.synchronize(self.device)
            if self.world > 1:
                self._dist.barrier(group=self.group)
            self.block = self.sum = self.total = self._acc = None
            self._release()
            self.mode = "closed"

    # ---- use ---------------------------------------------------------------------------------------------
    def attach(self, env):
        """Make ``env`` accumulate into this exchange's block (CUDA engines)."""
        env.use_stats_block(self.block, alternate=self.mode == "p2p")
        return env

    def flush(self):
        """Move what has accumulated in ``sum`` towards the global totals WITHOUT blocking the compute stream.
        p2p: nothing to do (every launch pushes its predecessor's sums).  collective: snapshot + zero on the compute
        stream, all-reduce on a side stream, double-buffered."""
        torch, dist = self._torch, self._dist
        if self.mode == "p2p":
            return
        if self.mode == "local":
            self.total.add_(self.sum)
            self.sum.zero_()
            return
        i = self._turn
        self._turn ^= 1
        snap = self._snap[i]
        if self.device.type != "cuda":
            snap.copy_(self.sum)
            self.sum.zero_()
            dist.all_reduce(snap, group=self.group)
            self.total.add_(snap)
            return
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self._free[i])
        snap.copy_(self.sum)
        self.sum.zero_()
        ready = torch.cuda.Event()
        ready.record(cur)
        with torch.cuda.stream(self._side):
            self._side.wait_event(ready)
            work = dist.all_reduce(snap, group=self.group, async_op=True)
            work.wait()  # orders the side stream after the collective; the host does not block
            self.total.add_(snap)
            self._free[i].record(self._side)

    def totals(self, reset: bool = False):
        """The 8 global sums over everything accumulated so far — a synchronisation point of all ranks."""
        torch, dist = self._torch, self._dist
        self.flush()
        if self.mode == "p2p":  # the last launch's sums are still in an accumulator: one tiny kernel moves them
            with torch.cuda.device(self.device):
                _cabi.check(_cabi.lib().frl_exchange_flush(C.c_void_p(self._own), _cabi.current_stream(self.device)),
                            "frl_exchange_flush")
        if self.device.type == "cuda":
            if self.mode == "collective":
                torch.cuda.current_stream(self.device).wait_stream(self._side)
            torch.cuda.synchronize(self.device)
        if self.mode == "p2p" and self.world > 1:
            dist.barrier(group=self.group)  # every rank's launches (and their pushes) have completed
        vals = self.total.tolist()
        if reset:
            if self.mode == "p2p" and self.world > 1:
                dist.barrier(group=self.group)  # everyone has read before anyone clears
            self.total.zero_()
            if self.device.type == "cuda":
                torch.cuda.synchronize(self.device)
            if self.mode == "p2p" and self.world > 1:
                dist.barrier(group=self.group)  # cleared everywhere before the next push
        return vals
