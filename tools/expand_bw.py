"""Host-side bandwidth probe for the dense-by-expansion path of step_host (GPU box): how fast do host threads rebuild
the dense observation (frl_expand_obs_host), how fast does the copy engine write dense rows into pinned memory, and
do the two add up when they run at the same time?

    python tools/expand_bw.py
"""
import ctypes as C
import os
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from finrl_b200 import _cabi  # noqa: E402

N, D, O, T = 1 << 20, 30, 301, 2500
lib = _cabi.lib()
tmpl = np.random.rand(T, O).astype(np.float32)
ep = torch.rand((N, 1 + D)).pin_memory()
sd = torch.randint(0, T, (N,), dtype=torch.int32).pin_memory()
out = torch.empty((N, O), dtype=torch.float32).pin_memory()
dev = torch.rand((N, O), device="cuda")
bytes_dense = N * O * 4


def expand(lo, hi, nt):
    rc = lib.frl_expand_obs_host(tmpl.ctypes.data, T, O, D, ep.data_ptr() + 4 * (1 + D) * lo, sd.data_ptr() + 4 * lo, hi - lo,
                                 out.data_ptr() + 4 * O * lo, nt)
    assert rc == 0


os.system("lscpu | egrep 'Model name|Socket|Core|Thread|NUMA|Flags' | cut -c1-400")
for nt in (4, 8, 12, 16):
    expand(0, N, nt)
    t = time.perf_counter()
    for _ in range(5):
        expand(0, N, nt)
    dt = (time.perf_counter() - t) / 5
    print(f"expand alone, {nt:2d} threads: {dt * 1e3:6.2f} ms  {bytes_dense / dt / 1e9:6.1f} GB/s written")

st = torch.cuda.Stream()
with torch.cuda.stream(st):
    out.copy_(dev, non_blocking=True)
st.synchronize()
t = time.perf_counter()
with torch.cuda.stream(st):
    for _ in range(5):
        out.copy_(dev, non_blocking=True)
st.synchronize()
dt = (time.perf_counter() - t) / 5
print(f"D2H dense alone: {dt * 1e3:6.2f} ms  {bytes_dense / dt / 1e9:6.1f} GB/s")

for frac in (0.15, 0.2, 0.25, 0.3):
    cut = int(N * frac) // 32 * 32
    for nt in (12, 16):
        best = 1e9
        for _ in range(4):
            t = time.perf_counter()
            with torch.cuda.stream(st):
                out[:cut].copy_(dev[:cut], non_blocking=True)
            expand(cut, N, nt)
            t_exp = time.perf_counter() - t
            st.synchronize()
            best = min(best, time.perf_counter() - t)
        print(f"hybrid: {frac:.2f} of the rows by DMA, the rest by {nt} threads: {best * 1e3:6.2f} ms (expand part {t_exp * 1e3:6.2f})"
              f"  {bytes_dense / best / 1e9:6.1f} GB/s  -> {N / best / 1e6:6.1f} M rows/s")
