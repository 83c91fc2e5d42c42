"""What does this GPU's HBM sustain for the traffic MIX of the env kernels?  MEASURED_PEAKS.json's figure is a copy
(50 % reads / 50 % writes); an env step is 75-95 % writes (the observation).  Plain torch kernels, CUDA events, buffers
far larger than L2:  fill (100 % writes), copy (50/50), read-only reduction, and a 1-read : 4-write broadcast copy.

    python tools/dram_mix.py
"""
import torch

dev = "cuda"
n = 1 << 28  # 1 GiB of float32 per unit


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


x = torch.empty(n, dtype=torch.float32, device=dev).normal_()
y = torch.empty(n, dtype=torch.float32, device=dev)
big = torch.empty(4 * n, dtype=torch.float32, device=dev)
t = timed(lambda: y.copy_(x))
print(f"copy   (50 % writes): {2 * 4 * n / t / 1e9:7.0f} GB/s")
t = timed(lambda: big.fill_(1.0))
print(f"fill  (100 % writes): {4 * 4 * n / t / 1e9:7.0f} GB/s")
t = timed(lambda: x.sum())
print(f"sum     (0 % writes): {4 * n / t / 1e9:7.0f} GB/s")
t = timed(lambda: big.view(4, n).copy_(x.view(1, n).expand(4, n)))
print(f"1 read : 4 writes (80 % writes, the mix of a trading-env step): {5 * 4 * n / t / 1e9:7.0f} GB/s  (DRAM traffic if the re-reads hit L2)")
small = torch.empty(n // 4, dtype=torch.float32, device=dev).normal_()
out = torch.empty(n, dtype=torch.float32, device=dev)
t = timed(lambda: out.view(4, n // 4).copy_(small.view(1, n // 4).expand(4, n // 4)))
print(f"same at a quarter of the size: {5 * n / t / 1e9:7.0f} GB/s")
