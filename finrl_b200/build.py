"""Build the C-ABI shared library (finrl_b200/libfinrl_b200.so) in-tree with nvcc for sm_100a.

The library travels to the GPU box with the repo snapshot (it is git-ignored, not gpurun-ignored).
nvcc cross-compiles without a GPU, so this also runs in the CPU-only build container.
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libfinrl_b200.so")
SOURCES = ["abi.cu", "trading.cu", "trading_small.cu", "trading_wide.cu", "nptrading.cu", "np_wide.cu", "portfolio.cu", "cashpenalty.cu", "preprocess.cu", "crypto.cu", "stoploss.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "finrl_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines=(), out: str = None) -> str:
    """Compile every .cu for sm_100a and link the shared library.  ``defines``/``out`` build a tuning
    variant (e.g. defines=["FRL_ST_STREAM=1"], out="gpurun_out/libvariant.so") without touching LIB."""
    lib_out = out or LIB
    if not force and out is None and not needs_build():
        return LIB
    objs = []
    procs = []
    tag = "" if out is None else "_" + os.path.basename(out).replace(".so", "")
    os.makedirs(os.path.join(PKG, "build"), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(PKG, "build", src.replace(".cu", tag + ".o"))
        cmd = [_nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-c",
               os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out:
            print(out, file=sys.stderr)
    link = [_nvcc(), "-shared", "--cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib_out, *objs]
    subprocess.check_call(link)
    return lib_out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
