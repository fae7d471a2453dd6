"""Build liblinr_b200.so (CUDA kernels + C ABI + host range coder) for sm_100a, in-tree.

    python linr-pcgc_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU.  The .so lands in linr-pcgc_b200/lib/ (git-ignored, travels with gpurun).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
SO = os.path.join(LIBDIR, "liblinr_b200.so")
SOURCES = ["coords.cu", "net.cu", "rc_host.cpp"]
HEADERS = ["common.cuh", "net_kernels.cuh", "prof.cuh", os.path.join("..", "..", "include", "linr_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--use_fast_math=false",
         "-Xcompiler", "-fPIC,-O3,-pthread", "-shared", "-Xptxas", "-warn-spills"]
FLAGS = [f for f in FLAGS if f != "--use_fast_math=false"]  # precise math: never pass fast-math


def stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return SO
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building liblinr_b200.so")
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
