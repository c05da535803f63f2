"""Build librrtmg_b200.so in-tree (mima_b200/lib/) with nvcc for sm_100a.

nvcc cross-compiles without a GPU.  The library is git-ignored but travels to the GPU box with the
repo snapshot.  -fmad=false: fused multiply-adds only where the kernels write fma() explicitly, so the
index/branch arithmetic (jp, jt, js, table indices) is bit-identical to an IEEE host evaluation.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "librrtmg_b200.so")
SOURCES = ["api.cu", "lw_kernels.cu", "sw_kernels.cu"]
HEADERS = ["rrtmg_dev.cuh", os.path.join("..", "..", "include", "rrtmg_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2",
    "-shared",
]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    env = dict(os.environ)
    # the image exports CC/CXX pointing at a gcc wrapper without OpenMP specs; nvcc only needs a host g++
    for host in ("/usr/bin/g++",):
        if os.path.exists(host):
            cmd[1:1] = ["-ccbin", host]
            break
    r = subprocess.run(cmd, cwd=CSRC, env=env, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building librrtmg_b200.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
