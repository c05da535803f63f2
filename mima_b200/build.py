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
# (source, FMA contraction).  setcoef/taumol decide table indices and branches from FP64 arithmetic and are
# compiled with -fmad=false so that those decisions match an IEEE host evaluation bit for bit (fused
# multiply-adds appear only where written as fma()); the solvers carry no such decisions and may contract.
SOURCES = [("api.cu", False), ("lw_kernels.cu", False), ("lw_column.cu", False), ("sw_kernels.cu", False), ("driver.cu", False),
           ("lw_solver.cu", True), ("sw_solver.cu", True), ("sw_column.cu", True)]
HEADERS = ["rrtmg_dev.cuh", "lw_bands.cuh", "sw_bands.cuh", "sw_twostream.cuh", os.path.join("..", "..", "include", "rrtmg_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2",
]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def source_hash() -> str:
    """sha256 over the translation units that hold kernels and the device headers they include (api.cu and the public header it
    alone includes are host code: C ABI, host pipelines, table set-up): profiles/traffic.json is stamped with it
    (tools/make_traffic.py), and bench.py reports the ncu-measured DRAM traffic only while the stamp matches the tree it runs."""
    import hashlib
    h = hashlib.sha256()
    for f in sorted([x for x, _ in SOURCES if x != "api.cu"] + [x for x in HEADERS if x.endswith(".cuh")]):
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in [x for x, _ in SOURCES] + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    cc = [nvcc()]
    # the image exports CC/CXX pointing at a gcc wrapper without OpenMP specs; nvcc only needs a host g++
    if os.path.exists("/usr/bin/g++"):
        cc += ["-ccbin", "/usr/bin/g++"]
    from concurrent.futures import ThreadPoolExecutor

    def compile_one(item):
        src, fmad = item
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        dev = ["-DRRTMG_B200_DEV_VARIANTS"] if os.environ.get("RRTMG_B200_DEV_VARIANTS") == "1" else []
        dev += os.environ.get("RRTMG_B200_DEFS", "").split()          # extra -D flags of an experiment build
        cmd = cc + NVCC_FLAGS + dev + ["-fmad=true" if fmad else "-fmad=false"] + (["-Xptxas", "-v"] if verbose else []) + \
            ["-c", "-o", obj, src]
        r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
        return obj, r

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    for obj, r in results:
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed building " + obj)
        if verbose:
            sys.stderr.write(r.stderr)
    r = subprocess.run(cc + ["-shared", "-o", LIB] + [o for o, _ in results], cwd=CSRC, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking librrtmg_b200.so")
    return LIB


# RRTMG_B200_DEV_VARIANTS=1 in the environment also compiles the earlier kernel forms (SW solver variants 0-3, the
# direct-load lw_rtrn) that tests/test_gpu_parity.py::test_sw_solver_variants_agree compares when they are present.
if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
