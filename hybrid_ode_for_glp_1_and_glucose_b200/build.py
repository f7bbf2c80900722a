"""Builds libhode.so (the C-ABI CUDA library) in-tree with plain nvcc for sm_100a.

    python -m hybrid_ode_for_glp_1_and_glucose_b200.build [--force] [--verbose]

No torch headers are involved: the library's boundary is include/hode.h.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libhode.so")
SOURCES = ["hode_api.cu", "hode_rollout_simt.cu", "hode_rollout_tc.cu", "hode_adjoint_simt.cu", "hode_adjoint_tc.cu",
           "hode_gen4gi.cu", "hode_train.cu", "hode_data.cu"]
HEADERS = ["hode_common.cuh", "hode_kernels.h", "hode_tcgen05.cuh", "hode_tc_mlp.cuh", os.path.join("..", "..", "include", "hode.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    objs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
        objs.append(obj)
    cmd = [_nvcc(), "-shared", "-o", LIB, *objs]  # cudart is linked statically (nvcc default)
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
