#!/usr/bin/env python
"""Cycle-level timeline of one main thread of rollout_bwd_tc_kernel (debug build, -DHODE_TIMELINE)."""
import ctypes, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CSRC = os.path.join(ROOT, "hybrid_ode_for_glp_1_and_glucose_b200", "csrc")
flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr"]
subprocess.run(["nvcc", *flags, "-DHODE_TIMELINE", *os.environ.get("HODE_TL_FLAGS", "").split(), "-c", os.path.join(CSRC, "hode_adjoint_tc.cu"), "-o",
                "/tmp/hode_adjoint_tc_tl.o"], check=True)
objs = [os.path.join(CSRC, f) for f in ("hode_api.o", "hode_rollout_simt.o", "hode_adjoint_simt.o", "hode_rollout_tc.o", "hode_gen4gi.o")]
lib = os.path.join(ROOT, "hybrid_ode_for_glp_1_and_glucose_b200", "libhode.so")
os.rename(lib, lib + ".bak")
try:
    subprocess.run(["nvcc", "-shared", "-o", lib, *objs, "/tmp/hode_adjoint_tc_tl.o"], check=True)
    import torch
    from hybrid_ode_for_glp_1_and_glucose_b200 import _lib, ops
    from hybrid_ode_for_glp_1_and_glucose_b200.synthetic import THETA_DEFAULT, cohort, random_mlp
    dev = torch.device("cuda:0")
    B = int(os.environ.get("TL_B", 32768))
    y0, t, ins = cohort(B, 61, seed=1000)
    W = random_mlp(64, 4, seed=1234, out_std=0.05)
    tt = lambda a: torch.from_numpy(a).to(dev)
    L = _lib.lib()
    L.hode_debug_timeline.restype = ctypes.c_int
    buf = np.zeros(2 * 16384, dtype=np.int64)
    g = torch.full((B, 61, 6), 1.0 / (B * 366), device=dev)
    for it in range(2):
        _, info, tape = ops.rollout(tt(y0), tt(t), {k: tt(v) for k, v in ins.items()}, tt(THETA_DEFAULT), tt(W),
                                    solver="dopri5", precision="tf32x3", device=dev, save_steps=True)
        ops.rollout_bwd(tape, g)
        torch.cuda.synchronize()
        n = L.hode_debug_timeline(buf.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), 16384)
    ev = buf[: 2 * n].reshape(n, 2)
    ids, clk = ev[:, 0], ev[:, 1]
    s = n // 5
    names = {}
    for a, b, d in zip(ids[s:-1], ids[s + 1:], np.diff(clk[s:])):
        names.setdefault((int(a), int(b)), []).append(int(d))
    n_steps = sum(1 for x in ids[s:] if x == 200) or max(1, sum(1 for x in ids[s:] if x in (300, 324)) // (24 if any(x == 300 for x in ids[s:]) else 6))
    print(f"events {n}, steps analysed {n_steps}")
    tot = 0.0
    for k, v in sorted(names.items()):
        v = np.array(v)
        print(f"{k[0]:3d}->{k[1]:3d}  n={len(v):5d}  median {np.median(v):8.0f}  mean {v.mean():8.0f}  p90 {np.percentile(v, 90):8.0f}   per-step {v.mean() * len(v) / max(n_steps, 1):9.0f}")
        tot += v.sum()
    print(f"total cycles per step: {tot / max(n_steps, 1):.0f}")
finally:
    os.replace(lib + ".bak", lib)
