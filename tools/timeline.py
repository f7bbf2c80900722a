#!/usr/bin/env python
"""Cycle-level timeline of one main thread of rollout_tc_kernel (debug build with -DHODE_TIMELINE).
Usage (GPU box): python tools/timeline.py  -> prints per-phase cycle statistics."""
import ctypes, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CSRC = os.path.join(ROOT, "hybrid_ode_for_glp_1_and_glucose_b200", "csrc")
flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr"]
subprocess.run(["nvcc", *flags, "-DHODE_TIMELINE", "-c", os.path.join(CSRC, "hode_rollout_tc.cu"), "-o",
                "/tmp/hode_rollout_tc_tl.o"], check=True)
objs = [os.path.join(CSRC, f) for f in ("hode_api.o", "hode_rollout_simt.o", "hode_adjoint_simt.o", "hode_adjoint_tc.o", "hode_gen4gi.o",
                                        "hode_train.o", "hode_data.o")]
PRECISION = os.environ.get("TL_PRECISION", "tf32x2bf16")
lib = os.path.join(ROOT, "hybrid_ode_for_glp_1_and_glucose_b200", "libhode.so")
os.rename(lib, lib + ".bak")
try:
    subprocess.run(["nvcc", "-shared", "-o", lib, *objs, "/tmp/hode_rollout_tc_tl.o"], check=True)
    import torch
    from hybrid_ode_for_glp_1_and_glucose_b200 import _lib, ops
    from hybrid_ode_for_glp_1_and_glucose_b200.synthetic import THETA_DEFAULT, cohort, random_mlp
    dev = torch.device("cuda:0")
    B = int(os.environ.get("TL_B", 131072))
    y0, t, ins = cohort(B, 61, seed=1000)
    W = random_mlp(64, 4, seed=1234, out_std=0.05)
    tt = lambda a: torch.from_numpy(a).to(dev)
    args = (tt(y0), tt(t), {k: tt(v) for k, v in ins.items()}, tt(THETA_DEFAULT), tt(W))
    L = _lib.lib()
    L.hode_debug_timeline.restype = ctypes.c_int
    buf = np.zeros(2 * 16384, dtype=np.int64)
    order = None
    for it in range(2):
        _, info = ops.rollout(*args, solver="dopri5", precision=PRECISION, device=dev, order=order)
        torch.cuda.synchronize()
        n = L.hode_debug_timeline(buf.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), 16384)
        if os.environ.get("TL_ORDER"):
            order = ops.launch_order(info)   # second pass: longest first
    ev = buf[: 2 * n].reshape(n, 2)
    ids, clk = ev[:, 0], ev[:, 1]
    # skip the first 20 % (start-up), analyse deltas between consecutive events
    s = n // 5
    names = {}
    for a, b, d in zip(ids[s:-1], ids[s + 1:], np.diff(clk[s:])):
        names.setdefault((int(a), int(b)), []).append(int(d))
    tot = 0.0
    rows = []
    for k, v in sorted(names.items()):
        v = np.array(v)
        rows.append((k, len(v), np.median(v), v.mean(), np.percentile(v, 90)))
    n_evals = sum(1 for x in ids[s:] if x == 0)
    print(f"events {n}, evaluations analysed {n_evals}")
    for k, cnt, med, mean, p90 in rows:
        print(f"{k[0]:3d}->{k[1]:3d}  n={cnt:5d}  median {med:8.0f}  mean {mean:8.0f}  p90 {p90:8.0f}   per-eval {mean * cnt / max(n_evals, 1):8.0f}")
        tot += mean * cnt
    print(f"total cycles per evaluation: {tot / max(n_evals, 1):.0f}")
finally:
    os.replace(lib + ".bak", lib)
