#!/usr/bin/env python
"""Error of the tensor-core MLP modes against the CPU oracle on the fixed-step parity cases (the 1e-5 bar of
tests/test_gpu_rollout.py::test_tc_rk4_parity) and their throughput on the bench cohort.
Usage (GPU box): python tools/tc_mode_errors.py > gpurun_out/tc_mode_errors.txt"""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import cohort, random_mlp, rel_err
from oracle import cpu_oracle as oracle
from hybrid_ode_for_glp_1_and_glucose_b200 import _lib, ops
if os.environ.get("HODE_LIB_PATH"):   # an experiment build from tools/build_variants.py
    _lib.LIB_PATH = os.environ["HODE_LIB_PATH"]
    print("library:", _lib.LIB_PATH)
ONLY_SPEED = bool(os.environ.get("ONLY_SPEED"))

dev = torch.device("cuda:0")
tt = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
modes = sys.argv[1:] or ["fp32", "tf32x3", "tf32x2bf16"]
for layers, out_std in (() if ONLY_SPEED else ((4, 0.0), (4, 0.05), (2, 0.0))):
    y0, t, ins = cohort(700, seed=21)
    W = random_mlp(64, layers, seed=22) if out_std == 0.0 else random_mlp(64, layers, seed=22, out_std=out_std)
    ref, _, _, _ = oracle.rollout(y0, t, ins, oracle.THETA_DEFAULT, W, 64, layers, solver="rk4", n_substeps=2, n_threads=8)
    truth, _, _, _ = oracle.rollout(y0, t, ins, oracle.THETA_DEFAULT, W, 64, layers, solver="rk4", n_substeps=2, n_threads=8, rhs="f64")
    for m in modes:
        tr, info = ops.rollout(tt(y0), tt(t), {k: tt(v) for k, v in ins.items()}, tt(oracle.THETA_DEFAULT), tt(W), 64, layers,
                               solver="rk4", n_substeps=2, precision=m, device=dev)
        tr = tr.cpu().numpy()
        print(f"rk4 64x{layers} out_std {out_std}: {m:11s} rel_err vs oracle(f32 RHS) {rel_err(tr, ref):.3e}   vs float64 truth {rel_err(tr, truth):.3e}"
              f"   [oracle vs truth {rel_err(ref, truth):.3e}]")
B = 262144
y0, t, ins = cohort(B, 61, seed=1000)
W = random_mlp(64, 4, seed=1234, out_std=0.05)
args = (tt(y0), tt(t), {k: tt(v) for k, v in ins.items()}, tt(oracle.THETA_DEFAULT), tt(W))
for m in modes:
    if m == "fp32":
        continue
    for _ in range(3):
        tr, info = ops.rollout(*args, solver="dopri5", precision=m, device=dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        tr, info = ops.rollout(*args, solver="dopri5", precision=m, device=dev)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    att = int(info.n_attempts.sum())
    print(f"dopri5 262144 traj: {m:11s} {ms:8.3f} ms  {att / ms / 1e3:8.1f} M trajectory-steps/s")
    order = ops.launch_order(info)
    for _ in range(2):
        ops.rollout(*args, solver="dopri5", precision=m, device=dev, order=order)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        ops.rollout(*args, solver="dopri5", precision=m, device=dev, order=order)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"   ... longest first (previous pass's attempt counts): {ms:8.3f} ms  {att / ms / 1e3:8.1f} M trajectory-steps/s")
