#!/usr/bin/env python
"""Measured error distributions behind the adaptive-parity bounds in tests/test_gpu_rollout.py and smoke():
per trajectory, e = max over (t, state) of |y - truth| / (atol + rtol |truth|) for the GPU kernels and for the oracle
(the reference's arithmetic), on the tests' own cohorts.  Prints p50 / p90 / max and the per-trajectory ratio."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybrid_ode_for_glp_1_and_glucose_b200 import ops
from hybrid_ode_for_glp_1_and_glucose_b200.synthetic import THETA_DEFAULT, cohort, random_mlp
from oracle import cpu_oracle as o
o.build()
dev = torch.device("cuda:0")
tt = lambda a: torch.from_numpy(np.ascontiguousarray(a))
def run(name, B, seed, W, kinks, prec):
    y0, t, ins = cohort(B, seed=seed)
    truth = o.rollout(y0, t, ins, THETA_DEFAULT, W, rhs="f64", rtol=1e-11, atol=1e-13, kinks="clip", n_threads=16)[0].astype(np.float64)
    orc = o.rollout(y0, t, ins, THETA_DEFAULT, W, kinks=kinks, n_threads=16)[0]
    tr, info = ops.rollout(tt(y0), tt(t), {k: tt(v) for k, v in ins.items()}, tt(THETA_DEFAULT), None if W is None else tt(W),
                           solver="dopri5", kinks=kinks, precision=prec, device=dev)
    tr = tr.cpu().numpy()
    sc = 1e-8 + 1e-6 * np.abs(truth)
    eg = (np.abs(tr - truth) / sc).max(axis=(1, 2)); ec = (np.abs(orc - truth) / sc).max(axis=(1, 2))
    eo = (np.abs(tr - orc.astype(np.float64)) / sc).max(axis=(1, 2))
    q = lambda v: f"p50 {np.percentile(v, 50):8.1f} p90 {np.percentile(v, 90):8.1f} p99 {np.percentile(v, 99):8.1f} max {v.max():8.1f}"
    print(f"{name:34s} gpu: {q(eg)} | oracle: {q(ec)} | gpu-vs-oracle: {q(eo)} | ratio p50 {np.median(eg)/np.median(ec):.2f} p90 {np.percentile(eg,90)/np.percentile(ec,90):.2f} max {eg.max()/ec.max():.2f}")
for prec in (sys.argv[1:] or ("fp32", "tf32x3", "tf32x2bf16", "f16bf16x2")):
    for kinks in ("clip", "scipy"):
        run(f"hybrid {prec} {kinks} (seed 8)", 256, 8, random_mlp(seed=9, out_std=0.02), kinks, prec)
    run(f"hybrid {prec} clip out_std .05 (s0)", 256, 0, random_mlp(seed=1, out_std=0.05), "clip", prec)
    run(f"hybrid {prec} clip 2048 (seed 11)", 2048, 11, random_mlp(seed=12, out_std=0.05), "clip", prec)
run("mech clip (seed 7)", 512, 7, None, "clip", "fp32")
run("mech scipy (seed 7)", 512, 7, None, "scipy", "fp32")
