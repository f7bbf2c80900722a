#!/usr/bin/env python
"""Latency of small cohorts through the default tensor-core rollout ('f16bf16x2') in its two launch shapes:
HODE_H16_TILES=3 (three 128-trajectory tiles per SM) against =2 (two tiles with helper warps, what the launch picks by
itself when the cohort fits two tiles per SM).  Also checks that the two shapes return bit-identical trajectories.
Usage (GPU box): python tools/time_small_cohorts.py > gpurun_out/small_cohorts.txt"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybrid_ode_for_glp_1_and_glucose_b200 import ops
from hybrid_ode_for_glp_1_and_glucose_b200.synthetic import THETA_DEFAULT, cohort, random_mlp

dev = torch.device("cuda:0")
tt = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
W = random_mlp(64, 4, seed=1234, out_std=0.05)
for B in (32, 128, 1024, 8192, 32768):
    y0, t, ins = cohort(B, 61, seed=1000)
    args = (tt(y0), tt(t), {k: tt(v) for k, v in ins.items()}, tt(THETA_DEFAULT), tt(W))
    res = {}
    for shape in ("3", "2"):
        os.environ["HODE_H16_TILES"] = shape
        tr, info = ops.rollout(*args, solver="dopri5", precision="f16bf16x2", device=dev)
        order = ops.launch_order(info)
        for _ in range(3):
            ops.rollout(*args, solver="dopri5", precision="f16bf16x2", device=dev, order=order)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            tr, info = ops.rollout(*args, solver="dopri5", precision="f16bf16x2", device=dev, order=order)
        e1.record(); torch.cuda.synchronize()
        res[shape] = (e0.elapsed_time(e1) / n, tr.clone(), int(info.n_attempts.sum()), int(info.n_attempts.max()))
    same = torch.equal(res["3"][1], res["2"][1])
    att, amax = res["3"][2], res["3"][3]
    print(f"B={B:6d} longest trajectory {amax:3d} attempts: three tiles {res['3'][0]:7.3f} ms ({att / res['3'][0] / 1e3:7.1f} M steps/s)   "
          f"two tiles + helpers {res['2'][0]:7.3f} ms ({att / res['2'][0] / 1e3:7.1f} M steps/s)   bit-identical: {same}")
os.environ.pop("HODE_H16_TILES", None)
