#!/usr/bin/env python
"""The default rollout ('f16bf16x2') in its launch shapes on one cohort: HODE_H16_TILES=3 (three tiles per SM) against 2 (two
tiles + a helper warpgroup per tile); bit-identity of the results is checked.  (profiles/r02_shared_helper_dead_end.txt was made
with a third shape, three tiles + one SHARED helper warpgroup — profiles/r02_shared_helper_variant.patch — which is not in the tree.)
Usage (GPU box): python tools/time_launch_shapes.py [B] > gpurun_out/launch_shapes.txt"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybrid_ode_for_glp_1_and_glucose_b200 import ops
from hybrid_ode_for_glp_1_and_glucose_b200.synthetic import THETA_DEFAULT, cohort, random_mlp

dev = torch.device("cuda:0")
tt = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
W = random_mlp(64, 4, seed=1234, out_std=0.05)
y0, t, ins = cohort(B, 61, seed=1000)
args = (tt(y0), tt(t), {k: tt(v) for k, v in ins.items()}, tt(THETA_DEFAULT), tt(W))
res = {}
for shape in ("3", "2"):
    os.environ["HODE_H16_TILES"] = shape
    for solver, kw in (("rk4", dict(n_substeps=2)), ("dopri5", {})):
        tr, info = ops.rollout(*args, solver=solver, precision="f16bf16x2", device=dev, **kw)
        torch.cuda.synchronize()
        order = ops.launch_order(info) if solver == "dopri5" else None
        times = {}
        for label, o in (("arrival", None), ("longest first", order)):
            if solver == "rk4" and o is None and label != "arrival":
                continue
            if label == "longest first" and o is None:
                continue
            for _ in range(2):
                ops.rollout(*args, solver=solver, precision="f16bf16x2", device=dev, order=o, **kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 5
            e0.record()
            for _ in range(n):
                tr2, info2 = ops.rollout(*args, solver=solver, precision="f16bf16x2", device=dev, order=o, **kw)
            e1.record(); torch.cuda.synchronize()
            times[label] = e0.elapsed_time(e1) / n
        att = int(info.n_attempts.sum())
        res[shape, solver] = tr
        print(f"B={B} shape {shape:2s} {solver:6s}: " + "  ".join(f"{k} {v:7.3f} ms ({att / v / 1e3:7.1f} M steps/s)" for k, v in times.items()), flush=True)
for solver in ("rk4", "dopri5"):
    print(solver, "bit-identical 3/2:", torch.equal(res["3", solver], res["2", solver]))
