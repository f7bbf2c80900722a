#!/usr/bin/env python
"""Scheduling facts of the tensor-core rollout (one GPU):
  1. CUDA-event split of a training-shaped step (rollout / recording rollout / adjoint), default precision;
  2. how well a MECHANISTIC-only pilot pass (no network, FP32 CUDA cores) predicts the hybrid pass's attempt counts,
     and what a launch order derived from it is worth on the first pass over a cohort;
  3. round latency of a tile as a function of the tiles resident on its SM (148 x 128 / 256 / 384 copies of the
     cohort's longest trajectory: every lane runs the same number of attempts)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from hybrid_ode_for_glp_1_and_glucose_b200 import ops
from hybrid_ode_for_glp_1_and_glucose_b200.synthetic import THETA_DEFAULT, cohort, random_mlp

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
tt = lambda a: torch.from_numpy(a).to(dev)
y0, t, ins = cohort(B, 61, seed=1000)
W = random_mlp(64, 4, seed=1234, out_std=0.05)
a = (tt(y0), tt(t), {k: tt(v) for k, v in ins.items()}, tt(THETA_DEFAULT), tt(W))
kw = dict(solver="dopri5", device=dev)


def timed(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


# ---- 1. split ------------------------------------------------------------------------------------------------
g = torch.full((B, 61, 6), 1.0 / (B * 366), device=dev)
ms_plain, (_, info) = timed(lambda: ops.rollout(*a, **kw))
att = (info.n_accept + info.n_reject).to(torch.float64)
A = float(att.sum())
order = ops.launch_order(info)
ms_ord, _ = timed(lambda: ops.rollout(*a, order=order, **kw))
ms_rec, (_, _, tape) = timed(lambda: ops.rollout(*a, save_steps=True, order=order, **kw))
ms_adj, _ = timed(lambda: ops.rollout_bwd(tape, g))
print(f"B={B} attempts/traj {A / B:.2f} max {int(att.max())}")
print(f"rollout arrival order   {ms_plain:8.3f} ms  {A / ms_plain / 1e3:8.1f} M steps/s")
print(f"rollout longest first   {ms_ord:8.3f} ms  {A / ms_ord / 1e3:8.1f} M steps/s")
print(f"recording rollout (ord) {ms_rec:8.3f} ms")
print(f"adjoint                 {ms_adj:8.3f} ms")
print(f"fwd + adjoint           {ms_rec + ms_adj:8.3f} ms  {A / (ms_rec + ms_adj) / 1e3:8.1f} M steps/s")

# ---- 2. mechanistic pilot ------------------------------------------------------------------------------------
ms_pilot, (_, pinfo) = timed(lambda: ops.rollout(a[0], a[1], a[2], a[3], None, **kw))
patt = (pinfo.n_accept + pinfo.n_reject).to(torch.float64)
c = torch.corrcoef(torch.stack([att, patt]))[0, 1].item()
porder = ops.launch_order(pinfo)
ms_pord, _ = timed(lambda: ops.rollout(*a, order=porder, **kw))
print(f"mechanistic pilot       {ms_pilot:8.3f} ms  attempts/traj {float(patt.mean()):.2f}  corr with hybrid {c:.3f}")
print(f"rollout, pilot's order  {ms_pord:8.3f} ms  {A / ms_pord / 1e3:8.1f} M steps/s   (+ pilot: {A / (ms_pord + ms_pilot) / 1e3:8.1f})")
for rt in (1e-3, 1e-4):
    ms_p2, (_, p2) = timed(lambda: ops.rollout(a[0], a[1], a[2], a[3], None, solver="dopri5", rtol=rt, atol=rt * 1e-2, device=dev))
    p2a = (p2.n_accept + p2.n_reject).to(torch.float64)
    c2 = torch.corrcoef(torch.stack([att, p2a]))[0, 1].item()
    o2 = ops.launch_order(p2)
    ms_o2, _ = timed(lambda: ops.rollout(*a, order=o2, **kw))
    print(f"  pilot rtol {rt:g}: {ms_p2:6.3f} ms corr {c2:.3f} -> rollout {ms_o2:8.3f} ms ({A / (ms_o2 + ms_p2) / 1e3:8.1f} M with the pilot)")

# ---- 3. round latency vs resident tiles -------------------------------------------------------------------------
imax = int(torch.argmax(att))
for tiles in (1, 2, 3):
    n = 148 * 128 * tiles
    idx = torch.full((n,), imax, device=dev, dtype=torch.long)
    b = (a[0][idx].contiguous(), a[1], {k: v[idx].contiguous() for k, v in a[2].items()}, a[3], a[4])
    ms, (_, inf) = timed(lambda: ops.rollout(*b, **kw))
    na = int((inf.n_accept + inf.n_reject)[0])
    print(f"{tiles} tile(s)/SM x 128 copies of the longest trajectory ({na} attempts): {ms:7.3f} ms = {ms * 1e3 / na:6.2f} us per round")
