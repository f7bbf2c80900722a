#!/usr/bin/env python
"""CUDA-event split of one training-shaped step: rollout with step recording vs. the adjoint call."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from hybrid_ode_for_glp_1_and_glucose_b200 import ops
from hybrid_ode_for_glp_1_and_glucose_b200.synthetic import THETA_DEFAULT, cohort, random_mlp
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
y0, t, ins = cohort(B, 61, seed=1000)
W = random_mlp(64, 4, seed=1234, out_std=0.05)
tt = lambda a: torch.from_numpy(a).to(dev)
a = (tt(y0), tt(t), {k: tt(v) for k, v in ins.items()}, tt(THETA_DEFAULT), tt(W))
g = torch.full((B, 61, 6), 1.0 / (B * 366), device=dev)
kw = dict(solver="dopri5", precision=(sys.argv[2] if len(sys.argv) > 2 else "auto"), device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
tf = tb = tp = 0.0
n = 20
for it in range(n + 3):
    ev[0].record()
    ops.rollout(*a, **kw)
    ev[1].record()
    _, info, tape = ops.rollout(*a, save_steps=True, **kw)
    ev[2].record()
    ops.rollout_bwd(tape, g)
    ev[3].record()
    torch.cuda.synchronize()
    if it >= 3:
        tp += ev[0].elapsed_time(ev[1]); tf += ev[1].elapsed_time(ev[2]); tb += ev[2].elapsed_time(ev[3])
att = float((info.n_accept.sum() + info.n_reject.sum()).item())
acc = float(info.n_accept.sum().item())
print(f"B={B} attempts/traj {att / B:.1f} accepted/traj {acc / B:.1f}")
print(f"rollout (no record)   {tp / n:8.3f} ms   {att / (tp / n * 1e-3) / 1e6:8.1f} M steps/s")
print(f"rollout (recording)   {tf / n:8.3f} ms")
print(f"adjoint call          {tb / n:8.3f} ms")
print(f"fwd + adjoint         {(tf + tb) / n:8.3f} ms   {att / ((tf + tb) / n * 1e-3) / 1e6:8.1f} M steps/s")
