import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from hybrid_ode_for_glp_1_and_glucose_b200 import ops
from hybrid_ode_for_glp_1_and_glucose_b200.synthetic import THETA_DEFAULT, cohort, random_mlp
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
y0, t, ins = cohort(B, 61, seed=1000)
W = random_mlp(64, 4, seed=1234, out_std=0.05)
tt = lambda a: torch.from_numpy(a).to(dev)
g = torch.full((B, 61, 6), 1.0 / (B * 366), device=dev)
_, info, tape = ops.rollout(tt(y0), tt(t), {k: tt(v) for k, v in ins.items()}, tt(THETA_DEFAULT), tt(W), solver="dopri5", precision="tf32x3", device=dev, save_steps=True)
torch.cuda.synchronize()
print("fwd ok", flush=True)
out = ops.rollout_bwd(tape, g)
torch.cuda.synchronize()
print("bwd ok", [float(o.abs().sum()) for o in out], flush=True)
