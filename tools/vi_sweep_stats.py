#!/usr/bin/env python
"""Where the fused posterior-predictive sweep (config 4: 64 parameter sets x 1 048 576 trajectories) spends its time: per (CTA, set)
the rounds its 384 lanes need when perfectly packed (sum of attempts / 384) against the attempts of its longest unit — the fused
mode gives every trajectory to ONE CTA for all sets and the CTA's lanes meet at a barrier between sets, so a straggler holds 384 lanes.
Usage (GPU box): python tools/vi_sweep_stats.py [B] > gpurun_out/vi_sweep_stats.txt"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from hybrid_ode_for_glp_1_and_glucose_b200 import _lib, ops
if os.environ.get("HODE_LIB_PATH"):   # an experiment build from tools/build_variants.py
    _lib.LIB_PATH = os.environ["HODE_LIB_PATH"]
    print("library:", _lib.LIB_PATH)
from hybrid_ode_for_glp_1_and_glucose_b200.synthetic import THETA_DEFAULT, cohort, random_mlp

dev = torch.device("cuda:0")
tt = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
S = 64
W = random_mlp(64, 4, seed=1234, out_std=0.05)
y0, t, ins = cohort(B, 61, seed=1000)
thS, WS = bench.vi_posterior_samples(THETA_DEFAULT, W, S)
args = (tt(y0), tt(t), {k: tt(v) for k, v in ins.items()}, tt(thS), tt(WS))
mean, std, info = ops.vi_predictive(*args, device=dev)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
mean, std, info = ops.vi_predictive(*args, device=dev)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
att = (info.n_accept + info.n_reject).to(torch.int64)            # [S, B]
st = torch.bincount(info.status.reshape(-1), minlength=5).tolist()
sms = torch.cuda.get_device_properties(dev).multi_processor_count
per = (B + sms - 1) // sms
pad = per * sms - B
a = torch.nn.functional.pad(att, (0, pad)).reshape(S, sms, per)   # [S, CTA, trajectories of the CTA]
packed = a.sum(dim=2).double() / 384.0                            # rounds when perfectly packed
longest = a.amax(dim=2).double()
print(f"B={B} S={S}: {ms:.1f} ms, {att.sum().item() / ms / 1e3:.1f} M trajectory-steps/s, attempts per unit mean {att.double().mean():.1f} "
      f"p50 {att.reshape(-1)[::97].double().quantile(0.5):.0f} p99 {att.reshape(-1)[::97].double().quantile(0.99):.0f} max {att.max().item()}; status histogram {st}")
print(f"per (CTA, set): packed rounds mean {packed.mean():.0f}; longest unit mean {longest.mean():.0f} p90 {longest.reshape(-1).quantile(0.9):.0f} max {longest.max():.0f}")
us = ms * 1e3 / S / packed.mean().item()
print(f"time per packed round: {us:.1f} us (a three-tile round with all lanes busy is ~33 us; every us above that is idle lanes)")
# lower bound of the set-synchronous schedule: a set lasts at least max(packed, longest) rounds on its CTA, the kernel the slowest CTA's sum
lb = torch.maximum(packed, longest).sum(dim=0).max().item()
print(f"set-synchronous lower bound: {lb:.0f} rounds on the slowest CTA = {lb * 33e-3:.0f} ms at 33 us per round; "
      f"without set barriers (packed only): {packed.sum(dim=0).max().item() * 33e-3:.0f} ms")
