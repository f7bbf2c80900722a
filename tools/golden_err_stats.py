#!/usr/bin/env python
"""Observed values behind the fixed bounds of tests/test_gpu_rollout.py's golden-fixture assertions (error in units of
the solver's local tolerance, atol + rtol |ref|): printed so that the bounds can be set a small factor above them."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import golden, golden_inputs, scaled_err, cohort
from hybrid_ode_for_glp_1_and_glucose_b200 import ops
from oracle import cpu_oracle as o
o.build()
dev = torch.device("cuda:0")
def run(d_y0, d_t, ins, theta, W, **kw):
    tin = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in (ins or {}).items()}
    kw.setdefault("precision", "fp32")
    tr, info = ops.rollout(torch.from_numpy(d_y0), torch.from_numpy(d_t), tin, torch.from_numpy(theta), None if W is None else torch.from_numpy(W), device=dev, **kw)
    return tr.cpu().numpy()
y0, t, ins = cohort(512, seed=7)
truth = o.rollout(y0, t, ins, o.THETA_DEFAULT, None, rhs="f64", rtol=1e-11, atol=1e-13, kinks="clip", n_threads=16)[0]
orc = o.rollout(y0, t, ins, o.THETA_DEFAULT, None, kinks="clip", n_threads=16)[0]
tr = run(y0, t, ins, o.THETA_DEFAULT, None, solver="dopri5", kinks="clip")
print("mech clip cohort: gpu vs truth", scaled_err(tr, truth.astype(np.float64)), " gpu vs oracle", scaled_err(tr, orc.astype(np.float64)), " oracle vs truth", scaled_err(orc, truth.astype(np.float64)))
d = golden("rollout_const_T2")
tr = run(d["y0"], d["t"], golden_inputs(d), d["theta"], d["W"], solver="rk45", kinks="scipy")
print("const_T2: vs rk45", scaled_err(tr, d["out_rk45"]), " vs dop853", scaled_err(tr, d["out_dopri5"]), " ref rk45 vs dop853", scaled_err(d["out_rk45"], d["out_dopri5"]))
trx = run(d["y0"], d["t"], golden_inputs(d), d["theta"], d["W"], solver="rk45", kinks="scipy", precision="tf32x3")
print("const_T2 tf32x3: vs rk45", scaled_err(trx, d["out_rk45"]), " vs dop853", scaled_err(trx, d["out_dopri5"]))
d = golden("rollout_fig2")
tr = run(d["y0"], d["t"], golden_inputs(d), d["theta"], None, solver="dopri5", kinks="clip")
print("fig2: vs rk45", scaled_err(tr, d["out_rk45"]), " vs dop853", scaled_err(tr, d["out_dopri5"]), " ref rk45 vs dop853", scaled_err(d["out_rk45"], d["out_dopri5"]))
d = golden("rollout_4gi_mech")
truth = o.rollout(d["y0"], d["t"], golden_inputs(d), d["theta"], None, rhs="f64", rtol=1e-11, atol=1e-13, kinks="clip")[0]
tr = run(d["y0"], d["t"], golden_inputs(d), d["theta"], None, solver="dopri5", kinks="clip")
orc = o.rollout(d["y0"], d["t"], golden_inputs(d), d["theta"], None, kinks="clip")[0]
print("4gi_mech: gpu vs truth", scaled_err(tr, truth.astype(np.float64)), " oracle vs truth", scaled_err(orc, truth.astype(np.float64)))
