"""Parameter sweeps over the mechanistic parameters: one trajectory per parameter set, all in ONE launch.

Replaces the loop of the reference's sensitivity analysis (plots/plot_all.py:124-224): 16 384 Saltelli parameter
sets, each applied with `setattr(model.ode_core, name, ...)` followed by one `model.forward` call on the same initial
state and inputs, and three scalar outputs per run (:183-187).  Here the parameter sets become the rows of a
`theta` table [P,17] handed to the rollout in per-trajectory-theta mode (hode_rollout_fwd_ex), so the sweep runs on
the tensor-core kernel with 128 parameter sets per tile instead of one 128-row tile per set.
SALib (sampling, Sobol indices) stays on the host and is not part of this package.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

from . import ops
from .ode_core import THETA_NAMES

OUTPUT_NAMES = ("Glucose AUC", "Insulin Peak", "GLP-1 Response")   # reference plots/plot_all.py:167


def theta_table(model, param_names: Sequence[str], param_samples: torch.Tensor) -> torch.Tensor:
    """[P,17] table: the model's current ODE parameters with the named columns replaced by `param_samples` [P,n]."""
    samples = torch.as_tensor(param_samples, dtype=torch.float32)
    if samples.dim() != 2 or samples.shape[1] != len(param_names):
        raise ValueError(f"param_samples must be [P,{len(param_names)}], got {tuple(samples.shape)}")
    table = model.ode_core.theta().detach().to(torch.float32).cpu().repeat(samples.shape[0], 1)
    for j, name in enumerate(param_names):
        if name not in THETA_NAMES:
            raise KeyError(f"'{name}' is not an ODE parameter (reference models/ode_core.py:44-71)")
        table[:, THETA_NAMES.index(name)] = samples[:, j]
    return table


def sweep(model, table: torch.Tensor, initial_state: torch.Tensor, time_points: torch.Tensor,
          external_inputs: Optional[Dict[str, torch.Tensor]] = None, solver: str = "dopri5", rtol: float = 1e-6,
          atol: float = 1e-8, out_state_mask: int = 0, **kernel_opts):
    """Trajectories [P,T,nc] of ONE scenario (initial_state [6], inputs [T] or [1,T]) under every row of `table`."""
    dev = model._cuda_device(initial_state, time_points, table)
    P = table.shape[0]
    y0 = initial_state.reshape(1, 6).to(dev).expand(P, 6)
    ext = None
    if external_inputs:
        ext = {k: (v.reshape(1, -1).to(dev).expand(P, -1) if torch.as_tensor(v).numel() > 1 else torch.as_tensor(v).reshape(1).expand(P))
               for k, v in external_inputs.items()}
    _, W = model.packed_parameters(None)
    traj, info = ops.rollout(y0, time_points, ext, table.to(dev), None if W is None else W.to(dev),
                             hidden=model.nn_residual.hidden_dim, layers=model.nn_residual.n_layers, solver=solver,
                             rtol=rtol, atol=atol, n_substeps=kernel_opts.get("n_substeps", model.rk4_substeps),
                             kinks=kernel_opts.get("kinks", model.kinks),
                             precision=kernel_opts.get("precision", model.precision),
                             max_steps=kernel_opts.get("max_steps", 0), device=dev, theta_per_traj=True,
                             out_state_mask=out_state_mask)
    model.last_info = info
    return traj, info


def sobol_outputs(model, param_names: Sequence[str], param_samples: torch.Tensor,
                  initial_state: Optional[torch.Tensor] = None, time_points: Optional[torch.Tensor] = None,
                  external_inputs: Optional[Dict[str, torch.Tensor]] = None, **kernel_opts) -> torch.Tensor:
    """The reference's three outputs per parameter set, [P,3] (plots/plot_all.py:164-187): glucose AUC (trapezoid,
    dx = 5/60), insulin peak, mean GLP-1 from the meal on.  Defaults are the reference's scenario: state
    [5,60,80,0,0,1], 61 points over 5 h, 75 mmol of glucose at index 6, no tVNS."""
    if initial_state is None:
        initial_state = torch.tensor([5.0, 60.0, 80.0, 0.0, 0.0, 1.0])
    if time_points is None:
        time_points = torch.linspace(0, 5, 61)
    if external_inputs is None:
        meal = torch.zeros(time_points.numel())
        meal[6] = 75.0
        external_inputs = {"meal": meal, "tVNS": torch.zeros(time_points.numel())}
    table = theta_table(model, param_names, param_samples)
    traj, _ = sweep(model, table, initial_state, time_points, external_inputs, out_state_mask=0b001011, **kernel_opts)
    g, ins, glp = traj[..., 0], traj[..., 1], traj[..., 2]          # columns 0, 1, 3 of the state
    auc = (5.0 / 60.0) * (g.sum(dim=1) - 0.5 * (g[:, 0] + g[:, -1]))   # np.trapz(y, dx)
    return torch.stack([auc, ins.max(dim=1).values, glp[:, 6:].mean(dim=1)], dim=1)
