"""torch.autograd.Function wrappers over the C ABI — the "custom op" of BASELINE.json's
north_star.  Forward and backward both run in libhode.so; nothing here computes numbers.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _lib, ops


def _dev(*tensors) -> torch.device:
    for t in tensors:
        if torch.is_tensor(t) and t.is_cuda:
            return t.device
    raise _lib.HodeError("libhode kernels run on CUDA devices only (no CPU fallback): pass CUDA "
                         "tensors or construct the model with device='cuda'")


class _Rhs(torch.autograd.Function):
    """f_physio + g_NN for a batch of states (reference models/hybrid_ode_nn.py:108-134)."""

    @staticmethod
    def forward(ctx, t, state, theta, W, meal, tvns, gd, hidden, layers, part):
        device = _dev(state, theta, W)
        inputs = {"meal": meal, "tVNS": tvns, "GD": gd}
        out = ops.rhs(t, state, inputs, theta, W, hidden, layers, device=device, part=part)
        ctx.save_for_backward(torch.as_tensor(t), state, theta,
                              W if W is not None else torch.empty(0),
                              *(x if x is not None else torch.empty(0) for x in (meal, tvns, gd)))
        ctx.meta = (hidden, layers, part, W is not None, meal is not None, tvns is not None,
                    gd is not None, device)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        t, state, theta, W, meal, tvns, gd = ctx.saved_tensors
        hidden, layers, part, has_W, has_meal, has_tvns, has_gd, device = ctx.meta
        inputs = {"meal": meal if has_meal else None, "tVNS": tvns if has_tvns else None,
                  "GD": gd if has_gd else None}
        g_state, g_theta, g_W = ops.rhs_vjp(t, state, inputs, theta, W if has_W else None,
                                            grad_out, hidden, layers, device=device, part=part)
        return (None, g_state, g_theta if ctx.needs_input_grad[2] else None,
                g_W if (has_W and ctx.needs_input_grad[3]) else None,
                None, None, None, None, None, None)


def rhs(t, state, inputs: Optional[Dict[str, torch.Tensor]], theta, W, hidden, layers):
    inputs = inputs or {}
    return _Rhs.apply(t, state, theta, W, inputs.get("meal"), inputs.get("tVNS"),
                      inputs.get("GD"), hidden, layers, 0)


def nn_only(t, state, tvns, W, hidden, layers):
    dummy_theta = torch.zeros(_lib.N_THETA, dtype=torch.float32, device=_dev(state, W))
    return _Rhs.apply(t, state, dummy_theta, W, None, tvns, None, hidden, layers, 1)


class _Rollout(torch.autograd.Function):
    """Batched IVP solve with a through-solver gradient: forward = hode_rollout_fwd with the
    accepted steps recorded, backward = hode_rollout_bwd (discrete adjoint, step sizes frozen).
    The reference's forward returns a graph-free tensor (models/hybrid_ode_nn.py:248); this is
    the gradient path BASELINE.json's north_star item (3) adds."""

    @staticmethod
    def forward(ctx, y0, theta, W, t_obs, meal, tvns, gd, opts):
        device = _dev(y0, theta, W, t_obs)
        inputs = {"meal": meal, "tVNS": tvns, "GD": gd}
        traj, info, tape = ops.rollout(y0, t_obs, inputs, theta, W, device=device, save_steps=True,
                                       **{k: v for k, v in opts.items() if k != "info"})
        opts["info"] = info
        ctx.tape = tape
        ctx.has_W = W is not None
        return traj

    @staticmethod
    def backward(ctx, grad_traj):
        g_y0, g_theta, g_W = ops.rollout_bwd(ctx.tape, grad_traj, need_y0=ctx.needs_input_grad[0])
        return (g_y0 if ctx.needs_input_grad[0] else None,
                g_theta if ctx.needs_input_grad[1] else None,
                g_W if (ctx.has_W and ctx.needs_input_grad[2]) else None,
                None, None, None, None, None)


def rollout(y0, t_obs, inputs: Optional[Dict[str, torch.Tensor]], theta, W, **opts):
    """Differentiable rollout; returns (traj, RolloutInfo).  opts: hidden, layers, solver, rtol,
    atol, n_substeps, kinks, precision, max_steps, max_saved_steps."""
    inputs = inputs or {}
    traj = _Rollout.apply(y0, theta, W, t_obs, inputs.get("meal"), inputs.get("tVNS"),
                          inputs.get("GD"), opts)
    return traj, opts.get("info")
