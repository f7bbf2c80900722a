"""Host-side mirror of the reference's residual network module (models/nn_residual.py).

Owns the parameters under the reference's `network.{0,2,4,...}.{weight,bias}` names and
packs them into the flat W layout of include/hode.h.  The forward arithmetic runs inside
libhode (csrc/hode_rollout_simt.cu::mlp_eval_*); only ReLU without dropout is compiled
into the kernels, other activations are rejected explicitly.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn


class NNResidual(nn.Module):
    """9 -> hidden x n_layers -> 6 ReLU MLP whose output layer starts at zero."""

    def __init__(self, input_dim: int = 9, hidden_dim: int = 64, output_dim: int = 6,
                 n_layers: int = 4, activation: str = "relu", dropout: float = 0.0):
        super().__init__()
        if input_dim != 9 or output_dim != 6:
            raise ValueError("libhode compiles the reference feature layout: input_dim=9, "
                             "output_dim=6")
        if activation != "relu" or dropout != 0.0:
            raise NotImplementedError("the CUDA kernels implement ReLU without dropout (the "
                                      "configuration every reference call site uses)")
        if not (1 <= hidden_dim <= 128 and 1 <= n_layers <= 8):
            raise NotImplementedError("hidden_dim must be in [1,128] and n_layers in [1,8]")
        self.input_dim, self.hidden_dim, self.output_dim = input_dim, hidden_dim, output_dim
        self.n_layers, self.dropout = n_layers, dropout
        self.activation = nn.ReLU()
        widths = [input_dim] + [hidden_dim] * n_layers
        mods = []
        for fan_in, fan_out in zip(widths[:-1], widths[1:]):
            mods += [nn.Linear(fan_in, fan_out), self.activation]
        mods.append(nn.Linear(hidden_dim, output_dim))
        self.network = nn.Sequential(*mods)
        self._initialize_zero_output()

    def _initialize_zero_output(self) -> None:
        # reference models/nn_residual.py:83-98: zero head, Xavier-normal(gain 0.1) body
        linears = self.linears()
        with torch.no_grad():
            for lin in linears[:-1]:
                nn.init.xavier_normal_(lin.weight, gain=0.1)
                lin.bias.zero_()
            linears[-1].weight.zero_()
            linears[-1].bias.zero_()

    def linears(self):
        return [m for m in self.network if isinstance(m, nn.Linear)]

    def packed(self, overrides: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
        """Flat parameter vector in named_parameters() order (the W layout of include/hode.h).
        Differentiable w.r.t. the module parameters."""
        parts = []
        for name, p in self.named_parameters():
            v = overrides.get(name) if overrides else None
            parts.append((p if v is None else v.to(p.device)).reshape(-1))
        return torch.cat(parts)

    def is_identically_zero(self) -> bool:
        last = self.linears()[-1]
        return bool((last.weight == 0).all() and (last.bias == 0).all())

    def forward(self, t: torch.Tensor, state: torch.Tensor, glp1: torch.Tensor,
                tvns: torch.Tensor) -> torch.Tensor:
        """g_NN alone.  `glp1` must be state[...,3] as at every reference call site
        (models/hybrid_ode_nn.py:125-129); the kernels read it from the state."""
        from . import autograd_ops
        squeeze = state.dim() == 1
        st = state.unsqueeze(0) if squeeze else state
        out = autograd_ops.nn_only(t, st, tvns, self.packed(), self.hidden_dim, self.n_layers)
        return out.squeeze(0) if squeeze else out

    def regularization_loss(self, l2_weight: float = 1e-4, sparsity_weight: float = 0.0):
        # reference models/nn_residual.py:198-223 (weights only, no biases)
        if l2_weight <= 0:
            return 0.0
        return l2_weight * sum(lin.weight.pow(2).sum() for lin in self.linears())
