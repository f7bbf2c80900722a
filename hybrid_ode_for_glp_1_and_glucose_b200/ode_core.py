"""Host-side mirror of the reference's mechanistic ODE module (models/ode_core.py).

Holds the 17 scalar parameters as float32 buffers under the reference's names (so
`state_dict` keys match) and evaluates f_physio through libhode's `hode_rhs`
kernel.  The arithmetic itself lives in csrc/hode_common.cuh::rhs_mech.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

# name -> default, in buffer-registration order = the theta layout of include/hode.h
# (values: reference models/ode_core.py:44-71)
THETA_SPEC = (
    ("a_GI", 0.0104), ("k_I", 0.025), ("rho", 0.003), ("G_b", 5.0), ("I_b", 60.0),
    ("E_max", 0.1), ("EC_50", 50.0), ("Glu_b", 80.0),
    ("V_max", 9.0), ("K_m", 7.0), ("k_L", 0.02),
    ("k_GE0", 0.01), ("IGD_50", 1000.0), ("g", 2.0),
    ("p_7", 0.05), ("p_8", 0.001), ("p_9", 0.01),
)
THETA_NAMES = tuple(n for n, _ in THETA_SPEC)
STATE_NAMES = ("Glucose", "Insulin", "Glucagon", "GLP1", "GE", "FFA")


class ODECore(nn.Module):
    """6-state GLP-1 / glucose mechanistic right-hand side, evaluated on the GPU."""

    def __init__(self, params: Optional[Dict[str, float]] = None):
        super().__init__()
        values = dict(THETA_SPEC)
        extra = {}
        for k, v in (params or {}).items():
            (values if k in values else extra)[k] = v
        for name in THETA_NAMES:
            self.register_buffer(name, torch.tensor(float(values[name]), dtype=torch.float32))
        # the reference registers unknown keys as extra buffers too (dict.update + loop)
        for name, v in extra.items():
            self.register_buffer(name, torch.tensor(float(v), dtype=torch.float32))

    def theta(self, overrides: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
        """The packed [17] parameter vector (include/hode.h order)."""
        vals = []
        for name in THETA_NAMES:
            v = overrides.get(name) if overrides else None
            v = getattr(self, name) if v is None else v
            vals.append(torch.as_tensor(v, dtype=torch.float32).reshape(()).to(self.a_GI.device))
        return torch.stack(vals)

    def forward(self, t: torch.Tensor, state: torch.Tensor,
                external_inputs: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
        from . import autograd_ops
        squeeze = state.dim() == 1
        st = state.unsqueeze(0) if squeeze else state
        out = autograd_ops.rhs(t, st, external_inputs, self.theta(), None, 0, 0)
        return out.squeeze(0) if squeeze else out

    def get_steady_state(self, external_inputs=None) -> torch.Tensor:
        ss = torch.zeros(6)
        ss[0], ss[1], ss[2], ss[5] = self.G_b, self.I_b, self.Glu_b, 1.0
        return ss

    def check_mass_balance(self, state: torch.Tensor, derivatives: torch.Tensor):
        G, I = state[..., 0], state[..., 1]
        return {"non_negative": (state >= 0).all(),
                "glucose_range": (G >= 2.0) & (G <= 30.0),
                "insulin_range": (I >= 0.0) & (I <= 1000.0)}
