"""TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Float64 PyTorch restatement of the reference's right-hand side and of explicit Runge-Kutta
stepping over a GIVEN step sequence, used as the gradient oracle: autograd through the
unrolled steps is the definition of the discrete adjoint that hode_rollout_bwd implements.

  rhs()      reference models/ode_core.py:122-161 (f_physio) + models/nn_residual.py:136-146
             (g_NN) + models/hybrid_ode_nn.py:125-134 (their sum and the feature vector)
  inputs_at  reference models/hybrid_ode_nn.py:217-231 (searchsorted-left + lerp + clamping)
  rollout_on_steps  SciPy RK45 = Dormand-Prince 5(4): scipy/integrate/_ivp/rk.py:14-71
             (rk_step), :538-565 (tableau, dense-output matrix P), :178-180 (dense output);
             classical RK4 for solver='rk4' (not in the reference)

Pinned to the reference by tests/test_oracle.py: rhs() against the reference-generated
fixtures tests/golden/rhs_*.npz and its autograd against tests/golden/rhs_vjp_*.npz.
Only tests/ import this module.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

F64 = torch.float64

# Dormand-Prince 5(4) (published tableau; SciPy rk.py:538-565)
DP_C = [0.0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0]
DP_A = [
    [],
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
]
DP_B = [35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84]
DP_P = [
    [1, -8048581381 / 2820520608, 8663915743 / 2820520608, -12715105075 / 11282082432],
    [0, 0, 0, 0],
    [0, 131558114200 / 32700410799, -68118460800 / 10900136933, 87487479700 / 32700410799],
    [0, -1754552775 / 470086768, 14199869525 / 1410260304, -10690763975 / 1880347072],
    [0, 127303824393 / 49829197408, -318862633887 / 49829197408, 701980252875 / 199316789632],
    [0, -282668133 / 205662961, 2019193451 / 616988883, -1453857185 / 822651844],
    [0, 40617522 / 29380423, -110615467 / 29380423, 69997945 / 29380423],
]


def unpack_mlp(W: torch.Tensor, hidden: int, layers: int):
    """Flat W (include/hode.h layout) -> [(weight [out,in], bias [out]), ...]."""
    out, off, n_in = [], 0, 9
    for l in range(layers + 1):
        n_out = 6 if l == layers else hidden
        w = W[off: off + n_out * n_in].reshape(n_out, n_in)
        off += n_out * n_in
        b = W[off: off + n_out]
        off += n_out
        out.append((w, b))
        n_in = n_out
    assert off == W.numel()
    return out


def rhs(t, y, meal, tvns, gd, theta, W, hidden: int = 64, layers: int = 4, part: str = "full"):
    """f_physio + g_NN for states y [...,6]; every argument a tensor (or float) of y's dtype."""
    (a_GI, k_I, rho, G_b, I_b, E_max, EC_50, Glu_b, V_max, K_m, k_L, k_GE0, IGD_50, g, p_7, p_8,
     p_9) = [theta[..., i] for i in range(17)]
    G, I, Glu, GLP1, GE, FFA = [y[..., i] for i in range(6)]
    meal = torch.as_tensor(meal, dtype=y.dtype)
    tvns = torch.as_tensor(tvns, dtype=y.dtype)
    gd = torch.as_tensor(gd, dtype=y.dtype)
    dI = (1.0 + rho * GLP1) * a_GI * (G - G_b) - k_I * (I - I_b)
    dGlu = -(E_max * GLP1 / (EC_50 + GLP1)) * (Glu - Glu_b)
    dGLP1 = V_max * G / (K_m + G) - k_L * GLP1
    gdg = torch.pow(gd + torch.zeros_like(G), g)
    k_GE = k_GE0 * (1.0 - gdg / (torch.pow(IGD_50, g) + gdg))
    dFFA = -p_7 * FFA - p_8 * I * FFA + p_9 * G * FFA
    dG = meal - 0.01 * (I - I_b) + 0.005 * (Glu - Glu_b) - k_GE * G
    f = torch.stack([dG, dI, dGlu, dGLP1, torch.zeros_like(G), dFFA], dim=-1)
    if W is None:
        return f
    tt = torch.as_tensor(t, dtype=y.dtype) + torch.zeros_like(G)
    x = torch.cat([tt.unsqueeze(-1), y, GLP1.unsqueeze(-1), (tvns + torch.zeros_like(G)).unsqueeze(-1)],
                  dim=-1)
    params = unpack_mlp(W, hidden, layers)
    for i, (w, b) in enumerate(params):
        x = x @ w.T + b
        if i + 1 < len(params):
            x = torch.relu(x)
    return x if part == "nn" else f + x


def inputs_at(t_obs: np.ndarray, values: Optional[np.ndarray], t: float) -> float:
    """One input channel at time t: scalar -> constant; [T] series -> the reference's lerp."""
    if values is None:
        return 0.0
    v = np.asarray(values)
    if v.ndim == 0:
        return float(v)
    idx = int(np.searchsorted(t_obs, np.float32(t), side="left"))
    if idx == 0:
        return float(v[0])
    if idx >= len(t_obs):
        return float(v[-1])
    t1, t2 = float(t_obs[idx - 1]), float(t_obs[idx])
    alpha = (float(np.float32(t)) - t1) / (t2 - t1)
    return float(v[idx - 1]) + alpha * (float(v[idx]) - float(v[idx - 1]))


def rollout_on_steps(y0: torch.Tensor, t_obs: np.ndarray, inputs: Dict[str, Optional[np.ndarray]],
                     theta: torch.Tensor, W: Optional[torch.Tensor], hidden: int, layers: int,
                     step_starts: Sequence[float], solver: str = "dopri5") -> torch.Tensor:
    """Integrate ONE trajectory over the given accepted-step start times (the last step ends at
    t_obs[-1]) and return the states at t_obs [T,6], differentiable w.r.t. y0/theta/W."""
    t_obs = np.asarray(t_obs, dtype=np.float64)
    T = len(t_obs)

    def f(t, y):
        return rhs(t, y, inputs_at(t_obs, inputs.get("meal"), t), inputs_at(t_obs, inputs.get("tVNS"), t),
                   inputs_at(t_obs, inputs.get("GD"), t), theta, W, hidden, layers)

    y = y0.to(F64)
    rows: List[Optional[torch.Tensor]] = [None] * T
    ei = 0
    while ei < T and t_obs[ei] <= t_obs[0]:
        rows[ei] = y
        ei += 1
    starts = list(step_starts) + [float(t_obs[-1])]
    if solver == "rk4":
        # n_substeps equal steps per observation interval; outputs at the interval ends
        nsub = (len(starts) - 1) // (T - 1)
        for n in range(len(starts) - 1):
            t, h = starts[n], starts[n + 1] - starts[n]
            k1 = f(t, y)
            k2 = f(t + 0.5 * h, y + 0.5 * h * k1)
            k3 = f(t + 0.5 * h, y + 0.5 * h * k2)
            k4 = f(t + h, y + h * k3)
            y = y + (h / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)
            if (n + 1) % nsub == 0:
                rows[(n + 1) // nsub] = y
        return torch.stack(rows)
    for n in range(len(starts) - 1):
        t, t_new = starts[n], starts[n + 1]
        h = t_new - t
        k = [f(t, y)]
        for i in range(1, 6):
            ys = y + h * sum(a * kk for a, kk in zip(DP_A[i], k))
            k.append(f(t_new if DP_C[i] == 1.0 else t + DP_C[i] * h, ys))
        y_new = y + h * sum(b * kk for b, kk in zip(DP_B, k))
        k.append(f(t_new, y_new))
        while ei < T and t_obs[ei] <= t_new:
            te = t_obs[ei]
            if te == t_new:
                rows[ei] = y_new
            else:
                x = (te - t) / h
                poly = sum(sum(DP_P[i][j] * x ** (j + 1) for j in range(4)) * k[i] for i in range(7))
                rows[ei] = y + h * poly
            ei += 1
        y = y_new
    assert ei == T, "step sequence does not cover the observation grid"
    return torch.stack(rows)
