"""Tensor-level entry points over the C ABI (include/hode.h).

PyTorch is plumbing here: it owns device memory and the stream; every number is computed
by libhode.so.  All compute functions require CUDA tensors and raise otherwise.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import HodeError

CHANNELS = ("meal", "tVNS", "GD")
SOLVERS = {
    # the reference maps 'dopri5' -> SciPy DOP853 (models/hybrid_ode_nn.py:174-181); here, as
    # BASELINE.json's north_star specifies, 'dopri5' is Dormand-Prince 5(4) (= SciPy 'RK45').
    "dopri5": _lib.SOLVER_DOPRI5, "rk45": _lib.SOLVER_DOPRI5, "rk4": _lib.SOLVER_RK4,
    # ... and 'dop853' is that DOP853 (scipy rk.py:568-720): float32 RHS / float64 stepping like the reference,
    # FP32 CUDA-core kernels, forward only — for callers who want the reference's actual default integrator
    "dop853": _lib.SOLVER_DOP853,
}
PRECISIONS = {"fp32": _lib.MLP_FP32, "tf32x3": _lib.MLP_TF32X3, "tf32": _lib.MLP_TF32, "tf32bf16": _lib.MLP_TF32BF16,
              "tf32x2bf16": _lib.MLP_TF32X2BF16, "f16bf16x2": _lib.MLP_F16BF16X2}


def default_precision(hidden: int, layers: int) -> str:
    """'auto' resolves to the tcgen05 kernels (split-precision passes, float32-equivalent accuracy) whenever
    the network has the shape both tensor-core kernels are compiled for (64 wide, <= 4 hidden layers: the
    reference's default, models/nn_residual.py:28-36), and to the FP32 CUDA-core kernels for any other shape.
    The tensor-core default is 'f16bf16x2' (three 128-trajectory tiles per SM; hi parts in FP16, remainders in BF16 /
    scaled FP16: three kind::f16 passes = 1.5 TF32-pass equivalents; 0.8e-6 against the oracle on the fixed-step parity
    cases — 'tf32x2bf16' 1.0-1.2e-6, 'tf32x3' 1.0e-6, both selectable: 14 % / 27 % slower).  Its products are
    float32-equivalent while |activations| and |weights| stay below FP16's 65504 (include/hode.h HODE_MLP_F16BF16X2);
    a network that leaves that range should ask for 'tf32x2bf16'.  Gradients of every tensor-core mode are computed by
    the 3xTF32 adjoint.  precision='fp32' is the bit-conservative parity mode."""
    return "f16bf16x2" if hidden == 64 and 1 <= layers <= 4 else "fp32"


def _mlp_mode(precision: str, hidden: int, layers: int) -> int:
    if precision == "auto":
        precision = default_precision(hidden, layers)
    if precision not in PRECISIONS:
        raise HodeError(f"precision '{precision}' is not one of {sorted(PRECISIONS) + ['auto']}")
    return PRECISIONS[precision]
KINKS = {"scipy": _lib.KINK_SCIPY, "clip": _lib.KINK_CLIP}


@dataclass
class RolloutInfo:
    status: torch.Tensor      # int32 [B] or [S,B]
    n_accept: torch.Tensor    # int32, accepted steps
    n_reject: torch.Tensor    # int32, rejected attempts

    @property
    def n_attempts(self) -> torch.Tensor:
        return self.n_accept + self.n_reject


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(device: torch.device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _f32c(t: torch.Tensor, device: torch.device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=torch.float32, non_blocking=True).contiguous()


def _require_cuda(device: torch.device) -> None:
    if device.type != "cuda":
        raise HodeError("libhode kernels run on CUDA devices only (no CPU fallback); got device "
                        f"'{device}'")


def prepare(y0: torch.Tensor, t_obs: torch.Tensor, inputs: Optional[Dict[str, torch.Tensor]],
            theta: torch.Tensor, W: Optional[torch.Tensor], hidden: int, layers: int,
            device: torch.device, theta_per_traj: bool = False):
    """Normalise shapes/dtypes, fill a hode_cfg.  Pure host logic (testable without a GPU)."""
    if y0.dim() != 2 or y0.shape[1] != _lib.N_STATE:
        raise ValueError(f"initial_state must be [B,6], got {tuple(y0.shape)}")
    B = y0.shape[0]
    if t_obs.dim() == 2:
        if t_obs.shape[0] != B:
            # reference models/hybrid_ode_nn.py:192-196: a leading dim != B is squeezed away
            t_obs = t_obs.squeeze()
            if t_obs.dim() != 1:
                t_obs = t_obs.flatten()
    elif t_obs.dim() != 1:
        raise ValueError(f"t_span must be [T] or [B,T], got {tuple(t_obs.shape)}")
    T = t_obs.shape[-1]
    cfg = _lib.new_cfg()
    cfg.n_traj, cfg.n_obs = B, T
    cfg.t_per_traj = 1 if t_obs.dim() == 2 else 0
    bufs = {"y0": _f32c(y0, device), "t_obs": _f32c(t_obs, device)}
    for ch, name in enumerate(CHANNELS):
        v = None if not inputs else inputs.get(name)
        if v is None:
            cfg.in_mode[ch] = _lib.IN_ABSENT
            bufs[name] = None
            continue
        if not torch.is_tensor(v):
            v = torch.as_tensor(v, dtype=torch.float32)
        if v.dim() == 2:
            if tuple(v.shape) != (B, T):
                raise ValueError(f"input '{name}' must be [B,T]=({B},{T}), got {tuple(v.shape)}")
            cfg.in_mode[ch] = _lib.IN_SERIES
        else:
            # reference models/hybrid_ode_nn.py:230-231: anything not 2-D is per-trajectory
            v = v.reshape(-1)
            if v.numel() == 1:
                v = v.expand(B)
            if v.numel() != B:
                raise ValueError(f"input '{name}' must have B={B} entries, got {v.numel()}")
            cfg.in_mode[ch] = _lib.IN_CONST
        bufs[name] = _f32c(v, device)
    theta2 = theta if theta.dim() == 2 else theta.unsqueeze(0)
    if theta2.shape[1] != _lib.N_THETA:
        raise ValueError(f"theta must have 17 entries per set, got {tuple(theta.shape)}")
    S = theta2.shape[0]
    if theta_per_traj:
        if S != B:
            raise ValueError(f"theta_per_traj: theta must be [B,17] = ({B},17), got {tuple(theta.shape)}")
        S = 1
    cfg.n_samples = S
    bufs["theta"] = _f32c(theta2, device)
    if W is None:
        cfg.mlp = _lib.MLP_NONE
        bufs["W"] = None
    else:
        P = _expected_param_count(hidden, layers)
        W2 = W if W.dim() == 2 else W.unsqueeze(0)
        if tuple(W2.shape) != (S, P):
            raise ValueError(f"W must be [{S},{P}] for hidden={hidden}, layers={layers}; got "
                             f"{tuple(W2.shape)}")
        cfg.mlp = _lib.MLP_FP32
        cfg.nn_hidden, cfg.nn_layers = hidden, layers
        bufs["W"] = _f32c(W2, device)
    return cfg, bufs


def _expected_param_count(hidden: int, layers: int) -> int:
    return 9 * hidden + hidden + (layers - 1) * (hidden * hidden + hidden) + hidden * 6 + 6


def workspace_bytes(cfg) -> Tuple[int, int]:
    fwd, bwd = ctypes.c_size_t(0), ctypes.c_size_t(0)
    _lib.check(_lib.lib().hode_workspace_bytes(ctypes.byref(cfg), ctypes.byref(fwd),
                                               ctypes.byref(bwd)), "hode_workspace_bytes")
    return int(fwd.value), int(bwd.value)


@dataclass
class RolloutTape:
    """What hode_rollout_bwd needs from a forward pass run with save_steps=1."""
    cfg: object
    bufs: Dict[str, Optional[torch.Tensor]]
    workspace: torch.Tensor
    squeeze_s: bool


def rollout(y0: torch.Tensor, t_obs: torch.Tensor, inputs: Optional[Dict[str, torch.Tensor]],
            theta: torch.Tensor, W: Optional[torch.Tensor], hidden: int = 64, layers: int = 4,
            solver: str = "dopri5", rtol: float = 1e-6, atol: float = 1e-8, n_substeps: int = 4,
            kinks: str = "clip", precision: str = "auto", max_steps: int = 0,
            device: Optional[torch.device] = None, save_steps: bool = False,
            max_saved_steps: int = 0, theta_per_traj: bool = False, order: Optional[torch.Tensor] = None,
            out_state_mask: int = 0):
    """Batched IVP solve on the GPU (hode_rollout_fwd / hode_rollout_fwd_ex).

    Returns traj [B,T,6] (or [S,B,T,6] when theta is [S,17]) and a RolloutInfo; with
    save_steps=True also a RolloutTape for rollout_bwd().
    theta_per_traj: theta is [B,17], one mechanistic parameter set per trajectory (parameter sweeps; the network is
    shared).  order: int32 [B] launch order of the trajectories (see launch_order()).  out_state_mask: bit i set =
    keep state column i (traj is then [..., T, popcount])."""
    device = torch.device(device) if device is not None else y0.device
    _require_cuda(device)
    if solver.lower() not in SOLVERS:
        raise HodeError(f"solver '{solver}' is not implemented on the GPU path; available: "
                        f"{sorted(SOLVERS)}")
    squeeze_s = theta.dim() == 1 or theta_per_traj
    cfg, bufs = prepare(y0, t_obs, inputs, theta, W, hidden, layers, device, theta_per_traj)
    cfg.solver = SOLVERS[solver.lower()]
    cfg.rtol, cfg.atol = float(rtol), float(atol)
    cfg.n_substeps = int(n_substeps)
    cfg.max_steps = int(max_steps)
    cfg.kink_mode = KINKS[kinks]
    cfg.save_steps = 1 if save_steps else 0
    cfg.max_saved_steps = int(max_saved_steps)
    if cfg.mlp != _lib.MLP_NONE:
        if cfg.solver == _lib.SOLVER_DOP853 and precision == "auto":
            precision = "fp32"   # the DOP853 kernels are the FP32 ones
        cfg.mlp = _mlp_mode(precision, hidden, layers)
    B, T, S = cfg.n_traj, cfg.n_obs, cfg.n_samples
    nc = bin(out_state_mask & 0x3F).count("1") if (out_state_mask & 0x3F) not in (0, 0x3F) else 6
    d_order = None
    if order is not None:
        d_order = order.detach().to(device=device, dtype=torch.int32).contiguous()
        if d_order.numel() != B:
            raise ValueError(f"order must have B={B} entries, got {d_order.numel()}")
    opts = _lib.new_fwd_opts(theta_per_traj, None if d_order is None else d_order.data_ptr(), out_state_mask & 0x3F)
    with torch.cuda.device(device):
        traj = torch.empty((S, B, T, nc), dtype=torch.float32, device=device)
        status = torch.empty((S, B), dtype=torch.int32, device=device)
        counters = torch.empty((2, S, B), dtype=torch.int32, device=device)
        ws_bytes = workspace_bytes(cfg)[0]
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device) if ws_bytes else None
        rc = 0 if B == 0 else _lib.lib().hode_rollout_fwd_ex(
            ctypes.byref(cfg), ctypes.byref(opts), _ptr(bufs["y0"]), _ptr(bufs["t_obs"]), _ptr(bufs["meal"]),
            _ptr(bufs["tVNS"]), _ptr(bufs["GD"]), _ptr(bufs["theta"]), _ptr(bufs["W"]),
            _ptr(traj), _ptr(status), _ptr(counters), _ptr(ws), ws_bytes, _stream(device))
    _lib.check(rc, "hode_rollout_fwd_ex")
    if save_steps and B > 0 and cfg.solver == _lib.SOLVER_DOPRI5 and max_saved_steps <= 0:
        # The record capacity was left to the library (max(128, 2 (T - 1)) accepted steps per trajectory).  A
        # trajectory that needs more is reported as ST_REC_OVERFLOW, zero-padded and left without gradient — unlike
        # the same call without save_steps.  Re-run with a larger capacity instead of returning that (one status
        # read-back per differentiable rollout; pass max_saved_steps explicitly to skip the check).
        cap = int(_lib.lib().hode_step_record_capacity(ctypes.byref(cfg)))
        limit = max_steps if max_steps > 0 else 100000
        if bool((status == _lib.ST_REC_OVERFLOW).any()) and cap < limit:
            return rollout(y0, t_obs, inputs, theta, W, hidden, layers, solver, rtol, atol, n_substeps, kinks,
                           precision, max_steps, device, save_steps, min(4 * cap, limit), theta_per_traj, order,
                           out_state_mask)
    if squeeze_s:
        out = traj[0], RolloutInfo(status[0], counters[0, 0], counters[1, 0])
    else:
        out = traj, RolloutInfo(status, counters[0], counters[1])
    if save_steps:
        return out + (RolloutTape(cfg, bufs, ws, squeeze_s),)
    return out


def launch_order(info: RolloutInfo) -> torch.Tensor:
    """int32 [B] launch order for the next rollout of the same cohort: trajectories by descending attempt count of
    the pass `info` came from (stable).  Training re-integrates a cohort every epoch (reference
    train/train_hybrid.py:225-275); handing the long trajectories out first removes the tail that bounds small
    cohorts on the tensor-core rollout."""
    att = (info.n_accept + info.n_reject).reshape(-1, info.n_accept.shape[-1])
    att = att.sum(dim=0) if att.shape[0] > 1 else att[0]
    return torch.argsort(att, descending=True, stable=True).to(torch.int32)


def rollout_bwd(tape: RolloutTape, grad_traj: torch.Tensor, need_y0: bool = True
                ) -> Tuple[Optional[torch.Tensor], torch.Tensor, Optional[torch.Tensor]]:
    """Discrete adjoint of a rollout (hode_rollout_bwd): (grad_y0 [S,B,6], grad_theta [S,17],
    grad_W [S,P] or None); the S axis is dropped when the forward's theta was 1-D."""
    cfg, bufs = tape.cfg, tape.bufs
    device = bufs["y0"].device
    _require_cuda(device)
    B, T, S = cfg.n_traj, cfg.n_obs, cfg.n_samples
    P = 0 if bufs["W"] is None else bufs["W"].shape[1]
    g = _f32c(grad_traj.reshape(S, B, T, 6), device)
    with torch.cuda.device(device):
        g_y0 = torch.empty((S, B, 6), dtype=torch.float32, device=device) if need_y0 else None
        g_theta = torch.empty((S, _lib.N_THETA), dtype=torch.float32, device=device)
        g_W = torch.empty((S, P), dtype=torch.float32, device=device) if P else None
        fwd_bytes, bwd_bytes = workspace_bytes(cfg)
        bws = torch.empty(max(bwd_bytes, 16), dtype=torch.uint8, device=device)
        if B == 0:
            g_theta.zero_()
            if g_W is not None:
                g_W.zero_()
            rc = 0
        else:
            rc = _lib.lib().hode_rollout_bwd(
                ctypes.byref(cfg), _ptr(bufs["y0"]), _ptr(bufs["t_obs"]), _ptr(bufs["meal"]),
                _ptr(bufs["tVNS"]), _ptr(bufs["GD"]), _ptr(bufs["theta"]), _ptr(bufs["W"]), _ptr(g),
                _ptr(g_y0), _ptr(g_theta), _ptr(g_W), _ptr(tape.workspace), fwd_bytes, _ptr(bws),
                bwd_bytes, _stream(device))
    _lib.check(rc, "hode_rollout_bwd")
    if tape.squeeze_s:
        return (None if g_y0 is None else g_y0[0], g_theta[0], None if g_W is None else g_W[0])
    return g_y0, g_theta, g_W


def data_loss_step(y0: torch.Tensor, t_obs: torch.Tensor, inputs: Optional[Dict[str, torch.Tensor]],
                   theta: torch.Tensor, W: Optional[torch.Tensor], obs: torch.Tensor, hidden: int = 64,
                   layers: int = 4, solver: str = "dopri5", rtol: float = 1e-6, atol: float = 1e-8,
                   n_substeps: int = 4, kinks: str = "clip", precision: str = "auto", max_steps: int = 0,
                   device: Optional[torch.device] = None, max_saved_steps: int = 0, need_y0: bool = False):
    """loss = mean((rollout - obs)^2) and its gradients in ONE library call (hode_loss_fused_fwd_bwd):
    returns (loss [S] or scalar, grad_y0 or None, grad_theta, grad_W or None, traj, RolloutInfo)."""
    device = torch.device(device) if device is not None else y0.device
    _require_cuda(device)
    if solver.lower() not in SOLVERS:
        raise HodeError(f"solver '{solver}' is not implemented on the GPU path; available: {sorted(SOLVERS)}")
    squeeze_s = theta.dim() == 1
    cfg, bufs = prepare(y0, t_obs, inputs, theta, W, hidden, layers, device)
    cfg.solver = SOLVERS[solver.lower()]
    cfg.rtol, cfg.atol = float(rtol), float(atol)
    cfg.n_substeps = int(n_substeps)
    cfg.max_steps = int(max_steps)
    cfg.kink_mode = KINKS[kinks]
    cfg.save_steps = 1
    cfg.max_saved_steps = int(max_saved_steps)
    if cfg.mlp != _lib.MLP_NONE:
        cfg.mlp = _mlp_mode(precision, hidden, layers)
    B, T, S = cfg.n_traj, cfg.n_obs, cfg.n_samples
    P = 0 if bufs["W"] is None else bufs["W"].shape[1]
    o = _f32c(obs.reshape(B, T, 6), device)
    with torch.cuda.device(device):
        traj = torch.empty((S, B, T, 6), dtype=torch.float32, device=device)
        g_traj = torch.empty((S, B, T, 6), dtype=torch.float32, device=device)
        status = torch.empty((S, B), dtype=torch.int32, device=device)
        counters = torch.empty((2, S, B), dtype=torch.int32, device=device)
        loss = torch.empty(S, dtype=torch.float32, device=device)
        g_y0 = torch.empty((S, B, 6), dtype=torch.float32, device=device) if need_y0 else None
        g_theta = torch.empty((S, _lib.N_THETA), dtype=torch.float32, device=device)
        g_W = torch.empty((S, P), dtype=torch.float32, device=device) if P else None
        fwd_bytes, bwd_bytes = workspace_bytes(cfg)
        bwd_bytes = max(bwd_bytes, 4 * S * ((B * T * 6 + 4095) // 4096), 16)
        ws = torch.empty(max(fwd_bytes, 16), dtype=torch.uint8, device=device)
        bws = torch.empty(bwd_bytes, dtype=torch.uint8, device=device)
        rc = _lib.lib().hode_loss_fused_fwd_bwd(
            ctypes.byref(cfg), _ptr(bufs["y0"]), _ptr(bufs["t_obs"]), _ptr(bufs["meal"]), _ptr(bufs["tVNS"]),
            _ptr(bufs["GD"]), _ptr(bufs["theta"]), _ptr(bufs["W"]), _ptr(o), _ptr(traj), _ptr(status),
            _ptr(counters), _ptr(loss), _ptr(g_traj), _ptr(g_y0), _ptr(g_theta), _ptr(g_W), _ptr(ws), fwd_bytes,
            _ptr(bws), bwd_bytes, _stream(device))
    _lib.check(rc, "hode_loss_fused_fwd_bwd")
    if squeeze_s:
        return (loss[0], None if g_y0 is None else g_y0[0], g_theta[0], None if g_W is None else g_W[0], traj[0],
                RolloutInfo(status[0], counters[0, 0], counters[1, 0]))
    return loss, g_y0, g_theta, g_W, traj, RolloutInfo(status, counters[0], counters[1])


PATIENT_TYPES = {"T2DM": 0, "HV": 1}


def generate_4gi(baselines: torch.Tensor, meal_rate: Optional[torch.Tensor] = None, n_obs: int = 61,
                 interval_hours: float = 5.0 / 60.0, patient_type: str = "T2DM", rtol: float = 0.0, atol: float = 0.0,
                 device: Optional[torch.device] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """The reference's 4GI simulator for a whole cohort in one launch (hode_generate_4gi; reference
    data/generate4GI.py FourGIModel.simulate).  baselines [N,5] (glucose, insulin, GLP-1, glucagon, GIP),
    meal_rate [N, n_obs-1] in mmol/h per sampling interval (None = no meals).
    Returns (concentrations [N, n_obs, 5], status [N])."""
    device = torch.device(device) if device is not None else baselines.device
    _require_cuda(device)
    if patient_type not in PATIENT_TYPES:
        raise HodeError(f"patient_type must be one of {sorted(PATIENT_TYPES)}")
    b = _f32c(baselines.reshape(-1, 5), device)
    N = b.shape[0]
    m = None
    if meal_rate is not None:
        m = _f32c(meal_rate.reshape(N, n_obs - 1), device)
    with torch.cuda.device(device):
        out = torch.empty((N, n_obs, 5), dtype=torch.float32, device=device)
        status = torch.empty(N, dtype=torch.int32, device=device)
        rc = _lib.lib().hode_generate_4gi(N, int(n_obs), float(interval_hours), PATIENT_TYPES[patient_type], float(rtol),
                                          float(atol), _ptr(b), _ptr(m), _ptr(out), _ptr(status), _stream(device))
    _lib.check(rc, "hode_generate_4gi")
    return out, status


def saved_steps(tape: RolloutTape):
    """(n [S*B] int32, t [max_saved, S*B] float64) views of the recorded accepted steps
    (diagnostics / tests: the step sequence the adjoint differentiates)."""
    cfg = tape.cfg
    units = cfg.n_samples * cfg.n_traj
    L = _lib.lib()
    max_saved = int(L.hode_step_record_capacity(ctypes.byref(cfg)))
    rec_bytes = 4 * int(L.hode_step_record_floats(ctypes.byref(cfg)))
    al = lambda x: (x + 255) // 256 * 256
    ws = tape.workspace
    n = ws[: units * 4].view(torch.int32)
    # one record per (unit, step): { t f64, h f32, pad, y[6], k1[6] (, k2..k6) } (csrc/hode_common.cuh step_rec)
    off_rec = al(units * 4)
    rec = ws[off_rec: off_rec + units * max_saved * rec_bytes].view(torch.float64).reshape(units, max_saved, rec_bytes // 8)
    return n, rec[:, :, 0].permute(1, 0)


def vi_predictive(y0: torch.Tensor, t_obs: torch.Tensor, inputs: Optional[Dict[str, torch.Tensor]],
                  theta: torch.Tensor, W: Optional[torch.Tensor], hidden: int = 64, layers: int = 4,
                  solver: str = "dopri5", rtol: float = 1e-6, atol: float = 1e-8, n_substeps: int = 4,
                  kinks: str = "clip", precision: str = "auto", max_steps: int = 0,
                  device: Optional[torch.device] = None
                  ) -> Tuple[torch.Tensor, torch.Tensor, RolloutInfo]:
    """Posterior-predictive mean and unbiased std over the S parameter sets theta [S,17] /
    W [S,P], reduced inside the rollout kernel (hode_vi_predictive); the [S,B,T,6] stack of
    reference inference/vi.py:306 is never materialised.  Returns (mean, std, info), [B,T,6]."""
    device = torch.device(device) if device is not None else y0.device
    _require_cuda(device)
    if solver.lower() not in SOLVERS:
        raise HodeError(f"solver '{solver}' is not implemented on the GPU path; available: "
                        f"{sorted(SOLVERS)}")
    if theta.dim() != 2:
        raise ValueError("vi_predictive() needs stacked parameter sets: theta [S,17], W [S,P]")
    cfg, bufs = prepare(y0, t_obs, inputs, theta, W, hidden, layers, device)
    cfg.solver = SOLVERS[solver.lower()]
    cfg.rtol, cfg.atol = float(rtol), float(atol)
    cfg.n_substeps = int(n_substeps)
    cfg.max_steps = int(max_steps)
    cfg.kink_mode = KINKS[kinks]
    if cfg.mlp != _lib.MLP_NONE:
        cfg.mlp = _mlp_mode(precision, hidden, layers)
    B, T, S = cfg.n_traj, cfg.n_obs, cfg.n_samples
    with torch.cuda.device(device):
        mean = torch.empty((B, T, 6), dtype=torch.float32, device=device)
        std = torch.empty((B, T, 6), dtype=torch.float32, device=device)
        status = torch.empty((S, B), dtype=torch.int32, device=device)
        counters = torch.empty((2, S, B), dtype=torch.int32, device=device)
        ws_bytes = workspace_bytes(cfg)[0]
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device) if ws_bytes else None
        rc = 0 if B == 0 else _lib.lib().hode_vi_predictive(
            ctypes.byref(cfg), _ptr(bufs["y0"]), _ptr(bufs["t_obs"]), _ptr(bufs["meal"]),
            _ptr(bufs["tVNS"]), _ptr(bufs["GD"]), _ptr(bufs["theta"]), _ptr(bufs["W"]),
            _ptr(mean), _ptr(std), _ptr(status), _ptr(counters), _ptr(ws), ws_bytes,
            _stream(device))
    _lib.check(rc, "hode_vi_predictive")
    return mean, std, RolloutInfo(status, counters[0], counters[1])


def rhs(t: torch.Tensor, state: torch.Tensor, inputs: Optional[Dict[str, torch.Tensor]],
        theta: torch.Tensor, W: Optional[torch.Tensor], hidden: int = 64, layers: int = 4,
        device: Optional[torch.device] = None, part: int = 0) -> torch.Tensor:
    """One batched evaluation of f_physio + g_NN on the GPU (hode_rhs). state [B,6] -> [B,6]."""
    device = torch.device(device) if device is not None else state.device
    _require_cuda(device)
    B = state.shape[0]
    t = torch.as_tensor(t, dtype=torch.float32)
    t = t.reshape(-1)
    if t.numel() == 1:
        t = t.expand(B)
    cfg, bufs = prepare(state, t, inputs, theta, W, hidden, layers, device)
    if cfg.n_samples != 1:
        raise ValueError("rhs() takes one parameter set")
    cfg.n_obs, cfg.t_per_traj = 1, 1  # `t` is one evaluation time per row, not a grid
    cfg.rhs_part = part
    with torch.cuda.device(device):
        out = torch.empty((B, 6), dtype=torch.float32, device=device)
        rc = _lib.lib().hode_rhs(
            ctypes.byref(cfg), _ptr(bufs["t_obs"]), _ptr(bufs["y0"]), _ptr(bufs["meal"]),
            _ptr(bufs["tVNS"]), _ptr(bufs["GD"]), _ptr(bufs["theta"]), _ptr(bufs["W"]),
            _ptr(out), _stream(device))
    _lib.check(rc, "hode_rhs")
    return out


def rhs_vjp(t, state, inputs, theta, W, grad_out, hidden=64, layers=4, device=None, part=0):
    """Vector-Jacobian product of rhs() (hode_rhs_vjp): returns (grad_state [B,6],
    grad_theta [17], grad_W [P] or None)."""
    device = torch.device(device) if device is not None else state.device
    _require_cuda(device)
    B = state.shape[0]
    t = torch.as_tensor(t, dtype=torch.float32).reshape(-1)
    if t.numel() == 1:
        t = t.expand(B)
    cfg, bufs = prepare(state, t, inputs, theta, W, hidden, layers, device)
    if cfg.n_samples != 1:
        raise ValueError("rhs_vjp() takes one parameter set")
    cfg.n_obs, cfg.t_per_traj = 1, 1
    cfg.rhs_part = part
    P = 0 if bufs["W"] is None else bufs["W"].shape[1]
    g = _f32c(grad_out.reshape(B, 6), device)
    with torch.cuda.device(device):
        g_state = torch.empty((B, 6), dtype=torch.float32, device=device)
        g_theta = torch.empty(_lib.N_THETA, dtype=torch.float32, device=device)
        g_W = torch.empty(P, dtype=torch.float32, device=device) if P else None
        bwd_bytes = workspace_bytes(cfg)[1]
        bws = torch.empty(max(bwd_bytes, 16), dtype=torch.uint8, device=device)
        rc = _lib.lib().hode_rhs_vjp(
            ctypes.byref(cfg), _ptr(bufs["t_obs"]), _ptr(bufs["y0"]), _ptr(bufs["meal"]),
            _ptr(bufs["tVNS"]), _ptr(bufs["GD"]), _ptr(bufs["theta"]), _ptr(bufs["W"]), _ptr(g),
            _ptr(g_state), _ptr(g_theta), _ptr(g_W), _ptr(bws), bwd_bytes, _stream(device))
    _lib.check(rc, "hode_rhs_vjp")
    return g_state, g_theta, g_W
