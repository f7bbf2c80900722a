"""TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-end of oracle/hode_oracle.c, the plain-C CPU restatement of the reference's
rollout path (reference models/hybrid_ode_nn.py:136-261 + SciPy RK45).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The product package never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libhode_oracle.so")

SOLVER_RK4, SOLVER_DOPRI5, SOLVER_DOP853 = 0, 1, 2
IN_ABSENT, IN_CONST, IN_SERIES = 0, 1, 2
MLP_NONE, MLP_FP32 = 0, 1
RHS_F32, RHS_F64 = 0, 1

THETA_NAMES = ["a_GI", "k_I", "rho", "G_b", "I_b", "E_max", "EC_50", "Glu_b", "V_max", "K_m",
               "k_L", "k_GE0", "IGD_50", "g", "p_7", "p_8", "p_9"]
THETA_DEFAULT = np.array([0.0104, 0.025, 0.003, 5.0, 60.0, 0.1, 50.0, 80.0, 9.0, 7.0, 0.02,
                          0.01, 1000.0, 2.0, 0.05, 0.001, 0.01], dtype=np.float32)


class HodeCfg(ctypes.Structure):
    """Mirror of `struct hode_cfg` in include/hode.h."""
    _fields_ = [
        ("struct_bytes", ctypes.c_int32), ("n_traj", ctypes.c_int32), ("n_obs", ctypes.c_int32),
        ("t_per_traj", ctypes.c_int32), ("in_mode", ctypes.c_int32 * 3),
        ("nn_hidden", ctypes.c_int32), ("nn_layers", ctypes.c_int32), ("mlp", ctypes.c_int32),
        ("n_samples", ctypes.c_int32), ("solver", ctypes.c_int32), ("n_substeps", ctypes.c_int32),
        ("max_steps", ctypes.c_int32), ("rtol", ctypes.c_double), ("atol", ctypes.c_double),
        ("save_steps", ctypes.c_int32), ("kink_mode", ctypes.c_int32),
        ("max_saved_steps", ctypes.c_int32), ("rhs_part", ctypes.c_int32),
    ]


def build(force: bool = False) -> str:
    """Compile oracle/hode_oracle.c with gcc (idempotent)."""
    src = os.path.join(_HERE, "hode_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "hode.h")
    coef = os.path.join(_HERE, "dop853_coef.h")
    stale = (not os.path.exists(_LIB_PATH)
             or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
             or os.path.getmtime(_LIB_PATH) < os.path.getmtime(coef)
             or (os.path.exists(hdr) and os.path.getmtime(_LIB_PATH) < os.path.getmtime(hdr)))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.hode_oracle_mlp_param_count.restype = ctypes.c_int64
        _lib.hode_oracle_mlp_param_count.argtypes = [ctypes.c_int32, ctypes.c_int32]
    return _lib


def mlp_param_count(hidden: int, layers: int) -> int:
    return int(lib().hode_oracle_mlp_param_count(hidden, layers))


def _f32(a) -> Optional[np.ndarray]:
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a: Optional[np.ndarray], ty=ctypes.c_float):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ty))


def make_cfg(B: int, T: int, t_obs: np.ndarray, inputs: Dict[str, Optional[np.ndarray]],
             hidden: int, layers: int, has_nn: bool, n_samples: int, solver: str, rtol: float,
             atol: float, n_substeps: int, max_steps: int, kinks: str = "scipy") -> HodeCfg:
    cfg = HodeCfg()
    cfg.struct_bytes = ctypes.sizeof(HodeCfg)
    cfg.n_traj, cfg.n_obs = B, T
    cfg.t_per_traj = 1 if t_obs.ndim == 2 else 0
    for ch, name in enumerate(("meal", "tVNS", "GD")):
        v = inputs.get(name)
        cfg.in_mode[ch] = IN_ABSENT if v is None else (IN_SERIES if v.ndim == 2 else IN_CONST)
    cfg.nn_hidden, cfg.nn_layers = hidden, layers
    cfg.mlp = MLP_FP32 if has_nn else MLP_NONE
    cfg.n_samples = n_samples
    cfg.solver = {"rk4": SOLVER_RK4, "dopri5": SOLVER_DOPRI5, "rk45": SOLVER_DOPRI5, "dop853": SOLVER_DOP853}[solver]
    cfg.n_substeps, cfg.max_steps = n_substeps, max_steps
    cfg.rtol, cfg.atol = rtol, atol
    cfg.kink_mode = {"scipy": 0, "clip": 1}[kinks]
    return cfg


def rollout(y0, t_obs, inputs=None, theta=None, W=None, hidden=64, layers=4, solver="dopri5",
            rtol=1e-6, atol=1e-8, n_substeps=1, max_steps=0, rhs="f32", n_threads=1,
            kinks="scipy") -> Tuple[np.ndarray, np.ndarray, np.ndarray, int]:
    """CPU oracle rollout.

    y0 [B,6]; t_obs [T] or [B,T]; inputs {'meal'|'tVNS'|'GD': [B] or [B,T]};
    theta [17] or [S,17]; W None (no NN) or [P] or [S,P] packed as include/hode.h says.
    Returns (traj [S,B,T,6] or [B,T,6] when theta is 1-D, status, counters [2,...], nfev).
    """
    y0 = _f32(np.atleast_2d(y0))
    t_obs = _f32(t_obs)
    B, T = y0.shape[0], t_obs.shape[-1]
    inputs = {k: _f32(v) for k, v in (inputs or {}).items() if v is not None}
    theta = THETA_DEFAULT if theta is None else _f32(theta)
    squeeze = theta.ndim == 1
    theta2 = np.ascontiguousarray(np.atleast_2d(theta))
    S = theta2.shape[0]
    W2 = None
    if W is not None:
        W2 = np.ascontiguousarray(np.atleast_2d(_f32(W)))
        assert W2.shape == (S, mlp_param_count(hidden, layers)), W2.shape
    cfg = make_cfg(B, T, t_obs, inputs, hidden, layers, W is not None, S, solver, rtol, atol,
                   n_substeps, max_steps, kinks)
    traj = np.zeros((S, B, T, 6), dtype=np.float32)
    status = np.zeros((S, B), dtype=np.int32)
    counters = np.zeros((2, S, B), dtype=np.int32)
    nfev = ctypes.c_int64(0)
    rc = lib().hode_oracle_rollout(
        ctypes.byref(cfg), ctypes.c_int(RHS_F64 if rhs == "f64" else RHS_F32),
        ctypes.c_int(n_threads), _ptr(y0), _ptr(t_obs), _ptr(inputs.get("meal")),
        _ptr(inputs.get("tVNS")), _ptr(inputs.get("GD")), _ptr(theta2), _ptr(W2), _ptr(traj),
        _ptr(status, ctypes.c_int32), _ptr(counters, ctypes.c_int32), ctypes.byref(nfev))
    if rc != 0:
        raise RuntimeError(f"hode_oracle_rollout returned {rc}")
    if squeeze:
        return traj[0], status[0], counters[:, 0], int(nfev.value)
    return traj, status, counters, int(nfev.value)


def rollout_with_steps(b, y0, t_obs, inputs=None, theta=None, W=None, hidden=64, layers=4,
                       rtol=1e-6, atol=1e-8, rhs="f32", cap=100000, kinks="scipy", solver="dopri5"):
    """Adaptive solve (DP5(4), or DOP853 with solver='dop853') of trajectory b that also returns the accepted-step
    log (t_n, h_n)."""
    y0 = _f32(np.atleast_2d(y0))
    t_obs = _f32(t_obs)
    B, T = y0.shape[0], t_obs.shape[-1]
    inputs = {k: _f32(v) for k, v in (inputs or {}).items() if v is not None}
    theta = THETA_DEFAULT if theta is None else _f32(theta)
    cfg = make_cfg(B, T, t_obs, inputs, hidden, layers, W is not None, 1, solver, rtol, atol,
                   1, 0, kinks)
    Wc = None if W is None else _f32(W)
    row = np.zeros((T, 6), dtype=np.float32)
    st_t = np.zeros(cap, dtype=np.float64)
    st_h = np.zeros(cap, dtype=np.float64)
    n = ctypes.c_int32(0)
    status = lib().hode_oracle_rollout_log(
        ctypes.byref(cfg), ctypes.c_int(RHS_F64 if rhs == "f64" else RHS_F32), ctypes.c_long(b),
        _ptr(y0), _ptr(t_obs), _ptr(inputs.get("meal")), _ptr(inputs.get("tVNS")),
        _ptr(inputs.get("GD")), _ptr(theta), _ptr(Wc), _ptr(row), _ptr(st_t, ctypes.c_double),
        _ptr(st_h, ctypes.c_double), ctypes.c_int32(cap), ctypes.byref(n))
    return row, status, st_t[: n.value].copy(), st_h[: n.value].copy()


def rhs_eval(t, state, inputs=None, theta=None, W=None, hidden=64, layers=4, rhs="f32"):
    """One batched evaluation of f_physio + g_NN; returns float64 [B,6]."""
    state = _f32(np.atleast_2d(state))
    B = state.shape[0]
    t = _f32(np.broadcast_to(np.asarray(t, dtype=np.float32), (B,)))
    inputs = {k: _f32(np.broadcast_to(np.asarray(v, dtype=np.float32), (B,)))
              for k, v in (inputs or {}).items() if v is not None}
    theta = THETA_DEFAULT if theta is None else _f32(theta)
    cfg = make_cfg(B, 1, t, inputs, hidden, layers, W is not None, 1, "rk4", 0, 0, 1, 0)
    Wc = None if W is None else _f32(W)
    out = np.zeros((B, 6), dtype=np.float64)
    rc = lib().hode_oracle_rhs(
        ctypes.byref(cfg), ctypes.c_int(RHS_F64 if rhs == "f64" else RHS_F32), _ptr(t),
        _ptr(state), _ptr(inputs.get("meal")), _ptr(inputs.get("tVNS")), _ptr(inputs.get("GD")),
        _ptr(theta), _ptr(Wc), _ptr(out, ctypes.c_double))
    if rc != 0:
        raise RuntimeError(f"hode_oracle_rhs returned {rc}")
    return out


def pack_mlp(weights_and_biases) -> np.ndarray:
    """[(W0,b0),(W1,b1),...] with W [out,in] -> flat float32 in include/hode.h order."""
    parts = []
    for w, b in weights_and_biases:
        parts.append(np.asarray(w, dtype=np.float32).reshape(-1))
        parts.append(np.asarray(b, dtype=np.float32).reshape(-1))
    return np.concatenate(parts)
