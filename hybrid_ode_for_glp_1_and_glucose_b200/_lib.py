"""ctypes binding of libhode.so (include/hode.h).

The library is the product: if it is missing or a symbol is absent this module raises —
there is no CPU or eager-PyTorch fallback for any compute entry point.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libhode.so")

ABI_VERSION = 4
N_STATE, N_THETA, NN_IN = 6, 17, 9
SOLVER_RK4, SOLVER_DOPRI5, SOLVER_DOP853 = 0, 1, 2
IN_ABSENT, IN_CONST, IN_SERIES = 0, 1, 2
MLP_NONE, MLP_FP32, MLP_TF32X3, MLP_TF32, MLP_TF32BF16, MLP_TF32X2BF16, MLP_F16BF16X2 = 0, 1, 2, 3, 4, 5, 6
KINK_SCIPY, KINK_CLIP = 0, 1
ST_OK, ST_STEP_TOO_SMALL, ST_MAX_STEPS, ST_NONFINITE, ST_REC_OVERFLOW = 0, 1, 2, 3, 4

STATUS_TEXT = {
    ST_OK: "ok",
    ST_STEP_TOO_SMALL: "Required step size is less than spacing between numbers.",
    ST_MAX_STEPS: "step budget exhausted",
    ST_NONFINITE: "state became non-finite",
    ST_REC_OVERFLOW: "accepted-step record capacity exhausted (max_saved_steps): zero-padded, no gradient",
}

EXPORTS = [
    "hode_version", "hode_last_error_string", "hode_mlp_param_count", "hode_workspace_bytes",
    "hode_rollout_fwd", "hode_rollout_bwd", "hode_vi_predictive", "hode_rhs", "hode_rhs_vjp",
    "hode_rollout_fwd_host", "hode_loss_fused_fwd_bwd", "hode_generate_4gi",
    "hode_step_record_floats", "hode_step_record_capacity", "hode_launch_count",
    "hode_rollout_fwd_ex", "hode_rollout_fwd_host_ex", "hode_train_step", "hode_train_step_workspace_bytes",
    "hode_window_count", "hode_window_dataset", "hode_eval_metrics",
]


class HodeCfg(ctypes.Structure):
    """Mirror of `struct hode_cfg` (include/hode.h)."""
    _fields_ = [
        ("struct_bytes", ctypes.c_int32), ("n_traj", ctypes.c_int32), ("n_obs", ctypes.c_int32),
        ("t_per_traj", ctypes.c_int32), ("in_mode", ctypes.c_int32 * 3),
        ("nn_hidden", ctypes.c_int32), ("nn_layers", ctypes.c_int32), ("mlp", ctypes.c_int32),
        ("n_samples", ctypes.c_int32), ("solver", ctypes.c_int32), ("n_substeps", ctypes.c_int32),
        ("max_steps", ctypes.c_int32), ("rtol", ctypes.c_double), ("atol", ctypes.c_double),
        ("save_steps", ctypes.c_int32), ("kink_mode", ctypes.c_int32),
        ("max_saved_steps", ctypes.c_int32), ("rhs_part", ctypes.c_int32),
    ]


class HodeFwdOpts(ctypes.Structure):
    """Mirror of `struct hode_fwd_opts` (include/hode.h)."""
    _fields_ = [("struct_bytes", ctypes.c_int32), ("theta_per_traj", ctypes.c_int32), ("order", ctypes.c_void_p),
                ("out_state_mask", ctypes.c_uint32), ("reserved", ctypes.c_uint32), ("prev_counters", ctypes.c_void_p)]


def new_fwd_opts(theta_per_traj: bool = False, order_ptr: Optional[int] = None, out_state_mask: int = 0,
                 prev_counters_ptr: Optional[int] = None) -> HodeFwdOpts:
    o = HodeFwdOpts()
    o.struct_bytes = ctypes.sizeof(HodeFwdOpts)
    o.theta_per_traj = 1 if theta_per_traj else 0
    o.order = order_ptr
    o.out_state_mask = int(out_state_mask)
    o.prev_counters = prev_counters_ptr   # host entries only: HOST pointer to a previous call's [2,B] counters
    return o


class HodeTrainCfg(ctypes.Structure):
    """Mirror of `struct hode_train_cfg` (include/hode.h)."""
    _fields_ = [("struct_bytes", ctypes.c_int32), ("n_physics", ctypes.c_int32), ("data_gradient", ctypes.c_int32),
                ("adam_step", ctypes.c_int32), ("lambda1", ctypes.c_float), ("lambda2", ctypes.c_float),
                ("grad_clip", ctypes.c_float), ("lr", ctypes.c_float), ("beta1", ctypes.c_float), ("beta2", ctypes.c_float),
                ("eps", ctypes.c_float), ("physics_dt", ctypes.c_float)]


class HodeError(RuntimeError):
    pass


_lib: Optional[ctypes.CDLL] = None

_P = ctypes.c_void_p


def lib() -> ctypes.CDLL:
    """Load libhode.so once; raise loudly when it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HodeError(
            f"{LIB_PATH} is missing: build it with "
            "`python -m hybrid_ode_for_glp_1_and_glucose_b200.build` (there is no fallback path)")
    L = ctypes.CDLL(LIB_PATH)
    for name in EXPORTS:
        if not hasattr(L, name):
            raise HodeError(f"libhode.so does not export {name}")
    L.hode_version.restype = ctypes.c_int
    L.hode_launch_count.restype = ctypes.c_int64
    L.hode_last_error_string.restype = ctypes.c_char_p
    L.hode_mlp_param_count.restype = ctypes.c_int64
    L.hode_mlp_param_count.argtypes = [ctypes.c_int32, ctypes.c_int32]
    L.hode_workspace_bytes.restype = ctypes.c_int
    L.hode_workspace_bytes.argtypes = [ctypes.POINTER(HodeCfg), ctypes.POINTER(ctypes.c_size_t),
                                       ctypes.POINTER(ctypes.c_size_t)]
    for fn in (L.hode_step_record_floats, L.hode_step_record_capacity):
        fn.restype = ctypes.c_int32
        fn.argtypes = [ctypes.POINTER(HodeCfg)]
    L.hode_rollout_fwd.restype = ctypes.c_int
    L.hode_rollout_fwd.argtypes = [ctypes.POINTER(HodeCfg)] + [_P] * 11 + [ctypes.c_size_t, _P]
    L.hode_rollout_fwd_ex.restype = ctypes.c_int
    L.hode_rollout_fwd_ex.argtypes = [ctypes.POINTER(HodeCfg), ctypes.POINTER(HodeFwdOpts)] + [_P] * 11 + [ctypes.c_size_t, _P]
    L.hode_rollout_fwd_host_ex.restype = ctypes.c_int
    L.hode_rollout_fwd_host_ex.argtypes = [ctypes.POINTER(HodeCfg), ctypes.POINTER(HodeFwdOpts)] + [_P] * 11
    L.hode_window_count.restype = ctypes.c_int
    L.hode_window_count.argtypes = [ctypes.c_int32] * 3
    L.hode_window_dataset.restype = ctypes.c_int
    L.hode_window_dataset.argtypes = [ctypes.c_int32] * 7 + [_P] * 9 + [ctypes.c_size_t, _P]
    L.hode_eval_metrics.restype = ctypes.c_int
    L.hode_eval_metrics.argtypes = [ctypes.c_int64, _P, _P, _P, ctypes.c_float, _P, ctypes.c_int32, _P, _P, ctypes.c_size_t, _P]
    L.hode_train_step_workspace_bytes.restype = ctypes.c_int
    L.hode_train_step_workspace_bytes.argtypes = [ctypes.POINTER(HodeCfg), ctypes.POINTER(HodeTrainCfg), ctypes.POINTER(ctypes.c_size_t)]
    L.hode_train_step.restype = ctypes.c_int
    L.hode_train_step.argtypes = [ctypes.POINTER(HodeCfg), ctypes.POINTER(HodeTrainCfg)] + [_P] * 16 + [ctypes.c_size_t, _P]
    L.hode_rollout_bwd.restype = ctypes.c_int
    L.hode_rollout_bwd.argtypes = ([ctypes.POINTER(HodeCfg)] + [_P] * 11
                                   + [_P, ctypes.c_size_t, _P, ctypes.c_size_t, _P])
    L.hode_loss_fused_fwd_bwd.restype = ctypes.c_int
    L.hode_loss_fused_fwd_bwd.argtypes = ([ctypes.POINTER(HodeCfg)] + [_P] * 16
                                          + [_P, ctypes.c_size_t, _P, ctypes.c_size_t, _P])
    L.hode_generate_4gi.restype = ctypes.c_int
    L.hode_generate_4gi.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_double, ctypes.c_int32, ctypes.c_double,
                                    ctypes.c_double, _P, _P, _P, _P, _P]
    L.hode_rhs_vjp.restype = ctypes.c_int
    L.hode_rhs_vjp.argtypes = [ctypes.POINTER(HodeCfg)] + [_P] * 12 + [ctypes.c_size_t, _P]
    L.hode_vi_predictive.restype = ctypes.c_int
    L.hode_vi_predictive.argtypes = [ctypes.POINTER(HodeCfg)] + [_P] * 12 + [ctypes.c_size_t, _P]
    L.hode_rhs.restype = ctypes.c_int
    L.hode_rhs.argtypes = [ctypes.POINTER(HodeCfg)] + [_P] * 9
    L.hode_rollout_fwd_host.restype = ctypes.c_int
    L.hode_rollout_fwd_host.argtypes = [ctypes.POINTER(HodeCfg)] + [_P] * 11
    if L.hode_version() != ABI_VERSION:
        raise HodeError(f"libhode.so ABI {L.hode_version()} != binding ABI {ABI_VERSION}")
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().hode_last_error_string().decode("utf-8", "replace")
        raise HodeError(f"{what} failed (rc={rc}): {msg}")


def mlp_param_count(hidden: int, layers: int) -> int:
    return int(lib().hode_mlp_param_count(hidden, layers))


def new_cfg() -> HodeCfg:
    cfg = HodeCfg()
    cfg.struct_bytes = ctypes.sizeof(HodeCfg)
    cfg.n_samples = 1
    cfg.n_substeps = 1
    cfg.rtol, cfg.atol = 1e-6, 1e-8
    return cfg
