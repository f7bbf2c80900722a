"""C-ABI checks that need no GPU: the library loads, exports every symbol include/hode.h
declares, agrees on the struct layout, and rejects bad arguments without launching."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "hode.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hode_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_documented_surface():
    names = declared_functions()
    for must in ("hode_rollout_fwd", "hode_rollout_bwd", "hode_vi_predictive", "hode_rhs",
                 "hode_rollout_fwd_host", "hode_workspace_bytes", "hode_version"):
        assert must in names


def test_library_exports_every_declared_symbol(built_lib):
    L = ctypes.CDLL(built_lib)
    for name in declared_functions():
        assert hasattr(L, name), f"libhode.so lacks {name}"


def test_binding_lists_every_declared_symbol(built_lib):
    from hybrid_ode_for_glp_1_and_glucose_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared_functions()
    assert _lib.lib().hode_version() == _lib.ABI_VERSION


def test_struct_layout_and_param_count(built_lib):
    from hybrid_ode_for_glp_1_and_glucose_b200 import _lib
    cfg = _lib.new_cfg()
    cfg.n_traj, cfg.n_obs = 4, 3
    assert ctypes.sizeof(_lib.HodeCfg) == 88
    assert _lib.HodeCfg.rtol.offset == 56 and _lib.HodeCfg.atol.offset == 64
    fwd, bwd = ctypes.c_size_t(1), ctypes.c_size_t(1)
    rc = _lib.lib().hode_workspace_bytes(ctypes.byref(cfg), ctypes.byref(fwd), ctypes.byref(bwd))
    assert rc == 0 and fwd.value == 0
    cfg.save_steps, cfg.solver, cfg.n_substeps = 1, _lib.SOLVER_RK4, 4
    rc = _lib.lib().hode_workspace_bytes(ctypes.byref(cfg), ctypes.byref(fwd), ctypes.byref(bwd))
    assert rc == 0 and fwd.value >= 4 * 8 * (4 + 8 + 4 + 24)
    assert _lib.mlp_param_count(64, 4) == 13510
    assert _lib.mlp_param_count(16, 2) == 9 * 16 + 16 + 16 * 16 + 16 + 16 * 6 + 6


def test_bad_arguments_are_rejected_before_any_launch(built_lib):
    from hybrid_ode_for_glp_1_and_glucose_b200 import _lib
    L = _lib.lib()
    cfg = _lib.new_cfg()
    cfg.struct_bytes = 12
    assert L.hode_workspace_bytes(ctypes.byref(cfg), None, None) == -2
    assert b"struct_bytes" in L.hode_last_error_string()
    cfg = _lib.new_cfg()
    cfg.n_traj, cfg.n_obs = 1, 0
    assert L.hode_workspace_bytes(ctypes.byref(cfg), None, None) == -2
    cfg.n_obs, cfg.solver = 2, 7
    assert L.hode_workspace_bytes(ctypes.byref(cfg), None, None) == -4
    cfg.solver = _lib.SOLVER_DOPRI5
    # NULL buffers: argument error, not a crash
    rc = L.hode_rollout_fwd(ctypes.byref(cfg), *([None] * 11), 0, None)
    assert rc == -1
    with pytest.raises(_lib.HodeError):
        _lib.check(rc, "hode_rollout_fwd")


def test_oracle_and_product_agree_on_the_struct(oracle):
    from hybrid_ode_for_glp_1_and_glucose_b200 import _lib
    assert ctypes.sizeof(oracle.HodeCfg) == ctypes.sizeof(_lib.HodeCfg)
    for (n1, _), (n2, _) in zip(oracle.HodeCfg._fields_, _lib.HodeCfg._fields_):
        assert n1 == n2


def test_network_shapes_that_cannot_fit_shared_memory_are_rejected(built_lib):
    """128 x 4 needs 239 KB for the FP32 kernels' weight image + activation columns (> 227 KB per CTA): the library says
    so (HODE_E_UNSUPPORTED) instead of failing at launch with 'invalid configuration argument'."""
    from hybrid_ode_for_glp_1_and_glucose_b200 import _lib
    L = _lib.lib()
    for H, layers, ok in ((64, 4, True), (128, 3, True), (128, 4, False), (128, 8, False), (96, 6, True)):
        cfg = _lib.new_cfg()
        cfg.n_traj, cfg.n_obs, cfg.mlp, cfg.nn_hidden, cfg.nn_layers = 4, 3, _lib.MLP_FP32, H, layers
        f, b = ctypes.c_size_t(), ctypes.c_size_t()
        rc = L.hode_workspace_bytes(ctypes.byref(cfg), ctypes.byref(f), ctypes.byref(b))
        assert (rc == 0) == ok, (H, layers, rc)
        if not ok:
            assert rc == -4 and b"shared memory" in L.hode_last_error_string()
    cfg = _lib.new_cfg()
    cfg.n_traj, cfg.n_obs, cfg.mlp, cfg.nn_hidden, cfg.nn_layers = 4, 3, _lib.MLP_TF32X3, 64, 7
    assert L.hode_workspace_bytes(ctypes.byref(cfg), None, None) == -4     # the tensor-core image of 7 layers does not fit
