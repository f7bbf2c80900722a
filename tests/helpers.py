"""Shared test helpers: golden loading, synthetic cohorts, tolerance metrics."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_inputs(d):
    ins = {}
    if "meal" in d.files:
        ins["meal"] = d["meal"]
    if "tvns" in d.files:
        ins["tVNS"] = d["tvns"]
    if "gd" in d.files:
        ins["GD"] = d["gd"]
    return ins


def scaled_err(a, ref, rtol=1e-6, atol=1e-8):
    """max |a-ref| / (atol + rtol |ref|): error in units of the solver's local tolerance."""
    return float((np.abs(np.asarray(a, np.float64) - ref) / (atol + rtol * np.abs(ref))).max())


def rel_err(a, ref, floor=1.0):
    """Relative error of trajectories [..., T, 6] in the max norm over time, per trajectory and
    state component: max_t |a-ref| / max_t |ref|  (floor=1).  A pure element-wise ratio is
    meaningless where a component crosses zero (e.g. the network-driven GE state)."""
    a = np.asarray(a, np.float64)
    ref = np.asarray(ref, np.float64)
    scale = np.abs(ref).max(axis=-2, keepdims=True)
    denom = np.maximum(np.abs(ref), floor * scale + 1e-30)
    return float((np.abs(a - ref) / denom).max())


def rel_err_report(a, ref):
    """Where the worst element is (for assertion messages)."""
    a = np.asarray(a, np.float64)
    ref = np.asarray(ref, np.float64)
    scale = np.abs(ref).max(axis=-2, keepdims=True)
    r = np.abs(a - ref) / np.maximum(np.abs(ref), 1.0 * scale + 1e-30)
    idx = np.unravel_index(np.argmax(r), r.shape)
    return f"worst {r[idx]:.3e} at {idx}: got {a[idx]!r} ref {ref[idx]!r} scale {scale[idx[:-2] + (0, idx[-1])]!r}"


from hybrid_ode_for_glp_1_and_glucose_b200.synthetic import cohort, random_mlp  # noqa: E402,F401
