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


def rel_err(a, ref, floor=1e-3):
    """max |a-ref| / max(|ref|, floor*scale_of_column)."""
    a = np.asarray(a, np.float64)
    ref = np.asarray(ref, np.float64)
    denom = np.maximum(np.abs(ref), floor * np.abs(ref).max(axis=tuple(range(ref.ndim - 1)),
                                                           keepdims=True) + 1e-30)
    return float((np.abs(a - ref) / denom).max())


from hybrid_ode_for_glp_1_and_glucose_b200.synthetic import cohort, random_mlp  # noqa: E402,F401
