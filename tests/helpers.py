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


def cohort(B, T=61, seed=0, horizon=5.0, meals=True, tvns=True):
    """4GI-shaped synthetic cohort: physiological-unit baselines,
    meal pulses at 0.5 h and 2.5 h, optional tVNS windows."""
    rng = np.random.default_rng(seed)
    y0 = np.stack([7.0 * rng.normal(1, 0.1, B), 50.0 * rng.normal(1, 0.15, B),
                   25.0 * rng.normal(1, 0.15, B), 10.0 * rng.normal(1, 0.15, B),
                   np.zeros(B), np.ones(B)], axis=1).astype(np.float32)
    t = np.linspace(0, horizon, T).astype(np.float32)
    ins = {}
    if meals:
        meal = np.zeros((B, T), dtype=np.float32)
        meal[:, T // 10] = rng.uniform(0.5, 1.5, B)
        meal[:, T // 2] = rng.uniform(0.3, 1.0, B)
        ins["meal"] = meal
    if tvns:
        tv = np.zeros((B, T), dtype=np.float32)
        on = rng.uniform(0, 1, B) > 0.5
        tv[on, T // 3: 2 * T // 3] = 1.0
        ins["tVNS"] = tv
    return y0, t, ins


def random_mlp(hidden=64, layers=4, seed=0, out_std=0.02, w_gain=1.0):
    """Packed MLP parameters (include/hode.h W layout) with a non-zero output layer."""
    rng = np.random.default_rng(seed)
    parts = []
    n_in = 9
    for l in range(layers + 1):
        n_out = 6 if l == layers else hidden
        if l == layers:
            w = rng.normal(0, out_std, (n_out, n_in))
            b = rng.normal(0, out_std, n_out)
        else:
            w = rng.normal(0, w_gain * np.sqrt(2.0 / (n_in + n_out)), (n_out, n_in))
            b = rng.normal(0, 0.05, n_out)
        parts += [w.reshape(-1), b]
        n_in = n_out
    return np.concatenate(parts).astype(np.float32)
