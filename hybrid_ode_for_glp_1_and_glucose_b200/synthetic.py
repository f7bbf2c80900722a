"""Synthetic 4GI-shaped cohorts and random networks for tests and benchmarks.

Statistics follow the reference's simulator: baselines G 7.0, I 50, Glu 25, GLP1 10 with
10-15 % between-subject spread (data/generate4GI.py:66-70,231-235), GE = 0 and FFA = 1
placeholders (train/train_hybrid.py:76-79), 61 samples over 5 h, meal pulses around 0.5 h and
2.5 h (data/generate4GI.py:279-280).  Nothing here is read from the reference at run time.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np


def cohort(B: int, T: int = 61, seed: int = 0, horizon: float = 5.0, meals: bool = True,
           tvns: bool = True) -> Tuple[np.ndarray, np.ndarray, Dict[str, np.ndarray]]:
    """(y0 [B,6], t [T], inputs {'meal','tVNS': [B,T]}) as float32 numpy arrays."""
    rng = np.random.default_rng(seed)
    y0 = np.empty((B, 6), dtype=np.float32)
    y0[:, 0] = 7.0 * rng.normal(1, 0.1, B)
    y0[:, 1] = 50.0 * rng.normal(1, 0.15, B)
    y0[:, 2] = 25.0 * rng.normal(1, 0.15, B)
    y0[:, 3] = 10.0 * rng.normal(1, 0.15, B)
    y0[:, 4] = 0.0
    y0[:, 5] = 1.0
    t = np.linspace(0, horizon, T).astype(np.float32)
    ins: Dict[str, np.ndarray] = {}
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


def clinical_cohort(B: int, T: int = 577, seed: int = 0) -> Tuple[np.ndarray, np.ndarray,
                                                                   Dict[str, np.ndarray]]:
    """mimic_clinical-shaped cohort: 48 h at 5 min with +-1 min jitter on
    a per-row grid, Poisson(4/day) meals of log-normal size, hourly Bernoulli tVNS windows."""
    rng = np.random.default_rng(seed)
    y0, _, _ = cohort(B, 2, seed=seed + 1, meals=False, tvns=False)
    base = np.arange(T, dtype=np.float64) * (5.0 / 60.0)
    t = np.sort(base[None, :] + rng.uniform(-1 / 60.0, 1 / 60.0, (B, T)), axis=1)
    t[:, 0] = 0.0
    meal = np.zeros((B, T), dtype=np.float32)
    n_meals = rng.poisson(4.0 * T * 5.0 / 60.0 / 24.0, B)
    for b in range(B):
        idx = rng.integers(1, T - 1, n_meals[b])
        meal[b, idx] = rng.lognormal(0.0, 0.5, n_meals[b])
    tv = np.repeat((rng.uniform(0, 1, (B, (T + 11) // 12)) > 0.8).astype(np.float32), 12, axis=1)
    return y0, t.astype(np.float32), {"meal": meal, "tVNS": tv[:, :T].copy()}


def random_mlp(hidden: int = 64, layers: int = 4, seed: int = 0, out_std: float = 0.02,
               w_gain: float = 1.0) -> np.ndarray:
    """Packed MLP parameters (include/hode.h W layout) with a NON-zero output layer, so the
    network actually contributes (a fresh reference model has a zero head)."""
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


THETA_DEFAULT = np.array([0.0104, 0.025, 0.003, 5.0, 60.0, 0.1, 50.0, 80.0, 9.0, 7.0, 0.02,
                          0.01, 1000.0, 2.0, 0.05, 0.001, 0.01], dtype=np.float32)


def meal_rate_from_events(meal_times, meal_sizes, n_obs: int = 61, interval_hours: float = 5.0 / 60.0,
                          n_subjects: int = 1) -> np.ndarray:
    """[n_subjects, n_obs-1] glucose input (mmol/h) per sampling interval, the way the reference spreads a
    meal over the interval that contains it (data/generate4GI.py:193-197: the last matching meal wins)."""
    t = np.arange(n_obs) * interval_hours
    rate = np.zeros(n_obs - 1, dtype=np.float32)
    for i in range(n_obs - 1):
        for mt, ms in zip(meal_times, meal_sizes):
            if t[i] <= mt < t[i + 1]:
                rate[i] = ms / (t[i + 1] - t[i])
    return np.tile(rate[None, :], (n_subjects, 1))


def fourgi_baselines(n_subjects: int, seed: int = 0) -> np.ndarray:
    """[n,5] per-subject baselines with generate_dataset's variability (data/generate4GI.py:64-70, :231-235):
    glucose 7.0 (cv 0.10), insulin 50, GLP-1 10, glucagon 25, GIP 20 (cv 0.15)."""
    rng = np.random.default_rng(seed)
    base = np.array([7.0, 50.0, 10.0, 25.0, 20.0])
    cv = np.array([0.1, 0.15, 0.15, 0.15, 0.15])
    return (base * (1.0 + cv * rng.normal(0, 1, (n_subjects, 5)))).astype(np.float32)


def fourgi_states(conc):
    """[N,T,5] generator output (glucose, insulin, GLP-1, glucagon, GIP) -> [N,T,6] model states in
    GlucoseDataset's column order [glucose, insulin, glucagon, GLP-1, ge = 0, ffa = 1]
    (train/train_hybrid.py:70-83).  Works on numpy arrays and torch tensors."""
    if isinstance(conc, np.ndarray):
        out = np.zeros(conc.shape[:-1] + (6,), dtype=conc.dtype)
    else:
        out = conc.new_zeros(conc.shape[:-1] + (6,))
    out[..., 0] = conc[..., 0]
    out[..., 1] = conc[..., 1]
    out[..., 2] = conc[..., 3]
    out[..., 3] = conc[..., 2]
    out[..., 5] = 1.0
    return out
