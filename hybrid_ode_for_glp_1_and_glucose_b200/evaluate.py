"""Evaluation metrics of the reference (eval/evaluate.py:26-288) with the reductions on the device.

`compute_rmse`, `compute_mae`, `compute_calibration_error` and `evaluate_model` keep the reference's signatures and
return values.  The reference moves every tensor to the host and reduces with numpy / scikit-learn; here one
streaming kernel (hode_eval_metrics) produces all the sums in a single pass over the device-resident predictions, and
the posterior-predictive loop of evaluate_model (:219-236) is the fused sweep (hode_vi_predictive).
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Union

import numpy as np
import torch

from . import _lib, ops

STATE_NAMES = ["Glucose", "Insulin", "Glucagon", "GLP1", "GE", "FFA"]   # reference eval/evaluate.py:204


def calibration_thresholds(n_bins: int = 10) -> np.ndarray:
    """The normalised-error threshold of every confidence level, computed the reference's way (:133-143): the conf-th
    percentile of |z| over 10 000 fresh standard-normal draws PER BIN from numpy's global generator (seed it to
    reproduce the reference's numbers)."""
    levels = np.linspace(0, 1, n_bins + 1)
    return np.array([np.percentile(np.abs(np.random.randn(10000)), levels[i] * 100) for i in range(n_bins)], dtype=np.float32)


def _sums(predictions: torch.Tensor, targets: torch.Tensor, uncertainties: Optional[torch.Tensor] = None,
          unc_const: float = 0.0, thresholds: Optional[np.ndarray] = None) -> np.ndarray:
    device = predictions.device
    ops._require_cuda(device)
    p = ops._f32c(predictions.reshape(-1, 6), device)
    t = ops._f32c(targets.reshape(-1, 6), device)
    u = None if uncertainties is None else ops._f32c(uncertainties.reshape(-1, 6), device)
    thr = None if thresholds is None else torch.from_numpy(np.asarray(thresholds, np.float32)).to(device)
    with torch.cuda.device(device):
        out = torch.empty(60, dtype=torch.float64, device=device)
        ws = torch.empty(192 * 1024, dtype=torch.uint8, device=device)
        rc = _lib.lib().hode_eval_metrics(p.shape[0], ops._ptr(p), ops._ptr(t), ops._ptr(u), float(unc_const), ops._ptr(thr),
                                          0 if thr is None else int(thr.numel()), ops._ptr(out), ops._ptr(ws), ws.numel(),
                                          ops._stream(device))
    _lib.check(rc, "hode_eval_metrics")
    return out.cpu().numpy(), p.shape[0]


def compute_rmse(predictions: torch.Tensor, targets: torch.Tensor, per_state: bool = False) -> Union[float, np.ndarray]:
    s, n = _sums(predictions, targets)
    return np.sqrt(s[0:6] / n) if per_state else float(np.sqrt(s[0:6].sum() / (6 * n)))


def compute_mae(predictions: torch.Tensor, targets: torch.Tensor, per_state: bool = False) -> Union[float, np.ndarray]:
    s, n = _sums(predictions, targets)
    return s[6:12] / n if per_state else float(s[6:12].sum() / (6 * n))


def compute_calibration_error(predictions: torch.Tensor, uncertainties: torch.Tensor, targets: torch.Tensor,
                              n_bins: int = 10) -> Dict[str, float]:
    thr = calibration_thresholds(n_bins)
    s, n = _sums(predictions, targets, uncertainties, thresholds=thr)
    m = 6 * n
    expected = np.linspace(0, 1, n_bins + 1)[:n_bins]
    observed = s[28: 28 + n_bins] / m
    return {"ece": float(np.mean(np.abs(expected - observed))), "msis": float(s[24] / m), "sharpness": float(s[25] / m),
            "coverage_95": float(s[26] / m), "mean_normalized_error": float(s[27] / m)}


def evaluate_model(model, test_loader, device: torch.device, use_variational: bool = False,
                   n_posterior_samples: int = 100) -> Dict[str, float]:
    """Comprehensive model evaluation (reference eval/evaluate.py:184-288)."""
    model.eval()
    preds, targs, uncs = [], [], []
    with torch.no_grad():
        for batch in test_loader:
            for key in batch:
                if isinstance(batch[key], torch.Tensor):
                    batch[key] = batch[key].to(device)
                elif isinstance(batch[key], dict):
                    for k, v in batch[key].items():
                        batch[key][k] = v.to(device)
            y0, targets, tpts = batch["initial_state"], batch["observations"], batch["time_points"]
            ext = batch.get("external_inputs", None)
            if use_variational and getattr(model, "variational_params", None) is not None:
                samples = [model.variational_params.sample(1)[0] for _ in range(n_posterior_samples)]
                p, u = model.predictive_with_param_samples(samples, y0, tpts, ext)
            else:
                p = model.forward(y0, tpts, ext)
                u = torch.ones_like(p) * 0.1
            preds.append(p); targs.append(targets); uncs.append(u)
    P, T_, U = torch.cat(preds, 0), torch.cat(targs, 0), torch.cat(uncs, 0)
    thr = calibration_thresholds(10) if use_variational else None
    s, n = _sums(P, T_, U if use_variational else None, thresholds=thr)
    metrics = {"rmse": float(np.sqrt(s[0:6].sum() / (6 * n))), "mae": float(s[6:12].sum() / (6 * n))}
    rmse_s, mae_s = np.sqrt(s[0:6] / n), s[6:12] / n
    for i, name in enumerate(STATE_NAMES):
        metrics[f"rmse_{name.lower()}"] = rmse_s[i]
        metrics[f"mae_{name.lower()}"] = mae_s[i]
    if use_variational:
        m = 6 * n
        expected = np.linspace(0, 1, 11)[:10]
        metrics.update({"ece": float(np.mean(np.abs(expected - s[28:38] / m))), "msis": float(s[24] / m),
                        "sharpness": float(s[25] / m), "coverage_95": float(s[26] / m),
                        "mean_normalized_error": float(s[27] / m)})
    # torch.std over (batch, time): unbiased
    target_std = np.sqrt(np.maximum(s[18:24] - s[12:18] ** 2 / n, 0.0) / max(n - 1, 1))
    metrics["nrmse"] = metrics["rmse"] / np.mean(target_std)
    for i, name in enumerate(STATE_NAMES):
        metrics[f"nrmse_{name.lower()}"] = rmse_s[i] / target_std[i]
    return metrics
