"""Device-resident mirror of the reference's GlucoseDataset (train/train_hybrid.py:43-155) for synthetic cohorts.

The reference reads a CSV, groups by subject, cuts sliding windows and z-scores them with pandas / numpy on the host.
For the cohorts this framework generates on the GPU (hode_generate_4gi: 262 144 - 1 048 576 subjects) that host pass
would dominate; `DeviceGlucoseDataset` does the windowing and normalisation in one library call (hode_window_dataset) and
serves batches as device tensors in the reference's item layout:
    {'initial_state' [6], 'observations' [L,6], 'time_points' [L], 'external_inputs': {'meal' [L], 'tVNS' [L]}}
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _lib, ops


class DeviceGlucoseDataset(torch.utils.data.Dataset):
    """states [N, n_t, 6] (GlucoseDataset's column order: glucose, insulin, glucagon, GLP-1, ge, ffa), inputs
    [N, n_t, n_in] with the meal indicator first (when present) and tVNS last, time [n_t] or [N, n_t] in hours."""

    def __init__(self, states: torch.Tensor, inputs: Optional[torch.Tensor], time: torch.Tensor,
                 sequence_length: int = 61, stride: int = 30, normalize: bool = True, has_meal: bool = True,
                 device: Optional[torch.device] = None):
        device = torch.device(device) if device is not None else states.device
        ops._require_cuda(device)
        self.sequence_length, self.stride, self.normalize, self.has_meal = sequence_length, stride, normalize, has_meal
        st = ops._f32c(states, device)
        N, n_t = st.shape[0], st.shape[1]
        if inputs is None:
            inputs = torch.zeros((N, n_t, 1))     # the reference adds a zero 'tvns' column when the file has none (:86-88)
            has_meal = self.has_meal = False
        inp = ops._f32c(inputs, device)
        tm = ops._f32c(time, device)
        n_in = inp.shape[2]
        L = _lib.lib()
        n_win = int(L.hode_window_count(n_t, sequence_length, stride))
        W = N * n_win
        with torch.cuda.device(device):
            self.observations = torch.empty((W, sequence_length, 6), dtype=torch.float32, device=device)
            self.initial_state = torch.empty((W, 6), dtype=torch.float32, device=device)
            self.inputs = torch.empty((W, sequence_length, n_in), dtype=torch.float32, device=device)
            self.time_points = torch.empty((W, sequence_length), dtype=torch.float32, device=device)
            ms = torch.empty(12, dtype=torch.float64, device=device)
            ws = torch.empty(64 * 1024, dtype=torch.uint8, device=device)
            rc = L.hode_window_dataset(N, n_t, n_in, sequence_length, stride, 1 if normalize else 0, 1 if tm.dim() == 2 else 0,
                                       ops._ptr(st), ops._ptr(inp), ops._ptr(tm), ops._ptr(self.observations),
                                       ops._ptr(self.initial_state), ops._ptr(self.inputs), ops._ptr(self.time_points), ops._ptr(ms),
                                       ops._ptr(ws), ws.numel(), ops._stream(device))
        _lib.check(rc, "hode_window_dataset")
        ms = ms.cpu().numpy()
        self.state_mean, self.state_std = ms[:6].copy(), ms[6:].copy()
        self.n_subjects, self.windows_per_subject = N, n_win

    def __len__(self) -> int:
        return self.observations.shape[0]

    def _ext(self, sel) -> Dict[str, torch.Tensor]:
        x = self.inputs[sel]
        meal = x[..., 0] if self.has_meal else torch.zeros_like(x[..., 0])
        return {"meal": meal, "tVNS": x[..., -1]}

    def __getitem__(self, idx: int) -> Dict[str, torch.Tensor]:
        return {"initial_state": self.initial_state[idx], "observations": self.observations[idx],
                "time_points": self.time_points[idx], "external_inputs": self._ext(idx)}

    def batch(self, indices: Sequence[int]) -> Dict[str, torch.Tensor]:
        """A collated batch (what the reference's DataLoader yields) gathered on the device."""
        sel = torch.as_tensor(indices, dtype=torch.long, device=self.observations.device)
        return {"initial_state": self.initial_state[sel], "observations": self.observations[sel],
                "time_points": self.time_points[sel], "external_inputs": self._ext(sel)}
