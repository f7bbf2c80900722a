"""Mean-field Gaussian variational family and posterior-predictive sweeps.

Host-side mirror of the reference's models/bayes.py.  Sampling and the KL term stay in
PyTorch on purpose: the reference draws `eps = randn_like(mean)` per tensor in
`param_shapes` insertion order (models/bayes.py:117-123), so keeping that exact call
sequence keeps seeded runs reproducible against the reference.  The expensive part — the
S x B rollouts — is one libhode launch (HybridODENN.forward_with_param_samples) or the
fused mean/std kernel (hode_vi_predictive).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn


class VariationalParameters(nn.Module):
    """q(psi) = prod N(mean, exp(log_std)^2) over a dict of named tensors."""

    def __init__(self, param_shapes: Dict[str, torch.Size],
                 prior_means: Optional[Dict[str, float]] = None,
                 prior_stds: Optional[Dict[str, float]] = None):
        super().__init__()
        self.param_shapes = param_shapes
        self.prior_means = prior_means or {}
        self.prior_stds = prior_stds or {}
        self.means = nn.ParameterDict()
        self.log_stds = nn.ParameterDict()
        for name, shape in param_shapes.items():
            mu0 = float(self.prior_means.get(name, 0.0))
            sd0 = float(self.prior_stds.get(name, 1.0))
            self.means[name] = nn.Parameter(torch.full(tuple(shape), mu0))
            # start at 10 % of the prior std (reference models/bayes.py:99-101)
            self.log_stds[name] = nn.Parameter(torch.full(tuple(shape), math.log(0.1 * sd0)))

    def sample(self, n_samples: int = 1) -> List[Dict[str, torch.Tensor]]:
        """Reparameterised draws psi = mu + eps * sigma, one dict per sample."""
        draws = []
        for _ in range(n_samples):
            draws.append({name: self.means[name] + torch.randn_like(self.means[name])
                          * self.log_stds[name].exp() for name in self.param_shapes})
        return draws

    def kl_divergence(self) -> torch.Tensor:
        """KL[q || prior], summed over every scalar (reference models/bayes.py:129-155)."""
        total = 0.0
        for name in self.param_shapes:
            mu, ls = self.means[name], self.log_stds[name]
            mu_p = self.prior_means.get(name, 0.0)
            sd_p = self.prior_stds.get(name, 1.0)
            term = (math.log(sd_p) - ls + (ls.exp().pow(2) + (mu - mu_p).pow(2))
                    / (2.0 * sd_p ** 2) - 0.5)
            total = total + term.sum()
        return total

    def get_flattened_params(self) -> Tuple[torch.Tensor, torch.Tensor]:
        names = sorted(self.param_shapes.keys())
        mu = torch.cat([self.means[n].flatten() for n in names])
        ls = torch.cat([self.log_stds[n].flatten() for n in names])
        return mu, ls


def bayes_loss(model, x_obs: torch.Tensor, noise_sigma: float = 1.0,
               n_samples: int = 5) -> torch.Tensor:
    """Negative ELBO with the reference's calling convention (models/bayes.py:15-62).

    In the reference this function cannot run: it calls
    `model.forward_with_params(psi_flat, x_obs)`, which lacks `t_span` and raises TypeError
    (SURVEY §0.6), and train/train_hybrid.py:452-461 falls back to point-estimate training.
    There is therefore no reference behaviour to reproduce ("parity unpinned").  The same
    TypeError is raised here so callers written against the reference (which catch it and fall
    back) behave identically; use inference.vi.VariationalInference.elbo for a working ELBO."""
    raise TypeError("forward() missing 1 required positional argument: 't_span' "
                    "(reference models/bayes.py:45 calls forward_with_params(psi, x_obs); use "
                    "VariationalInference.elbo)")


def compute_posterior_predictive(model, x_initial: torch.Tensor, t_span: torch.Tensor,
                                 external_inputs: Optional[Dict[str, torch.Tensor]] = None,
                                 n_samples: int = 100) -> Tuple[torch.Tensor, torch.Tensor]:
    """Posterior-predictive mean and (unbiased) std over `n_samples` draws
    (reference models/bayes.py:178-214), computed by one fused GPU sweep."""
    from .vi import predictive_from_samples
    samples = [model.sample_posterior(1)[0] for _ in range(n_samples)]
    return predictive_from_samples(model, samples, x_initial, t_span, external_inputs)
