"""Variational-inference driver (host-side mirror of the reference's inference/vi.py).

The Python training loop stays on the host as in the reference; every rollout it triggers
is a libhode launch, and the S-sample sweeps are a single launch instead of a Python loop.
"""
from __future__ import annotations

import logging
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

logger = logging.getLogger(__name__)


def predictive_from_samples(model, samples: List[Dict[str, torch.Tensor]], initial_state,
                            time_points, external_inputs=None, **kernel_opts
                            ) -> Tuple[torch.Tensor, torch.Tensor]:
    """mean / unbiased std over parameter samples (reference inference/vi.py:306-310)."""
    with torch.no_grad():
        return model.predictive_with_param_samples(samples, initial_state, time_points,
                                                   external_inputs, **kernel_opts)


class VariationalInference:
    """Mean-field VI trainer (reference inference/vi.py:19)."""

    def __init__(self, model, prior_params: Optional[Dict[str, Dict[str, float]]] = None,
                 learning_rate: float = 1e-3, device: Optional[torch.device] = None):
        self.model = model
        self.device = torch.device(device) if device is not None else torch.device(
            "cuda" if torch.cuda.is_available() else "cpu")
        self.learning_rate = learning_rate
        if getattr(model, "variational_params", None) is None:
            raise ValueError("Model must be initialized with use_variational=True")
        self.variational_params = model.variational_params
        self.optimizer = torch.optim.Adam(self.variational_params.parameters(), lr=learning_rate)
        self.history = {"elbo": [], "kl": [], "log_likelihood": []}

    def elbo(self, batch: Dict[str, torch.Tensor], n_samples: int = 5, noise_sigma: float = 1.0,
             reparam_gradient: Optional[bool] = None) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
        """ELBO = E_q[log p(x|psi)] - KL (reference inference/vi.py:60-118).  The n_samples
        rollouts run as one [S,B] launch.

        Default (`reparam_gradient` False): as in the reference the likelihood term carries no
        gradient (forward returns a graph-free tensor); only the KL term does.
        `reparam_gradient=True` (or `self.reparam_gradient = True`): the sweep records its steps
        and the likelihood back-propagates through hode_rollout_bwd into every sampled parameter
        set, i.e. into the means and log-stds via psi = mu + eps * sigma — the estimator
        SURVEY §8f row 2 asks for; same value, same sampling order."""
        y0, obs = batch["initial_state"], batch["observations"]
        t, ext = batch["time_points"], batch.get("external_inputs", None)
        kl = self.variational_params.kl_divergence()
        samples = [self.variational_params.sample(1)[0] for _ in range(n_samples)]
        if reparam_gradient is None:
            reparam_gradient = getattr(self, "reparam_gradient", False)
        if reparam_gradient and torch.is_grad_enabled():
            preds = self.model.forward_with_param_samples(samples, y0, t, ext, differentiable=True,
                                                          **getattr(self, "kernel_opts", {}))
        else:
            with torch.no_grad():
                preds = self.model.forward_with_param_samples(samples, y0, t, ext,
                                                              **getattr(self, "kernel_opts", {}))
        obs = obs.to(preds.device)
        sq = ((obs.unsqueeze(0) - preds) / noise_sigma).pow(2).sum(dim=(1, 2, 3))
        ll = (-0.5 * sq).sum() / n_samples
        ll = ll - 0.5 * obs.numel() * np.log(2 * np.pi * noise_sigma ** 2)
        elbo = ll - kl
        return elbo, {"elbo": elbo, "kl": kl, "log_likelihood": ll}

    def train_step(self, batch: Dict[str, torch.Tensor], n_samples: int = 5,
                   reparam_gradient: Optional[bool] = None) -> Dict[str, float]:
        self.optimizer.zero_grad()
        elbo, comp = self.elbo(batch, n_samples=n_samples, reparam_gradient=reparam_gradient)
        loss = -elbo
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.variational_params.parameters(), max_norm=5.0)
        self.optimizer.step()
        return {"loss": loss.item(), "elbo": elbo.item(), "kl": comp["kl"].item(),
                "log_likelihood": float(comp["log_likelihood"])}

    def train(self, train_loader, val_loader=None, epochs: int = 100, n_samples: int = 5,
              early_stopping_patience: int = 10, verbose: bool = True):
        """Epoch loop of the reference (inference/vi.py:157-260), statement for statement where it matters: early
        stopping and the best-state bookkeeping run ONLY with a validation loader; the history is appended after
        the early-stop check (a run that stops early does not record its last epoch); `best_state` is the live
        state_dict, as in the reference (:237) — its tensors alias the parameters, so the final `load_state_dict`
        leaves the last epoch's values in place, exactly as the reference does."""
        best_val_elbo, patience_counter = -float("inf"), 0
        for epoch in range(epochs):
            sums = {"elbo": 0.0, "kl": 0.0, "log_likelihood": 0.0}
            n = 0
            for batch in train_loader:
                m = self.train_step(self._to_device(batch), n_samples)
                for k in sums:
                    sums[k] += m[k]
                n += 1
            for k in sums:
                sums[k] /= max(n, 1)
            val_elbo = None
            if val_loader is not None:
                with torch.no_grad():
                    vals = [self.elbo(self._to_device(b), n_samples)[0].item() for b in val_loader]
                val_elbo = float(np.sum(vals)) / max(len(vals), 1)
                if val_elbo > best_val_elbo:
                    best_val_elbo, patience_counter = val_elbo, 0
                    self.best_state = self.variational_params.state_dict()
                else:
                    patience_counter += 1
                if patience_counter >= early_stopping_patience:
                    logger.info(f"Early stopping at epoch {epoch + 1}")
                    break
            for k in sums:
                self.history[k].append(sums[k])
            if verbose and (epoch + 1) % 10 == 0:
                logger.info(f"Epoch {epoch + 1}: Train ELBO={sums['elbo']:.4f}, KL={sums['kl']:.4f}, "
                            f"LL={sums['log_likelihood']:.4f}")
                if val_elbo is not None:
                    logger.info(f"  Val ELBO={val_elbo:.4f}")
        if val_loader is not None and hasattr(self, "best_state"):
            self.variational_params.load_state_dict(self.best_state)

    def _to_device(self, batch):
        out = {}
        for k, v in batch.items():
            if torch.is_tensor(v):
                out[k] = v.to(self.device)
            elif isinstance(v, dict):
                out[k] = {kk: vv.to(self.device) for kk, vv in v.items()}
            else:
                out[k] = v
        return out

    def sample_posterior(self, n_samples: int = 100) -> List[Dict[str, torch.Tensor]]:
        return self.variational_params.sample(n_samples)

    def posterior_predictive(self, initial_state: torch.Tensor, time_points: torch.Tensor,
                             external_inputs: Optional[Dict[str, torch.Tensor]] = None,
                             n_samples: int = 100) -> Tuple[torch.Tensor, torch.Tensor]:
        """Predictive mean / std (reference inference/vi.py:274-312).  Samples are drawn one
        at a time, in the reference's order, then swept in a single launch."""
        samples = [self.variational_params.sample(1)[0] for _ in range(n_samples)]
        return predictive_from_samples(self.model, samples, initial_state, time_points,
                                       external_inputs)

    def save_checkpoint(self, path: str) -> None:
        torch.save({"variational_params": self.variational_params.state_dict(),
                    "optimizer": self.optimizer.state_dict(), "history": self.history}, path)
        logger.info(f"Checkpoint saved to {path}")

    def load_checkpoint(self, path: str) -> None:
        ck = torch.load(path, map_location=self.device)
        self.variational_params.load_state_dict(ck["variational_params"])
        self.optimizer.load_state_dict(ck["optimizer"])
        self.history = ck["history"]
        logger.info(f"Checkpoint loaded from {path}")
