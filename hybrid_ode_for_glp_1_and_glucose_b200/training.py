"""Training loop of the reference with the whole batch step as ONE library call.

Mirrors train/train_hybrid.py:225-302 of the reference: `train_epoch(model, train_loader, optimizer, scheduler, config,
writer, epoch, device)` and `validate(model, val_loader, config, device)` keep their signatures, config keys and return
values.  What changes is the body of the batch loop: `model.loss(batch)`, `loss.backward()`, `clip_grad_norm_` and
`optimizer.step()` (Adam) become `hode_train_step` — rollout, physics re-solves, RHS, RHS-VJP, (optionally) the
discrete adjoint of the data term, L2 term, gradient-norm clipping and the Adam update, stream-ordered in one call,
optionally replayed as a CUDA graph.  The optimizer object stays the owner of the hyper-parameters (lr, betas, eps:
schedulers keep working) and of the Adam state (`state_dict()` checkpoints keep working): its `exp_avg` / `exp_avg_sq`
tensors are views of the packed moment buffers the kernel updates.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import torch

from . import _lib, ops
from ._lib import HodeError


class FusedTrainer:
    """Packed network parameters (the module's parameters become views of one flat buffer), Adam moments, workspace."""

    def __init__(self, model, optimizer: Optional[torch.optim.Optimizer] = None, lr: float = 1e-3,
                 betas=(0.9, 0.999), eps: float = 1e-8, data_gradient: bool = False, use_cuda_graph: bool = False):
        self.model = model
        self.optimizer = optimizer
        if optimizer is not None and not isinstance(optimizer, torch.optim.Adam):
            raise HodeError("the fused step implements torch.optim.Adam (what the reference trains with, "
                            "train/train_hybrid.py:470); use the unfused loop for other optimizers")
        self.defaults = dict(lr=lr, betas=tuple(betas), eps=eps)
        self.data_gradient = data_gradient
        self.use_cuda_graph = use_cuda_graph
        self.dev = model._cuda_device()
        params = [p for _, p in model.nn_residual.named_parameters()]
        with torch.no_grad():
            self.flat = torch.cat([p.detach().reshape(-1).to(self.dev, torch.float32) for p in params]).contiguous()
        self.grad = torch.zeros_like(self.flat)
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        self.t = 0
        off = 0
        for p in params:     # parameters, their .grad and their Adam moments all alias the packed buffers
            n = p.numel()
            p.data = self.flat[off: off + n].view(p.shape)
            p.grad = self.grad[off: off + n].view(p.shape)
            if optimizer is not None:
                st = optimizer.state[p]
                st["step"] = torch.tensor(0.0)
                st["exp_avg"] = self.m[off: off + n].view(p.shape)
                st["exp_avg_sq"] = self.v[off: off + n].view(p.shape)
            off += n
        self.params = params
        self.scalars = torch.zeros(9, dtype=torch.float32, device=self.dev)
        self._ws = None
        self._graph = None
        self._graph_key = None
        self.last_traj = None

    def _hyper(self):
        if self.optimizer is not None:
            g = self.optimizer.param_groups[0]
            if g.get("weight_decay", 0) != 0 or g.get("amsgrad", False):
                raise HodeError("the fused Adam update implements weight_decay=0, amsgrad=False")
            return float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"])
        d = self.defaults
        return float(d["lr"]), float(d["betas"][0]), float(d["betas"][1]), float(d["eps"])

    def step(self, batch: Dict[str, torch.Tensor], lambda1: float = 1.0, lambda2: float = 1.0,
             use_physics_loss: bool = True, grad_clip: float = 0.0, update: bool = True) -> Dict[str, torch.Tensor]:
        """One batch: loss (reference HybridODENN.loss semantics), gradient, clip, Adam.  Returns DEVICE scalars
        {'loss','data','physics','reg','grad_norm','clip_coef'} (no host synchronisation)."""
        m, dev = self.model, self.dev
        if m.use_variational and lambda2 > 0:
            raise HodeError("use_variational models take the reference's bayes_loss branch, which cannot run "
                            "(SURVEY §0.6); train them with VariationalInference")
        y0, obs = batch["initial_state"], batch["observations"]
        tpts, ext = batch["time_points"], batch.get("external_inputs", None)
        theta = m.ode_core.theta().to(dev)
        cfg, bufs = ops.prepare(y0, tpts, ext, theta, self.flat, m.nn_residual.hidden_dim, m.nn_residual.n_layers, dev)
        cfg.solver = _lib.SOLVER_DOPRI5
        cfg.rtol, cfg.atol = 1e-6, 1e-8                      # forward()'s defaults, as loss() calls it
        cfg.kink_mode = ops.KINKS[m.kinks]
        cfg.mlp = ops._mlp_mode(m.precision, m.nn_residual.hidden_dim, m.nn_residual.n_layers)
        B, T = cfg.n_traj, cfg.n_obs
        n_pts = min(20, len(tpts)) if (use_physics_loss and lambda1 > 0) else 0
        idx = torch.randperm(len(tpts))[:n_pts].to(torch.int32) if n_pts else torch.zeros(0, dtype=torch.int32)
        lr, b1, b2, eps = self._hyper()
        tc = _lib.HodeTrainCfg()
        tc.struct_bytes = ctypes.sizeof(_lib.HodeTrainCfg)
        tc.n_physics, tc.data_gradient = int(n_pts), 1 if self.data_gradient else 0
        tc.adam_step = self.t + 1 if update else 0
        tc.lambda1, tc.lambda2 = float(lambda1 if n_pts else 0.0), float(lambda2)
        tc.grad_clip = float(grad_clip)
        tc.lr, tc.beta1, tc.beta2, tc.eps = lr, b1, b2, eps
        tc.physics_dt = 0.1
        L = _lib.lib()
        need = ctypes.c_size_t(0)
        _lib.check(L.hode_train_step_workspace_bytes(ctypes.byref(cfg), ctypes.byref(tc), ctypes.byref(need)),
                   "hode_train_step_workspace_bytes")
        with torch.cuda.device(dev):
            if self._ws is None or self._ws.numel() < need.value:
                self._ws = torch.empty(max(need.value, 16), dtype=torch.uint8, device=dev)
            o = ops._f32c(obs.reshape(B, T, 6), dev)
            d_idx = idx.to(dev) if n_pts else None
            traj = torch.empty((B, T, 6), dtype=torch.float32, device=dev)
            status = torch.empty(B, dtype=torch.int32, device=dev)
            rc = L.hode_train_step(
                ctypes.byref(cfg), ctypes.byref(tc), ops._ptr(bufs["y0"]), ops._ptr(bufs["t_obs"]), ops._ptr(bufs["meal"]),
                ops._ptr(bufs["tVNS"]), ops._ptr(bufs["GD"]), ops._ptr(bufs["theta"]), ops._ptr(self.flat), ops._ptr(o),
                ops._ptr(d_idx), ops._ptr(self.m), ops._ptr(self.v), ops._ptr(self.grad), ops._ptr(self.scalars),
                ops._ptr(traj), ops._ptr(status), ops._ptr(self._ws), self._ws.numel(), ops._stream(dev))
        _lib.check(rc, "hode_train_step")
        if update:
            self.t += 1
            if self.optimizer is not None:
                for p in self.params:
                    self.optimizer.state[p]["step"] = torch.tensor(float(self.t))
        self.last_traj, self.last_status = traj, status
        s = self.scalars.clone()
        return {"loss": s[0], "data": s[1], "physics": s[2], "reg": s[3], "grad_norm": s[4], "clip_coef": s[5]}

    # ------------------------------------------------------------------ CUDA graph
    def capture(self, batch: Dict[str, torch.Tensor], lambda1: float = 1.0, lambda2: float = 1.0,
                use_physics_loss: bool = True, grad_clip: float = 0.0):
        """Capture one update on batches shaped like `batch` as a CUDA graph; returns replay(batch) -> device scalars.
        The inputs live in static buffers the replay copies into; the update count lives on the device (hode_train_cfg
        adam_step < 0), the physics indices are re-drawn per replay on the host like the reference draws them and copied
        into a static device buffer.  Hyper-parameters (lr, betas, lambdas, clip) are baked in: re-capture after a
        scheduler step."""
        m, dev = self.model, self.dev
        y0, obs = batch["initial_state"], batch["observations"]
        tpts, ext = batch["time_points"], batch.get("external_inputs", None) or {}
        theta = m.ode_core.theta().to(dev)
        cfg, bufs = ops.prepare(y0, tpts, ext, theta, self.flat, m.nn_residual.hidden_dim, m.nn_residual.n_layers, dev)
        cfg.solver = _lib.SOLVER_DOPRI5
        cfg.rtol, cfg.atol = 1e-6, 1e-8
        cfg.kink_mode = ops.KINKS[m.kinks]
        cfg.mlp = ops._mlp_mode(m.precision, m.nn_residual.hidden_dim, m.nn_residual.n_layers)
        B, T = cfg.n_traj, cfg.n_obs
        n_pts = min(20, len(tpts)) if (use_physics_loss and lambda1 > 0) else 0
        lr, b1, b2, eps = self._hyper()
        tc = _lib.HodeTrainCfg()
        tc.struct_bytes = ctypes.sizeof(_lib.HodeTrainCfg)
        tc.n_physics, tc.data_gradient, tc.adam_step = int(n_pts), 1 if self.data_gradient else 0, -1
        tc.lambda1, tc.lambda2, tc.grad_clip = float(lambda1 if n_pts else 0.0), float(lambda2), float(grad_clip)
        tc.lr, tc.beta1, tc.beta2, tc.eps, tc.physics_dt = lr, b1, b2, eps, 0.1
        L = _lib.lib()
        need = ctypes.c_size_t(0)
        _lib.check(L.hode_train_step_workspace_bytes(ctypes.byref(cfg), ctypes.byref(tc), ctypes.byref(need)),
                   "hode_train_step_workspace_bytes")
        static = dict(bufs)
        static["obs"] = ops._f32c(obs.reshape(B, T, 6), dev).clone()
        static["idx"] = torch.zeros(max(n_pts, 1), dtype=torch.int32, device=dev)
        static["traj"] = torch.empty((B, T, 6), dtype=torch.float32, device=dev)
        static["status"] = torch.empty(B, dtype=torch.int32, device=dev)
        static["ws"] = torch.empty(max(need.value, 16), dtype=torch.uint8, device=dev)
        self.scalars[6] = float(self.t)

        def launch():
            rc = L.hode_train_step(
                ctypes.byref(cfg), ctypes.byref(tc), ops._ptr(static["y0"]), ops._ptr(static["t_obs"]), ops._ptr(static["meal"]),
                ops._ptr(static["tVNS"]), ops._ptr(static["GD"]), ops._ptr(static["theta"]), ops._ptr(self.flat),
                ops._ptr(static["obs"]), ops._ptr(static["idx"]), ops._ptr(self.m), ops._ptr(self.v), ops._ptr(self.grad),
                ops._ptr(self.scalars), ops._ptr(static["traj"]), ops._ptr(static["status"]), ops._ptr(static["ws"]),
                static["ws"].numel(), ops._stream(dev))
            _lib.check(rc, "hode_train_step")

        def load(b):
            static["y0"].copy_(b["initial_state"].reshape(B, 6), non_blocking=True)
            static["t_obs"].copy_(b["time_points"].reshape(static["t_obs"].shape), non_blocking=True)
            static["obs"].copy_(b["observations"].reshape(B, T, 6), non_blocking=True)
            e = b.get("external_inputs", None) or {}
            for name in ops.CHANNELS:
                if static[name] is not None:
                    static[name].copy_(torch.as_tensor(e[name]).reshape(static[name].shape), non_blocking=True)
            if n_pts:
                static["idx"][:n_pts].copy_(torch.randperm(len(b["time_points"]))[:n_pts].to(torch.int32), non_blocking=True)

        graph = torch.cuda.CUDAGraph()
        # warm-up on a side stream (no parameter update: adam_step = 0), then capture the real step
        tc.adam_step = 0
        s_ = torch.cuda.Stream(device=dev)
        s_.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s_):
            load(batch)
            launch()
        torch.cuda.current_stream(dev).wait_stream(s_)
        tc.adam_step = -1
        with torch.cuda.graph(graph):
            launch()

        def replay(b):
            load(b)
            graph.replay()
            self.t += 1
            self.last_traj, self.last_status = static["traj"], static["status"]
            s = self.scalars.clone()
            return {"loss": s[0], "data": s[1], "physics": s[2], "reg": s[3], "grad_norm": s[4], "clip_coef": s[5]}
        self._graph = graph
        return replay


def _to_device(batch, device):
    for key in batch:
        if isinstance(batch[key], torch.Tensor):
            batch[key] = batch[key].to(device)
        elif isinstance(batch[key], dict):
            for k, v in batch[key].items():
                batch[key][k] = v.to(device)
    return batch


def _trainer_of(model, optimizer) -> FusedTrainer:
    tr = getattr(model, "_fused_trainer", None)
    if tr is None or (optimizer is not None and tr.optimizer is not optimizer):
        tr = FusedTrainer(model, optimizer, data_gradient=getattr(model, "differentiable", False))
        model._fused_trainer = tr
    return tr


def train_epoch(model, train_loader, optimizer, scheduler, config, writer, epoch, device):
    """Train for one epoch (reference train/train_hybrid.py:225-275), one hode_train_step per batch.  The per-batch
    `loss.item()` of the reference (its only use: the progress bar and TensorBoard) is read only when a writer is
    given; the epoch mean is accumulated on the device."""
    model.train()
    lambda1 = config["training"].get("lambda1", 1.0)
    lambda2 = config["training"].get("lambda2", 1.0)
    use_physics = not config["ablation"].get("no_physics", False)
    clip = config["training"].get("gradient_clip", 0)
    trainer = _trainer_of(model, optimizer)
    total = None
    n = 0
    for batch_idx, batch in enumerate(train_loader):
        batch = _to_device(batch, device)
        out = trainer.step(batch, lambda1=lambda1, lambda2=lambda2, use_physics_loss=use_physics,
                           grad_clip=clip if clip and clip > 0 else 0.0, update=True)
        total = out["loss"] if total is None else total + out["loss"]
        n += 1
        if writer is not None:
            writer.add_scalar("train/loss", out["loss"].item(), epoch * len(train_loader) + batch_idx)
    if scheduler is not None:
        scheduler.step()
    return float(total.item()) / max(n, 1) if total is not None else 0.0


def validate(model, val_loader, config, device):
    """Validate the model (reference train/train_hybrid.py:278-302): the same loss, no update."""
    model.eval()
    lambda1 = config["training"].get("lambda1", 1.0)
    lambda2 = config["training"].get("lambda2", 1.0)
    use_physics = not config["ablation"].get("no_physics", False)
    trainer = _trainer_of(model, None)
    total, n = None, 0
    with torch.no_grad():
        for batch in val_loader:
            batch = _to_device(batch, device)
            out = trainer.step(batch, lambda1=lambda1, lambda2=lambda2, use_physics_loss=use_physics, update=False)
            total = out["loss"] if total is None else total + out["loss"]
            n += 1
    return float(total.item()) / max(n, 1) if total is not None else 0.0
