"""Data-parallel plumbing for the rollout / gradient path: one process per GPU.

Every (trajectory, parameter-sample) unit is an independent initial-value problem (the
reference iterates them serially, models/hybrid_ode_nn.py:184), so the batch dimension shards
across ranks with NO collective on the data path: rollouts and posterior-predictive sweeps run
independently on each GPU (all S samples of a trajectory stay on one rank, so mean/std need no
exchange).  The only exchange step is the optimiser step: the gradients of the shared network
weights / ODE parameters and the loss or ELBO scalars are summed with ONE all-reduce over a
single packed float32 buffer (54-110 KB: latency-bound, NCCL over NVLink on GPUs, gloo in the
CPU tests).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of n trajectories owned by `rank`; sizes differ by at most 1."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(batch: Dict, rank: Optional[int] = None, world_size: Optional[int] = None) -> Dict:
    """This rank's slice of a collated batch {'initial_state' [B,6], 'observations' [B,T,6],
    'time_points' [T]|[B,T], 'external_inputs' {name: [B]|[B,T]}} (the dict the reference's
    loaders produce, train/train_hybrid.py:131-155).  A shared [T] time grid is replicated."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    B = batch["initial_state"].shape[0]
    lo, hi = shard_bounds(B, rank, world_size)

    def cut(v):
        if torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == B:
            return v[lo:hi]
        return v
    out = {}
    for k, v in batch.items():
        if isinstance(v, dict):
            out[k] = {kk: cut(vv) for kk, vv in v.items()}
        elif k == "time_points" and torch.is_tensor(v) and v.dim() == 1:
            out[k] = v
        else:
            out[k] = cut(v)
    return out


class PackedGradients:
    """One flat float32 buffer holding the gradients of `params` plus `n_scalars` extra slots
    (loss / ELBO / counts), reduced with a single all-reduce."""

    def __init__(self, params: Iterable[torch.Tensor], n_scalars: int = 0,
                 device: Optional[torch.device] = None):
        self.params: List[torch.Tensor] = [p for p in params]
        self.sizes = [p.numel() for p in self.params]
        self.n_scalars = n_scalars
        dev = device if device is not None else (self.params[0].device if self.params
                                                 else torch.device("cpu"))
        self.buffer = torch.zeros(sum(self.sizes) + n_scalars, dtype=torch.float32, device=dev)

    def pack(self, scalars: Sequence[float | torch.Tensor] = (), weight: float = 1.0) -> torch.Tensor:
        """Gradients (times `weight`) and scalars into the persistent buffer: one torch.cat, one copy."""
        if len(scalars) != self.n_scalars:
            raise ValueError(f"expected {self.n_scalars} scalars, got {len(scalars)}")
        dev = self.buffer.device
        parts = [(torch.zeros(n, dtype=torch.float32, device=dev) if p.grad is None
                  else p.grad.detach().reshape(-1).to(device=dev, dtype=torch.float32))
                 for p, n in zip(self.params, self.sizes)]
        if self.n_scalars:
            parts.append(torch.stack([(v.detach().to(device=dev, dtype=torch.float32).reshape(()) if torch.is_tensor(v)
                                       else torch.tensor(float(v), dtype=torch.float32, device=dev)) for v in scalars]))
        if parts:
            torch.cat(parts, out=self.buffer)
            if weight != 1.0:
                self.buffer[: sum(self.sizes)].mul_(weight)
        return self.buffer

    def unpack(self, scale: float = 1.0) -> torch.Tensor:
        """Write the (scaled) reduced gradients back into p.grad; returns the scalar slots."""
        off = 0
        for p, n in zip(self.params, self.sizes):
            g = self.buffer[off: off + n].reshape(p.shape) * scale
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += n
        return self.buffer[off:] * scale


def allreduce_gradients(params: Iterable[torch.Tensor], scalars: Sequence = (), average: bool = True,
                        packed: Optional[PackedGradients] = None) -> torch.Tensor:
    """Sum (or average) the gradients of `params` and the extra `scalars` over all ranks with a
    single all-reduce of one packed buffer; gradients are written back in place.  Returns the
    reduced scalars.  With one rank this is a no-op apart from returning the scalars."""
    rank, w = world()
    packed = packed or PackedGradients(params, len(scalars))
    buf = packed.pack(scalars)
    if w > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    return packed.unpack(1.0 / w if average else 1.0)


def sharded_loss_step(model, batch: Dict, optimizer, lambda1: float = 1.0, lambda2: float = 1.0,
                      gradient_clip: Optional[float] = None, use_physics_loss: bool = True) -> float:
    """One data-parallel optimiser step of train/train_hybrid.py:244-261: every rank evaluates model.loss on its shard
    of the batch; loss and gradients are combined with ONE packed all-reduce, each rank weighted by its shard size —
    the result is the loss / gradient of the global-batch means the reference computes (shards differ in size when B
    is not a multiple of the world size, and a rank whose shard is empty contributes nothing).  The physics indices
    are drawn from ONE generator state on every rank (rank 0's seed is broadcast), as a single process would draw
    them.  Then every rank applies the identical update."""
    rank, w = world()
    local = shard_batch(batch)
    n_local = int(local["initial_state"].shape[0])
    n_total = int(batch["initial_state"].shape[0])
    if w > 1:   # the same torch.randperm draw on every rank
        seed = torch.randint(0, 2 ** 31 - 1, (1,), dtype=torch.int64)
        dist.broadcast(seed, src=0)
        torch.manual_seed(int(seed.item()))
    optimizer.zero_grad()
    params = [p for p in model.parameters() if p.requires_grad]
    if n_local > 0:
        loss = model.loss(local, lambda1=lambda1, lambda2=lambda2, use_physics_loss=use_physics_loss)
        loss.backward()
        loss_v = loss.detach()
    else:
        loss_v = torch.zeros(())
    packed = PackedGradients(params, 2)
    buf = packed.pack([loss_v * n_local, float(n_local)], weight=float(n_local))
    if w > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    red = packed.unpack(1.0 / max(n_total, 1))
    if gradient_clip is not None and gradient_clip > 0:
        torch.nn.utils.clip_grad_norm_(params, gradient_clip)
    optimizer.step()
    return float(red[0])
