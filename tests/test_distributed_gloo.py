"""world_size-2 gloo tests (CPU) of the data-parallel host logic: trajectory sharding and the
single packed all-reduce of gradients + loss scalars."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world_size, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        from hybrid_ode_for_glp_1_and_glucose_b200 import HybridODENN, distributed as D
        torch.manual_seed(0)
        m = HybridODENN(nn_hidden=16, nn_layers=2, device="cpu")
        params = [p for p in m.parameters() if p.requires_grad]
        # rank-dependent gradients: p.grad = (rank + 1) * ones, one parameter left without grad
        for i, p in enumerate(params):
            p.grad = None if i == 1 else torch.full_like(p, float(rank + 1))
        red = D.allreduce_gradients(params, [10.0 * (rank + 1), 1.0], average=False)
        ok = all(torch.equal(p.grad, torch.full_like(p, 0.0 if i == 1 else 3.0))
                 for i, p in enumerate(params))
        ok = ok and red.tolist() == [30.0, 2.0]
        # averaged variant
        for p in params:
            p.grad = torch.full_like(p, float(rank))
        red = D.allreduce_gradients(params, [float(rank)], average=True)
        ok = ok and all(torch.equal(p.grad, torch.full_like(p, 0.5)) for p in params)
        ok = ok and abs(float(red[0]) - 0.5) < 1e-7
        # sharding covers the batch exactly once
        B, T = 11, 5
        batch = {"initial_state": torch.arange(B * 6.0).reshape(B, 6),
                 "observations": torch.zeros(B, T, 6), "time_points": torch.linspace(0, 1, T),
                 "external_inputs": {"meal": torch.arange(B * T * 1.0).reshape(B, T),
                                     "dose": torch.arange(B * 1.0)}}
        loc = D.shard_batch(batch)
        lo, hi = D.shard_bounds(B, rank, world_size)
        ok = ok and torch.equal(loc["initial_state"], batch["initial_state"][lo:hi])
        ok = ok and torch.equal(loc["external_inputs"]["meal"], batch["external_inputs"]["meal"][lo:hi])
        ok = ok and torch.equal(loc["external_inputs"]["dose"], batch["external_inputs"]["dose"][lo:hi])
        ok = ok and loc["time_points"].shape == (T,)
        # pack(weight): gradients scaled by the shard size, scalars untouched (sharded_loss_step's global-batch mean)
        for p in params:
            p.grad = torch.full_like(p, 2.0)
        pg = D.PackedGradients(params, 2)
        buf = pg.pack([5.0 * (rank + 1), float(rank + 1)], weight=float(rank + 1))
        dist.all_reduce(buf)
        sc = pg.unpack(1.0 / 3.0)
        ok = ok and all(torch.allclose(p.grad, torch.full_like(p, 2.0)) for p in params)   # (1*2 + 2*2) / 3
        ok = ok and abs(float(sc[0]) - 5.0) < 1e-6 and abs(float(sc[1]) - 1.0) < 1e-6
        sizes = [torch.zeros(1, dtype=torch.long) for _ in range(world_size)]
        dist.all_gather(sizes, torch.tensor([hi - lo]))
        ok = ok and int(sum(s.item() for s in sizes)) == B
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_packed_allreduce_and_sharding_world2():
    world_size = 2
    port = _free_port()
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_worker, args=(world_size, port, results), nprocs=world_size, join=True)
        assert dict(results) == {0: True, 1: True}


def test_shard_bounds_properties():
    from hybrid_ode_for_glp_1_and_glucose_b200.distributed import shard_bounds
    for n in (0, 1, 7, 8, 262144, 1048577):
        for w in (1, 2, 3, 8):
            cuts = [shard_bounds(n, r, w) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)
