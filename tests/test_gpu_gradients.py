"""GPU gradient parity: hode_rhs_vjp against autograd of the reference's own ode_residual
(golden fixtures), hode_rollout_bwd against autograd through the float64 restatement over the
kernel's own accepted steps, and HybridODENN.loss().backward() against the reference's.
Tolerance: 1e-4 relative (BASELINE.json north_star), in the max norm per gradient tensor."""
import numpy as np
import pytest
import torch

from helpers import cohort, golden, golden_inputs, random_mlp

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def dev(built_lib):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def relmax(a, ref):
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    return float(np.abs(a - ref).max() / (np.abs(ref).max() + 1e-30))


# ---------------------------------------------------------------------------- RHS VJP
@pytest.mark.parametrize("tag", ["mech", "nn64x4", "nn16x2", "nn32x3"])
def test_rhs_vjp_matches_reference_autograd(dev, tag):
    from hybrid_ode_for_glp_1_and_glucose_b200 import ops
    d = golden(f"rhs_vjp_{tag}")
    W = None if tag == "mech" else torch.from_numpy(d["W"])
    ins = {k: torch.from_numpy(v) for k, v in golden_inputs(d).items()}
    gs, gt, gW = ops.rhs_vjp(torch.from_numpy(d["t"]), torch.from_numpy(d["state"]), ins,
                             torch.from_numpy(d["theta"]), W, torch.from_numpy(d["grad_out"]),
                             int(d["hidden"]), int(d["layers"]), device=dev)
    assert relmax(gs.cpu().numpy(), d["grad_state"]) < TOL
    assert relmax(gt.cpu().numpy(), d["grad_theta"]) < TOL
    if W is not None:
        assert relmax(gW.cpu().numpy(), d["grad_W"]) < TOL


@pytest.mark.parametrize("hidden,layers,B", [(64, 4, 1000), (24, 3, 300), (128, 1, 130)])
def test_rhs_vjp_many_rows_vs_restatement_and_deterministic(dev, hidden, layers, B):
    from hybrid_ode_for_glp_1_and_glucose_b200 import ops
    from oracle import torch_restate as R
    y0, _, _ = cohort(B, 2, seed=5, meals=False, tvns=False)
    rng = np.random.default_rng(6)
    state = (y0 * rng.uniform(0.5, 1.5, y0.shape)).astype(np.float32)
    t = rng.uniform(0, 5, B).astype(np.float32)
    meal = rng.uniform(0, 2, B).astype(np.float32)
    tv = (rng.uniform(0, 1, B) > 0.5).astype(np.float32)
    gd = rng.uniform(0, 1500, B).astype(np.float32)
    g = rng.normal(0, 1, (B, 6)).astype(np.float32)
    W = random_mlp(hidden, layers, seed=7, out_std=0.05)
    theta = golden("rhs_mech")["theta"]
    tt = lambda a: torch.from_numpy(a)
    ins = {"meal": tt(meal), "tVNS": tt(tv), "GD": tt(gd)}
    out1 = ops.rhs_vjp(tt(t), tt(state), ins, tt(theta), tt(W), tt(g), hidden, layers, device=dev)
    out2 = ops.rhs_vjp(tt(t), tt(state), ins, tt(theta), tt(W), tt(g), hidden, layers, device=dev)
    for a, b in zip(out1, out2):
        assert torch.equal(a, b), "gradient reduction must be bit-reproducible"
    f64 = lambda a, rg=False: torch.tensor(np.asarray(a, np.float64), requires_grad=rg)
    s64, th64, W64 = f64(state, True), f64(theta, True), f64(W, True)
    out = R.rhs(f64(t), s64, f64(meal), f64(tv), f64(gd), th64, W64, hidden, layers)
    out.backward(f64(g))
    assert relmax(out1[0].cpu().numpy(), s64.grad.numpy()) < TOL
    assert relmax(out1[1].cpu().numpy(), th64.grad.numpy()) < TOL
    assert relmax(out1[2].cpu().numpy(), W64.grad.numpy()) < TOL


def test_ode_residual_autograd_through_module(dev):
    """The call site of the reference's loss: ode_residual(...).backward() fills .grad of every
    network parameter (tests/test_gradient_correctness.py:65 contract)."""
    from hybrid_ode_for_glp_1_and_glucose_b200 import HybridODENN
    d = golden("rhs_vjp_nn16x2")
    m = HybridODENN(nn_hidden=16, nn_layers=2, device=dev)
    with torch.no_grad():
        off = 0
        for _, p in m.nn_residual.named_parameters():
            p.copy_(torch.from_numpy(d["W"][off: off + p.numel()]).reshape(p.shape))
            off += p.numel()
    state = torch.from_numpy(d["state"]).to(dev).requires_grad_(True)
    ext = {"meal": torch.from_numpy(d["meal"]).to(dev), "tVNS": torch.from_numpy(d["tvns"]).to(dev),
           "GD": torch.from_numpy(d["gd"]).to(dev)}
    out = m.ode_residual(torch.from_numpy(d["t"]).to(dev), state, ext)
    out.backward(torch.from_numpy(d["grad_out"]).to(dev))
    gW = torch.cat([p.grad.reshape(-1) for _, p in m.nn_residual.named_parameters()]).cpu().numpy()
    assert relmax(gW, d["grad_W"]) < TOL
    assert relmax(state.grad.cpu().numpy(), d["grad_state"]) < TOL


# ---------------------------------------------------------------------------- discrete adjoint
def _adjoint_case(dev, B, T, hidden, layers, solver, kinks="clip", const_inputs=False, seed=0,
                  precision="fp32", **kw):
    from hybrid_ode_for_glp_1_and_glucose_b200 import ops
    from oracle import torch_restate as R
    y0, t, ins = cohort(B, T, seed=seed, horizon=2.0)
    if const_inputs:
        ins = {"meal": ins["meal"][:, T // 10].copy(), "GD": np.full(B, 300.0, np.float32)}
    W = random_mlp(hidden, layers, seed=seed + 1, out_std=0.05)
    theta = golden("rhs_mech")["theta"]
    rng = np.random.default_rng(seed + 2)
    g = rng.normal(0, 1, (B, T, 6)).astype(np.float32)
    tt = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    traj, info, tape = ops.rollout(tt(y0), tt(t), {k: tt(v) for k, v in ins.items()}, tt(theta), tt(W),
                                   hidden, layers, solver=solver, kinks=kinks, device=dev,
                                   save_steps=True, precision=precision, **kw)
    assert bool((info.status == 0).all())
    g_y0, g_theta, g_W = ops.rollout_bwd(tape, tt(g).to(dev))
    again = ops.rollout_bwd(tape, tt(g).to(dev))
    assert torch.equal(g_W, again[2]) and torch.equal(g_theta, again[1]), "must be bit-reproducible"
    n, st = ops.saved_steps(tape)
    n, st = n.cpu().numpy(), st.cpu().numpy()
    f64 = lambda a, rg=False: torch.tensor(np.asarray(a, np.float64), requires_grad=rg)
    th64, W64 = f64(theta, True), f64(W, True)
    gy_ref = np.zeros((B, 6))
    tr_ref = np.zeros((B, T, 6))
    for b in range(B):
        y64 = f64(y0[b], True)
        ins_b = {k: (v[b] if v.ndim == 2 else np.float64(v[b])) for k, v in ins.items()}
        tr = R.rollout_on_steps(y64, t, ins_b, th64, W64, hidden, layers, list(st[: n[b], b]), solver)
        (tr * f64(g[b])).sum().backward()
        gy_ref[b] = y64.grad.numpy()
        tr_ref[b] = tr.detach().numpy()
    # forward sanity: same steps -> same trajectory up to float32 round-off
    assert relmax(traj.cpu().numpy(), tr_ref) < 2e-5
    return (relmax(g_y0.cpu().numpy(), gy_ref), relmax(g_theta.cpu().numpy(), th64.grad.numpy()),
            relmax(g_W.cpu().numpy(), W64.grad.numpy()))


@pytest.mark.parametrize("hidden,layers", [(64, 4), (16, 2)])
def test_rollout_bwd_rk4_matches_autograd(dev, hidden, layers):
    errs = _adjoint_case(dev, 5, 9, hidden, layers, "rk4", n_substeps=2)
    assert max(errs) < TOL, errs


@pytest.mark.parametrize("kinks", ["clip", "scipy"])
def test_rollout_bwd_dopri5_matches_autograd(dev, kinks):
    errs = _adjoint_case(dev, 5, 13, 64, 4, "dopri5", kinks=kinks, seed=3)
    assert max(errs) < TOL, errs


def test_rollout_bwd_dopri5_const_inputs_small_net(dev):
    errs = _adjoint_case(dev, 4, 7, 32, 3, "dopri5", const_inputs=True, seed=5)
    assert max(errs) < TOL, errs


def test_rollout_bwd_after_tensor_core_forward(dev):
    """The adjoint runs over the steps a 3xTF32 tensor-core forward recorded."""
    errs = _adjoint_case(dev, 5, 13, 64, 4, "dopri5", seed=9, precision="tf32x3")
    assert max(errs) < TOL, errs


def test_rollout_bwd_many_blocks_and_parameter_sets(dev):
    """B spanning several CTAs and S = 3 parameter sets: per-set gradients equal separate launches;
    a failed trajectory contributes nothing."""
    from hybrid_ode_for_glp_1_and_glucose_b200 import ops
    B, T, S = 700, 9, 3
    y0, t, ins = cohort(B, T, seed=11, horizon=1.0)
    theta = np.tile(golden("rhs_mech")["theta"], (S, 1))
    theta[1, 8] = 7.5   # V_max
    theta[2, 1] = 0.03  # k_I
    W = np.stack([random_mlp(64, 4, seed=20 + s, out_std=0.05) for s in range(S)])
    g = np.random.default_rng(12).normal(0, 1, (S, B, T, 6)).astype(np.float32)
    tt = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    tin = {k: tt(v) for k, v in ins.items()}
    _, info, tape = ops.rollout(tt(y0), tt(t), tin, tt(theta), tt(W), solver="dopri5", device=dev,
                                save_steps=True)
    gy, gth, gW = ops.rollout_bwd(tape, tt(g).to(dev))
    for s in range(S):
        _, _, tp = ops.rollout(tt(y0), tt(t), tin, tt(theta[s]), tt(W[s]), solver="dopri5", device=dev,
                               save_steps=True)
        gy1, gth1, gW1 = ops.rollout_bwd(tp, tt(g[s]).to(dev))
        assert relmax(gy[s].cpu().numpy(), gy1.cpu().numpy()) < 1e-6
        assert relmax(gth[s].cpu().numpy(), gth1.cpu().numpy()) < 2e-5
        assert relmax(gW[s].cpu().numpy(), gW1.cpu().numpy()) < 2e-5
    # budget of 3 attempts: every trajectory fails -> all gradients are exactly zero
    _, info, tape = ops.rollout(tt(y0), tt(t), tin, tt(theta[0]), tt(W[0]), solver="dopri5", device=dev,
                                save_steps=True, max_steps=3)
    assert bool((info.status != 0).all())
    gy, gth, gW = ops.rollout_bwd(tape, tt(g[0]).to(dev))
    assert float(gy.abs().max()) == 0.0 and float(gth.abs().max()) == 0.0 and float(gW.abs().max()) == 0.0


# ---------------------------------------------------------------------------- module-level
def _load_W(m, W):
    with torch.no_grad():
        off = 0
        for _, p in m.nn_residual.named_parameters():
            p.copy_(torch.from_numpy(W[off: off + p.numel()]).reshape(p.shape))
            off += p.numel()


@pytest.mark.parametrize("tag", ["nn64x4", "nn16x2"])
def test_loss_backward_matches_reference(dev, tag):
    """model.loss(batch).backward() with the reference's semantics (physics residual + L2 carry the
    gradient; same torch seed -> same physics indices) against the reference's own numbers."""
    from hybrid_ode_for_glp_1_and_glucose_b200 import HybridODENN
    d = golden(f"loss_{tag}")
    m = HybridODENN(nn_hidden=int(d["hidden"]), nn_layers=int(d["layers"]), device=dev)
    _load_W(m, d["W"])
    m.kinks = "scipy"
    to = lambda a: torch.from_numpy(a).to(dev)
    batch = {"initial_state": to(d["y0"]), "observations": to(d["obs"]), "time_points": to(d["t"]),
             "external_inputs": {"meal": to(d["meal"]), "tVNS": to(d["tvns"])}}
    torch.manual_seed(int(d["seed"]))
    loss = m.loss(batch, lambda1=float(d["lambda1"]), lambda2=float(d["lambda2"]))
    loss.backward()
    assert abs(loss.item() - float(d["loss"])) <= 1e-4 * abs(float(d["loss"]))
    gW = torch.cat([p.grad.reshape(-1) for _, p in m.nn_residual.named_parameters()]).cpu().numpy()
    assert relmax(gW, d["grad_W"]) < TOL
    for name, p in m.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), name


def test_loss_stacked_physics_launch_equals_the_loop(dev):
    """loss() runs the physics re-solves as one stacked launch (fused_physics); the reference-shaped
    Python loop must give the same value and the same gradients on the same physics indices."""
    from hybrid_ode_for_glp_1_and_glucose_b200 import HybridODENN
    d = golden("loss_nn64x4")
    to = lambda a: torch.from_numpy(a).to(dev)
    batch = {"initial_state": to(d["y0"]), "observations": to(d["obs"]), "time_points": to(d["t"]),
             "external_inputs": {"meal": to(d["meal"]), "tVNS": to(d["tvns"])}}
    out = {}
    for fused in (True, False):
        m = HybridODENN(nn_hidden=64, nn_layers=4, device=dev)
        _load_W(m, d["W"])
        m.fused_physics = fused
        torch.manual_seed(7)
        loss = m.loss(batch, lambda1=1.0, lambda2=1.0)
        loss.backward()
        out[fused] = (loss.item(), torch.cat([p.grad.reshape(-1) for _, p in m.nn_residual.named_parameters()]).cpu().numpy())
    assert abs(out[True][0] - out[False][0]) <= 1e-6 * abs(out[False][0])
    assert relmax(out[True][1], out[False][1]) < 1e-5


def test_differentiable_forward_trains_the_data_term(dev):
    """differentiable=True: the data-MSE term, which carries no gradient in the reference, now
    reaches the network and initial state; one Adam step lowers it."""
    from hybrid_ode_for_glp_1_and_glucose_b200 import HybridODENN
    y0, t, ins = cohort(64, 13, seed=31, horizon=1.0)
    m = HybridODENN(device=dev)
    m.differentiable = True
    to = lambda a: torch.from_numpy(a).to(dev)
    ext = {k: to(v) for k, v in ins.items()}
    with torch.no_grad():
        target = m(to(y0), to(t), ext) * 1.02
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    y0t = to(y0).requires_grad_(True)
    losses = []
    for _ in range(3):
        opt.zero_grad()
        loss = torch.nn.functional.mse_loss(m(y0t, to(t), ext), target)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    head = m.nn_residual.linears()[-1]
    assert head.weight.grad is not None and float(head.weight.grad.abs().max()) > 0
    assert y0t.grad is not None and float(y0t.grad.abs().max()) > 0
    assert losses[-1] < losses[0]


# ---------------------------------------------------------------------------- tensor-core adjoint
@pytest.mark.parametrize("precision", ["tf32x3", "tf32x2bf16", "f16bf16x2"])
@pytest.mark.parametrize("solver,layers", [("rk4", 4), ("dopri5", 4), ("rk4", 2), ("dopri5", 1)])
def test_tensor_core_adjoint_matches_autograd(dev, solver, layers, precision):
    """The tensor-core precisions route hode_rollout_bwd to the tcgen05 adjoint (hode_adjoint_tc.cu); a rollout made
    in the three-tile mode ('tf32x2bf16', the default) hands its step records to the same 3xTF32 adjoint."""
    kw = dict(n_substeps=2) if solver == "rk4" else {}
    errs = _adjoint_case(dev, 6, 11, 64, layers, solver, seed=21 + layers, precision=precision, **kw)
    assert max(errs) < TOL, errs


def test_tensor_core_adjoint_scipy_kinks_and_constant_inputs(dev):
    """The tensor-core adjoint away from its cached-input fast path: kinks='scipy' (steps may straddle input
    kinks, every stage looks its interval up) and per-trajectory constant inputs with GD present."""
    errs = _adjoint_case(dev, 6, 11, 64, 3, "dopri5", kinks="scipy", seed=31, precision="tf32x3")
    assert max(errs) < TOL, errs
    errs = _adjoint_case(dev, 5, 9, 64, 4, "dopri5", const_inputs=True, seed=33, precision="tf32x3")
    assert max(errs) < TOL, errs


def test_tensor_core_adjoint_parameter_sets_dopri5(dev):
    """S = 3 parameter sets with adaptive steps (per-set step counts, per-set sort and schedule): every set's
    gradients equal a single-set launch of the same inputs up to summation order."""
    from hybrid_ode_for_glp_1_and_glucose_b200 import ops
    B, T, S = 700, 13, 3
    y0, t, ins = cohort(B, T, seed=71, horizon=1.0)
    theta = np.tile(golden("rhs_mech")["theta"], (S, 1))
    theta[1, 8] = 8.0
    theta[2, 0] *= 1.1
    W = np.stack([random_mlp(64, 4, seed=80 + s, out_std=0.05) for s in range(S)])
    g = np.random.default_rng(72).normal(0, 1, (S, B, T, 6)).astype(np.float32)
    tt = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    tin = {k: tt(v) for k, v in ins.items()}
    kw = dict(solver="dopri5", precision="tf32x3", device=dev, save_steps=True)
    _, info, tape = ops.rollout(tt(y0), tt(t), tin, tt(theta), tt(W), **kw)
    assert bool((info.status == 0).all())
    gy, gth, gW = ops.rollout_bwd(tape, tt(g))
    for s in range(S):
        _, info1, tape1 = ops.rollout(tt(y0), tt(t), tin, tt(theta[s]), tt(W[s]), **kw)
        assert torch.equal(info1.n_accept, info.n_accept[s])
        y1, th1, W1 = ops.rollout_bwd(tape1, tt(g[s]))
        assert relmax(gy[s].cpu().numpy(), y1.cpu().numpy()) < 1e-5
        assert relmax(gth[s].cpu().numpy(), th1.cpu().numpy()) < 1e-5
        assert relmax(gW[s].cpu().numpy(), W1.cpu().numpy()) < 1e-5


def test_tensor_core_adjoint_equals_fp32_adjoint_at_scale(dev):
    """Fixed-step rollouts take identical steps on both paths, so the two adjoints must agree:
    B spans several tiles per CTA is exercised with a small grid by S = 2 parameter sets."""
    from hybrid_ode_for_glp_1_and_glucose_b200 import ops
    B, T, S = 1500, 7, 2
    y0, t, ins = cohort(B, T, seed=41, horizon=0.6)
    ins["GD"] = np.linspace(0, 800, B).astype(np.float32)
    theta = np.tile(golden("rhs_mech")["theta"], (S, 1))
    theta[1, 8] = 8.0
    W = np.stack([random_mlp(64, 4, seed=50 + s, out_std=0.05) for s in range(S)])
    g = np.random.default_rng(42).normal(0, 1, (S, B, T, 6)).astype(np.float32)
    tt = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    tin = {k: tt(v) for k, v in ins.items()}
    out = {}
    for prec in ("fp32", "tf32x3"):
        _, info, tape = ops.rollout(tt(y0), tt(t), tin, tt(theta), tt(W), solver="rk4", n_substeps=2,
                                    precision=prec, device=dev, save_steps=True)
        assert bool((info.status == 0).all())
        out[prec] = ops.rollout_bwd(tape, tt(g).to(dev))
        again = ops.rollout_bwd(tape, tt(g).to(dev))
        assert all(torch.equal(a, b) for a, b in zip(out[prec], again)), "bit-reproducible"
    for a, b, name in zip(out["tf32x3"], out["fp32"], ("y0", "theta", "W")):
        for s in range(S):
            assert relmax(a[s].cpu().numpy(), b[s].cpu().numpy()) < TOL, (name, s)


def test_config3_gradient_properties_at_size(dev):
    """32 768 trajectories per GPU (configs/default.yaml on 8 GPUs), tensor-core forward + adjoint:
    exact properties of the discrete adjoint that need no oracle."""
    from hybrid_ode_for_glp_1_and_glucose_b200 import ops
    B, T = 32768, 61
    y0, t, ins = cohort(B, T, seed=77)
    W = random_mlp(64, 4, seed=1234, out_std=0.05)
    theta = golden("rhs_mech")["theta"]
    tt = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    _, info, tape = ops.rollout(tt(y0), tt(t), {k: tt(v) for k, v in ins.items()}, tt(theta), tt(W),
                                solver="dopri5", precision="tf32x3", device=dev, save_steps=True)
    assert bool((info.status == 0).all())
    # (1) the first observation IS the initial state: d traj[:,0,:] / d y0 = identity, no parameter gradient
    g = torch.zeros((B, T, 6), device=dev)
    g[:, 0, :] = torch.randn((B, 6), device=dev, generator=torch.Generator(dev).manual_seed(1))
    gy, gth, gW = ops.rollout_bwd(tape, g)
    assert torch.equal(gy, g[:, 0, :]) and float(gth.abs().max()) == 0.0 and float(gW.abs().max()) == 0.0
    # (2) linearity in the cotangent, exact for a power-of-two scale
    g = torch.randn((B, T, 6), device=dev, generator=torch.Generator(dev).manual_seed(2)) / (B * T)
    a = ops.rollout_bwd(tape, g)
    b = ops.rollout_bwd(tape, 4.0 * g)
    for x, y in zip(a, b):
        assert torch.equal(4.0 * x, y)
    assert all(bool(torch.isfinite(x).all()) for x in a) and float(a[2].abs().max()) > 0
    # (3) GE has zero mechanistic derivative and enters the network: its y0-gradient is finite and the
    #     FFA row of theta (p_7..p_9) receives gradient
    assert float(a[1][14:].abs().max()) > 0


def test_tensor_core_adjoint_repeatable_over_many_launches(dev):
    """Several tiles per CTA, every launch scheduled from scratch: 15 fwd + adjoint passes must return
    bit-identical gradients.  (Regression test for a phase race between the main and helper warps of a
    tile that showed up as an occasional hang or a wrong gradient, never in single-tile cases.)"""
    from hybrid_ode_for_glp_1_and_glucose_b200 import ops
    B, T = 32768, 61
    y0, t, ins = cohort(B, T, seed=1000)
    W = random_mlp(64, 4, seed=1234, out_std=0.05)
    theta = golden("rhs_mech")["theta"]
    tt = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    args = (tt(y0), tt(t), {k: tt(v) for k, v in ins.items()}, tt(theta), tt(W))
    g = torch.randn((B, T, 6), device=dev, generator=torch.Generator(dev).manual_seed(3)) / (B * T)
    ref = None
    for _ in range(15):
        _, info, tape = ops.rollout(*args, solver="dopri5", precision="tf32x3", device=dev, save_steps=True)
        out = ops.rollout_bwd(tape, g)
        torch.cuda.synchronize()
        if ref is None:
            ref = [o.clone() for o in out]
            assert bool((info.status == 0).all()) and all(bool(torch.isfinite(o).all()) for o in out)
        else:
            assert all(torch.equal(a, b) for a, b in zip(ref, out))


@pytest.mark.parametrize("precision,S", [("tf32x3", 1), ("fp32", 2)])
def test_fused_data_loss_step_equals_the_composition(dev, precision, S):
    """hode_loss_fused_fwd_bwd = hode_rollout_fwd + mean squared residual + hode_rollout_bwd: identical
    trajectories and gradients, loss equal to torch's mse_loss."""
    from hybrid_ode_for_glp_1_and_glucose_b200 import ops
    B, T = 700, 13
    y0, t, ins = cohort(B, T, seed=5, horizon=1.0)
    theta = np.tile(golden("rhs_mech")["theta"], (S, 1)) if S > 1 else golden("rhs_mech")["theta"]
    W = np.stack([random_mlp(64, 4, seed=60 + s, out_std=0.05) for s in range(S)]) if S > 1 else random_mlp(64, 4, seed=60, out_std=0.05)
    obs = (np.repeat(y0[:, None, :], T, 1) * (1 + 0.1 * np.random.default_rng(6).normal(0, 1, (B, T, 6)))).astype(np.float32)
    tt = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    args = (tt(y0), tt(t), {k: tt(v) for k, v in ins.items()}, tt(theta), tt(W))
    kw = dict(solver="dopri5", precision=precision, device=dev)
    loss, g_y0, g_theta, g_W, traj, info = ops.data_loss_step(*args, tt(obs), need_y0=True, **kw)
    traj2, info2, tape = ops.rollout(*args, save_steps=True, **kw)
    assert torch.equal(traj, traj2) and torch.equal(info.status, info2.status)
    resid = traj2 - tt(obs)
    ref_loss = (resid.double() ** 2).mean(dim=tuple(range(resid.dim() - 3, resid.dim())))
    assert float((loss.double() - ref_loss).abs().max()) <= 1e-6 * float(ref_loss.abs().max())
    r_y0, r_theta, r_W = ops.rollout_bwd(tape, resid * (2.0 / (B * T * 6)))   # one fp32 multiply, as in the kernel
    assert torch.equal(g_theta, r_theta) and torch.equal(g_W, r_W) and torch.equal(g_y0, r_y0)


def test_tensor_core_adjoint_more_sets_than_the_sort_key_holds(dev):
    """S > 4096 parameter sets: the composite sort key (set index above 20 bits of step count) no longer fits
    32 bits and the schedule falls back to index order with scanned tile costs.  The gradients of a few sets
    must equal those of the same sets launched on their own (sorted path)."""
    from hybrid_ode_for_glp_1_and_glucose_b200 import ops
    B, T, S = 3, 5, 4100
    y0, t, ins = cohort(B, T, seed=91, horizon=0.4)
    rng = np.random.default_rng(92)
    theta = np.tile(golden("rhs_mech")["theta"], (S, 1)) * (1 + 0.05 * rng.normal(0, 1, (S, 17))).astype(np.float32)
    W0 = random_mlp(64, 2, seed=93, out_std=0.05)
    W = (W0[None, :] * (1 + 0.1 * rng.normal(0, 1, (S, 1)))).astype(np.float32)
    g = rng.normal(0, 1, (S, B, T, 6)).astype(np.float32)
    tt = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    tin = {k: tt(v) for k, v in ins.items()}
    kw = dict(hidden=64, layers=2, solver="dopri5", precision="tf32x3", device=dev, save_steps=True, max_saved_steps=32)
    _, info, tape = ops.rollout(tt(y0), tt(t), tin, tt(theta.astype(np.float32)), tt(W), **kw)
    assert bool((info.status == 0).all())
    gy, gth, gW = ops.rollout_bwd(tape, tt(g))
    assert bool(torch.isfinite(gW).all()) and bool(torch.isfinite(gth).all())
    for s in (0, 1, 2049, 4099):
        _, _, tape1 = ops.rollout(tt(y0), tt(t), tin, tt(theta[s].astype(np.float32)), tt(W[s]), **kw)
        y1, th1, W1 = ops.rollout_bwd(tape1, tt(g[s]))
        assert relmax(gy[s].cpu().numpy(), y1.cpu().numpy()) < 1e-5, s
        assert relmax(gth[s].cpu().numpy(), th1.cpu().numpy()) < 1e-5, s
        assert relmax(gW[s].cpu().numpy(), W1.cpu().numpy()) < 1e-5, s


def test_tensor_core_adjoint_per_trajectory_time_grids(dev):
    """Per-row observation grids (t_span [B,T], the clinical-shaped configuration) through the tensor-core
    adjoint, against float64 autograd over the recorded steps of every trajectory."""
    from hybrid_ode_for_glp_1_and_glucose_b200 import ops
    from oracle import torch_restate as R
    B, T = 5, 9
    y0, t, ins = cohort(B, T, seed=101, horizon=1.5)
    rng = np.random.default_rng(102)
    tb = np.tile(t[None, :], (B, 1)).astype(np.float64)
    tb[:, 1:-1] += rng.uniform(-0.02, 0.02, (B, T - 2))          # jittered interior points, still increasing
    tb = np.sort(tb, axis=1).astype(np.float32)
    W = random_mlp(64, 4, seed=103, out_std=0.05)
    theta = golden("rhs_mech")["theta"]
    g = rng.normal(0, 1, (B, T, 6)).astype(np.float32)
    tt = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    traj, info, tape = ops.rollout(tt(y0), tt(tb), {k: tt(v) for k, v in ins.items()}, tt(theta), tt(W),
                                   solver="dopri5", precision="tf32x3", device=dev, save_steps=True)
    assert bool((info.status == 0).all())
    g_y0, g_theta, g_W = ops.rollout_bwd(tape, tt(g).to(dev))
    n, st = ops.saved_steps(tape)
    n, st = n.cpu().numpy(), st.cpu().numpy()
    f64 = lambda a, rg=False: torch.tensor(np.asarray(a, np.float64), requires_grad=rg)
    th64, W64 = f64(theta, True), f64(W, True)
    gy_ref = np.zeros((B, 6))
    for b in range(B):
        y64 = f64(y0[b], True)
        ins_b = {k: (v[b] if v.ndim == 2 else np.float64(v[b])) for k, v in ins.items()}
        tr = R.rollout_on_steps(y64, tb[b], ins_b, th64, W64, 64, 4, list(st[: n[b], b]), "dopri5")
        (tr * f64(g[b])).sum().backward()
        gy_ref[b] = y64.grad.numpy()
    assert relmax(g_y0.cpu().numpy(), gy_ref) < TOL
    assert relmax(g_theta.cpu().numpy(), th64.grad.numpy()) < TOL
    assert relmax(g_W.cpu().numpy(), W64.grad.numpy()) < TOL


def test_config5_clinical_shape_adjoint_properties(dev):
    """mimic_clinical-shaped cohort (T = 577 over 48 h, per-row jittered grids, irregular meal / dose events):
    the tensor-core adjoint over 100+ recorded steps per trajectory and a non-default record capacity — exact
    properties that need no oracle (linearity in the cotangent for a power-of-two scale, the t0 row is the
    identity, finiteness)."""
    from hybrid_ode_for_glp_1_and_glucose_b200 import ops
    from hybrid_ode_for_glp_1_and_glucose_b200.synthetic import clinical_cohort
    B, T = 384, 577
    y0, t, ins = clinical_cohort(B, T, seed=5)
    # as in test_config5_clinical_shape_long_horizon_per_row_grids: start insulin / glucagon near their set points and
    # keep the residual small, so that every trajectory stays physiological over 48 h
    rng = np.random.default_rng(5)
    y0[:, 1] = 60.0 * rng.normal(1, 0.1, B)
    y0[:, 2] = 80.0 * rng.normal(1, 0.05, B)
    W = random_mlp(seed=17, out_std=0.001)
    theta = golden("rhs_mech")["theta"]
    tt = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    args = (tt(y0), tt(t), {k: tt(v) for k, v in ins.items()}, tt(theta), tt(W))
    _, info, tape = ops.rollout(*args, solver="dopri5", precision="tf32x3", device=dev, save_steps=True,
                                max_saved_steps=8192)
    assert bool((info.status == 0).all()), (torch.unique(info.status).tolist(), int(info.n_accept.max()))
    assert int(info.n_accept.max()) > 100      # long records (the 5 h cohorts take ~33 steps)
    g = torch.zeros((B, T, 6), device=dev)
    g[:, 0, :] = torch.randn((B, 6), device=dev, generator=torch.Generator(dev).manual_seed(1))
    gy, gth, gW = ops.rollout_bwd(tape, g)
    assert torch.equal(gy, g[:, 0, :]) and float(gth.abs().max()) == 0.0 and float(gW.abs().max()) == 0.0
    g = torch.randn((B, T, 6), device=dev, generator=torch.Generator(dev).manual_seed(2)) / (B * T)
    a = ops.rollout_bwd(tape, g)
    b = ops.rollout_bwd(tape, 0.5 * g)
    for x, y in zip(a, b):
        assert bool(torch.isfinite(x).all()) and torch.equal(0.5 * x, y)
    assert float(a[2].abs().max()) > 0 and float(a[1].abs().max()) > 0
