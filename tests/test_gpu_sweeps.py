"""GPU tests of the hode_rollout_fwd_ex options: per-trajectory theta (the Sobol sweep of reference
plots/plot_all.py:124-224), the launch-order hint, and the output-state mask — against the existing S-set path and the
CPU oracle."""
import numpy as np
import pytest
import torch

from helpers import cohort, random_mlp, rel_err, rel_err_report

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(built_lib):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def _saltelli_like(P, seed=0):
    """P parameter sets inside the reference's Sobol bounds (plots/plot_all.py:138-147)."""
    names = ["a_GI", "k_I", "rho", "E_max", "V_max", "K_m", "k_L"]
    lo = np.array([0.008, 0.02, 0.002, 0.08, 7.0, 5.5, 0.015])
    hi = np.array([0.012, 0.03, 0.004, 0.12, 11.0, 8.5, 0.025])
    u = np.random.default_rng(seed).uniform(0, 1, (P, 7))
    return names, (lo + u * (hi - lo)).astype(np.float32)


@pytest.mark.parametrize("precision,solver", [("tf32x3", "dopri5"), ("fp32", "dopri5"), ("tf32x3", "rk4"),
                                              ("tf32x2bf16", "dopri5"), ("tf32x2bf16", "rk4"),
                                              ("f16bf16x2", "dopri5"), ("f16bf16x2", "rk4")])
def test_theta_per_trajectory_equals_the_parameter_set_sweep(dev, oracle, precision, solver):
    """theta [P,17] in per-trajectory mode == theta [P,17] as P parameter sets over ONE trajectory (the round-1 path,
    one 128-row tile per set on the tensor cores), bit for bit; and == the oracle on a few rows."""
    from hybrid_ode_for_glp_1_and_glucose_b200 import HybridODENN, ops, sensitivity
    P = 300
    names, samples = _saltelli_like(P)
    m = HybridODENN(device=dev)
    torch.manual_seed(1)
    with torch.no_grad():
        for p in m.nn_residual.parameters():
            p.copy_(0.05 * torch.randn_like(p))
    table = sensitivity.theta_table(m, names, torch.from_numpy(samples))
    y0 = torch.tensor([5.0, 60.0, 80.0, 0.0, 0.0, 1.0])
    t = torch.linspace(0, 5, 61)
    meal = torch.zeros(61); meal[6] = 75.0
    ext = {"meal": meal, "tVNS": torch.zeros(61)}
    kw = dict(solver=solver, precision=precision, n_substeps=2)
    traj, info = sensitivity.sweep(m, table, y0, t, ext, **kw)
    assert traj.shape == (P, 61, 6) and bool((info.status == 0).all())
    _, W = m.packed_parameters(None)
    sets, info2 = ops.rollout(y0.reshape(1, 6), t, {k: v.reshape(1, -1) for k, v in ext.items()}, table.to(dev),
                              W.to(dev).reshape(1, -1).expand(P, -1).contiguous(), device=dev, **kw)
    assert torch.equal(traj, sets[:, 0])
    assert torch.equal(info.n_accept.reshape(-1), info2.n_accept.reshape(-1))
    rows = [0, 17, 299]
    for r in rows:
        if solver == "rk4":
            ref, _, _, _ = oracle.rollout(y0.numpy()[None], t.numpy(), {k: v.numpy()[None] for k, v in ext.items()},
                                          table[r].numpy(), W.detach().cpu().numpy(), solver="rk4", n_substeps=2)
            assert rel_err(traj[r].cpu().numpy()[None], ref) < 1e-5, rel_err_report(traj[r].cpu().numpy()[None], ref)


def test_sobol_outputs_match_a_per_set_loop(dev):
    """The three outputs of the reference's sensitivity analysis (glucose AUC by np.trapz, insulin peak, mean GLP-1
    after the meal; plots/plot_all.py:183-187) from the one-launch sweep == the same formulas applied to per-set
    forward() calls of the drop-in class with the ODE buffers overwritten the way the reference does (:170-174)."""
    from hybrid_ode_for_glp_1_and_glucose_b200 import HybridODENN, sensitivity
    names, samples = _saltelli_like(24, seed=3)
    m = HybridODENN(device=dev)
    out = sensitivity.sobol_outputs(m, names, torch.from_numpy(samples)).cpu().numpy()
    assert out.shape == (24, 3)
    y0 = torch.tensor([5.0, 60.0, 80.0, 0.0, 0.0, 1.0], device=dev)
    t = torch.linspace(0, 5, 61, device=dev)
    meal = torch.zeros(61, device=dev); meal[6] = 75.0
    ext = {"meal": meal.unsqueeze(0), "tVNS": torch.zeros(61, device=dev).unsqueeze(0)}
    for i in range(0, 24, 5):
        for name, value in zip(names, samples[i]):
            setattr(m.ode_core, name, torch.tensor(float(value), device=dev))
        tr = m.forward(y0.unsqueeze(0), t, ext).squeeze(0).cpu().numpy()
        ref = np.array([np.trapezoid(tr[:, 0], dx=5 / 60), tr[:, 1].max(), tr[6:, 3].mean()])
        np.testing.assert_allclose(out[i], ref, rtol=2e-6)


def test_launch_order_and_state_mask_do_not_change_results(dev):
    from hybrid_ode_for_glp_1_and_glucose_b200 import ops
    y0, t, ins = cohort(3000, seed=51)
    W = random_mlp(seed=52, out_std=0.05)
    from hybrid_ode_for_glp_1_and_glucose_b200.synthetic import THETA_DEFAULT
    tt = lambda a: torch.from_numpy(a)
    args = (tt(y0), tt(t), {k: tt(v) for k, v in ins.items()}, tt(THETA_DEFAULT), tt(W))
    a, info = ops.rollout(*args, solver="dopri5", precision="tf32x3", device=dev)
    order = ops.launch_order(info)
    att = (info.n_accept + info.n_reject).cpu().numpy()
    assert sorted(order.cpu().numpy().tolist()) == list(range(3000))
    assert (np.diff(att[order.cpu().numpy()]) <= 0).all(), "longest first"
    b, info_b = ops.rollout(*args, solver="dopri5", precision="tf32x3", device=dev, order=order)
    assert torch.equal(a, b) and torch.equal(info.n_accept, info_b.n_accept)
    for prec in ("tf32x3", "fp32"):
        full, _ = ops.rollout(*args, solver="dopri5", precision=prec, device=dev)
        part, _ = ops.rollout(*args, solver="dopri5", precision=prec, device=dev, out_state_mask=0b001011)
        assert part.shape == (3000, 61, 3)
        assert torch.equal(part, full[..., [0, 1, 3]])
    with pytest.raises(Exception):
        ops.rollout(*args, solver="dopri5", device=dev, order=order[:10])


def test_host_entry_with_state_mask(dev):
    """hode_rollout_fwd_host_ex: only the selected state columns travel back to the host."""
    import ctypes
    from hybrid_ode_for_glp_1_and_glucose_b200 import _lib, ops
    from hybrid_ode_for_glp_1_and_glucose_b200.synthetic import THETA_DEFAULT
    B = 20000
    y0, t, ins = cohort(B, seed=53)
    W = random_mlp(seed=54, out_std=0.05)
    tt = lambda a: torch.from_numpy(a)
    ref, _ = ops.rollout(tt(y0), tt(t), {k: tt(v) for k, v in ins.items()}, tt(THETA_DEFAULT), tt(W), solver="dopri5",
                         precision="tf32x3", device=dev)
    cfg, _ = ops.prepare(tt(y0), tt(t), {k: tt(v) for k, v in ins.items()}, tt(THETA_DEFAULT), tt(W), 64, 4, torch.device("cpu"))
    cfg.solver, cfg.mlp = _lib.SOLVER_DOPRI5, _lib.MLP_TF32X3
    cfg.kink_mode = _lib.KINK_CLIP
    out = np.empty((B, 61, 2), dtype=np.float32)
    st = np.empty(B, dtype=np.int32)
    opts = _lib.new_fwd_opts(out_state_mask=0b000101)
    p = lambda a: ctypes.c_void_p(a.ctypes.data)
    rc = _lib.lib().hode_rollout_fwd_host_ex(ctypes.byref(cfg), ctypes.byref(opts), p(y0), p(t), p(ins["meal"]), p(ins["tVNS"]), None,
                                             p(THETA_DEFAULT), p(W), p(out), p(st), None,
                                             ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, "hode_rollout_fwd_host_ex")
    assert (st == 0).all()
    assert np.array_equal(out, ref.cpu().numpy()[..., [0, 2]])
