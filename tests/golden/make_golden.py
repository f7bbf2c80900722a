"""Generate the golden fixtures in tests/golden/ by RUNNING THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference, which does not exist on the GPU
box):  python tests/golden/make_golden.py
The .npz files it writes are committed; tests never import the reference.

Every case stores the exact inputs handed to the reference and the outputs it returned:
  rhs_*      HybridODENN.ode_residual / ODECore.forward   (models/hybrid_ode_nn.py:108,
             models/ode_core.py:81), inputs from tests/test_ode_jacobians.py:68-75,143-150,183-185
  rollout_*  HybridODENN.forward                            (models/hybrid_ode_nn.py:136)
             with solver='rk45' (SciPy RK45 = DP5(4)) and 'dopri5' (SciPy DOP853)
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("HODE_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.modules.setdefault("arviz", types.ModuleType("arviz"))  # inference/mcmc.py:11 imports it

from models.hybrid_ode_nn import HybridODENN  # noqa: E402
from train.train_hybrid import GlucoseDataset  # noqa: E402

torch.set_num_threads(1)
CPU = torch.device("cpu")


def pack_W(model):
    return np.concatenate([p.detach().cpu().numpy().reshape(-1).astype(np.float32)
                           for _, p in model.nn_residual.named_parameters()])


def theta_of(model):
    return np.array([float(b) for _, b in model.ode_core.named_buffers()], dtype=np.float32)


def make_model(hidden=64, layers=4, seed=0, out_std=0.0, ode_params=None):
    torch.manual_seed(seed)
    m = HybridODENN(ode_params=ode_params, nn_hidden=hidden, nn_layers=layers, device=CPU)
    if out_std > 0:
        with torch.no_grad():
            last = m.nn_residual.network[-1]
            last.weight.normal_(0.0, out_std)
            last.bias.normal_(0.0, out_std)
            # give hidden biases some life too so ReLU patterns are non-trivial
            for layer in m.nn_residual.network[:-1]:
                if isinstance(layer, torch.nn.Linear):
                    layer.bias.normal_(0.0, 0.05)
                    layer.weight.mul_(10.0)  # undo the gain-0.1 init: O(1) activations
    return m


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, {k: getattr(v, "shape", None) for k, v in arrays.items()})


def rhs_cases():
    states = np.array([
        [5.0, 100.0, 50.0, 20.0, 0.0, 1.0],          # tests/test_ode_jacobians.py:68-75
        [8.0, 150.0, 40.0, 30.0, 0.5, 1.2],          # :143-150
        [0.1, 1.0, 1.0, 0.1, 0.0, 0.1],              # :183-185 (extreme low)
        [30.0, 1000.0, 200.0, 100.0, 1.0, 5.0],      # extreme high
        [5.0, 60.0, 80.0, 0.0, 0.0, 1.0],            # steady state
    ], dtype=np.float32)
    t = np.array([0.0, 1.0, 0.5, 2.0, 3.0], dtype=np.float32)
    meal = np.array([0.0, 10.0, 0.0, 5.0, 0.0], dtype=np.float32)
    tvns = np.array([0.0, 0.0, 1.0, 1.0, 0.0], dtype=np.float32)
    gd = np.array([0.0, 0.0, 0.0, 500.0, 1500.0], dtype=np.float32)
    for tag, hidden, layers, std in (("mech", 64, 4, 0.0), ("nn64x4", 64, 4, 0.05),
                                     ("nn16x2", 16, 2, 0.05), ("nn32x3", 32, 3, 0.1)):
        m = make_model(hidden, layers, seed=1, out_std=std)
        ext = {"meal": torch.tensor(meal), "tVNS": torch.tensor(tvns), "GD": torch.tensor(gd)}
        with torch.no_grad():
            out_b = m.ode_residual(torch.tensor(t), torch.tensor(states), ext).numpy()
            out_1 = np.stack([
                m.ode_residual(torch.tensor(t[i]), torch.tensor(states[i]),
                               {k: v[i] for k, v in ext.items()}).numpy()
                for i in range(len(t))])
        save(f"rhs_{tag}", t=t, state=states, meal=meal, tvns=tvns, gd=gd, theta=theta_of(m),
             W=pack_W(m), hidden=hidden, layers=layers, out_batched=out_b, out_single=out_1)


class StepLog:
    """Records (t_n, h_n) of every accepted step SciPy's RK45 takes inside the reference's
    forward(), by wrapping RungeKutta._step_impl (scipy/integrate/_ivp/rk.py:111)."""

    def __enter__(self):
        import scipy.integrate._ivp.rk as rk
        self.rk, self.orig, self.log = rk, rk.RungeKutta._step_impl, []
        outer = self

        def patched(solver):
            t0 = solver.t
            ok = outer.orig(solver)
            outer.log.append((t0, solver.t - t0, solver.nfev))
            return ok
        rk.RungeKutta._step_impl = patched
        return self

    def __exit__(self, *a):
        self.rk.RungeKutta._step_impl = self.orig


def run_forward(m, y0, t, ext, solver, rtol=1e-6, atol=1e-8):
    ext_t = {k: torch.tensor(v) for k, v in ext.items()} if ext else None
    with torch.no_grad():
        return m.forward(torch.tensor(y0), torch.tensor(t), ext_t, solver=solver, rtol=rtol,
                         atol=atol).numpy()


def rollout_fig2():
    # plots/plot_all.py:164-187 scenario
    y0 = np.array([[5.0, 60.0, 80.0, 0.0, 0.0, 1.0]], dtype=np.float32)
    t = np.linspace(0, 5, 61).astype(np.float32)
    meal = np.zeros((1, 61), dtype=np.float32)
    meal[0, 6] = 75.0
    tv = np.zeros((1, 61), dtype=np.float32)
    m = make_model()
    with StepLog() as sl:
        o45 = run_forward(m, y0, t, {"meal": meal, "tVNS": tv}, "rk45")
    steps = np.array(sl.log, dtype=np.float64)
    o853 = run_forward(m, y0, t, {"meal": meal, "tVNS": tv}, "dopri5")
    save("rollout_fig2", y0=y0, t=t, meal=meal, tvns=tv, theta=theta_of(m), W=pack_W(m),
         hidden=64, layers=4, out_rk45=o45, out_dopri5=o853, steps_rk45=steps)


def windows_4gi():
    ds = GlucoseDataset(os.path.join(REF, "data/4gi_dataset.csv"), 61, 30, True)
    items = [ds[i] for i in range(len(ds))]
    y0 = np.stack([it["initial_state"].numpy() for it in items])
    obs = np.stack([it["observations"].numpy() for it in items])
    t = np.stack([it["time_points"].numpy() for it in items])
    meal = np.stack([it["external_inputs"]["meal"].numpy() for it in items])
    tv = np.stack([it["external_inputs"]["tVNS"].numpy() for it in items])
    return y0, obs, t, meal, tv, ds


def rollout_4gi():
    y0, obs, t, meal, tv, ds = windows_4gi()
    for tag, std in (("mech", 0.0), ("nn", 0.05)):
        m = make_model(seed=2, out_std=std)
        out = {s: run_forward(m, y0, t, {"meal": meal, "tVNS": tv}, s) for s in ("rk45", "dopri5")}
        save(f"rollout_4gi_{tag}", y0=y0, t=t, meal=meal, tvns=tv, obs=obs, theta=theta_of(m),
             W=pack_W(m), hidden=64, layers=4, state_mean=ds.state_mean, state_std=ds.state_std,
             out_rk45=out["rk45"], out_dopri5=out["dopri5"])


def rollout_physio():
    """Physiological-unit cohort (SURVEY §8d config 2/3 statistics), shared t, hybrid net,
    small nets, constant inputs (the physics re-solve pattern, models/hybrid_ode_nn.py:320)."""
    rng = np.random.default_rng(3)
    B = 6
    y0 = np.stack([7.0 * rng.normal(1, 0.1, B), 50.0 * rng.normal(1, 0.15, B),
                   25.0 * rng.normal(1, 0.15, B), 10.0 * rng.normal(1, 0.15, B),
                   np.zeros(B), np.ones(B)], axis=1).astype(np.float32)
    t = np.linspace(0, 5, 61).astype(np.float32)
    meal = np.zeros((B, 61), dtype=np.float32)
    meal[:, 6] = 1.0
    meal[:, 30] = rng.uniform(0.5, 1.5, B)
    tv = np.zeros((B, 61), dtype=np.float32)
    tv[::2, 20:40] = 1.0
    for tag, hidden, layers in (("nn64x4", 64, 4), ("nn16x2", 16, 2)):
        m = make_model(hidden, layers, seed=4, out_std=0.02)
        out = run_forward(m, y0, t, {"meal": meal, "tVNS": tv}, "rk45")
        save(f"rollout_physio_{tag}", y0=y0, t=t, meal=meal, tvns=tv, theta=theta_of(m),
             W=pack_W(m), hidden=hidden, layers=layers, out_rk45=out)
    # constant inputs, T=2 local-time window
    m = make_model(seed=5, out_std=0.02)
    t2 = np.array([0.0, 0.1], dtype=np.float32)
    mc = rng.uniform(0, 2, B).astype(np.float32)
    tc = (rng.uniform(0, 1, B) > 0.5).astype(np.float32)
    with StepLog() as sl:
        out = run_forward(m, y0[:1], t2, {"meal": mc[:1], "tVNS": tc[:1]}, "rk45")
    steps = np.array(sl.log, dtype=np.float64)
    out = run_forward(m, y0, t2, {"meal": mc, "tVNS": tc}, "rk45")
    out8 = run_forward(m, y0, t2, {"meal": mc, "tVNS": tc}, "dopri5")
    save("rollout_const_T2", y0=y0, t=t2, meal=mc, tvns=tc, theta=theta_of(m), W=pack_W(m),
         hidden=64, layers=4, out_rk45=out, out_dopri5=out8, steps_rk45_traj0=steps)
    # per-row time grids with jitter + non-default ODE parameters
    m = make_model(seed=6, out_std=0.02, ode_params={"k_L": 0.05, "V_max": 7.0, "rho": 0.01})
    T = 25
    tj = np.sort(np.linspace(0, 2, T)[None, :] + rng.uniform(-0.02, 0.02, (B, T)), axis=1)
    tj = tj.astype(np.float32)
    mj = (rng.uniform(0, 1, (B, T)) > 0.85).astype(np.float32) * rng.lognormal(0, 0.5, (B, T)).astype(np.float32)
    out = run_forward(m, y0, tj, {"meal": mj}, "rk45")
    save("rollout_perrow_t", y0=y0, t=tj, meal=mj, theta=theta_of(m), W=pack_W(m), hidden=64,
         layers=4, out_rk45=out)


def rhs_vjp_cases():
    """Autograd of the reference's own ode_residual (the backward of the physics-residual loss
    term, models/hybrid_ode_nn.py:327-330).  ODE parameters are buffers in the reference
    (models/ode_core.py:78-79); to obtain their cotangents they are swapped for leaf tensors,
    exactly what forward_with_params does with sampled values (models/hybrid_ode_nn.py:407-411)."""
    d0 = np.load(os.path.join(HERE, "rhs_mech.npz"))
    rng = np.random.default_rng(11)
    g_out = rng.normal(0, 1, (5, 6)).astype(np.float32)
    for tag, hidden, layers, std in (("mech", 64, 4, 0.0), ("nn64x4", 64, 4, 0.05),
                                     ("nn16x2", 16, 2, 0.05), ("nn32x3", 32, 3, 0.1)):
        m = make_model(hidden, layers, seed=1, out_std=std)
        names = [n for n, _ in m.ode_core.named_buffers()]
        for n in names:
            m.ode_core._buffers[n] = m.ode_core._buffers[n].clone().requires_grad_(True)
        state = torch.tensor(d0["state"], requires_grad=True)
        ext = {"meal": torch.tensor(d0["meal"]), "tVNS": torch.tensor(d0["tvns"]),
               "GD": torch.tensor(d0["gd"])}
        out = m.ode_residual(torch.tensor(d0["t"]), state, ext)
        out.backward(torch.tensor(g_out))
        g_theta = np.array([float(m.ode_core._buffers[n].grad) for n in names], dtype=np.float32)
        g_W = np.concatenate([(p.grad if p.grad is not None else torch.zeros_like(p)).numpy().reshape(-1)
                              for _, p in m.nn_residual.named_parameters()]).astype(np.float32)
        save(f"rhs_vjp_{tag}", t=d0["t"], state=d0["state"], meal=d0["meal"], tvns=d0["tvns"],
             gd=d0["gd"], theta=theta_of(m), W=pack_W(m), hidden=hidden, layers=layers,
             grad_out=g_out, out=out.detach().numpy(), grad_state=state.grad.numpy(),
             grad_theta=g_theta, grad_W=g_W)


def loss_cases():
    """HybridODENN.loss + backward of the reference on a collated batch (train/train_hybrid.py:247-252):
    loss value and the gradient of every network parameter, with the torch seed that fixes the
    physics-index draw (models/hybrid_ode_nn.py:301)."""
    y0, obs, t, meal, tv, ds = windows_4gi()
    B = 4
    # physiological-unit cohort instead of the z-scored windows: keeps every solve well-conditioned
    rng = np.random.default_rng(21)
    y0 = np.stack([7.0 * rng.normal(1, 0.1, B), 50.0 * rng.normal(1, 0.15, B),
                   25.0 * rng.normal(1, 0.15, B), 10.0 * rng.normal(1, 0.15, B),
                   np.zeros(B), np.ones(B)], axis=1).astype(np.float32)
    T = 13
    tt = np.tile(np.linspace(0, 1, T).astype(np.float32), (B, 1))
    ml = np.zeros((B, T), dtype=np.float32)
    ml[:, 3] = rng.uniform(0.5, 1.5, B)
    tvn = np.zeros((B, T), dtype=np.float32)
    tvn[::2, 5:9] = 1.0
    obs = (y0[:, None, :] * (1 + 0.1 * rng.normal(0, 1, (B, T, 6)))).astype(np.float32)
    for tag, hidden, layers in (("nn64x4", 64, 4), ("nn16x2", 16, 2)):
        m = make_model(hidden, layers, seed=7, out_std=0.02)
        batch = {"initial_state": torch.tensor(y0), "observations": torch.tensor(obs),
                 "time_points": torch.tensor(tt),
                 "external_inputs": {"meal": torch.tensor(ml), "tVNS": torch.tensor(tvn)}}
        torch.manual_seed(123)
        loss = m.loss(batch, lambda1=1.0, lambda2=0.5)
        loss.backward()
        g_W = np.concatenate([p.grad.numpy().reshape(-1) for _, p in m.nn_residual.named_parameters()])
        save(f"loss_{tag}", y0=y0, obs=obs, t=tt, meal=ml, tvns=tvn, theta=theta_of(m), W=pack_W(m),
             hidden=hidden, layers=layers, seed=123, lambda1=1.0, lambda2=0.5,
             loss=np.float64(loss.item()), grad_W=g_W.astype(np.float32))


def vi_cases():
    """The reference's variational family and VI driver on a tiny problem: VariationalParameters.sample under a
    fixed seed, kl_divergence, get_flattened_params (models/bayes.py:103-175), VariationalInference.elbo and
    posterior_predictive (inference/vi.py:60-118,274-312).  The reference integrates with its default solver
    ('dopri5' -> SciPy DOP853, rtol 1e-6): the mirrors are compared at 1e-4."""
    from inference.vi import VariationalInference
    prior = {"ode_a_GI": {"mean": 0.0104, "std": 0.002}, "ode_k_I": {"mean": 0.025, "std": 0.005},
             "ode_rho": {"mean": 0.003, "std": 0.001}, "ode_E_max": {"mean": 0.1, "std": 0.02},
             "ode_EC_50": {"mean": 50.0, "std": 5.0}, "ode_V_max": {"mean": 9.0, "std": 2.0},
             "ode_K_m": {"mean": 7.0, "std": 1.5}, "ode_k_L": {"mean": 0.02, "std": 0.005}}   # configs/4gi_vi.yaml:26-33 (+ EC_50)
    torch.manual_seed(3)
    m = HybridODENN(nn_hidden=16, nn_layers=2, use_variational=True, prior_params=prior, device=CPU)
    vp = m.variational_params
    names = list(vp.param_shapes.keys())
    with torch.no_grad():   # a posterior whose network does something and whose widths differ per tensor
        for n in names:
            if n.startswith("nn_"):
                vp.means[n].normal_(0.0, 0.08)
                vp.log_stds[n].fill_(float(np.log(0.01)))
            vp.log_stds[n].add_(0.2 * torch.randn_like(vp.log_stds[n]))
    flat = lambda d: np.concatenate([d[n].detach().numpy().reshape(-1).astype(np.float32) for n in names])
    sizes = np.array([int(np.prod(vp.param_shapes[n])) if len(vp.param_shapes[n]) else 1 for n in names])
    torch.manual_seed(5)
    smp = vp.sample(2)
    kl = float(vp.kl_divergence())
    mu_f, ls_f = vp.get_flattened_params()
    rng = np.random.default_rng(31)
    B, T = 2, 7
    y0 = np.stack([7.0 * rng.normal(1, 0.1, B), 50.0 * rng.normal(1, 0.15, B), 25.0 * rng.normal(1, 0.15, B),
                   10.0 * rng.normal(1, 0.15, B), np.zeros(B), np.ones(B)], axis=1).astype(np.float32)
    t = np.linspace(0, 0.5, T).astype(np.float32)
    meal = np.zeros((B, T), dtype=np.float32)
    meal[:, 2] = 1.0
    tv = np.zeros((B, T), dtype=np.float32)
    tv[0, 3:] = 1.0
    obs = (y0[:, None, :] * (1 + 0.05 * rng.normal(0, 1, (B, T, 6)))).astype(np.float32)
    batch = {"initial_state": torch.tensor(y0), "observations": torch.tensor(obs), "time_points": torch.tensor(t),
             "external_inputs": {"meal": torch.tensor(meal), "tVNS": torch.tensor(tv)}}
    vi = VariationalInference(m, device=CPU)
    torch.manual_seed(7)
    elbo, comp = vi.elbo(batch, n_samples=2, noise_sigma=0.7)
    torch.manual_seed(9)
    mean, std = vi.posterior_predictive(batch["initial_state"], batch["time_points"], batch["external_inputs"], n_samples=3)
    save("vi_bayes", names=np.array(names), sizes=sizes, prior_names=np.array(sorted(prior)),
         prior_mean=np.array([prior[k]["mean"] for k in sorted(prior)]), prior_std=np.array([prior[k]["std"] for k in sorted(prior)]),
         means=flat(vp.means), log_stds=flat(vp.log_stds), sample_seed=5, sample0=flat(smp[0]), sample1=flat(smp[1]),
         kl=np.float64(kl), flat_mu=mu_f.detach().numpy(), flat_log_std=ls_f.detach().numpy(),
         y0=y0, t=t, meal=meal, tvns=tv, obs=obs, hidden=16, layers=2,
         elbo_seed=7, noise_sigma=0.7, elbo=np.float64(float(elbo)), elbo_kl=np.float64(float(comp["kl"])),
         elbo_ll=np.float64(float(comp["log_likelihood"])), pp_seed=9, pp_mean=mean.numpy(), pp_std=std.numpy())


def data_eval_cases():
    """GlucoseDataset (train/train_hybrid.py:43-155) on a small synthetic CSV, and the metric functions of
    eval/evaluate.py:26-181 on random predictions: the reference's own outputs for the device mirrors
    (hode_window_dataset / hode_eval_metrics)."""
    import tempfile
    import pandas as pd
    from eval.evaluate import compute_calibration_error, compute_mae, compute_rmse
    rng = np.random.default_rng(41)
    N, n_t = 3, 100
    rows = []
    st = np.empty((N, n_t, 4), dtype=np.float32)
    meal = np.zeros((N, n_t), dtype=np.float32)
    tv = np.zeros((N, n_t), dtype=np.float32)
    for sidx in range(N):
        base = np.array([7.0, 50.0, 25.0, 10.0]) * (1 + 0.1 * rng.normal(0, 1, 4))
        st[sidx] = (base[None, :] * (1 + 0.2 * rng.normal(0, 1, (n_t, 4)))).astype(np.float32)
        meal[sidx, rng.integers(1, n_t - 1, 4)] = 1.0
        tv[sidx, 40:70] = float(sidx % 2)
        for j in range(n_t):
            rows.append({"subject_id": sidx, "time_minutes": 5.0 * j, "glucose_mmol_L": float(st[sidx, j, 0]),
                         "insulin_pmol_L": float(st[sidx, j, 1]), "glucagon_pmol_L": float(st[sidx, j, 2]),
                         "glp1_pmol_L": float(st[sidx, j, 3]), "meal_indicator": float(meal[sidx, j]), "tvns": float(tv[sidx, j])})
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "cohort.csv")
        pd.DataFrame(rows).to_csv(path, index=False)
        out = {}
        for tag, L_, stride, norm in (("a", 61, 30, True), ("b", 25, 7, True), ("c", 61, 30, False)):
            ds = GlucoseDataset(path, L_, stride, norm)
            items = [ds[i] for i in range(len(ds))]
            out[f"obs_{tag}"] = np.stack([it["observations"].numpy() for it in items])
            out[f"init_{tag}"] = np.stack([it["initial_state"].numpy() for it in items])
            out[f"time_{tag}"] = np.stack([it["time_points"].numpy() for it in items])
            out[f"meal_{tag}"] = np.stack([it["external_inputs"]["meal"].numpy() for it in items])
            out[f"tvns_{tag}"] = np.stack([it["external_inputs"]["tVNS"].numpy() for it in items])
            out[f"mean_{tag}"], out[f"std_{tag}"] = np.asarray(ds.state_mean, np.float64), np.asarray(ds.state_std, np.float64)
            out[f"cfg_{tag}"] = np.array([L_, stride, int(norm)])
    states6 = np.concatenate([st, np.zeros((N, n_t, 1), np.float32), np.ones((N, n_t, 1), np.float32)], axis=2)
    # metrics
    pred = (rng.normal(0, 1, (40, 61, 6)) * np.array([2, 50, 20, 10, 0.1, 0.3]) + np.array([7, 50, 25, 10, 0, 1])).astype(np.float32)
    targ = (pred + rng.normal(0, 1, pred.shape) * np.array([0.5, 10, 5, 2, 0.05, 0.1])).astype(np.float32)
    unc = (np.abs(rng.normal(1, 0.3, pred.shape)) * np.array([0.5, 10, 5, 2, 0.05, 0.1])).astype(np.float32)
    tp, tt_, tu = torch.tensor(pred), torch.tensor(targ), torch.tensor(unc)
    np.random.seed(77)
    cal = compute_calibration_error(tp, tu, tt_)
    save("data_eval", states=states6, meal=meal, tvns=tv, time_hours=(np.arange(n_t) * 5.0 / 60.0).astype(np.float32),
         pred=pred, target=targ, unc=unc, rmse=np.float64(compute_rmse(tp, tt_)), mae=np.float64(compute_mae(tp, tt_)),
         rmse_state=compute_rmse(tp, tt_, per_state=True), mae_state=compute_mae(tp, tt_, per_state=True),
         cal_seed=77, cal_keys=np.array(sorted(cal)), cal_vals=np.array([cal[k] for k in sorted(cal)], dtype=np.float64),
         target_std=tt_.std(dim=(0, 1)).numpy(), **out)


if __name__ == "__main__":
    only = sys.argv[1:]
    for fn in (rhs_cases, rollout_fig2, rollout_4gi, rollout_physio, rhs_vjp_cases, loss_cases, vi_cases, data_eval_cases):
        if not only or fn.__name__ in only:
            fn()
