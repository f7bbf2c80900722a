"""Host-side logic that needs no GPU: the drop-in surface (constructor, attribute names,
state_dict keys, parameter packing order), input normalisation, and the loud failure when
no CUDA device is available."""
import numpy as np
import pytest
import torch

import hybrid_ode_for_glp_1_and_glucose_b200 as hode
from hybrid_ode_for_glp_1_and_glucose_b200 import _lib, ops
from helpers import golden

CPU = torch.device("cpu")


def test_state_dict_keys_match_the_reference():
    # 63 keys with VI
    m = hode.HybridODENN(use_variational=True, device=CPU)
    keys = list(m.state_dict())
    assert len(keys) == 63
    ode = ["a_GI", "k_I", "rho", "G_b", "I_b", "E_max", "EC_50", "Glu_b", "V_max", "K_m", "k_L",
           "k_GE0", "IGD_50", "g", "p_7", "p_8", "p_9"]
    assert keys[:17] == [f"ode_core.{n}" for n in ode]
    assert keys[17:27] == [f"nn_residual.network.{i}.{p}" for i in (0, 2, 4, 6, 8)
                           for p in ("weight", "bias")]
    assert "variational_params.means.ode_a_GI" in keys
    assert "variational_params.log_stds.nn_network_8_bias" in keys
    assert m.n_states == 6 and m.state_names[3] == "GLP1" and m.use_variational
    assert hode.HybridODENN(device=CPU).variational_params is None


def test_fresh_model_reproduces_reference_defaults():
    d = golden("rhs_mech")
    m = hode.HybridODENN(device=CPU)
    assert np.array_equal(m.ode_core.theta().numpy(), d["theta"])
    m2 = hode.HybridODENN(ode_params={"k_L": 0.05, "V_max": 7.0}, device=CPU)
    th = m2.ode_core.theta().numpy()
    assert th[10] == np.float32(0.05) and th[8] == np.float32(7.0) and th[0] == np.float32(0.0104)
    # zero-initialised head, Xavier body with gain 0.1 (reference nn_residual.py:83-98)
    lin = m.nn_residual.linears()
    assert len(lin) == 5 and lin[0].weight.shape == (64, 9) and lin[-1].weight.shape == (6, 64)
    assert float(lin[-1].weight.detach().abs().max()) == 0 and float(lin[-1].bias.detach().abs().max()) == 0
    assert 0 < float(lin[1].weight.std()) < 0.03
    assert m.nn_residual.is_identically_zero()
    assert m.packed_parameters()[1] is None


def test_packing_order_is_named_parameters_order():
    m = hode.HybridODENN(nn_hidden=16, nn_layers=2, device=CPU)
    with torch.no_grad():
        for i, (_, p) in enumerate(m.nn_residual.named_parameters()):
            p.fill_(float(i + 1))
    W = m.nn_residual.packed().detach().numpy()
    sizes = [16 * 9, 16, 16 * 16, 16, 6 * 16, 6]
    assert W.shape == (sum(sizes),) == (_lib.mlp_param_count(16, 2),)
    off = 0
    for i, n in enumerate(sizes):
        assert (W[off:off + n] == i + 1).all()
        off += n
    # overrides by the reference's 'nn_<clean name>' / 'ode_<name>' keys
    th, W2 = m.packed_parameters({"ode_rho": torch.tensor(0.5),
                                  "nn_network_2_bias": torch.zeros(16)})
    assert float(th[2]) == 0.5 and (W2.detach()[16 * 9 + 16 + 256:16 * 9 + 16 + 256 + 16] == 0).all()
    # packed() is differentiable w.r.t. the module parameters
    m.nn_residual.packed().sum().backward()
    assert all(p.grad is not None for p in m.nn_residual.parameters())


def test_prepare_normalises_inputs_like_the_reference():
    y0 = torch.zeros(4, 6)
    t = torch.linspace(0, 1, 5)
    th = torch.zeros(17)
    cfg, bufs = ops.prepare(y0, t, {"meal": torch.ones(4, 5), "tVNS": torch.ones(4)}, th, None,
                            64, 4, CPU)
    assert (cfg.n_traj, cfg.n_obs, cfg.t_per_traj, cfg.n_samples) == (4, 5, 0, 1)
    assert list(cfg.in_mode) == [_lib.IN_SERIES, _lib.IN_CONST, _lib.IN_ABSENT]
    assert cfg.mlp == _lib.MLP_NONE
    cfg, _ = ops.prepare(y0, t.repeat(4, 1), None, th.repeat(3, 1), torch.zeros(3, 13510), 64, 4, CPU)
    assert cfg.t_per_traj == 1 and cfg.n_samples == 3 and cfg.mlp == _lib.MLP_FP32
    # a [1,T] grid for B != 1 is squeezed (reference models/hybrid_ode_nn.py:192-196)
    cfg, bufs = ops.prepare(y0, t.unsqueeze(0), None, th, None, 64, 4, CPU)
    assert cfg.t_per_traj == 0 and bufs["t_obs"].shape == (5,)
    with pytest.raises(ValueError):
        ops.prepare(y0, t, {"meal": torch.ones(3, 5)}, th, None, 64, 4, CPU)
    with pytest.raises(ValueError):
        ops.prepare(y0, t, None, th, torch.zeros(100), 64, 4, CPU)
    with pytest.raises(ValueError):
        ops.prepare(torch.zeros(4, 5), t, None, th, None, 64, 4, CPU)


def test_no_cpu_fallback():
    m = hode.HybridODENN(device=CPU)
    with pytest.raises(hode.HodeError):
        m(torch.zeros(2, 6), torch.linspace(0, 1, 5))
    with pytest.raises(hode.HodeError):
        m.ode_residual(torch.tensor(0.0), torch.zeros(2, 6))
    with pytest.raises(hode.HodeError):
        ops.rollout(torch.zeros(2, 6), torch.linspace(0, 1, 5), None, torch.zeros(17), None)
    with pytest.raises(NotImplementedError):
        hode.NNResidual(activation="tanh")


def test_variational_parameters_follow_the_reference_formulas():
    torch.manual_seed(0)
    shapes = {"ode_a": torch.Size([]), "nn_w": torch.Size([3, 2])}
    vp = hode.VariationalParameters(shapes, {"ode_a": 0.5}, {"ode_a": 2.0})
    assert float(vp.means["ode_a"]) == 0.5
    assert abs(float(vp.log_stds["ode_a"]) - np.log(0.2)) < 1e-7      # 10 % of the prior std
    assert abs(float(vp.log_stds["nn_w"][0, 0]) - np.log(0.1)) < 1e-7
    # KL of N(mu, s) from N(mu_p, s_p), reference models/bayes.py:146-153
    kl = float(vp.kl_divergence())
    k_a = np.log(2.0) - np.log(0.2) + (0.2 ** 2) / (2 * 4.0) - 0.5
    k_w = 6 * (0.0 - np.log(0.1) + (0.1 ** 2) / 2 - 0.5)
    assert abs(kl - (k_a + k_w)) < 1e-5
    # sampling consumes randn_like per tensor in insertion order (reference :117-123)
    torch.manual_seed(1)
    s = vp.sample(1)[0]
    torch.manual_seed(1)
    e_a = torch.randn(())
    e_w = torch.randn(3, 2)
    assert torch.allclose(s["ode_a"], 0.5 + e_a * 0.2) and torch.allclose(s["nn_w"], e_w * 0.1)
    mu, ls = vp.get_flattened_params()
    assert mu.shape == (7,) and ls.shape == (7,)
    with pytest.raises(TypeError):
        hode.bayes_loss(None, torch.zeros(1))


def _vi_model_from_fixture(d, device):
    """The fixture's posterior loaded into our mirror (same constructor arguments as the reference run)."""
    prior = {str(k): {"mean": float(m), "std": float(sd)} for k, m, sd in zip(d["prior_names"], d["prior_mean"], d["prior_std"])}
    m = hode.HybridODENN(nn_hidden=int(d["hidden"]), nn_layers=int(d["layers"]), use_variational=True, prior_params=prior,
                         device=device)
    vp = m.variational_params
    names = [str(n) for n in d["names"]]
    assert list(vp.param_shapes.keys()) == names, "same parameter names in the same (insertion) order as the reference"
    off = 0
    with torch.no_grad():
        for n, sz in zip(names, d["sizes"]):
            vp.means[n].copy_(torch.from_numpy(d["means"][off: off + sz]).reshape(vp.means[n].shape))
            vp.log_stds[n].copy_(torch.from_numpy(d["log_stds"][off: off + sz]).reshape(vp.log_stds[n].shape))
            off += int(sz)
    return m, names


def test_variational_parameters_match_reference_outputs():
    """VariationalParameters.sample under the reference's seed, kl_divergence and get_flattened_params against
    outputs of the reference itself (tests/golden/make_golden.py::vi_cases -> vi_bayes.npz; models/bayes.py:103-175)."""
    from helpers import golden
    d = golden("vi_bayes")
    m, names = _vi_model_from_fixture(d, torch.device("cpu"))
    vp = m.variational_params
    torch.manual_seed(int(d["sample_seed"]))
    smp = vp.sample(2)
    for k, key in enumerate(("sample0", "sample1")):
        got = np.concatenate([smp[k][n].detach().numpy().reshape(-1) for n in names])
        np.testing.assert_array_equal(got.astype(np.float32), d[key])        # same draws, same arithmetic: bit-identical
    assert abs(float(vp.kl_divergence()) - float(d["kl"])) <= 1e-6 * abs(float(d["kl"]))
    mu, ls = vp.get_flattened_params()
    np.testing.assert_array_equal(mu.detach().numpy(), d["flat_mu"])
    np.testing.assert_array_equal(ls.detach().numpy(), d["flat_log_std"])


def test_generator_host_helpers():
    """Meal spreading and the state-column mapping of the cohort generator (no GPU): the reference puts a meal of
    size s at time m into the sampling interval [t_i, t_i+1) that contains m, as the rate s / (t_i+1 - t_i)
    (data/generate4GI.py:193-197); GlucoseDataset orders the states glucose, insulin, glucagon, GLP-1, ge, ffa
    (train/train_hybrid.py:70-83)."""
    import numpy as np
    from hybrid_ode_for_glp_1_and_glucose_b200.synthetic import fourgi_baselines, fourgi_states, meal_rate_from_events
    r = meal_rate_from_events([0.5, 2.5], [75, 50], 61, 5 / 60, 3)
    assert r.shape == (3, 60)
    nz = np.nonzero(r[0])[0].tolist()
    assert nz == [6, 30]
    assert abs(r[0, 6] - 75 / (5 / 60)) < 1e-3 and abs(r[0, 30] - 50 / (5 / 60)) < 1e-3
    assert np.array_equal(r[0], r[2])
    assert float(meal_rate_from_events([5.0], [10], 61, 5 / 60)[0].sum()) == 0.0      # a meal at t_end falls in no interval
    conc = np.arange(2 * 3 * 5, dtype=np.float32).reshape(2, 3, 5)
    st = fourgi_states(conc)
    assert st.shape == (2, 3, 6)
    assert np.array_equal(st[..., 0], conc[..., 0]) and np.array_equal(st[..., 1], conc[..., 1])
    assert np.array_equal(st[..., 2], conc[..., 3]) and np.array_equal(st[..., 3], conc[..., 2])
    assert float(np.abs(st[..., 4]).max()) == 0.0 and float(st[..., 5].min()) == 1.0
    b = fourgi_baselines(1000, seed=1)
    assert b.shape == (1000, 5) and abs(float(b[:, 0].mean()) - 7.0) < 0.1 and abs(float(b[:, 1].std()) / 50.0 - 0.15) < 0.02
