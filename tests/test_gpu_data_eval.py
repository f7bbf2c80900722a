"""GPU tests of the data formats either side of the path against outputs of the reference itself
(tests/golden/data_eval.npz, made by tests/golden/make_golden.py::data_eval_cases): GlucoseDataset's windowing and
z-scoring (hode_window_dataset) and the evaluation metrics (hode_eval_metrics)."""
import numpy as np
import pytest
import torch

from helpers import golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(built_lib):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_device_dataset_matches_reference_glucose_dataset(dev, tag):
    from hybrid_ode_for_glp_1_and_glucose_b200.data import DeviceGlucoseDataset
    d = golden("data_eval")
    L, stride, norm = [int(x) for x in d[f"cfg_{tag}"]]
    inputs = np.stack([d["meal"], d["tvns"]], axis=2)
    ds = DeviceGlucoseDataset(torch.from_numpy(d["states"]), torch.from_numpy(inputs), torch.from_numpy(d["time_hours"]),
                              sequence_length=L, stride=stride, normalize=bool(norm), device=dev)
    assert len(ds) == d[f"obs_{tag}"].shape[0]
    np.testing.assert_allclose(ds.state_mean, d[f"mean_{tag}"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(ds.state_std, d[f"std_{tag}"], rtol=1e-10, atol=1e-12)
    # normalised values: float64 arithmetic rounded once to float32 on both sides; (x - mean) / 1e-6 for the constant
    # placeholder columns amplifies the last-bit difference of the statistics, hence the absolute floor
    np.testing.assert_allclose(ds.observations.cpu().numpy(), d[f"obs_{tag}"], rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(ds.initial_state.cpu().numpy(), d[f"init_{tag}"], rtol=2e-6, atol=2e-6)
    assert np.array_equal(ds.time_points.cpu().numpy(), d[f"time_{tag}"])
    item = ds[len(ds) - 1]
    assert np.array_equal(item["external_inputs"]["meal"].cpu().numpy(), d[f"meal_{tag}"][-1])
    assert np.array_equal(item["external_inputs"]["tVNS"].cpu().numpy(), d[f"tvns_{tag}"][-1])
    b = ds.batch([0, len(ds) - 1])
    assert b["observations"].shape == (2, L, 6) and torch.equal(b["initial_state"], b["observations"][:, 0])


def test_metrics_match_reference_evaluate(dev):
    from hybrid_ode_for_glp_1_and_glucose_b200 import evaluate
    d = golden("data_eval")
    p, t, u = (torch.from_numpy(d[k]).to(dev) for k in ("pred", "target", "unc"))
    assert abs(evaluate.compute_rmse(p, t) - float(d["rmse"])) <= 1e-6 * float(d["rmse"])
    assert abs(evaluate.compute_mae(p, t) - float(d["mae"])) <= 1e-6 * float(d["mae"])
    np.testing.assert_allclose(evaluate.compute_rmse(p, t, per_state=True), d["rmse_state"], rtol=1e-6)
    np.testing.assert_allclose(evaluate.compute_mae(p, t, per_state=True), d["mae_state"], rtol=1e-6)
    np.random.seed(int(d["cal_seed"]))
    cal = evaluate.compute_calibration_error(p, u, t)
    for k, v in zip(d["cal_keys"], d["cal_vals"]):
        assert abs(cal[str(k)] - float(v)) <= 1e-5 * max(abs(float(v)), 1e-3), k


def test_evaluate_model_point_estimate(dev):
    """evaluate_model on a loader of two batches: the metric dictionary of the reference (eval/evaluate.py:184-288) with
    the predictions of the drop-in class; checked against numpy on the same predictions."""
    from hybrid_ode_for_glp_1_and_glucose_b200 import HybridODENN, evaluate
    from helpers import cohort
    y0, t, ins = cohort(24, 13, seed=71, horizon=1.0)
    rng = np.random.default_rng(72)
    obs = (y0[:, None, :] * (1 + 0.1 * rng.normal(0, 1, (24, 13, 6)))).astype(np.float32)
    tt = lambda a: torch.from_numpy(a)
    loader = [{"initial_state": tt(y0[i: i + 12]), "observations": tt(obs[i: i + 12]), "time_points": tt(t),
               "external_inputs": {k: tt(v[i: i + 12]) for k, v in ins.items()}} for i in (0, 12)]
    m = HybridODENN(device=dev)
    met = evaluate.evaluate_model(m, loader, dev)
    with torch.no_grad():
        pred = torch.cat([m(b["initial_state"], b["time_points"], b["external_inputs"]) for b in loader]).cpu().numpy()
    assert abs(met["rmse"] - np.sqrt(np.mean((pred - obs) ** 2))) < 1e-5 * met["rmse"]
    assert abs(met["mae_glucose"] - np.mean(np.abs(pred[..., 0] - obs[..., 0]))) < 1e-5 * met["mae_glucose"]
    tstd = torch.from_numpy(obs).std(dim=(0, 1)).numpy()
    assert abs(met["nrmse"] - met["rmse"] / np.mean(tstd)) < 1e-5 * met["nrmse"]
    assert abs(met["nrmse_insulin"] - np.sqrt(np.mean((pred[..., 1] - obs[..., 1]) ** 2)) / tstd[1]) < 1e-5 * met["nrmse_insulin"]
    assert "ece" not in met
