"""Golden fixtures for SciPy DOP853 — what the reference's solver='dopri5' / 'dop853' really runs
(models/hybrid_ode_nn.py:174-181) — produced by RUNNING THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_dop853.py
Adds, next to the `out_dopri5` arrays make_golden.py already stores, the accepted-step log (t_n, h_n, nfev) of
SciPy's DOP853 inside the reference's forward() for two scenarios, so that the oracle's controller, error norm
and dense output are pinned step by step and not only through the final trajectories.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (sets up the reference import path and the arviz stub)


def main():
    # Fig-2 scenario (plots/plot_all.py:164-187), mechanistic (zero-initialised output layer)
    y0 = np.array([[5.0, 60.0, 80.0, 0.0, 0.0, 1.0]], dtype=np.float32)
    t = np.linspace(0, 5, 61).astype(np.float32)
    meal = np.zeros((1, 61), dtype=np.float32)
    meal[0, 6] = 75.0
    tv = np.zeros((1, 61), dtype=np.float32)
    m = mg.make_model()
    with mg.StepLog() as sl:
        out = mg.run_forward(m, y0, t, {"meal": meal, "tVNS": tv}, "dop853")
    mg.save("dop853_fig2", y0=y0, t=t, meal=meal, tvns=tv, theta=mg.theta_of(m), W=mg.pack_W(m), hidden=64, layers=4,
            out_dop853=out, steps_dop853=np.array(sl.log, dtype=np.float64))
    # smooth inputs, hybrid network, the 61-point grid: dense output exercised at every observation time
    rng = np.random.default_rng(11)
    B = 4
    y0 = np.stack([7.0 * rng.normal(1, 0.1, B), 50.0 * rng.normal(1, 0.15, B), 25.0 * rng.normal(1, 0.15, B),
                   10.0 * rng.normal(1, 0.15, B), np.zeros(B), np.ones(B)], axis=1).astype(np.float32)
    mc = rng.uniform(0, 2, B).astype(np.float32)
    tc = (rng.uniform(0, 1, B) > 0.5).astype(np.float32)
    m = mg.make_model(seed=5, out_std=0.02)
    with mg.StepLog() as sl:
        out0 = mg.run_forward(m, y0[:1], t, {"meal": mc[:1], "tVNS": tc[:1]}, "dop853")
    steps = np.array(sl.log, dtype=np.float64)
    out = mg.run_forward(m, y0, t, {"meal": mc, "tVNS": tc}, "dopri5")   # 'dopri5' -> DOP853 in the reference
    assert np.array_equal(out[:1], out0)
    mg.save("dop853_smooth", y0=y0, t=t, meal=mc, tvns=tc, theta=mg.theta_of(m), W=mg.pack_W(m), hidden=64, layers=4,
            out_dop853=out, steps_dop853_traj0=steps)


if __name__ == "__main__":
    main()
