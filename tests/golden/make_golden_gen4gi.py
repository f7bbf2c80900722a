"""Golden fixtures for the cohort generator, produced by RUNNING THE REFERENCE's FourGIModel
(data/generate4GI.py:6-211).  Build container only:  python tests/golden/make_golden_gen4gi.py

Each case stores the baselines, the meal events and what `FourGIModel.simulate` returned
(time grid + 5 concentration series); `generate_dataset`'s baseline perturbation is reproduced by
setting the BSL* attributes directly, so no random state is involved."""
import os
import sys
import types

import numpy as np

REF = os.environ.get("HODE_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(REF, "data"))
for mod in ("matplotlib", "matplotlib.pyplot"):     # data/generate4GI.py:4 imports pyplot for its demo
    sys.modules.setdefault(mod, types.ModuleType(mod))

from generate4GI import FourGIModel  # noqa: E402


def run(patient_type, baselines, meal_times, meal_sizes, duration=5, interval=5):
    outs = []
    for b in baselines:
        m = FourGIModel(patient_type=patient_type)
        m.BSLglc, m.BSLins, m.BSLglp, m.BSLglg, m.BSLgip = [float(x) for x in b]
        t, glc, ins, glp, glg, gip = m.simulate(duration, interval, list(meal_times), list(meal_sizes))
        outs.append(np.stack([glc, ins, glp, glg, gip], axis=-1))
    return np.asarray(t, np.float64), np.asarray(outs, np.float64)


def main():
    rng = np.random.default_rng(2024)
    base = np.array([7.0, 50.0, 10.0, 25.0, 20.0])
    cv = np.array([0.1, 0.15, 0.15, 0.15, 0.15])          # generate_dataset :231-235
    for name, ptype, n, mt, ms, dur in (
            ("gen4gi_t2dm", "T2DM", 6, [0.5, 2.5], [75, 50], 5),      # the dataset's own meal plan (:279-280)
            ("gen4gi_hv", "HV", 4, [1.0, 3.0], [75, 50], 5),          # generate_dataset defaults (:226)
            ("gen4gi_nomeal", "T2DM", 2, [], [], 2),                  # steady state
    ):
        baselines = base * (1 + cv * rng.normal(0, 1, (n, 5)))
        baselines[0] = base
        t, out = run(ptype, baselines, mt, ms, dur)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), patient_type=ptype, baselines=baselines,
                            meal_times=np.asarray(mt, np.float64), meal_sizes=np.asarray(ms, np.float64),
                            t=t, out=out)
        print(name, out.shape, "glucose range", out[..., 0].min(), out[..., 0].max())


if __name__ == "__main__":
    main()
