"""B200-native batched hybrid ODE-NN integrator: a drop-in for the rollout / gradient path
of OliverDOU776/Hybrid-ODE-for-GLP-1-and-Glucose.

    from hybrid_ode_for_glp_1_and_glucose_b200 import HybridODENN

The compute lives in libhode.so (csrc/, C ABI in include/hode.h); this package is the
host-side mirror of the reference's Python interface for that path.
"""
from . import _lib, distributed, ops, sensitivity, training
from ._lib import HodeError
from .bayes import VariationalParameters, bayes_loss, compute_posterior_predictive
from .hybrid_ode_nn import HybridODENN
from .nn_residual import NNResidual
from .ode_core import ODECore
from .vi import VariationalInference

__all__ = ["HybridODENN", "ODECore", "NNResidual", "VariationalParameters", "bayes_loss",
           "compute_posterior_predictive", "VariationalInference", "HodeError", "ops", "distributed", "sensitivity", "training"]
