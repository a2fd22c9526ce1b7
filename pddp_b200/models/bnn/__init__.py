"""Bayesian neural network dynamics models (namespaced like pddp.models.bnn)."""
from .losses import gaussian_log_likelihood
from .modules import BDropout, BSequential, CDropout, bayesian_model, bnn_dynamics_model_factory

__all__ = ["BDropout", "BSequential", "CDropout", "bayesian_model", "bnn_dynamics_model_factory",
           "gaussian_log_likelihood"]
