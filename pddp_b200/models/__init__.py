"""Dynamics models (namespaced like pddp.models)."""
from .base import DynamicsModel, KnownDynamicsModel, geometry_of  # noqa: F401
from . import bnn  # noqa: F401

__all__ = ["DynamicsModel", "bnn"]
