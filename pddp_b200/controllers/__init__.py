"""Controllers (namespaced like pddp.controllers)."""
from .base import Controller
from .ilqr import iLQRController, iLQRState
from .pddp import PDDPController, _apply_controller, _concat_datasets  # noqa: F401
from .ilqr import Q, _control_law, _trajectory_cost, backward, forward  # noqa: F401

__all__ = ["Controller", "iLQRController", "PDDPController"]
