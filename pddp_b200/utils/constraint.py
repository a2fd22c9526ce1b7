"""Control constraints.  `clamp` and the box QP live INSIDE the kernels on the hot path (pddp_backward, the
linearise / rollout kernels); what is here are the host-side names of pddp/utils/constraint.py."""
import torch

from .. import _lib  # noqa: F401


class BOXQP_RESULTS(object):
    """Result codes of the projected-Newton box QP.  ref: pddp/utils/constraint.py:23-32"""
    HESSIAN_NOT_PD = -1
    NO_DESCENT_DIRECTION = 0
    MAX_ITER_EXCEEDED = 1
    MAX_LS_ITER_EXCEEDED = 2
    NO_BOUNDS = 3
    IMPROVEMENT_TOO_SMALL = 4
    GRADIENT_TOO_SMALL = 5
    ALL_DIMENSIONS_CLAMPED = 6


def clamp(u, min_bounds, max_bounds):
    """ref: pddp/utils/constraint.py:146-147"""
    return torch.min(torch.max(u, min_bounds), max_bounds)


def constrain(u, min_bounds, max_bounds):
    """tanh squash of u into [min_bounds, max_bounds].  ref: pddp/utils/constraint.py:35-48"""
    diff = (max_bounds - min_bounds) / 2.0
    mean = (max_bounds + min_bounds) / 2.0
    return diff * u.tanh() + mean


def boxqp(x0, Q, c, lower, upper, **kwargs):
    """ref: pddp/utils/constraint.py:150-266.  The box QP is solved inside `pddp_backward` (one thread / warp per
    problem, csrc/backward.cu `boxqp1`, csrc/backward_nu.cu), where the reference calls it (ilqr.py:649-651); there
    is no stand-alone entry point for it in the C ABI."""
    raise NotImplementedError("pddp_b200: boxqp runs inside pddp_backward (ilqr.backward with u_min / u_max); "
                              "it has no stand-alone entry point")
