"""Common utilities, namespaced like pddp.utils (angular, constraint, encoding, evaluation, gaussian_variable)."""
from . import angular, constraint, encoding, gaussian_variable  # noqa: F401
from . import evaluation  # noqa: F401  (after encoding: it builds solvers)

__all__ = ["angular", "constraint", "encoding", "evaluation", "gaussian_variable"]
