"""Cost functions (namespaced like pddp.costs)."""
from .base import Cost
from .quadratic import QRCost

__all__ = ["Cost", "QRCost"]
