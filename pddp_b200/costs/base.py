"""Cost interface (mirror of pddp/costs/base.py)."""
import torch

from ..utils.encoding import StateEncoding


class Cost(torch.nn.Module):
    """ref: pddp/costs/base.py:21-122.  A cost the kernels can evaluate exposes `constants()` (QRCostConstants)
    and `geometry()`; operator-built AggregateCosts (base.py:125-181) are not on the hot path and not built."""

    def forward(self, z, u, i, terminal=False, encoding=StateEncoding.DEFAULT, **kwargs):
        raise NotImplementedError

    def constants(self):
        raise NotImplementedError("pddp_b200: the kernels evaluate QRCost-family costs (Q, R, Q_term, x_goal, u_goal)")

    def geometry(self):
        raise NotImplementedError
