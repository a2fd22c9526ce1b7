"""Base controller (mirror of pddp/controllers/base.py)."""
import torch

from ..utils.encoding import StateEncoding


class Controller(torch.nn.Module):
    """ref: pddp/controllers/base.py:21-71"""

    def fit(self, U, encoding=StateEncoding.DEFAULT, **kwargs):
        raise NotImplementedError

    def forward(self, z, i, encoding=StateEncoding.DEFAULT, **kwargs):
        raise NotImplementedError
