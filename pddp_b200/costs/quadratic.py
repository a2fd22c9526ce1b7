"""Quadratic cost (mirror of pddp/costs/quadratic.py).

`QRCost` holds Q, R, Q_term, x_goal, u_goal like the reference; the kernels evaluate the expected
cost of the (angle-augmented) Gaussian state and its gradient / Hessian from these constants."""
import torch

from .. import _lib
from ..solver import QRCostConstants, cached_solver
from ..utils.encoding import StateEncoding
from .base import Cost


class QRCost(Cost):
    """E[L] = tr(Q Sigma) + (mu-x_goal)^T Q (mu-x_goal) + (u-u_goal)^T R (u-u_goal)
    ref: pddp/costs/quadratic.py:24-99.  `state_size` / `angular_indices` say how the encoded state
    is augmented before Q is applied (the reference's example costs do this in their forward)."""
    state_size = None
    angular_indices = ()

    def __init__(self, Q, R, Q_term=None, x_goal=0.0, u_goal=0.0, state_size=None, angular_indices=None):
        super().__init__()
        Q_term = Q if Q_term is None else Q_term
        self.Q = torch.nn.Parameter(torch.as_tensor(Q).clone(), requires_grad=False)
        self.R = torch.nn.Parameter(torch.as_tensor(R).clone(), requires_grad=False)
        self.Q_term = torch.nn.Parameter(torch.as_tensor(Q_term).clone(), requires_grad=False)
        DA, nu = self.Q.shape[0], self.R.shape[0]
        self.x_goal = torch.nn.Parameter(torch.as_tensor(x_goal, dtype=self.Q.dtype).expand(DA).clone(),
                                         requires_grad=False)
        self.u_goal = torch.nn.Parameter(torch.as_tensor(u_goal, dtype=self.Q.dtype).expand(nu).clone(),
                                         requires_grad=False)
        if state_size is not None:
            self.state_size = state_size
        if angular_indices is not None:
            self.angular_indices = tuple(int(i) for i in angular_indices)

    def geometry(self):
        from ..models.base import geometry_of
        if self.state_size is None:
            raise NotImplementedError("pddp_b200: QRCost needs state_size / angular_indices (kernel geometry)")
        return geometry_of(self.state_size, self.angular_indices)

    def constants(self):
        return QRCostConstants(self.Q.data, self.R.data, self.Q_term.data, self.x_goal.data, self.u_goal.data)

    def forward(self, z, u, i, terminal=False, encoding=StateEncoding.DEFAULT, **kwargs):
        """Expected cost of z:[..., nz] (u:[..., nu] unless terminal), evaluated on the GPU."""
        zz = z.reshape(-1, z.shape[-1])
        _lib.require_cuda(zz, "z")
        B = zz.shape[0]
        s = cached_solver(None, self, encoding, B, 1, zz.dtype, zz.device, layout=_lib.PROBLEM_MAJOR)
        uu = torch.zeros(B, 1, self.R.shape[0], dtype=zz.dtype, device=zz.device) if u is None else u.reshape(
            -1, u.shape[-1]).expand(B, -1).reshape(B, 1, -1)
        Z = zz.detach().unsqueeze(1).expand(B, 2, -1)
        s.store("Z", Z)
        s.store("U", uu.detach())
        s.cost_only()
        L = s.view("L")[:, 1 if terminal else 0, 0].clone()
        return L[0] if z.dim() == 1 else L.reshape(z.shape[:-1])
