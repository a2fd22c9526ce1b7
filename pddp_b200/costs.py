"""Cost interface (mirror of pddp/costs/base.py, pddp/costs/quadratic.py) and the example costs.

`QRCost` holds Q, R, Q_term, x_goal, u_goal like the reference; the kernels evaluate the expected
cost of the (angle-augmented) Gaussian state and its gradient / Hessian from these constants."""
import math

import torch

from . import _lib
from .encoding import StateEncoding
from .solver import BatchedSolver, KnownDynamics, QRCostConstants


class Cost(torch.nn.Module):
    """ref: pddp/costs/base.py:21-122"""

    def forward(self, z, u, i, terminal=False, encoding=StateEncoding.DEFAULT, **kwargs):
        raise NotImplementedError

    def constants(self):
        raise NotImplementedError


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
        from .models import geometry_of
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
        geo = self.geometry()
        s = BatchedSolver(KnownDynamics(geo, [0.0] * 8), self.constants(), encoding, B, 1, dtype=zz.dtype,
                          device=zz.device, layout=_lib.PROBLEM_MAJOR)
        uu = torch.zeros(B, 1, self.R.shape[0], dtype=zz.dtype, device=zz.device) if u is None else u.reshape(B, 1, -1)
        Z = zz.unsqueeze(1).expand(B, 2, -1)
        s.store("Z", Z)
        s.store("U", uu)
        s.cost_only()
        L = s.view("L")[:, 1 if terminal else 0, 0].clone()
        return L[0] if z.dim() == 1 else L.reshape(z.shape[:-1])


def _augmented_goal(x, ang):
    from .models import _augment
    return _augment(torch.as_tensor(x, dtype=torch.float32), list(ang))


class PendulumCost(QRCost):
    """ref: pddp/examples/pendulum/cost.py:32-88"""

    def __init__(self, pendulum_length=0.5):
        l = pendulum_length
        Q = torch.zeros(3, 3)
        Q[0, 0] = 1.0
        Q[0, 1] = Q[1, 0] = l
        Q[1, 1] = Q[2, 2] = l ** 2
        super().__init__(Q, 0.1 * torch.eye(1), 100 * torch.eye(3), _augmented_goal([math.pi, 0.0], (0,)),
                         state_size=2, angular_indices=(0,))


class CartpoleCost(QRCost):
    """ref: pddp/examples/cartpole/cost.py:32-87"""

    def __init__(self, pole_length=0.5):
        l = pole_length
        Q = torch.zeros(5, 5)
        Q[0, 0] = 1.0
        Q[0, 3] = Q[3, 0] = l
        Q[3, 3] = Q[4, 4] = l ** 2
        super().__init__(Q, 0.1 * torch.eye(1), torch.eye(5), _augmented_goal([0.0, 0.0, math.pi, 0.0], (2,)),
                         state_size=4, angular_indices=(2,))


class RendezvousCost(QRCost):
    """ref: pddp/examples/rendezvous/cost.py:29-43 (||x_0 - x_1||^2 + velocities, R = 0.1 I, Q_term = Q)"""

    def __init__(self):
        Q = torch.eye(8)
        Q[0, 2] = Q[2, 0] = -1
        Q[1, 3] = Q[3, 1] = -1
        super().__init__(Q, 0.1 * torch.eye(4), state_size=8, angular_indices=())


class DoubleCartpoleCost(QRCost):
    """ref: pddp/examples/double_cartpole/cost.py:32-96"""

    def __init__(self, pole1_length=0.6, pole2_length=0.6):
        C = torch.tensor([[1, -pole1_length, 0, -pole2_length, 0], [0, 0, pole1_length, 0, pole2_length]])
        Q = torch.zeros(8, 8)
        dims = torch.tensor([0, 4, 5, 6, 7])
        Q[dims[:, None], dims[None, :]] = C.t().mm(C)
        super().__init__(Q, 0.1 * torch.eye(1), 100 * torch.eye(8), _augmented_goal(torch.zeros(6), (2, 4)),
                         state_size=6, angular_indices=(2, 4))
