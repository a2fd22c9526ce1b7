"""Dynamics-model interface (mirror of pddp/models/base.py).

A model object is a plain torch.nn.Module holding constants / weights exactly like the reference's;
what the CUDA kernels consume is its `descriptor()` (flat constants for the closed-form models,
weights + persistent dropout masks + eps_in[0] for the BNN).  Nothing here computes dynamics on the
CPU: `forward()` runs one step of the device rollout kernel.
"""
import torch

from .. import _lib
from ..solver import KnownDynamics, cached_solver
from ..utils.encoding import StateEncoding


def geometry_of(state_size, angular_indices):
    """Maps (D, angular dims) to a kernel geometry id; unsupported geometries raise."""
    ang = tuple(int(i) for i in angular_indices)
    for geo, (D, _, a, _) in _lib.GEO_INFO.items():
        if D == state_size and a == ang:
            return geo
    raise NotImplementedError("pddp_b200 has kernels for the pendulum (D=2, angle 0), cartpole (D=4, angle 2), "
                              "double-cartpole (D=6, angles 2,4) and rendezvous (D=8, no angles) state geometries; "
                              "got D=%d angles=%s" % (state_size, ang))


class DynamicsModel(torch.nn.Module):
    """ref: pddp/models/base.py:23-83 (same members; class-level sizes are plain class attributes)."""
    action_size = None
    state_size = None
    angular_indices = torch.tensor([]).long()
    non_angular_indices = torch.tensor([]).long()

    def fit(self, X, U, dX, quiet=False, **kwargs):
        raise NotImplementedError

    def descriptor(self):
        """KnownDynamics / BNNDynamics consumed by pddp_b200.solver.BatchedSolver."""
        raise NotImplementedError

    def reset_parameters(self, initializer=torch.nn.init.normal_):
        """ref: pddp/models/base.py:27-40"""
        for p in self.parameters():
            if p.requires_grad:
                initializer(p)
        return self

    def forward(self, z, u, i, encoding=StateEncoding.DEFAULT, identical_inputs=False, **kwargs):
        """Next encoded state distribution(s) z' for z:[..., nz], u:[..., nu] (on the GPU).

        One step of the device rollout kernel with zero gains (k = K = 0, alpha = 1)."""
        squeeze = z.dim() == 1
        zz = z.reshape(-1, z.shape[-1])
        uu = u.reshape(-1, u.shape[-1]).expand(zz.shape[0], -1)
        _lib.require_cuda(zz, "z")
        if getattr(self, "is_bnn", False) and int(i) != 0:
            raise NotImplementedError("pddp_b200: a stand-alone BNN forward is only defined for step 0 (particles "
                                      "drawn from eps_in[0]); later steps depend on the particle cache that the "
                                      "controller's rollout carries on the device")
        kwargs.pop("resample", None)
        s = cached_solver(self, None, encoding, zz.shape[0], 1, zz.dtype, zz.device, model_opts=kwargs)
        s.set_problem(zz, uu.unsqueeze(1), alphas=torch.ones(1))
        s.view("Z")[:, 0] = zz.detach()
        s.view("k").zero_()
        s.view("K").zero_()
        s.rollout(use_active=False, use_bw_status=False)
        out = s.view("Z_new")[:, 1].clone()
        return out[0] if squeeze else out.reshape(*z.shape[:-1], -1)


def _DA(model):
    return model.state_size + len(model.angular_indices)


class KnownDynamicsModel(DynamicsModel):
    """Closed-form example dynamics (pddp/examples/<problem>/model.py): constants are Parameters like the
    reference's; the step itself is a device function (csrc/core.cuh, csrc/known_lq.cu)."""
    _param_order = ()
    action_size = 1

    def fit(self, X, U, dX, quiet=False, **kwargs):
        pass            # ref: examples/*/model.py -- known models have nothing to fit

    def descriptor(self):
        geo = geometry_of(self.state_size, self.angular_indices.tolist())
        return KnownDynamics(geo, [float(getattr(self, n).detach()) for n in self._param_order])
