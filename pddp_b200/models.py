"""Dynamics-model interface (mirror of pddp/models/base.py) and the supported model families.

A model object is a plain torch.nn.Module holding constants / weights exactly like the reference's;
what the CUDA kernels consume is its `descriptor()` (flat constants for the closed-form models,
weights + persistent dropout masks + eps_in[0] for the BNN).  Nothing here computes dynamics on the
CPU: `forward()` runs one step of the device rollout kernel.
"""
import math
from collections import OrderedDict

import torch
from torch.nn import Parameter

from . import _lib
from .encoding import StateEncoding
from .solver import BatchedSolver, BNNDynamics, KnownDynamics, QRCostConstants


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
        nu = int(self.action_size)
        dummy = QRCostConstants(torch.zeros(_DA(self), _DA(self)), torch.zeros(nu, nu), torch.zeros(_DA(self), _DA(self)),
                                torch.zeros(_DA(self)))
        s = BatchedSolver(self.descriptor(), dummy, encoding, zz.shape[0], 1, dtype=zz.dtype, device=zz.device)
        s.set_problem(zz, uu.unsqueeze(1), alphas=torch.ones(1))
        s.view("Z")[:, 0] = zz
        s.rollout(use_active=False, use_bw_status=False)
        out = s.view("Z_new")[:, 1].clone()
        return out[0] if squeeze else out.reshape(*z.shape[:-1], -1)


def _DA(model):
    return model.state_size + len(model.angular_indices)


class _KnownModel(DynamicsModel):
    _param_order = ()
    action_size = 1

    def fit(self, X, U, dX, quiet=False, **kwargs):
        pass            # ref: examples/*/model.py -- known models have nothing to fit

    def descriptor(self):
        geo = geometry_of(self.state_size, self.angular_indices.tolist())
        return KnownDynamics(geo, [float(getattr(self, n).detach()) for n in self._param_order])


class PendulumDynamicsModel(_KnownModel):
    """ref: pddp/examples/pendulum/model.py:33-119 (state [theta, theta'], action [torque])."""
    state_size = 2
    angular_indices = torch.tensor([0]).long()
    non_angular_indices = torch.tensor([1]).long()
    _param_order = ("dt", "m", "l", "mu", "g")

    def __init__(self, dt, m=1.0, l=1.0, mu=0.1, g=9.80665):
        super().__init__()
        self.dt = Parameter(torch.tensor(dt), requires_grad=False)
        for n, v in (("m", m), ("l", l), ("mu", mu), ("g", g)):
            setattr(self, n, Parameter(torch.tensor(v), requires_grad=True))


class CartpoleDynamicsModel(_KnownModel):
    """ref: pddp/examples/cartpole/model.py:30-141 (state [x, x', theta, theta'], action [F])."""
    state_size = 4
    angular_indices = torch.tensor([2]).long()
    non_angular_indices = torch.tensor([0, 1, 3]).long()
    _param_order = ("dt", "mc", "mp", "l", "mu", "g")

    def __init__(self, dt, mc=0.5, mp=0.5, l=0.5, mu=0.1, g=9.82):
        super().__init__()
        self.dt = Parameter(torch.tensor(dt), requires_grad=False)
        for n, v in (("mc", mc), ("mp", mp), ("l", l), ("mu", mu), ("g", g)):
            setattr(self, n, Parameter(torch.tensor(v), requires_grad=True))


class DoubleCartpoleDynamicsModel(_KnownModel):
    """ref: pddp/examples/double_cartpole/model.py:33-195."""
    state_size = 6
    angular_indices = torch.tensor([2, 4]).long()
    non_angular_indices = torch.tensor([0, 1, 3, 5]).long()
    _param_order = ("dt", "mc", "mp1", "mp2", "l1", "l2", "mu", "g")

    def __init__(self, dt, mc=0.5, mp1=0.5, mp2=0.5, l1=0.6, l2=0.6, mu=0.1, g=9.80665):
        super().__init__()
        self.dt = Parameter(torch.tensor(dt), requires_grad=False)
        for n, v in (("mc", mc), ("mp1", mp1), ("mp2", mp2), ("l1", l1), ("l2", l2), ("mu", mu), ("g", g)):
            setattr(self, n, Parameter(torch.tensor(v), requires_grad=True))


class RendezvousDynamicsModel(_KnownModel):
    """ref: pddp/examples/rendezvous/model.py:25-115 (state [x0, y0, x1, y1, and their velocities],
    action [Fx0, Fy0, Fx1, Fy1]; linear dynamics, the whole covariance is passed through)."""
    state_size = 8
    action_size = 4
    angular_indices = torch.tensor([]).long()
    non_angular_indices = torch.arange(8).long()
    _param_order = ("dt", "m", "alpha")

    def __init__(self, dt, m=1.0, alpha=0.1):
        super().__init__()
        self.dt = Parameter(torch.tensor(dt), requires_grad=False)
        for n, v in (("m", m), ("alpha", alpha)):
            setattr(self, n, Parameter(torch.tensor(v), requires_grad=True))


# ---------------------------------------------------------------------------------------------
# BNN (MC-dropout) dynamics
# ---------------------------------------------------------------------------------------------
class CDropoutMask(torch.nn.Module):
    """Eval-mode concrete-dropout mask: sigmoid((logit_p + log r - log(1-r)) / temperature), drawn
    once per (particles, width) and kept (ref: pddp/models/bnn/modules.py:486-583)."""

    def __init__(self, rate=0.5, temperature=0.1):
        super().__init__()
        self.logit_p = Parameter(torch.tensor(math.log((1 - rate) / rate)))   # logit of keep-probability 1-rate
        self.temperature = temperature
        self.mask = None

    def draw(self, P, H, generator=None):
        r = torch.rand(P, H, generator=generator)
        self.mask = torch.sigmoid((self.logit_p.detach().cpu() + r.log() - (1 - r).log()) / self.temperature)
        return self.mask


def bnn_dynamics_model_factory(state_size, action_size, hidden_features, angular_indices=None,
                               non_angular_indices=None, **kwargs):
    """ref: pddp/models/bnn/modules.py:44-391.  Returns a BNNDynamicsModel class."""
    if action_size != 1 or len(hidden_features) != 2:
        raise NotImplementedError("pddp_b200: BNN kernels need action_size == 1 and two hidden layers")
    ang = [] if angular_indices is None else [int(i) for i in angular_indices]
    geo = geometry_of(state_size, ang)
    DA = state_size + len(ang)
    _state_size, _ang = state_size, ang

    class BNNDynamicsModel(DynamicsModel):
        state_size = _state_size
        action_size = 1
        angular_indices = torch.tensor(_ang).long()
        non_angular_indices = torch.tensor([i for i in range(_state_size) if i not in _ang]).long()
        is_bnn = True

        def __init__(self, n_particles=100):
            super().__init__()
            dims = [DA + 1] + list(hidden_features) + [2 * _state_size]
            layers = OrderedDict()
            for li, (din, dout) in enumerate(zip(dims[:-1], dims[1:])):
                name = "fc_out" if li == len(dims) - 2 else "fc_%d" % li
                lin = torch.nn.Linear(din, dout)
                torch.nn.init.xavier_normal_(lin.weight, gain=torch.nn.init.calculate_gain("relu"))
                torch.nn.init.uniform_(lin.bias, -0.1, 0.1)                    # ref: modules.py:797-801
                layers[name] = lin
                if name != "fc_out":
                    layers["drop_%d" % li] = CDropoutMask()
            self.model = torch.nn.Sequential(layers)
            self.n_particles = n_particles
            for n, v in (("X_mean", 0.0), ("X_std", 1.0), ("X_std_inv", 1.0), ("dX_mean", 0.0), ("dX_std", 1.0),
                         ("dX_std_inv", 1.0)):
                self.register_buffer(n, torch.tensor(v))                        # ref: modules.py:93-98
            self.eps_in = {}
            self.eps_out = {}

        def resample(self, generator=None):
            """New eps_in[0] and dropout masks (ref: modules.py:281-285; draw order eps, drop_0, drop_1)."""
            P = self.n_particles
            eps = torch.randn(P, _state_size, generator=generator)
            self.eps_in = {0: (eps - eps.mean(0)) / eps.std(0)}                 # ref: modules.py:321-329
            self.eps_out = {}
            for name, mod in self.model._modules.items():
                if isinstance(mod, CDropoutMask):
                    mod.draw(P, self.model._modules["fc_" + name.split("_")[1]].out_features, generator)

        def load_reference(self, ref_model):
            """Copies weights, normalisation buffers, the dropout masks and eps_in[0] a reference
            BNNDynamicsModel object currently holds (bit-exact mask / particle indexing)."""
            for name in ("fc_0", "fc_1", "fc_out"):
                src = getattr(ref_model.model, name)
                dst = getattr(self.model, name)
                dst.weight.data.copy_(src.weight.data)
                dst.bias.data.copy_(src.bias.data)
            for li in (0, 1):
                drop = getattr(ref_model.model, "drop_%d" % li)
                mask = getattr(drop, "concrete_noise", None)
                getattr(self.model, "drop_%d" % li).mask = (drop.noise if mask is None else mask).detach().clone()
            self.eps_in = {int(i): e.detach().clone() for i, e in ref_model.eps_in.items()}
            self.eps_out = {int(i): e.detach().clone() for i, e in getattr(ref_model, "eps_out", {}).items()}
            self.n_particles = self.eps_in[0].shape[0]
            for n in ("X_mean", "X_std", "X_std_inv", "dX_mean", "dX_std", "dX_std_inv"):
                setattr(self, n, getattr(ref_model, n).detach().clone())
            return self

        def fit(self, X, U, dX, n_iter=500, batch_size=128, learning_rate=1e-4, normalize=True, quiet=False, **kw):
            """Host-side PyTorch training (outside the hot path): Adam on the Gaussian
            log-likelihood of dX with fresh Bernoulli-relaxed masks per batch."""
            from .encoding import StateEncoding as _E  # noqa: F401
            Xa = _augment(X, _ang)
            X_ = torch.cat([Xa, U], -1)
            if normalize:
                self.X_mean, self.X_std = X_.mean(0).detach(), X_.std(0).detach()
                self.X_std_inv = self.X_std.reciprocal()
                self.dX_mean, self.dX_std = dX.mean(0).detach(), dX.std(0).detach()
                self.dX_std_inv = self.dX_std.reciprocal()
            opt = torch.optim.Adam([p for p in self.parameters() if p.requires_grad], learning_rate, amsgrad=True)
            lin = [m for m in self.model if isinstance(m, torch.nn.Linear)]
            for _ in range(n_iter):
                idx = torch.randint(0, X_.shape[0], (min(batch_size, X_.shape[0]),))
                h = (X_[idx] - self.X_mean) * self.X_std_inv
                for l in lin[:-1]:
                    r = torch.rand(h.shape[0], l.out_features)
                    h = torch.relu(l(h) * torch.sigmoid((r.log() - (1 - r).log()) / 0.1))
                mean, log_std = lin[-1](h).split([_state_size, _state_size], -1)
                mean = mean * self.dX_std + self.dX_mean
                log_std = log_std + self.dX_std.log()
                loss = (0.5 * ((dX[idx] - mean) / log_std.exp()) ** 2 + log_std).sum(-1).mean()
                opt.zero_grad()
                loss.backward()
                opt.step()
            self.resample()

        def descriptor(self, model_opts=None, N=None):
            """BNNDynamics for the kernels.  model_opts selects how input particles are formed
            (ref: modules.py:320-358): infer_noise_variables=False needs eps_in[i] for every step
            i < N; steps the model has not seen yet are drawn here, in step order, the way the
            reference draws them on first use (modules.py:321-329)."""
            opts = model_opts or {}
            if 0 not in self.eps_in:
                self.resample()
            mode, eps_in = _lib.BNN_INPUT_INFER, None
            if not opts.get("sample_input_distribution", True):
                mode = _lib.BNN_INPUT_MEAN
            elif not opts.get("infer_noise_variables", True):
                mode = _lib.BNN_INPUT_RESAMPLE
                if N is None:
                    raise ValueError("infer_noise_variables=False needs the horizon N to lay out eps_in")
                for i in range(N):
                    if i not in self.eps_in:
                        eps = torch.randn(self.n_particles, _state_size)
                        self.eps_in[i] = (eps - eps.mean(0)) / eps.std(0)
                eps_in = torch.stack([self.eps_in[i] for i in range(N)])
            eps_out = None
            if opts.get("use_predicted_std", False):        # ref: modules.py:242-262, eps_out[i] drawn on first use
                if N is None:
                    raise ValueError("use_predicted_std=True needs the horizon N to lay out eps_out")
                for i in range(N):
                    if i not in self.eps_out:
                        eps = torch.randn(self.n_particles, _state_size)
                        self.eps_out[i] = (eps - eps.mean(0)) / eps.std(0)
                eps_out = torch.stack([self.eps_out[i] for i in range(N)])
            m = self.model
            vec = lambda b: None if b.dim() == 0 else b
            return BNNDynamics(geo, [m.fc_0.weight, m.fc_1.weight, m.fc_out.weight],
                               [m.fc_0.bias, m.fc_1.bias, m.fc_out.bias], [m.drop_0.mask, m.drop_1.mask],
                               self.eps_in[0], vec(self.X_mean), vec(self.X_std_inv), vec(self.dX_mean),
                               vec(self.dX_std), input_mode=mode, eps_in=eps_in, eps_out=eps_out,
                               independent_noise=bool(opts.get("independent_noise", False)))

    return BNNDynamicsModel


def _augment(x, ang):
    """[x_nonang, sin a1, cos a1, ...]   ref: pddp/utils/angular.py:251-286"""
    if not ang:
        return x
    non = [i for i in range(x.shape[-1]) if i not in ang]
    a = x[..., ang]
    sc = torch.stack([a.sin(), a.cos()], -1).reshape(*x.shape[:-1], 2 * len(ang))
    return torch.cat([x[..., non], sc], -1)


def check_model_opts(model, model_opts):
    """Options the kernels implement: infer_noise_variables, sample_input_distribution, use_predicted_std and
    independent_noise either way (BNNDynamics.input_mode / eps_out); resample=True (fresh noise on every call) is
    not: all noise is data on this path."""
    if not getattr(model, "is_bnn", False):
        return
    want = dict(resample=False)
    for k, v in model_opts.items():
        if k in want and bool(v) != want[k]:
            raise NotImplementedError("pddp_b200: model option %s=%r is not built (SURVEY 8f rank 2); supported: %r"
                                      % (k, v, want))
