"""CPU oracle: a from-scratch restatement of the reference's iteration hot path.

TEST INFRASTRUCTURE ONLY -- this file is the checker, never the product.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import it.
`pddp_b200/` must never import anything under `oracle/`.

Parity pin: every function here is checked against the UNMODIFIED reference (anassinator/pddp run
through oracle/refshim.py under torch 2.11) by oracle/make_golden.py, which also writes the golden
fixtures in tests/golden/*.npz that `tests/test_oracle_golden.py` replays on any machine.
The reference's own tests hold no golden vectors for this path, only known-answer
pins (Hessian of QRCost == Q+Q^T, encoded sizes, U^T U == C, zero-variance augmentation) which
tests/test_oracle_known_answers.py replays against this file.

Arithmetic is torch CPU (fp32 or fp64) with reverse-mode autograd for the Jacobians/Hessians, the
same mechanism the reference uses (pddp/utils/evaluation.py:134-288), so that derivative
conventions (e.g. the symmetrised Cholesky gradient) are identical.

Citations `ref:` are relative to /root/reference/pddp/.
"""
import math
from dataclasses import dataclass, field
from typing import Optional, Sequence

import torch

# ref: utils/encoding.py:25-33
FULL_COVARIANCE_MATRIX = 0
UPPER_TRIANGULAR_CHOLESKY = 1
DEFAULT = 1
VARIANCE_ONLY = 2
STANDARD_DEVIATION_ONLY = 3
IGNORE_UNCERTAINTY = 4

# ref: controllers/ilqr.py:35-64
UNDEFINED, ACCEPTED, REJECTED, NOT_PD, MAX_REG, CONVERGED = range(6)


class NotPositiveDefinite(RuntimeError):
    pass


# --------------------------------------------------------------------------------------
# State encodings (ref: utils/encoding.py)
# --------------------------------------------------------------------------------------
def encoded_size(D, enc):
    """ref: utils/encoding.py:46-67"""
    return {FULL_COVARIANCE_MATRIX: D + D * D,
            UPPER_TRIANGULAR_CHOLESKY: (3 * D + D * D) // 2,
            VARIANCE_ONLY: 2 * D, STANDARD_DEVIATION_ONLY: 2 * D,
            IGNORE_UNCERTAINTY: D}[enc]


def state_size(nz, enc):
    """ref: utils/encoding.py:70-96"""
    if enc == FULL_COVARIANCE_MATRIX:
        return int(0.5 * (-1 + math.sqrt(1 + 4 * nz)))
    if enc == UPPER_TRIANGULAR_CHOLESKY:
        return int(0.5 * (-3 + math.sqrt(9 + 8 * nz)))
    if enc in (VARIANCE_ONLY, STANDARD_DEVIATION_ONLY):
        return nz // 2
    return nz


def _triu(D):
    return torch.triu_indices(D, D)  # row-major order of the upper triangle (ref: encoding.py:129)


def chol_upper(C, jitter=1e-12, max_jitter=10.0):
    """Upper Cholesky factor U (U^T U = C + jitter I) with the reference's escalating jitter.

    ref: utils/encoding.py:536-564 -- jitter starts at 1e-12 and is ALWAYS added; on failure it is
    multiplied by 10 until it exceeds 10, then the error propagates.  Batched inputs are
    factorised row by row there; here rows that fail are retried individually (same result).
    """
    eye = torch.eye(C.shape[-1], dtype=C.dtype)
    if C.dim() == 2:
        j = jitter
        while True:
            L, info = torch.linalg.cholesky_ex(C + j * eye)
            if int(info) == 0:
                return L.mT
            j *= 10
            if j > max_jitter:
                raise NotPositiveDefinite("cholesky: jitter exceeded")
    return torch.stack([chol_upper(c, jitter, max_jitter) for c in C.reshape(-1, *C.shape[-2:])]
                       ).reshape(C.shape)


def split(z, enc, D=None):
    if D is None:
        D = state_size(z.shape[-1], enc)
    return z[..., :D], z[..., D:], D


def _unflatten_triu(flat, D):
    iu = _triu(D)
    U = torch.zeros(*flat.shape[:-1], D, D, dtype=flat.dtype)
    U[..., iu[0], iu[1]] = flat
    return U


def decode_mean(z, enc, D=None):
    return split(z, enc, D)[0]


def decode_covar(z, enc, D=None):
    """ref: utils/encoding.py:159-216"""
    m, other, D = split(z, enc, D)
    if enc == FULL_COVARIANCE_MATRIX:
        return other.reshape(*z.shape[:-1], D, D)
    if enc == UPPER_TRIANGULAR_CHOLESKY:
        U = _unflatten_triu(other, D)
        return U.mT @ U
    if enc == VARIANCE_ONLY:
        return torch.diag_embed(other)
    if enc == STANDARD_DEVIATION_ONLY:
        return torch.diag_embed(other ** 2)
    return (1e-6 * torch.eye(D, dtype=z.dtype)).expand(*z.shape[:-1], D, D)


def decode_var(z, enc, D=None):
    """ref: utils/encoding.py:219-258"""
    m, other, D = split(z, enc, D)
    if enc == FULL_COVARIANCE_MATRIX:
        return other[..., ::D + 1]
    if enc == UPPER_TRIANGULAR_CHOLESKY:
        return (_unflatten_triu(other, D) ** 2).sum(-2)
    if enc == VARIANCE_ONLY:
        return other
    if enc == STANDARD_DEVIATION_ONLY:
        return other ** 2
    return (1e-6 * torch.ones(D, dtype=z.dtype)).expand(*z.shape[:-1], D)


def decode_covar_sqrt(z, enc, D=None):
    """ref: utils/encoding.py:304-362"""
    m, other, D = split(z, enc, D)
    if enc == FULL_COVARIANCE_MATRIX:
        return chol_upper(other.reshape(*z.shape[:-1], D, D))
    if enc == UPPER_TRIANGULAR_CHOLESKY:
        return _unflatten_triu(other, D)
    if enc == VARIANCE_ONLY:
        return torch.diag_embed(other.sqrt())
    if enc == STANDARD_DEVIATION_ONLY:
        return torch.diag_embed(other)
    return (1e-3 * torch.eye(D, dtype=z.dtype)).expand(*z.shape[:-1], D, D)


def encode(M, C=None, V=None, S=None, enc=DEFAULT):
    """ref: utils/encoding.py:99-141"""
    if enc == IGNORE_UNCERTAINTY:
        return M
    D = M.shape[-1]

    def var():
        if V is not None:
            return V
        if S is not None:
            return S ** 2
        return torch.diagonal(C, dim1=-2, dim2=-1)

    def cov():
        return C if C is not None else torch.diag_embed(var())

    if enc == FULL_COVARIANCE_MATRIX:
        other = cov().reshape(*M.shape[:-1], D * D)
    elif enc == UPPER_TRIANGULAR_CHOLESKY:
        iu = _triu(D)
        other = chol_upper(cov())[..., iu[0], iu[1]]
    elif enc == VARIANCE_ONLY:
        other = var()
    elif enc == STANDARD_DEVIATION_ONLY:
        other = S if S is not None else var().sqrt()
    else:
        raise NotImplementedError(enc)
    return torch.cat([M, other], -1)


# --------------------------------------------------------------------------------------
# Angular augmentation (ref: utils/angular.py)
# --------------------------------------------------------------------------------------
def augment_state(x, ang, nonang):
    """x -> [x_nonang, sin a1, cos a1, sin a2, ...]   ref: utils/angular.py:251-286"""
    if len(ang) == 0:
        return x
    a = x[..., list(ang)]
    sc = torch.stack([a.sin(), a.cos()], -1).reshape(*x.shape[:-1], 2 * len(ang))
    return torch.cat([x[..., list(nonang)], sc], -1)


def augment_moments(m, c, ang, nonang):
    """Gaussian (m, c) pushed through [x_na, sin, cos] in closed form.

    ref: utils/angular.py:161-248 (_augment_covar).  Note the cross block uses c[d, j] with d the
    ANGULAR row and j the non-angular column (c^T . Ca), and the lower-left block is the transpose
    of the upper-right one -- relevant for FULL_COVARIANCE where c_ij and c_ji are separate inputs.
    """
    ang, nonang = list(ang), list(nonang)
    if not ang:
        return m, c
    na, Dna = len(ang), len(nonang)
    mi = m[..., ang]
    ci = c[..., ang, :][..., :, ang]
    cii = torch.diagonal(ci, dim1=-2, dim2=-1)
    damp = (-0.5 * cii).exp()
    s, co = damp * mi.sin(), damp * mi.cos()
    Ma = torch.stack([s, co], -1).reshape(*m.shape[:-1], 2 * na)
    lq = -0.5 * (cii.unsqueeze(-1) + cii.unsqueeze(-2))
    q = lq.exp()
    ep, em = (lq + ci).exp() - q, (lq - ci).exp() - q
    dm = mi.unsqueeze(-1) - mi.unsqueeze(-2)
    sm = mi.unsqueeze(-1) + mi.unsqueeze(-2)
    U1, U2 = ep * dm.sin(), em * sm.sin()
    U3, U4 = ep * dm.cos(), em * sm.cos()
    ss, cc, sc = 0.5 * (U3 - U4), 0.5 * (U3 + U4), 0.5 * (U1 + U2)
    # interleave: Va[2i, 2j]=ss, Va[2i+1,2j+1]=cc, Va[2i,2j+1]=sc, Va[2i+1,2j]=sc^T
    top = torch.stack([ss, sc], -1).reshape(*ss.shape[:-1], 2 * na)           # rows 2i
    bot = torch.stack([sc.mT, cc], -1).reshape(*ss.shape[:-1], 2 * na)        # rows 2i+1
    Va = torch.stack([top, bot], -2).reshape(*ss.shape[:-2], 2 * na, 2 * na)
    M = torch.cat([m[..., nonang], Ma], -1)
    if Dna == 0:
        return M, Va
    # Ca[d, 2i] = cos-moment, Ca[d, 2i+1] = -sin-moment for d == ang[i]
    Vna = c[..., nonang, :][..., :, nonang]
    crows = c[..., ang, :][..., :, nonang]                      # [.., na, Dna]  c[ang_i, j]
    cross_s = crows * co.unsqueeze(-1)                          # col 2i
    cross_c = -crows * s.unsqueeze(-1)                          # col 2i+1
    cross = torch.stack([cross_s, cross_c], -2).reshape(*crows.shape[:-2], 2 * na, Dna).mT
    Ctop = torch.cat([Vna, cross], -1)
    Cbot = torch.cat([cross.mT, Va], -1)
    return M, torch.cat([Ctop, Cbot], -2)


def _augment_var(m, v, ang, nonang):
    """ref: utils/angular.py:87-158 (diagonal variant)."""
    ang, nonang = list(ang), list(nonang)
    if not ang:
        return m, v
    mi, vi = m[..., ang], v[..., ang]
    damp = (-0.5 * vi).exp()
    Ma = torch.stack([damp * mi.sin(), damp * mi.cos()], -1).reshape(*m.shape[:-1], 2 * len(ang))
    q = (-vi).exp()
    U3 = (1.0 - q)                       # exp(lq+v) - q with lq=-v ; cos(0) = 1
    U4 = ((-2 * vi).exp() - q) * (2 * mi).cos()
    Va = 0.5 * torch.stack([U3 - U4, U3 + U4], -1).reshape(*m.shape[:-1], 2 * len(ang))
    return torch.cat([m[..., nonang], Ma], -1), torch.cat([v[..., nonang], Va], -1)


def augment_encoded_state(z, ang, nonang, enc, D):
    """ref: utils/angular.py:47-84"""
    if enc == IGNORE_UNCERTAINTY:
        return augment_state(z, ang, nonang)
    m = decode_mean(z, enc, D)
    if enc in (FULL_COVARIANCE_MATRIX, UPPER_TRIANGULAR_CHOLESKY):
        M, C = augment_moments(m, decode_covar(z, enc, D), ang, nonang)
        return encode(M, C=C, enc=enc)
    M, V = _augment_var(m, decode_var(z, enc, D), ang, nonang)
    return encode(M, V=V, enc=enc)


# --------------------------------------------------------------------------------------
# Costs (ref: costs/quadratic.py:60-99 and examples/*/cost.py)
# --------------------------------------------------------------------------------------
@dataclass
class QRCostSpec:
    """Constants of a (possibly angle-augmented) QRCost.  Q/Q_term act on the augmented state."""
    Q: torch.Tensor
    R: torch.Tensor
    Q_term: torch.Tensor
    x_goal: torch.Tensor
    u_goal: torch.Tensor
    D: int
    ang: Sequence[int] = ()
    nonang: Sequence[int] = ()

    def to(self, dtype):
        return QRCostSpec(self.Q.to(dtype), self.R.to(dtype), self.Q_term.to(dtype),
                          self.x_goal.to(dtype), self.u_goal.to(dtype), self.D,
                          tuple(self.ang), tuple(self.nonang))


def cost_value(spec, z, u, terminal, enc):
    """E[l] = dx^T Q dx + du^T R du + sum(C * Q^T), on the augmented encoded state."""
    za = augment_encoded_state(z, spec.ang, spec.nonang, enc, spec.D)
    Q = spec.Q_term if terminal else spec.Q
    dx = decode_mean(za, enc) - spec.x_goal
    val = ((dx @ Q) * dx).sum(-1)
    if not terminal:
        du = u - spec.u_goal
        val = val + ((du @ spec.R) * du).sum(-1)
    if enc != IGNORE_UNCERTAINTY:
        C = decode_covar(za, enc)
        val = val + (C * Q.mT).sum((-2, -1))
    return val


# --------------------------------------------------------------------------------------
# Known dynamics (ref: examples/*/model.py)
# --------------------------------------------------------------------------------------
@dataclass
class KnownDynamicsSpec:
    kind: str                     # "pendulum" | "cartpole" | "double_cartpole"
    params: dict
    D: int
    nu: int
    ang: Sequence[int]
    nonang: Sequence[int]


def pendulum_spec(dt, m=1.0, l=1.0, mu=0.1, g=9.80665):
    return KnownDynamicsSpec("pendulum", dict(dt=dt, m=m, l=l, mu=mu, g=g), 2, 1, (0,), (1,))


def cartpole_spec(dt, mc=0.5, mp=0.5, l=0.5, mu=0.1, g=9.82):
    return KnownDynamicsSpec("cartpole", dict(dt=dt, mc=mc, mp=mp, l=l, mu=mu, g=g), 4, 1, (2,),
                             (0, 1, 3))


def double_cartpole_spec(dt, mc=0.5, mp1=0.5, mp2=0.5, l1=0.6, l2=0.6, mu=0.1, g=9.80665):
    return KnownDynamicsSpec("double_cartpole",
                             dict(dt=dt, mc=mc, mp1=mp1, mp2=mp2, l1=l1, l2=l2, mu=mu, g=g), 6, 1,
                             (2, 4), (0, 1, 3, 5))


def _pendulum(p, x, u):
    """ref: examples/pendulum/model.py:84-119"""
    th, thd, tq = x[..., 0], x[..., 1], u[..., 0]
    ml = p["m"] * p["l"]
    acc = 3 * (tq - p["mu"] * thd - 0.5 * ml * p["g"] * th.sin()) / (ml * p["l"])
    return torch.stack([th + thd * p["dt"], thd + acc * p["dt"]], -1)


def _cartpole(p, x, u):
    """ref: examples/cartpole/model.py:88-141"""
    pos, vel, th, thd, F = x[..., 0], x[..., 1], x[..., 2], x[..., 3], u[..., 0]
    mc, mp, l, mu, g, dt = (p[k] for k in ("mc", "mp", "l", "mu", "g", "dt"))
    s, c = th.sin(), th.cos()
    a0 = mp * l * thd ** 2 * s
    a1 = g * s
    a2 = F - mu * vel
    a3 = 4 * (mc + mp) - 3 * mp * c ** 2
    thdd = -3 * (a0 * c + 2 * ((mc + mp) * a1 + a2 * c)) / (l * a3)
    acc = (2 * a0 + 3 * mp * a1 * c + 4 * a2) / a3
    nvel, nthd = vel + acc * dt, thd + thdd * dt
    return torch.stack([pos + nvel * dt, nvel, th + nthd * dt, nthd], -1)


def _double_cartpole(p, x, u):
    """ref: examples/double_cartpole/model.py:100-195 (A sol = b, 3x3)."""
    pos, vel, t1, t1d, t2, t2d, F = (x[..., 0], x[..., 1], x[..., 2], x[..., 3], x[..., 4],
                                     x[..., 5], u[..., 0])
    mc, mp1, mp2, l1, l2, mu, g, dt = (p[k] for k in ("mc", "mp1", "mp2", "l1", "l2", "mu", "g",
                                                      "dt"))
    s1, c1, s2, c2 = t1.sin(), t1.cos(), t2.sin(), t2.cos()
    sd, cd = (t1 - t2).sin(), (t1 - t2).cos()
    a0 = mp2 + 2 * mc
    a1 = mc * l2
    a2 = l1 * t1d ** 2
    a3 = a1 * t2d ** 2
    one = torch.ones_like(pos)
    A = torch.stack([
        torch.stack([2 * (mp1 + mp2 + mc) * one, -a0 * l1 * c1, -a1 * c2], -1),
        torch.stack([-3 * a0 * c1, (2 * a0 + 2 * mc) * l1 * one, 3 * a1 * cd], -1),
        torch.stack([-3 * c2, 3 * l1 * cd, 2 * l2 * one], -1)], -2)
    b = torch.stack([2 * F - 2 * mu * vel - a0 * a2 * s1 - a3 * s2,
                     3 * a0 * g * s1 - 3 * a3 * sd,
                     3 * a2 * sd + 3 * g * s2], -1).unsqueeze(-1)
    sol = torch.linalg.solve(A, b).squeeze(-1)
    nvel, n1, n2 = vel + sol[..., 0] * dt, t1d + sol[..., 1] * dt, t2d + sol[..., 2] * dt
    return torch.stack([pos + nvel * dt, nvel, t1 + n1 * dt, n1, t2 + n2 * dt, n2], -1)


def rendezvous_spec(dt, m=1.0, alpha=0.1):
    return KnownDynamicsSpec("rendezvous", dict(dt=dt, m=m, alpha=alpha), 8, 4, (), tuple(range(8)))


def _rendezvous(p, x, u):
    """ref: examples/rendezvous/model.py:79-115 -- two point masses with friction; the reference's
    `_acceleration` returns v (1 - alpha dt / m) + u dt / m and the step adds it times dt."""
    dt, m, al = p["dt"], p["m"], p["alpha"]
    pos, vel = x[..., :4], x[..., 4:]
    acc = vel * (1 - al * dt / m) + u * dt / m
    return torch.cat([pos + vel * dt, vel + acc * dt], -1)


_KNOWN = {"pendulum": _pendulum, "cartpole": _cartpole, "double_cartpole": _double_cartpole,
          "rendezvous": _rendezvous}


def known_step(spec, z, u, enc, carry=None, i=None):
    """Mean through the ODE step, variance passed through unchanged; the
    rendezvous model passes the whole covariance through (ref: rendezvous/model.py:98,112)."""
    mean = _KNOWN[spec.kind](spec.params, decode_mean(z, enc, spec.D), u)
    if spec.kind == "rendezvous":
        return encode(mean, C=decode_covar(z, enc, spec.D), enc=enc), None
    return encode(mean, V=decode_var(z, enc, spec.D), enc=enc), None


# --------------------------------------------------------------------------------------
# BNN dynamics, eval mode (ref: models/bnn/modules.py:200-264, 287-386, 774-789)
# --------------------------------------------------------------------------------------
@dataclass
class BNNSpec:
    """Everything the eval-mode BNN forward reads.  Masks and eps0 are DATA (drawn once by the
    model object), so there is no RNG on this path."""
    weights: list                 # [(W[out,in], b[out]), ...]  hidden layers then output layer
    masks: list                   # one [P, H_l] multiplicative mask per hidden layer
    eps0: torch.Tensor            # [P, D] standardised input noise of step 0
    D: int
    nu: int
    ang: Sequence[int] = ()
    nonang: Sequence[int] = ()
    X_mean: Optional[torch.Tensor] = None      # [Da+nu] or None (0)
    X_std_inv: Optional[torch.Tensor] = None   # [Da+nu] or None (1)
    dX_mean: Optional[torch.Tensor] = None     # [D] or None (0)
    dX_std: Optional[torch.Tensor] = None      # [D] or None (1)
    # input particles (ref: modules.py:320-358).  "infer": eps of step i > 0 is inferred from the previous
    # output particles (infer_noise_variables=True, the default); "resample": eps_in[i] of every step
    # (infer_noise_variables=False); "mean": every particle starts at the mean
    # (sample_input_distribution=False)
    input_mode: str = "infer"
    eps_in: Optional[torch.Tensor] = None      # [N, P, D], eps_in[0] == eps0 ("resample" only)
    # use_predicted_std=True (ref: modules.py:242-262): dx += exp(log_std) * eps_out[i], log_std = the second
    # half of the output layer + log(dX_std); independent_noise detaches exp(log_std)
    eps_out: Optional[torch.Tensor] = None     # [N, P, D] or None (use_predicted_std=False)
    independent_noise: bool = False

    @property
    def P(self):
        return self.eps0.shape[0]

    def to(self, dtype):
        c = lambda t: None if t is None else t.to(dtype)
        return BNNSpec([(W.to(dtype), b.to(dtype)) for W, b in self.weights],
                       [m.to(dtype) for m in self.masks], self.eps0.to(dtype), self.D, self.nu,
                       tuple(self.ang), tuple(self.nonang), c(self.X_mean), c(self.X_std_inv),
                       c(self.dX_mean), c(self.dX_std), self.input_mode, c(self.eps_in), c(self.eps_out),
                       self.independent_noise)


def bnn_particles(spec, X, u, i=None):
    """Particles X:[R,P,D], u:[R,nu] -> next particles [R,P,D].

    ref: models/bnn/modules.py:200-264; masks are [P,H], shared over the leading dim R.
    """
    a = augment_state(X, spec.ang, spec.nonang)
    a = torch.cat([a, u.unsqueeze(-2).expand(*X.shape[:-1], spec.nu)], -1)
    if spec.X_mean is not None:
        a = (a - spec.X_mean) * spec.X_std_inv
    h = a
    for (W, b), mask in zip(spec.weights[:-1], spec.masks):
        h = torch.relu((h @ W.mT + b) * mask)
    W, b = spec.weights[-1]
    out = h @ W.mT + b
    dx, log_std = out[..., :spec.D], out[..., spec.D:]
    if spec.dX_std is not None:
        dx = dx * spec.dX_std + spec.dX_mean
        log_std = log_std + spec.dX_std.log()
    if spec.eps_out is not None:                         # use_predicted_std (ref: modules.py:242-262)
        noise_std = log_std.exp()
        if spec.independent_noise:
            noise_std = noise_std.detach()
        dx = dx + noise_std * spec.eps_out[i]
    return X + dx


def bnn_step(spec, z, u, enc, carry=None, i=None):
    """One moment-matched BNN step on rows z:[R,nz], u:[R,nu]; carry = previous particles [R,P,D].

    ref: models/bnn/modules.py:287-386 with sample_input_distribution=True,
    infer_noise_variables=True.  Step 0 (carry None) uses eps0; later steps infer
    eps = (X_prev - m) U^-1 and DETACH it, so gradients flow through m and U only.
    """
    squeeze = z.dim() == 1
    if squeeze:
        z, u = z.unsqueeze(0), u.unsqueeze(0)
    D, P = spec.D, spec.P
    m = decode_mean(z, enc, D)
    Uc = decode_covar_sqrt(z, enc, D)                                    # [R,D,D] upper
    if spec.input_mode == "mean":
        eps = torch.zeros(z.shape[0], P, D, dtype=z.dtype)
    elif spec.input_mode == "resample":
        eps = spec.eps_in[i].unsqueeze(0).expand(z.shape[0], P, D)
    elif carry is None:
        eps = spec.eps0.unsqueeze(0).expand(z.shape[0], P, D)
    else:
        delta = (carry - m.unsqueeze(-2)).detach()                       # [R,P,D]
        eps = torch.linalg.solve_triangular(Uc.detach(), delta, upper=True, left=False)
    X = m.unsqueeze(-2) + eps @ Uc
    Xn = bnn_particles(spec, X, u, i)
    M = Xn.mean(-2)
    if enc in (FULL_COVARIANCE_MATRIX, UPPER_TRIANGULAR_CHOLESKY):
        d = Xn - M.unsqueeze(-2)
        out = encode(M, C=d.mT @ d / (P - 1), enc=enc)                   # ref: utils/particles.py:136
    else:
        out = encode(M, S=Xn.std(-2), enc=enc)
    if squeeze:
        return out[0], Xn.detach()
    return out, Xn.detach()


def step_fn(dyn):
    return bnn_step if isinstance(dyn, BNNSpec) else known_step


# --------------------------------------------------------------------------------------
# Linearisation (ref: controllers/ilqr.py:393-486, utils/evaluation.py:134-288)
# --------------------------------------------------------------------------------------
def clamp(u, lo, hi):
    return torch.min(torch.max(u, lo), hi)


def cost_derivatives(cost, z, u, terminal, enc):
    """value, gradient and full Hessian of the cost w.r.t. [z,u] by replicating the input row
    n times and back-propagating an identity matrix (ref: utils/evaluation.py:201-239)."""
    nz = z.shape[-1]
    zu = z if terminal else torch.cat([z, u], -1)
    n = zu.shape[0]
    rows = zu.detach().repeat(n, 1).requires_grad_()
    val = cost_value(cost, rows[:, :nz], None if terminal else rows[:, nz:], terminal, enc)
    g, = torch.autograd.grad(val, rows, torch.ones(n, dtype=z.dtype), create_graph=True)
    if g.requires_grad:
        H, = torch.autograd.grad(g, rows, torch.eye(n, dtype=z.dtype), allow_unused=True)
        H = torch.zeros(n, n, dtype=z.dtype) if H is None else H
    else:
        H = torch.zeros(n, n, dtype=z.dtype)
    g0, l = g[0].detach(), val[0].detach()
    if terminal:
        return l, g0, None, H.detach(), None, None
    return l, g0[:nz], g0[nz:], H[:nz, :nz].detach(), H[nz:, :nz].detach(), H[nz:, nz:].detach()


def dynamics_derivatives(dyn, z, u, enc, carry, i=None):
    """z' and d z'/d[z,u] via nz replicated rows (ref: utils/evaluation.py:242-288)."""
    nz = z.shape[-1]
    zu = torch.cat([z, u], -1).detach()
    rows = zu.expand(nz, -1).clone().requires_grad_()
    c = None if carry is None else carry[:1].expand(nz, -1, -1)
    zn, carry = step_fn(dyn)(dyn, rows[:, :nz], rows[:, nz:], enc, c, i=i)
    J, = torch.autograd.grad(zn, rows, torch.eye(nz, dtype=z.dtype))
    return zn[0].detach(), J[:, :nz], J[:, nz:], (None if carry is None else carry[:1])


def linearize(z0, U, dyn, cost, enc, u_min=None, u_max=None):
    """Nominal rollout with derivatives.  ref: controllers/ilqr.py:393-486."""
    N, nu = U.shape
    nz = z0.shape[-1]
    o = dict(dtype=z0.dtype)
    Z = torch.empty(N + 1, nz, **o)
    F_z, F_u = torch.empty(N, nz, nz, **o), torch.empty(N, nz, nu, **o)
    L, L_z, L_u = torch.empty(N + 1, **o), torch.empty(N + 1, nz, **o), torch.empty(N, nu, **o)
    L_zz = torch.empty(N + 1, nz, nz, **o)
    L_uz, L_uu = torch.empty(N, nu, nz, **o), torch.empty(N, nu, nu, **o)
    Z[0] = z0
    carry = None
    for t in range(N):
        u = U[t] if u_min is None or u_max is None else clamp(U[t], u_min, u_max)
        L[t], L_z[t], L_u[t], L_zz[t], L_uz[t], L_uu[t] = cost_derivatives(cost, Z[t], u, False,
                                                                           enc)
        Z[t + 1], F_z[t], F_u[t], carry = dynamics_derivatives(dyn, Z[t], u, enc, carry, i=t)
    L[N], L_z[N], _, L_zz[N], _, _ = cost_derivatives(cost, Z[N], None, True, enc)
    return Z, F_z, F_u, L, L_z, L_u, L_zz, L_uz, L_uu


# --------------------------------------------------------------------------------------
# Backward Riccati pass (ref: controllers/ilqr.py:489-674, default V_zz_reg=False branch)
# --------------------------------------------------------------------------------------
def q_terms(F_z, F_u, L_z, L_u, L_zz, L_uz, L_uu, V_z, V_zz):
    """ref: controllers/ilqr.py:489-526"""
    Q_z = L_z + F_z.mT @ V_z
    Q_u = L_u + F_u.mT @ V_z
    Q_zz = L_zz + F_z.mT @ V_zz @ F_z
    Q_uz = L_uz + F_u.mT @ V_zz @ F_z
    Q_uu = L_uu + F_u.mT @ V_zz @ F_u
    return Q_z, Q_u, 0.5 * (Q_zz + Q_zz.mT), Q_uz, 0.5 * (Q_uu + Q_uu.mT)


def boxqp(x0, Q, c, lower, upper, max_iter=100, min_grad=1e-8, tol=1e-8, step_dec=0.6,
          min_step=1e-22, armijo=0.1):
    """Projected-Newton box QP (Tassa).  ref: utils/constraint.py:150-266.
    Returns (x, result, U_free, free_mask)."""
    n = Q.shape[0]
    clamped = torch.zeros(n, dtype=torch.bool)
    free = ~clamped
    Ufree = torch.zeros(n, dtype=Q.dtype)
    x = clamp(x0, lower, upper).clone()
    x[torch.isinf(x)] = 0.0
    f = 0.5 * x @ Q @ x + x @ c
    result, old_f = 0, 0
    for it in range(max_iter):
        if result != 0:
            break
        if it > 0 and (old_f - f) < tol * abs(old_f):
            result = 4
            break
        old_f = f
        g = Q @ x + c
        prev = clamped
        clamped = ((x == lower) & (g > 0)) | ((x == upper) & (g < 0))
        free = ~clamped
        if bool(clamped.all()):
            result = 6
            break
        if it == 0 or bool((prev != clamped).any()):
            Lf, info = torch.linalg.cholesky_ex(Q[free][:, free])
            if int(info) != 0:
                result = -1
                break
            Ufree = Lf.mT
        if g[free].norm() < min_grad:
            result = 5
            break
        g_cl = Q @ (x * clamped.to(x.dtype)) + c
        search = torch.zeros_like(x)
        search[free] = -torch.cholesky_solve(g_cl[free].unsqueeze(-1), Ufree,
                                             upper=True).flatten() - x[free]
        sdotg = (search * g).sum()
        step = 1.0
        xc = clamp(x + step * search, lower, upper)
        fc = 0.5 * xc @ Q @ xc + xc @ c
        while (fc - old_f) / (step * sdotg) < armijo:
            step *= step_dec
            xc = clamp(x + step * search, lower, upper)
            fc = 0.5 * xc @ Q @ xc + xc @ c
            if step < min_step:
                result = 2
                break
        x, f = xc, fc
    return x, result, Ufree, free


def backward_pass(Z, F_z, F_u, L, L_z, L_u, L_zz, L_uz, L_uu, reg=0.0, u_min=None, u_max=None,
                  U=None, symmetric_eig=False):
    """ref: controllers/ilqr.py:626-672.  Eigen-clip Q_uu (e<0 -> 1e-12), add reg; gains from the
    regularised inverse, value update with the UN-regularised Q_uu.  Raises NotPositiveDefinite
    where the reference raises RuntimeError."""
    N, nu = L_u.shape
    nz = Z.shape[1]
    V_z, V_zz = L_z[-1], L_zz[-1]
    k = torch.zeros(N, nu, dtype=Z.dtype)
    K = torch.zeros(N, nu, nz, dtype=Z.dtype)
    for t in range(N - 1, -1, -1):
        Q_z, Q_u, Q_zz, Q_uz, Q_uu = q_terms(F_z[t], F_u[t], L_z[t], L_u[t], L_zz[t], L_uz[t],
                                             L_uu[t], V_z, V_zz)
        if not torch.isfinite(Q_uu).all():
            raise NotPositiveDefinite("Q_uu has NaN/Inf")   # torch.linalg.eig raises (shim)
        if symmetric_eig:
            # orthonormal eigenvectors.  The reference's general `eig` (LAPACK geev) returns
            # non-orthogonal vectors for a REPEATED eigenvalue, and (E/e)E^T is then a rounding-
            # dependent matrix instead of the regularised inverse (stock RendezvousCost: Q_uu is a
            # multiple of I).  Where the eigenvalues are distinct the two agree.
            e, E = torch.linalg.eigh(Q_uu)
            e = e.clone()
        else:
            w, E = torch.linalg.eig(Q_uu)
            e, E = w.real.clone(), E.real
        e[e < 0] = 1e-12
        e = e + reg
        if u_min is None or u_max is None:
            kK = -((E / e) @ E.mT) @ torch.cat([Q_u.unsqueeze(1), Q_uz], -1)
            if torch.isnan(kK).any():
                raise NotPositiveDefinite("non-positive definite matrix")
            k[t], K[t] = kK[:, 0], kK[:, 1:]
        else:
            Q_uu_reg = (E * e) @ E.mT
            warm = k[t + 1] if t < N - 1 else k[-1]
            k[t], result, Uf, free = boxqp(warm, Q_uu_reg, Q_u, u_min - U[t], u_max - U[t])
            if result < 1:
                raise NotPositiveDefinite("boxqp failed: %d" % result)
            if bool(free.any()):
                idx = free.nonzero().flatten()
                K[t, idx] = -torch.cholesky_solve(Q_uz[idx], Uf, upper=True)
        V_z = Q_z + K[t].mT @ Q_u + K[t].mT @ Q_uu @ k[t] + Q_uz.mT @ k[t]
        V_zz = Q_zz + K[t].mT @ Q_uu @ K[t] + K[t].mT @ Q_uz + Q_uz.mT @ K[t]
        V_zz = 0.5 * (V_zz + V_zz.mT)
    return k, K


# --------------------------------------------------------------------------------------
# Rollout with parallel line search + trajectory cost (ref: controllers/ilqr.py:677-723, 764-791)
# --------------------------------------------------------------------------------------
@torch.no_grad()
def rollout(dyn, Z, U, k, K, alphas, enc, u_min=None, u_max=None):
    """All alphas at once: Z_new:[N+1,A,nz], U_new:[N,A,nu]."""
    N, A = U.shape[0], alphas.numel()
    Z_new = torch.empty(N + 1, A, Z.shape[1], dtype=Z.dtype)
    U_new = torch.empty(N, A, U.shape[1], dtype=Z.dtype)
    Z_new[0] = Z[0]
    al = alphas.reshape(-1, 1).to(Z.dtype)
    carry = None
    f = step_fn(dyn)
    for t in range(N):
        du = al * k[t] + (Z_new[t] - Z[t]) @ K[t].mT
        u = U[t] + du
        if u_min is not None and u_max is not None:
            u = clamp(u, u_min, u_max)
        Z_new[t + 1], carry = f(dyn, Z_new[t], u, enc, carry, i=t)
        U_new[t] = u
    return Z_new, U_new


@torch.no_grad()
def trajectory_cost(cost, Z, U, enc):
    """J[a] = sum_t l(Z[t,a],U[t,a]) + l_f(Z[N,a]);  Z:[N+1,(A,)nz]."""
    run = cost_value(cost, Z[:-1], U, False, enc)
    return run.sum(0) + cost_value(cost, Z[-1], None, True, enc)


# --------------------------------------------------------------------------------------
# Controller state machine (ref: controllers/ilqr.py:102-316, 364-390)
# --------------------------------------------------------------------------------------
def fit_alphas(dtype=torch.float32, n=10):
    """ref: controllers/ilqr.py:282  (1.025 ** -(j^2), j < 10)"""
    return (1.025 ** (-torch.arange(float(n)) ** 2)).to(dtype)


class Regulariser:
    """Tassa schedule.  ref: controllers/ilqr.py:364-390"""

    def __init__(self):
        self.mu_min, self.delta0 = 1e-6, 2.0
        self.reset()

    def reset(self):
        self.mu, self.delta = 0.0, self.delta0

    def decrease(self):
        self.delta = min(1.0, self.delta) / self.delta0
        self.mu *= self.delta
        if self.mu <= self.mu_min:
            self.mu = 0.0

    def increase(self, max_reg):
        self.delta = max(1.0, self.delta) * self.delta0
        self.mu = max(self.mu_min, self.mu * self.delta)
        return self.mu < max_reg


class ILQR:
    """Single-problem iLQR/PDDP solver (the reference's unit of work)."""

    def __init__(self, dyn, cost, enc):
        self.dyn, self.cost, self.enc = dyn, cost, enc
        self.reg = Regulariser()
        self.Z = self.U = self.K = None

    def _try(self, lin, Z, U, J_opt, alphas, u_min, u_max, max_reg, tol):
        """ref: controllers/ilqr.py:102-181"""
        try:
            k, K = backward_pass(*lin, reg=self.reg.mu, u_min=u_min, u_max=u_max, U=U)
        except (NotPositiveDefinite, RuntimeError):
            return (NOT_PD if self.reg.increase(max_reg) else MAX_REG), Z, U, J_opt
        Zb, Ub = rollout(self.dyn, Z, U, k, K, alphas, self.enc, u_min, u_max)
        Jb = trajectory_cost(self.cost, Zb, Ub, self.enc)
        a = int(Jb.argmin())
        if Jb[a] < J_opt:
            self.Z, self.U, self.K = Zb[:, a].clone(), Ub[:, a].clone(), K
            self.reg.decrease()
            state = CONVERGED if (J_opt - Jb[a]).abs() / J_opt < tol else ACCEPTED
            return state, self.Z, self.U, Jb[a]
        return (REJECTED if self.reg.increase(max_reg) else MAX_REG), Z, U, J_opt

    def step(self, z0, U, alphas, u_min=None, u_max=None, max_reg=1e10, tol=5e-6, trace=None):
        """ref: controllers/ilqr.py:183-235 -- linearise ONCE, retry backward+rollout."""
        lin = linearize(z0, U, self.dyn, self.cost, self.enc, u_min, u_max)
        Z, J_opt = lin[0], lin[3].sum()
        state = UNDEFINED
        while state in (UNDEFINED, NOT_PD, REJECTED):
            state, Z, U, J_opt = self._try(lin, Z, U, J_opt, alphas, u_min, u_max, max_reg, tol)
            if trace is not None:
                trace.append((state, float(J_opt), self.reg.mu))
        return state

    def fit(self, z0, U, n_iterations=50, tol=5e-6, max_reg=1e10, u_min=None, u_max=None,
            alphas=None, trace=None):
        """ref: controllers/ilqr.py:237-316"""
        self.U = U.detach().clone()
        self.reg.reset()
        alphas = fit_alphas(U.dtype) if alphas is None else alphas
        state = UNDEFINED
        for _ in range(n_iterations):
            state = self.step(z0, self.U, alphas, u_min, u_max, max_reg, tol, trace)
            if state in (CONVERGED, MAX_REG):
                break
        return self.Z, self.U, state


# --------------------------------------------------------------------------------------
# BNN training (ref: models/bnn/modules.py:131-198 fit loop, 434-447 / 517-530 / 749-766 regulariser,
# models/bnn/losses.py:20-38 likelihood).  Restated with explicit forward / backward formulas (no autograd)
# so that it documents exactly what the device trainer computes; pinned by tests/golden/train_*.npz, which the
# reference's own fit() produced (oracle/make_golden_train.py).
# --------------------------------------------------------------------------------------
def bnn_train(p, X_, dX, batch_idx, noise, hidden, D, dropout, lr, reg_scale, reg=1.0, rate=0.5, temperature=0.1,
              X_mean=None, X_std_inv=None, dX_mean=None, dX_std=None, betas=(0.9, 0.999), eps=1e-8):
    """p: flat [W0|b0|W1|b1|W2|b2|logit_p0|logit_p1]; X_: [n, K0] augmented state + action; batch_idx [T, bs]
    (-1 = empty); noise [T, bs, H0+H1] uniforms.  Returns (p_final, grads of step 0, loss of every step)."""
    H0, H1 = hidden
    K0, OUT, n = X_.shape[1], 2 * D, X_.shape[0]
    sizes = [H0 * K0, H0, H1 * H0, H1, OUT * H1, OUT, 1, 1]
    p = p.clone()
    m, v, vmax = torch.zeros_like(p), torch.zeros_like(p), torch.zeros_like(p)
    losses, grads0 = [], None
    for it in range(batch_idx.shape[0]):
        W0, b0, W1, b1, W2, b2, lp0, lp1 = [t.reshape(s) for t, s in zip(
            p.split(sizes), [(H0, K0), (H0,), (H1, H0), (H1,), (OUT, H1), (OUT,), (), ()])]
        rows = batch_idx[it]
        live = rows >= 0
        rows, r = rows[live].long(), noise[it][live]
        nb = rows.numel()
        a0 = X_[rows] if X_mean is None else (X_[rows] - X_mean) * X_std_inv
        if dropout == 0:
            keep = [torch.sigmoid(lp0), torch.sigmoid(lp1)]
            mk = lambda u, lp: torch.sigmoid((lp + u.log() - (1 - u).log()) / temperature)
            m0, m1 = mk(r[:, :H0], lp0), mk(r[:, H0:], lp1)
        else:
            keep = [torch.tensor(1 - rate, dtype=p.dtype)] * 2
            m0, m1 = (r[:, :H0] < keep[0]).to(p.dtype), (r[:, H0:] < keep[1]).to(p.dtype)
        pre0 = a0 @ W0.T + b0
        h0 = torch.relu(pre0 * m0)
        pre1 = h0 @ W1.T + b1
        h1 = torch.relu(pre1 * m1)
        out = h1 @ W2.T + b2
        sc = torch.ones(D, dtype=p.dtype) if dX_std is None else dX_std
        sh = torch.zeros(D, dtype=p.dtype) if dX_mean is None else dX_mean
        mean, log_std = out[:, :D] * sc + sh, out[:, D:] + sc.log()
        q = (mean - dX[rows]) * (-log_std).exp()
        nll = 0.5 * (q ** 2).sum(-1) + log_std.sum(-1) + 0.5 * math.log(2 * math.pi)
        # the regulariser's keep-probability is ALWAYS 1 - rate: CDropout.regularization sets p.data = sigmoid(logit_p)
        # and then calls BDropout.regularization, whose first statement rebinds self.p = 1 - self.rate
        # (modules.py:443, 526-527), so the learned logit_p never reaches it
        pr = torch.tensor(1 - rate, dtype=p.dtype)
        regv = reg * (pr * (W1 ** 2).sum() + (b1 ** 2).sum()) + reg * (pr * (W2 ** 2).sum() + (b2 ** 2).sum())
        if dropout == 0:
            regv = regv - 2 * (-(1 - pr) * (1 - pr).log() - pr * pr.log())
        losses.append(nll.mean() + reg_scale * regv / n)
        # backward
        dout = torch.cat([q * (-log_std).exp() * sc, 1 - q ** 2], -1) / nb
        dz1 = (dout @ W2) * (pre1 * m1 > 0)
        dp1 = dz1 * m1
        dz0 = (dp1 @ W1) * (pre0 * m0 > 0)
        dp0 = dz0 * m0
        rs = reg_scale / n
        g = [dp0.T @ a0, dp0.sum(0), dp1.T @ h0 + rs * reg * pr * 2 * W1, dp1.sum(0) + rs * reg * 2 * b1,
             dout.T @ h1 + rs * reg * pr * 2 * W2, dout.sum(0) + rs * reg * 2 * b2]
        if dropout == 0:
            g += [(dz0 * pre0 * m0 * (1 - m0) / temperature).sum().reshape(1), (dz1 * pre1 * m1 * (1 - m1) / temperature).sum().reshape(1)]
        else:
            g += [torch.zeros(1, dtype=p.dtype)] * 2
        g = torch.cat([t.reshape(-1) for t in g])
        if grads0 is None:
            grads0 = g.clone()
        m = betas[0] * m + (1 - betas[0]) * g
        v = betas[1] * v + (1 - betas[1]) * g * g
        vmax = torch.maximum(vmax, v)
        step = (lr / (1 - betas[0] ** (it + 1))) * m / (vmax.sqrt() / math.sqrt(1 - betas[1] ** (it + 1)) + eps)
        if dropout != 0:
            step[-2:] = 0
        p = p - step
    return p, grads0, torch.stack(losses)
