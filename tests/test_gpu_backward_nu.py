"""GPU parity of the general-action-size backward pass (pddp_backward with nu > 1: eigen-clipping by
Jacobi rotations, n-dimensional projected-Newton box QP) against the reference's golden gains, and of
the same kernel forced onto the nu == 1 fixtures (PDDP_FORCE_BACKWARD_NU=1) against the reference and
the specialised nu == 1 kernels.

ref: pddp/controllers/ilqr.py:626-672, pddp/utils/constraint.py:150-266."""
import os
import subprocess
import sys

import pytest
import torch

import pddp_oracle as O
from golden_util import Fixture, all_tags, rel_err

pytestmark = pytest.mark.gpu
RDV = [t for t in all_tags() if t.startswith("known_rendezvous")]


def tol(fx):
    return 1e-5 if fx.dtype == torch.float64 else 1e-3


@pytest.mark.parametrize("tag", RDV)
def test_backward_matches_reference(tag):
    from pddp_b200 import controllers as C
    fx = Fixture(tag)
    lo, hi = fx.bounds
    cu = lambda t: None if t is None else t.cuda()
    k, K = C.backward(*[cu(t) for t in fx.lin()], reg=fx.reg, u_min=cu(lo), u_max=cu(hi), U=cu(fx.t("U")))
    assert k.shape == fx.t("k").shape and K.shape == fx.t("K").shape
    assert rel_err(k.cpu(), fx.t("k")) <= tol(fx) * 10
    assert rel_err(K.cpu(), fx.t("K")) <= tol(fx) * 10


@pytest.mark.parametrize("nz,nu,bounded", [(5, 2, False), (5, 2, True), (9, 3, True), (14, 4, False), (30, 4, True),
                                           (30, 2, False)])
def test_random_problems_against_the_oracle(nz, nu, bounded):
    """Random well-conditioned linear-quadratic data with distinct Q_uu eigenvalues, tight bounds so
    that the box QP clamps a varying subset of the controls; B problems in one launch."""
    from pddp_b200 import _lib
    from pddp_b200.solver import BatchedSolver  # noqa: F401  (loads the library)
    import ctypes as C
    g = torch.Generator().manual_seed(100 * nz + nu)
    B, N, dt = 5, 7, torch.float64
    r = lambda *s: torch.randn(*s, generator=g, dtype=dt)
    F_z = torch.eye(nz, dtype=dt) + 0.1 * r(B, N, nz, nz)
    F_u = 0.5 * r(B, N, nz, nu)
    sym = lambda M: M @ M.mT
    L_zz = sym(r(B, N + 1, nz, nz)) / nz + 0.1 * torch.eye(nz, dtype=dt)
    L_uu = sym(r(B, N, nu, nu)) + torch.diag(torch.arange(1, nu + 1, dtype=dt))
    L_uz = 0.05 * r(B, N, nu, nz)
    L_z, L_u, U = r(B, N + 1, nz), r(B, N, nu), 0.3 * r(B, N, nu)
    lo = -0.4 * torch.ones(nu, dtype=dt) if bounded else None
    hi = 0.5 * torch.ones(nu, dtype=dt) if bounded else None
    reg = 0.3
    shape = _lib.Shape(_lib.F64, _lib.PROBLEM_MAJOR, 0, 0, B, N, nz, nu)
    dev = lambda t: None if t is None else t.contiguous().cuda()
    bufs = [dev(t.reshape(B, t.shape[1], -1)) for t in (F_z, F_u, L_z, L_u, L_zz, L_uz, L_uu)]
    mu = torch.full((B,), reg, dtype=torch.float64, device="cuda")
    k = torch.zeros(B, N, nu, dtype=dt, device="cuda")
    K = torch.zeros(B, N, nu * nz, dtype=dt, device="cuda")
    status = torch.full((B,), -1, dtype=torch.int32, device="cuda")
    Ud, lod, hid = dev(U), dev(lo), dev(hi)
    p = _lib.ptr
    _lib.check(_lib.load().pddp_backward(C.byref(shape), *[p(t) for t in bufs], p(mu), p(Ud), p(lod), p(hid), None,
                                         p(k), p(K), p(status), _lib.stream_ptr()), "backward")
    torch.cuda.synchronize()
    assert status.cpu().tolist() == [0] * B
    clamped_rows = 0
    for b in range(B):
        Z = torch.zeros(N + 1, nz, dtype=dt)
        ok_, oK = O.backward_pass(Z, F_z[b], F_u[b], None, L_z[b], L_u[b], L_zz[b], L_uz[b], L_uu[b], reg=reg,
                                  u_min=lo, u_max=hi, U=U[b])
        assert rel_err(k[b].cpu(), ok_) <= 1e-7, b
        assert rel_err(K[b].cpu().reshape(N, nu, nz), oK) <= 1e-7, b
        clamped_rows += int((oK.abs().sum(-1) == 0).sum())
    if bounded:
        assert clamped_rows > 0          # the box QP really clamped something


def test_general_kernel_on_the_scalar_fixtures():
    """PDDP_FORCE_BACKWARD_NU=1 routes nu == 1 through the general kernel: it must reproduce the reference's
    gains on every nu == 1 fixture (run in a subprocess: the switch is read once per process)."""
    code = r'''
import sys, torch
sys.path.insert(0, "tests"); sys.path.insert(0, "oracle")
from golden_util import Fixture, all_tags, rel_err
from pddp_b200 import controllers as C
n = 0
for tag in all_tags():
    fx = Fixture(tag)
    if fx.nu != 1 or fx.dtype != torch.float64:
        continue
    lo, hi = fx.bounds
    cu = lambda t: None if t is None else t.cuda()
    k, K = C.backward(*[cu(t) for t in fx.lin()], reg=fx.reg, u_min=cu(lo), u_max=cu(hi), U=cu(fx.t("U")))
    assert rel_err(k.cpu(), fx.t("k")) <= 1e-4 and rel_err(K.cpu(), fx.t("K")) <= 1e-4, tag
    n += 1
assert n >= 20
print("ok", n)
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PDDP_FORCE_BACKWARD_NU="1", PYTHONPATH=root)
    out = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr


def test_stock_rendezvous_cost_degenerate_quu():
    """RendezvousCost as shipped (R = 0.1 I) makes Q_uu a multiple of the identity: the reference's
    (E/e)E^T with LAPACK's non-orthogonal eigenvectors of a repeated eigenvalue is rounding noise, the
    kernel returns the regularised inverse (oracle with orthonormal eigenvectors).  Whole API path:
    examples.rendezvous.RendezvousDynamicsModel + examples.rendezvous.RendezvousCost -> forward / backward."""
    import pddp_b200 as P
    from pddp_b200 import controllers as C
    torch.manual_seed(3)
    N, enc = 9, O.IGNORE_UNCERTAINTY
    model, cost = P.examples.rendezvous.RendezvousDynamicsModel(0.1).double(), P.examples.rendezvous.RendezvousCost().double()
    z0 = torch.tensor([-10.0, -10.0, 10.0, 10.0, 0.0, -5.0, 5.0, 0.0], dtype=torch.float64)
    U = 0.1 * torch.randn(N, 4, dtype=torch.float64)
    lin = C.forward(z0.cuda(), U.cuda(), model, cost, enc)
    ospec = O.QRCostSpec(cost.Q.data, cost.R.data, cost.Q_term.data, cost.x_goal.data, cost.u_goal.data, 8, (),
                         tuple(range(8)))
    # (the model's constants are fp32 Parameters widened to fp64, exactly like the reference's)
    odyn = O.rendezvous_spec(float(model.dt), float(model.m), float(model.alpha))
    olin = O.linearize(z0, U, odyn, ospec, enc)
    for got, want in zip(lin, olin):
        assert rel_err(got.cpu(), want) <= 1e-9
    k, K = C.backward(*lin, reg=0.5)
    ok_, oK = O.backward_pass(*olin, reg=0.5, symmetric_eig=True)
    assert rel_err(k.cpu(), ok_) <= 1e-9 and rel_err(K.cpu(), oK) <= 1e-9
