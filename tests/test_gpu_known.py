"""GPU parity (through the C ABI) of the known-dynamics path against the reference's golden
outputs: linearise, backward, rollout + line search, and whole `fit` runs.

Tolerances are the north-star ones: 1e-5 relative in fp64, 1e-3 relative in fp32 (relative to the
tensor's scale, see golden_util.rel_err)."""
import pytest
import torch

from golden_util import Fixture, LIN_NAMES, all_tags, rel_err

pytestmark = pytest.mark.gpu
KNOWN = [t for t in all_tags() if t.startswith("known_")]


def tol(fx):
    return 1e-5 if fx.dtype == torch.float64 else 1e-3


def layouts_for(fx):
    from pddp_b200 import _lib
    return [_lib.PROBLEM_MAJOR, _lib.BATCH_INNER]


@pytest.mark.parametrize("tag", KNOWN)
@pytest.mark.parametrize("layout", [0, 1])
def test_linearize_backward_rollout(tag, layout):
    from gpu_util import solver_from_fixture, tile
    fx = Fixture(tag)
    B = 3
    s = solver_from_fixture(fx, B=B, layout=layout)
    lo, hi = fx.bounds
    s.set_problem(tile(fx.t("z0"), B), tile(fx.t("U"), B), lo, hi, alphas=fx.t("alphas"))
    s.linearize()
    torch.cuda.synchronize()
    for name, want in zip(LIN_NAMES, fx.lin()):
        got = s.matrices(name).cpu()
        for b in range(B):
            assert rel_err(got[b].reshape(want.shape), want) <= tol(fx), (name, b)
    assert rel_err(s.J_opt.cpu()[0], fx.t("L").sum()) <= tol(fx)

    # backward on the REFERENCE's linearisation (isolates the kernel under test)
    for name, want in zip(LIN_NAMES, fx.lin()):
        s.store(name, tile(want.reshape(want.shape[0], -1), B))
    s.mu.fill_(fx.reg)
    s.backward()
    torch.cuda.synchronize()
    assert s.bw_status.cpu().tolist() == [0] * B
    k, K = s.matrices("k").cpu(), s.matrices("K").cpu()
    for b in range(B):
        assert rel_err(k[b], fx.t("k")) <= tol(fx) * 10
        assert rel_err(K[b], fx.t("K")) <= tol(fx) * 10

    # rollout + line search from the reference's gains
    s.store("k", tile(fx.t("k"), B))
    s.store("K", tile(fx.t("K").reshape(fx.N, -1), B))
    s.rollout()
    torch.cuda.synchronize()
    J = s.J_all.cpu()[:, :fx.t("J").numel()]
    want_amin = int(fx.t("J").argmin())
    for b in range(B):
        assert rel_err(J[b], fx.t("J")) <= tol(fx) * 10
        assert int(s.amin.cpu()[b]) == want_amin
        assert rel_err(s.view("Z_new").cpu()[b], fx.t("Z_new")[:, want_amin]) <= tol(fx) * 10
        assert rel_err(s.view("U_new").cpu()[b], fx.t("U_new")[:, want_amin]) <= tol(fx) * 10


@pytest.mark.parametrize("tag", [t for t in KNOWN if Fixture(t).has("fit_trace")])
def test_fit_matches_reference(tag):
    """Per-problem state machine: same final state, trajectory and controls as the reference's
    controller.fit() (which retries NOT_PD / REJECTED steps with larger regularisation)."""
    from gpu_util import solver_from_fixture, tile
    fx = Fixture(tag)
    B = 2
    s = solver_from_fixture(fx, B=B)
    lo, hi = fx.bounds
    states = []
    Z, U, state = s.fit(tile(fx.t("z0"), B), tile(fx.t("U"), B), n_iterations=int(fx.raw["fit_iters"]),
                        u_min=lo, u_max=hi,
                        on_pass=lambda i, sv: states.append(int(sv.state[0].item())))
    want = fx.raw["fit_trace"]
    assert states == [int(x) for x in want[:, 0]]
    assert state.cpu().tolist() == [int(fx.raw["fit_state"])] * B
    loose = 1e3 if fx.dtype == torch.float64 else 10     # iterated: errors compound
    assert rel_err(Z.cpu()[0], fx.t("fit_Z")) <= tol(fx) * loose
    assert rel_err(U.cpu()[0], fx.t("fit_U")) <= tol(fx) * loose
    assert abs(float(s.mu[0]) - want[-1, 2]) <= 1e-12 * max(1.0, want[-1, 2])


def test_not_pd_is_reported_per_problem():
    """A problem whose Q_uu goes non-finite reports PDDP_STATUS_NOT_PD while its neighbours in
    the batch are unaffected (the reference raises RuntimeError, ilqr.py:639-640)."""
    from gpu_util import solver_from_fixture, tile
    fx = Fixture("known_pendulum_ign_f64")
    B = 4
    s = solver_from_fixture(fx, B=B)
    s.set_problem(tile(fx.t("z0"), B), tile(fx.t("U"), B))
    for name, want in zip(LIN_NAMES, fx.lin()):
        s.store(name, tile(want.reshape(want.shape[0], -1), B))
    s.view("L_uu")[2, 3, 0] = float("nan")
    s.mu.fill_(fx.reg)
    s.backward()
    torch.cuda.synchronize()
    assert s.bw_status.cpu().tolist() == [0, 0, 1, 0]
    assert rel_err(s.matrices("K").cpu()[3], fx.t("K")) <= 1e-9


def test_loud_failures():
    from pddp_b200.solver import BatchedSolver, KnownDynamics, QRCostConstants
    from pddp_b200 import _lib
    import pddp_b200
    cost = QRCostConstants(torch.eye(3), torch.eye(1), torch.eye(3), torch.zeros(3))
    with pytest.raises(RuntimeError):
        BatchedSolver(KnownDynamics(_lib.GEO_PENDULUM, [0.1, 1, 1, 0.1, 9.8]), cost, 4, 1, 5,
                      device="cpu")
    # an unsupported configuration fails loudly through the C ABI's error code: hidden width > 256
    from pddp_b200.solver import BNNDynamics
    H, P = 300, 4
    dyn = BNNDynamics(_lib.GEO_PENDULUM, [torch.zeros(H, 4), torch.zeros(H, H), torch.zeros(4, H)],
                      [torch.zeros(H), torch.zeros(H), torch.zeros(4)], [torch.ones(P, H), torch.ones(P, H)],
                      torch.zeros(P, 2))
    with pytest.raises((RuntimeError, ValueError), match="hidden widths"):
        s = BatchedSolver(dyn, cost, 4, 1, 5)
        s.set_problem(torch.zeros(1, 2).cuda(), torch.zeros(1, 5, 1).cuda())
        s.linearize()
