"""GPU parity (through the C ABI) of the BNN path against the reference's golden outputs:
particle propagation + moment matching + forward-mode linearisation, and the line-search rollout.

The dropout masks and eps_in[0] in the fixtures are the ones the reference model object drew; they
are passed to the kernels as data, so mask / particle indexing is compared exactly (a permuted
mask row would change every output)."""
import pytest
import torch

from golden_util import Fixture, LIN_NAMES, all_tags, rel_err

pytestmark = pytest.mark.gpu
BNN = [t for t in all_tags() if t.startswith("bnn_")]


def tol(fx):
    return 1e-5 if fx.dtype == torch.float64 else 1e-3


@pytest.mark.parametrize("tag", BNN)
def test_linearize(tag):
    from gpu_util import solver_from_fixture, tile
    fx = Fixture(tag)
    B = 3
    s = solver_from_fixture(fx, B=B)
    lo, hi = fx.bounds
    s.set_problem(tile(fx.t("z0"), B), tile(fx.t("U"), B), lo, hi, alphas=fx.t("alphas"))
    s.linearize()
    torch.cuda.synchronize()
    assert s.lin_status.cpu().tolist() == [0] * B
    for name, want in zip(LIN_NAMES, fx.lin()):
        got = s.matrices(name).cpu()
        for b in range(B):
            assert rel_err(got[b].reshape(want.shape), want) <= tol(fx), (name, b)


@pytest.mark.parametrize("tag", BNN)
def test_backward_and_rollout(tag):
    from gpu_util import solver_from_fixture, tile
    fx = Fixture(tag)
    B = 2
    s = solver_from_fixture(fx, B=B)
    lo, hi = fx.bounds
    s.set_problem(tile(fx.t("z0"), B), tile(fx.t("U"), B), lo, hi, alphas=fx.t("alphas"))
    for name, want in zip(LIN_NAMES, fx.lin()):
        s.store(name, tile(want.reshape(want.shape[0], -1), B))
    s.mu.fill_(fx.reg)
    s.backward()
    torch.cuda.synchronize()
    assert s.bw_status.cpu().tolist() == [0] * B
    for b in range(B):
        assert rel_err(s.matrices("k").cpu()[b], fx.t("k")) <= tol(fx) * 10
        assert rel_err(s.matrices("K").cpu()[b], fx.t("K")) <= tol(fx) * 10
    s.store("k", tile(fx.t("k"), B))
    s.store("K", tile(fx.t("K").reshape(fx.N, -1), B))
    s.rollout()
    torch.cuda.synchronize()
    want_amin = int(fx.t("J").argmin())
    for b in range(B):
        assert rel_err(s.J_all.cpu()[b], fx.t("J")) <= tol(fx) * 10
        assert int(s.amin.cpu()[b]) == want_amin
        assert rel_err(s.view("Z_new").cpu()[b], fx.t("Z_new")[:, want_amin]) <= tol(fx) * 10
        assert rel_err(s.view("U_new").cpu()[b], fx.t("U_new")[:, want_amin]) <= tol(fx) * 10


def test_fit_matches_reference():
    from gpu_util import solver_from_fixture, tile
    fx = Fixture("bnn_cartpole_ut_small_f64")
    B = 2
    s = solver_from_fixture(fx, B=B)
    states = []
    Z, U, state = s.fit(tile(fx.t("z0"), B), tile(fx.t("U"), B), n_iterations=int(fx.raw["fit_iters"]),
                        on_pass=lambda i, sv: states.append(int(sv.state[0].item())))
    want = fx.raw["fit_trace"]
    assert states == [int(x) for x in want[:, 0]]
    assert rel_err(Z.cpu()[0], fx.t("fit_Z")) <= 1e-5
    assert rel_err(U.cpu()[0], fx.t("fit_U")) <= 1e-5


def test_mask_rows_are_per_particle():
    """Permuting the particle order of (mask0, mask1, eps0) together leaves z' unchanged
    (moments are permutation invariant); permuting only the masks changes it."""
    from gpu_util import solver_from_fixture, tile, dynamics_from_fixture, cost_from_fixture
    from pddp_b200.solver import BatchedSolver, BNNDynamics
    fx = Fixture("bnn_cartpole_ut_small_f64")
    base = solver_from_fixture(fx, B=1)
    base.set_problem(tile(fx.t("z0"), 1), tile(fx.t("U"), 1))
    base.linearize()
    perm = torch.randperm(int(fx.raw["P"]), generator=torch.Generator().manual_seed(0))
    d = dynamics_from_fixture(fx).tensors
    both = BNNDynamics(base.geo, [d["W0"], d["W1"], d["W2"]], [d["b0"], d["b1"], d["b2"]],
                       [d["mask0"][perm], d["mask1"][perm]], d["eps0"][perm])
    only_masks = BNNDynamics(base.geo, [d["W0"], d["W1"], d["W2"]], [d["b0"], d["b1"], d["b2"]],
                             [d["mask0"][perm], d["mask1"][perm]], d["eps0"])
    outs = []
    for dyn in (both, only_masks):
        s = BatchedSolver(dyn, cost_from_fixture(fx), fx.enc, 1, fx.N, dtype=fx.dtype)
        s.set_problem(tile(fx.t("z0"), 1), tile(fx.t("U"), 1))
        s.linearize()
        outs.append(s.view("Z").cpu())
    torch.cuda.synchronize()
    ref = base.view("Z").cpu()
    assert rel_err(outs[0], ref) <= 1e-12
    assert rel_err(outs[1], ref) > 1e-8
