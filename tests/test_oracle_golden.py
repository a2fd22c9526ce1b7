"""The oracle (oracle/pddp_oracle.py) replayed against the committed reference outputs.

The fixtures were produced by the UNMODIFIED reference (oracle/make_golden.py); this is the pin
that makes the oracle trustworthy on machines where /root/reference does not exist."""
import pytest
import torch

import pddp_oracle as O
from golden_util import Fixture, LIN_NAMES, all_tags, rel_err

TAGS = all_tags()


def tol(fx, loose=1.0):
    return (1e-9 if fx.dtype == torch.float64 else 2e-4) * loose


def test_fixtures_exist():
    assert len(TAGS) >= 18


@pytest.mark.parametrize("tag", TAGS)
def test_linearize(tag):
    fx = Fixture(tag)
    lo, hi = fx.bounds
    out = O.linearize(fx.t("z0"), fx.t("U"), fx.dyn_spec(), fx.cost_spec(), fx.enc, lo, hi)
    for name, got, want in zip(LIN_NAMES, out, fx.lin()):
        assert rel_err(got, want) <= tol(fx), name


@pytest.mark.parametrize("tag", TAGS)
def test_backward_rollout_cost(tag):
    fx = Fixture(tag)
    lo, hi = fx.bounds
    k, K = O.backward_pass(*fx.lin(), reg=fx.reg, u_min=lo, u_max=hi, U=fx.t("U"))
    assert rel_err(k, fx.t("k")) <= tol(fx, 10)
    assert rel_err(K, fx.t("K")) <= tol(fx, 10)
    Zb, Ub = O.rollout(fx.dyn_spec(), fx.t("Z"), fx.t("U"), fx.t("k"), fx.t("K"), fx.t("alphas"),
                       fx.enc, lo, hi)
    assert rel_err(Zb, fx.t("Z_new")) <= tol(fx, 10)
    assert rel_err(Ub, fx.t("U_new")) <= tol(fx, 10)
    J = O.trajectory_cost(fx.cost_spec(), fx.t("Z_new"), fx.t("U_new"), fx.enc)
    assert rel_err(J, fx.t("J")) <= tol(fx, 10)


@pytest.mark.parametrize("tag", [t for t in TAGS if Fixture(t).has("fit_trace")])
def test_fit_state_machine(tag):
    """Same sequence of iLQRState transitions, regularisation values and accepted costs."""
    fx = Fixture(tag)
    lo, hi = fx.bounds
    trace = []
    solver = O.ILQR(fx.dyn_spec(), fx.cost_spec(), fx.enc)
    Z, U, state = solver.fit(fx.t("z0"), fx.t("U"), n_iterations=int(fx.raw["fit_iters"]),
                             u_min=lo, u_max=hi, trace=trace)
    want = fx.raw["fit_trace"]
    assert [t[0] for t in trace] == [int(s) for s in want[:, 0]]
    assert state == int(fx.raw["fit_state"])
    for (s, J, mu), w in zip(trace, want):
        assert abs(J - w[1]) <= tol(fx, 1e4) * max(1.0, abs(w[1]))  # iterated: errors compound
        assert abs(mu - w[2]) <= 1e-12 * max(1.0, abs(w[2]))
    assert rel_err(Z, fx.t("fit_Z")) <= tol(fx, 1e3)
    assert rel_err(U, fx.t("fit_U")) <= tol(fx, 1e3)
