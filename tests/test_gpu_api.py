"""The reference-shaped public API (pddp_b200.controllers / models / costs) on the GPU against the
reference's golden outputs -- these read like the reference's own tests/controllers/test_ilqr.py
but check VALUES, which the reference's tests never pinned."""
import pytest
import torch

from golden_util import Fixture, LIN_NAMES, rel_err

pytestmark = pytest.mark.gpu


def build(fx):
    import pddp_b200 as P
    models, costs, examples = P.models, P.costs, P.examples
    if fx.name == "rendezvous":     # the fixtures use RendezvousCost's Q with a non-degenerate R (oracle/make_golden.py)
        assert torch.equal(examples.rendezvous.RendezvousCost().Q.data.double(), fx.t("Q").double())
        cost = costs.QRCost(fx.t("Q"), fx.t("R"), state_size=8, angular_indices=()).to(fx.dtype)
    else:
        cost = {"pendulum": examples.pendulum.PendulumCost, "cartpole": examples.cartpole.CartpoleCost,
                "double_cartpole": examples.double_cartpole.DoubleCartpoleCost}[fx.name]().to(fx.dtype)
    if fx.is_bnn:
        ang = {"cartpole": ([2], [0, 1, 3]), "double_cartpole": ([2, 4], [0, 1, 3, 5])}[fx.name]
        hidden = [int(h) for h in fx.raw["hidden"]]
        model = models.bnn.bnn_dynamics_model_factory(fx.D, 1, hidden, *ang)(n_particles=int(fx.raw["P"])).to(fx.dtype)
        with torch.no_grad():
            for name, i in (("fc_0", 0), ("fc_1", 1), ("fc_out", 2)):
                getattr(model.model, name).weight.copy_(fx.t("W%d" % i))
                getattr(model.model, name).bias.copy_(fx.t("b%d" % i))
        model.model.drop_0.mask, model.model.drop_1.mask = fx.t("mask0"), fx.t("mask1")
        model.eps_in = {0: fx.t("eps0")}
        opts = {"use_predicted_std": False, "infer_noise_variables": True}
        if fx.input_mode == "resample":       # ref: modules.py:320-358 -- eps_in[i] of every step, as the reference drew them
            model.eps_in = {i: e for i, e in enumerate(fx.t("eps_in"))}
            opts["infer_noise_variables"] = False
        elif fx.input_mode == "mean":
            opts["sample_input_distribution"] = False
        if fx.has("eps_out"):                 # ref: modules.py:242-262
            model.eps_out = {i: e for i, e in enumerate(fx.t("eps_out"))}
            opts["use_predicted_std"] = True
            opts["independent_noise"] = bool(int(fx.raw["independent_noise"]))
    else:
        cls = {"pendulum": examples.pendulum.PendulumDynamicsModel, "cartpole": examples.cartpole.CartpoleDynamicsModel,
               "double_cartpole": examples.double_cartpole.DoubleCartpoleDynamicsModel,
               "rendezvous": examples.rendezvous.RendezvousDynamicsModel}[fx.name]
        model = cls(**fx.known_params()).to(fx.dtype)
        opts = {}
    return model, cost, opts


TAGS = ["known_pendulum_ign_f64", "known_cartpole_ut_bounded_f64", "known_double_cartpole_full_f64",
        "bnn_cartpole_ut_small_f64", "bnn_double_cartpole_full_small_f64", "bnn_cartpole_ut_bounded_f64",
        "bnn_cartpole_ut_resample_small_f64", "bnn_cartpole_ut_mean_small_f64",
        "known_rendezvous_ign_bounded_f64", "known_rendezvous_ut_f64",
        "bnn_cartpole_ut_pstd_small_f64", "bnn_cartpole_ut_pstd_indep_small_f64"]


@pytest.mark.parametrize("tag", TAGS)
def test_module_level_functions(tag):
    from pddp_b200 import controllers as C
    fx = Fixture(tag)
    model, cost, opts = build(fx)
    lo, hi = fx.bounds
    cu = lambda t: None if t is None else t.cuda()
    out = C.forward(cu(fx.t("z0")), cu(fx.t("U")), model, cost, fx.enc, True, opts, {}, u_min=cu(lo), u_max=cu(hi))
    for name, got, want in zip(LIN_NAMES, out, fx.lin()):
        assert got.shape == want.shape, name
        assert rel_err(got.cpu(), want) <= 1e-5, name
    k, K = C.backward(*[cu(t) for t in fx.lin()], reg=fx.reg, u_min=cu(lo), u_max=cu(hi), U=cu(fx.t("U")))
    assert k.shape == fx.t("k").shape and K.shape == fx.t("K").shape
    assert rel_err(k.cpu(), fx.t("k")) <= 1e-5 and rel_err(K.cpu(), fx.t("K")) <= 1e-5
    Zb, Ub = C._control_law(model, cu(fx.t("Z")), cu(fx.t("U")), cu(fx.t("k")), cu(fx.t("K")), cu(fx.t("alphas")),
                            fx.enc, opts, u_min=cu(lo), u_max=cu(hi))
    assert Zb.shape == fx.t("Z_new").shape and Ub.shape == fx.t("U_new").shape
    assert rel_err(Zb.cpu(), fx.t("Z_new")) <= 1e-5 and rel_err(Ub.cpu(), fx.t("U_new")) <= 1e-5
    J = C._trajectory_cost(cost, cu(fx.t("Z_new")), cu(fx.t("U_new")), fx.enc, {})
    assert rel_err(J.cpu(), fx.t("J")) <= 1e-5


def test_backward_raises_like_the_reference():
    from pddp_b200 import controllers as C
    fx = Fixture("known_pendulum_ign_f64")
    lin = [t.cuda() for t in fx.lin()]
    lin[8] = lin[8].clone()
    lin[8][2, 0, 0] = float("nan")
    with pytest.raises(RuntimeError):
        C.backward(*lin, reg=fx.reg)


@pytest.mark.parametrize("tag", ["known_pendulum_ign_f64", "known_pendulum_ign_bounded_f64",
                                 "bnn_cartpole_ut_small_f64", "known_rendezvous_ign_bounded_f64",
                                 "known_rendezvous_ign_f64"])
def test_controller_fit_single_problem(tag):
    """controller.fit(U) with the reference call shape: env supplies the state, callbacks fire
    once per attempt with (iteration, state, Z, U, J_opt)."""
    import pddp_b200 as P
    fx = Fixture(tag)
    model, cost, opts = build(fx)
    lo, hi = fx.bounds

    class Env:
        def get_state(self):
            return P.GaussianVariable(fx.t("z0")[:fx.D], var=1e-2 * torch.ones(fx.D, dtype=fx.dtype))

    ctrl = P.iLQRController(Env(), model, cost, model_opts=opts)
    seen = []
    Z, U, state = ctrl.fit(fx.t("U").cuda(), encoding=fx.enc, n_iterations=int(fx.raw["fit_iters"]), quiet=True,
                           u_min=None if lo is None else lo.cuda(), u_max=None if hi is None else hi.cuda(),
                           on_iteration=lambda i, s, Z_, U_, J: seen.append((int(s), float(J), ctrl._mu)))
    want = fx.raw["fit_trace"]
    assert isinstance(state, P.iLQRState) and int(state) == int(fx.raw["fit_state"])
    assert [s[0] for s in seen] == [int(x) for x in want[:, 0]]
    for (s, J, mu), w in zip(seen, want):
        assert abs(J - w[1]) <= 1e-4 * max(1.0, abs(w[1]))
        assert abs(mu - w[2]) <= 1e-12 * max(1.0, abs(w[2]))
    assert Z.shape == fx.t("fit_Z").shape and U.shape == fx.t("fit_U").shape
    assert rel_err(Z.cpu(), fx.t("fit_Z")) <= 1e-2 and rel_err(U.cpu(), fx.t("fit_U")) <= 1e-2
    # feedback law of the fitted controller (ref: ilqr.py:339-354)
    u = ctrl(Z[3], 3, fx.enc)
    assert torch.allclose(u, U[3])


def test_batched_fit_and_mpc_step():
    """[B, N, nu] controls: every problem keeps its own regularisation / state; an MPC call returns
    the first action and shifts the nominal controls (ref: ilqr.py:355-362)."""
    import pddp_b200 as P
    fx = Fixture("known_pendulum_ign_f64")
    model, cost, opts = build(fx)
    B = 5
    g = torch.Generator().manual_seed(0)
    z0 = fx.t("z0").unsqueeze(0) + 0.05 * torch.randn(B, 2, generator=g, dtype=fx.dtype)
    U = fx.t("U").unsqueeze(0) + 0.05 * torch.randn(B, fx.N, 1, generator=g, dtype=fx.dtype)
    z0[0], U[0] = fx.t("z0"), fx.t("U")
    ctrl = P.iLQRController(None, model, cost)
    Z, Uo, state = ctrl.fit(U.cuda(), encoding=fx.enc, n_iterations=int(fx.raw["fit_iters"]), z0=z0.cuda())
    assert Z.shape == (B, fx.N + 1, 2) and state.shape == (B,)
    assert int(state[0]) == int(fx.raw["fit_state"])
    assert rel_err(Z.cpu()[0], fx.t("fit_Z")) <= 1e-2
    before = ctrl._U_nominal.clone()
    u = ctrl(z0.cuda(), 0, fx.enc, mpc=True)
    assert u.shape == (B, 1)
    assert torch.equal(ctrl._U_nominal[:, -1], ctrl._U_nominal[:, -2])
    assert ctrl._U_nominal.shape == before.shape


def test_model_and_cost_forward():
    """model(z, u, i, encoding) and cost(z, u, i, terminal, encoding) evaluate on the device and
    match the reference trajectory / cost path step by step."""
    fx = Fixture("known_cartpole_ut_f64")
    model, cost, _ = build(fx)
    Z, U, L = fx.t("Z").cuda(), fx.t("U").cuda(), fx.t("L")
    zn = model(Z[:-1], U, 0, fx.enc)
    assert rel_err(zn.cpu(), fx.t("Z")[1:]) <= 1e-9
    l = cost(Z[:-1], U, 0, terminal=False, encoding=fx.enc)
    lt = cost(Z[-1], None, 0, terminal=True, encoding=fx.enc)
    assert rel_err(l.cpu(), L[:-1]) <= 1e-9 and rel_err(lt.cpu(), L[-1]) <= 1e-9
