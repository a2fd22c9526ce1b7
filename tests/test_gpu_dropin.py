"""The drop-in claim, tested the way a user of the reference would exercise it: the package is installed under
the name `pddp` (`pddp_b200.install_as("pddp")`) and code written against the REFERENCE's module paths runs on
CUDA tensors -- the flow of the reference's `examples/cartpole.py` (minus matplotlib / rendering), the shape
contract of `tests/controllers/test_ilqr.py::test_forward_backward`, and `tests/utils/test_evaluation.py`'s
batch == row-wise check, here additionally pinned to the oracle's autograd VALUES (the reference's tests only
compare the two spellings with each other)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pddp():
    import pddp_b200
    return pddp_b200.install_as("pddp")


@pytest.fixture()
def cuda_default():
    torch.set_default_device("cuda")
    yield
    torch.set_default_device(None)          # back to "no default-device mode" (not the same as "cpu")


def test_reference_module_paths_resolve(pddp):
    import importlib
    for path in ("pddp", "pddp.controllers", "pddp.controllers.ilqr", "pddp.controllers.pddp", "pddp.costs",
                 "pddp.costs.quadratic", "pddp.envs", "pddp.models", "pddp.models.bnn", "pddp.models.bnn.modules",
                 "pddp.models.bnn.losses", "pddp.utils", "pddp.utils.encoding", "pddp.utils.angular",
                 "pddp.utils.constraint", "pddp.utils.evaluation", "pddp.utils.gaussian_variable", "pddp.examples",
                 "pddp.examples.cartpole", "pddp.examples.pendulum", "pddp.examples.double_cartpole",
                 "pddp.examples.rendezvous"):
        assert importlib.import_module(path) is not None, path
    from pddp.controllers.ilqr import Q, StateEncoding, backward, forward, iLQRController  # noqa: F401
    from pddp.utils.encoding import decode_covar_sqrt, decode_mean, decode_std, decode_var  # noqa: F401
    from pddp.utils.evaluation import batch_eval_cost, batch_eval_dynamics, eval_cost, eval_dynamics  # noqa: F401
    from pddp.models.bnn import BDropout, CDropout, bayesian_model, bnn_dynamics_model_factory  # noqa: F401
    import pddp_b200
    assert pddp.controllers.ilqr is pddp_b200.controllers.ilqr          # aliases, not second copies
    assert pddp.StateEncoding.DEFAULT == 1 and pddp.GaussianVariable is pddp_b200.utils.gaussian_variable.GaussianVariable


def test_cartpole_example_script_flow(pddp, cuda_default):
    """ref: examples/cartpole.py:17-23, 119-181 -- same calls, same keyword arguments; smaller trial / iteration
    counts so the test stays short.  Everything (data collection on the device environment, BNN training,
    the iLQR iterations on the learned model, MPC trials, the final feedback rollout) runs on the GPU."""
    torch.manual_seed(0)
    N, DT = 25, 0.1
    ENCODING = pddp.StateEncoding.DEFAULT
    UMIN, UMAX = torch.tensor([-10.0]), torch.tensor([10.0])
    J_hist, trials = [], []

    def on_trial(trial, X, U):
        trials.append((trial, tuple(X.shape), tuple(U.shape)))

    def on_iteration(iteration, state, Z, U, J_opt):
        J_hist.append(float(J_opt))
        mean_ = pddp.utils.encoding.decode_mean(Z, ENCODING)
        std_ = pddp.utils.encoding.decode_std(Z, ENCODING)
        assert mean_.shape == (N + 1, 4) and std_.shape == (N + 1, 4)

    cost = pddp.examples.cartpole.CartpoleCost()
    env = pddp.examples.cartpole.CartpoleEnv(dt=DT, render=False)
    model_class = pddp.examples.cartpole.CartpoleDynamicsModel
    real_model = model_class(DT)
    model = pddp.models.bnn.bnn_dynamics_model_factory(
        env.state_size, env.action_size, [200, 200], model_class.angular_indices,
        model_class.non_angular_indices)(n_particles=100)
    U = (UMAX - UMIN) * torch.rand(N, model.action_size) + UMIN
    controller = pddp.controllers.PDDPController(
        env, model, cost,
        model_opts={"use_predicted_std": False, "infer_noise_variables": True},
        training_opts={"n_iter": 300, "learning_rate": 1e-3})
    controller.train()
    w_before = model.model.fc_1.weight.detach().clone()
    with _maybe_warns():
        Z, U, state = controller.fit(U, encoding=ENCODING, n_iterations=4, on_iteration=on_iteration,
                                     on_trial=on_trial, max_trials=4, u_min=UMIN, u_max=UMAX, quiet=True)
    assert Z.shape == (N + 1, 14) and U.shape == (N, 1) and isinstance(state, pddp.controllers.ilqr.iLQRState)
    assert Z.is_cuda and U.is_cuda and bool(torch.isfinite(U).all())
    assert bool((U >= UMIN).all()) and bool((U <= UMAX).all())
    assert len(trials) == 2 + 2 and trials[0][1] == (N, 4) and trials[-1][1] == (2 * N, 4)   # 2 initial + 2 MPC trials
    assert len(J_hist) >= 2 and all(math.isfinite(j) for j in J_hist)
    assert not torch.equal(w_before, model.model.fc_1.weight.detach())                         # the BNN was trained
    assert model.X_mean.shape == (6,) and model.dX_std.shape == (4,)                           # normalisation fitted
    # the ground-truth model under the optimised controls (examples/utils.py rollout)
    z, real_Z = Z[0], [Z[0]]
    for i in range(N):
        z = real_model(z, U[i], i, ENCODING)
        real_Z.append(z)
    assert torch.stack(real_Z).shape == Z.shape
    for i in range(N):                                   # examples/cartpole.py:173-176
        z = env.get_state().encode(ENCODING)
        u = controller(z, i, ENCODING)
        assert u.shape == (1,)
        env.apply(u)
    env.close()


class _maybe_warns:
    """fit may or may not warn 'exceeded max regularization term' on an untrained model: accept both."""

    def __enter__(self):
        import warnings
        self._cm = warnings.catch_warnings()
        self._cm.__enter__()
        warnings.simplefilter("ignore")

    def __exit__(self, *exc):
        return self._cm.__exit__(*exc)


PROBLEMS = ["cartpole", "pendulum", "rendezvous", "double_cartpole"]
CLASSES = {"cartpole": "Cartpole", "pendulum": "Pendulum", "rendezvous": "Rendezvous", "double_cartpole": "DoubleCartpole"}


def _setup(pddp, name, encoding, N, dtype=torch.float64):
    mod = getattr(pddp.examples, name)
    model = getattr(mod, CLASSES[name] + "DynamicsModel")(0.1).to(dtype)
    cost = getattr(mod, CLASSES[name] + "Cost")().to(dtype)
    z0 = pddp.utils.gaussian_variable.GaussianVariable.random(model.state_size, dtype=dtype).encode(encoding)
    U = torch.randn(N, model.action_size, requires_grad=True, dtype=dtype)
    return z0, U, model, cost


@pytest.mark.parametrize("N", [1, 3])
@pytest.mark.parametrize("encoding", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("name", PROBLEMS)
def test_forward_backward_shapes(pddp, cuda_default, name, encoding, N):
    """ref: tests/controllers/test_ilqr.py:49-92 (same calls, same shape assertions, same reg escalation)."""
    from pddp.controllers.ilqr import Q, backward, forward
    encoding = pddp.StateEncoding(encoding)
    torch.manual_seed(N * 10 + int(encoding))
    z0, U, model, cost = _setup(pddp, name, encoding, N)
    nz, nu = z0.shape[-1], model.action_size
    Z, F_z, F_u, L, L_z, L_u, L_zz, L_uz, L_uu = forward(z0, U, model, cost, encoding)
    assert Z.shape == (N + 1, nz) and F_z.shape == (N, nz, nz) and F_u.shape == (N, nz, nu)
    assert L.shape == torch.Size([N + 1]) and L_z.shape == (N + 1, nz) and L_u.shape == (N, nu)
    assert L_zz.shape == (N + 1, nz, nz) and L_uz.shape == (N, nu, nz) and L_uu.shape == (N, nu, nu)
    Q_z, Q_u, Q_zz, Q_uz, Q_uu = Q(F_z[0], F_u[0], L_z[0], L_u[0], L_zz[0], L_uz[0], L_uu[0], L_z[-1], L_zz[-1])
    assert Q_z.shape == (nz,) and Q_u.shape == (nu,) and Q_zz.shape == (nz, nz)
    assert Q_uz.shape == (nu, nz) and Q_uu.shape == (nu, nu)
    reg = 1.0
    while reg <= 1e10:
        try:
            k, K = backward(Z, F_z, F_u, L, L_z, L_u, L_zz, L_uz, L_uu, reg=reg)
            break
        except RuntimeError:
            reg *= 10
    assert k.shape == (N, nu) and K.shape == (N, nu, nz)


@pytest.mark.filterwarnings("ignore:exceeded max regularization term")
@pytest.mark.parametrize("name", PROBLEMS)
def test_fit_terminates(pddp, cuda_default, name):
    """ref: tests/controllers/test_ilqr.py:95-106"""
    torch.manual_seed(3)
    mod = getattr(pddp.examples, name)
    env = getattr(mod, CLASSES[name] + "Env")()
    _, U, model, cost = _setup(pddp, name, pddp.StateEncoding.DEFAULT, 3, torch.float32)
    controller = pddp.controllers.iLQRController(env, model, cost)
    Z, U, state = controller.fit(U, encoding=pddp.StateEncoding.DEFAULT, quiet=True)
    assert state.is_terminal()


def _oracle_specs(name, dtype):
    import pddp_oracle as O
    from golden_util import GEOMETRY, SPEC_FN
    import bench
    D, nu, ang, nonang = GEOMETRY[name]
    Q, R, Qt, goal = bench.cost_constants(name, dtype)
    if name in ("pendulum", "cartpole", "double_cartpole"):      # the example costs hold fp32 sin(pi) in x_goal
        import pddp_b200 as P
        goal = getattr(getattr(P.examples, name), CLASSES[name] + "Cost")().x_goal.data.to(dtype).cpu()
    cost = O.QRCostSpec(Q, R, Qt, goal, torch.zeros(nu, dtype=dtype), D, ang, nonang)
    return O, SPEC_FN[name](0.1), cost


@pytest.mark.parametrize("approximate", [False, True])
@pytest.mark.parametrize("terminal", [False, True])
@pytest.mark.parametrize("encoding", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("name", PROBLEMS)
def test_eval_cost(pddp, name, encoding, terminal, approximate):
    """ref: tests/utils/test_evaluation.py:33-87 + the oracle's autograd values (evaluation.py:23-94, 134-239)."""
    from pddp.utils.evaluation import batch_eval_cost, eval_cost
    dtype = torch.float64
    torch.manual_seed(encoding * 7 + terminal)
    O, _, ocost = _oracle_specs(name, dtype)
    mod = getattr(pddp.examples, name)
    model = getattr(mod, CLASSES[name] + "DynamicsModel")(0.1)
    cost = getattr(mod, CLASSES[name] + "Cost")().to(dtype)
    n, m = model.state_size, model.action_size
    z = pddp.GaussianVariable.random(n, dtype=dtype).encode(pddp.StateEncoding(encoding))
    u = None if terminal else torch.randn(m, dtype=dtype)
    cu = lambda t: None if t is None else t.cuda()
    got = eval_cost(cost, cu(z), cu(u), 0, terminal, pddp.StateEncoding(encoding), approximate)
    bat = batch_eval_cost(cost, cu(z), cu(u), 0, terminal, pddp.StateEncoding(encoding), approximate)
    l, l_z, l_u, l_zz, l_uz, l_uu = got
    assert l.shape == torch.Size([]) and l_z.shape == z.shape and l_zz.shape == (z.shape[0], z.shape[0])
    if terminal:
        assert l_u is None and l_uz is None and l_uu is None
    else:
        assert l_u.shape == u.shape and l_uz.shape == (m, z.shape[0]) and l_uu.shape == (m, m)
    want = O.cost_derivatives(ocost, z, u, terminal, encoding)
    if approximate:                                                # Gauss-Newton outer products (evaluation.py:72-78)
        wl, wz, wu = want[0], want[1], want[2]
        want = (wl, wz, wu, wz.view(-1, 1) @ wz.view(1, -1), None if terminal else wu.view(-1, 1) @ wz.view(1, -1),
                None if terminal else wu.view(-1, 1) @ wu.view(1, -1))
    for a, b, w in zip(got, bat, want):
        if a is None:
            continue
        assert torch.allclose(a, b, 1e-3, 1e-3)                    # the reference's own assertion
        scale = max(1.0, float(w.abs().max()))
        assert float((a.cpu() - w).abs().max()) <= 1e-5 * scale


@pytest.mark.parametrize("encoding", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("name", PROBLEMS)
def test_eval_dynamics(pddp, name, encoding):
    """ref: tests/utils/test_evaluation.py:90-114 + the oracle's autograd Jacobians (evaluation.py:97-131, 242-288)."""
    from pddp.utils.evaluation import batch_eval_dynamics, eval_dynamics
    dtype = torch.float64
    torch.manual_seed(encoding + 11)
    O, odyn, _ = _oracle_specs(name, dtype)
    mod = getattr(pddp.examples, name)
    model = getattr(mod, CLASSES[name] + "DynamicsModel")(0.1).to(dtype)
    x = pddp.GaussianVariable.random(model.state_size, dtype=dtype)
    u = torch.randn(model.action_size, dtype=dtype)
    z = x.encode(pddp.StateEncoding(encoding))
    z_next, d_dz, d_du = eval_dynamics(model, z.cuda(), u.cuda(), 0, pddp.StateEncoding(encoding))
    b_next, b_dz, b_du = batch_eval_dynamics(model, z.cuda(), u.cuda(), 0, pddp.StateEncoding(encoding))
    assert z_next.shape == z.shape and d_dz.shape == (z.shape[0], z.shape[0]) and d_du.shape == (z.shape[0], u.shape[0])
    assert b_next.allclose(z_next, 1e-3, 1e-3) and b_dz.allclose(d_dz, 1e-3, 1e-3) and b_du.allclose(d_du, 1e-3, 1e-3)
    wn, wz, wu, _ = O.dynamics_derivatives(odyn, z, u, encoding, None, i=0)
    for a, w in ((z_next, wn), (d_dz, wz), (d_du, wu)):
        assert float((a.cpu() - w).abs().max()) <= 1e-5 * max(1.0, float(w.abs().max()))


def test_gains_of_the_last_accepted_step_survive_max_reg(pddp):
    """the reference stores self._K only on an accepted step (ilqr.py:166-171).  A fit that ends in
    MAX_REG after rejected retries must leave the accepted step's gains in controller._K, not those of the
    last (huge-mu) backward pass.  Checked against the oracle's controller, which keeps K the same way."""
    import pddp_oracle as O
    dtype = torch.float64
    torch.manual_seed(5)
    _, odyn, ocost = _oracle_specs("pendulum", dtype)
    model = pddp.examples.pendulum.PendulumDynamicsModel(0.1).to(dtype)
    cost = pddp.examples.pendulum.PendulumCost().to(dtype)
    N = 20
    z0 = torch.tensor([0.1, -0.2], dtype=dtype)
    U = 0.1 * torch.randn(N, 1, dtype=dtype)
    enc = O.IGNORE_UNCERTAINTY
    solver = O.ILQR(odyn, ocost, enc)
    trace = []
    # tol = 0: never "converged"; at the numerical optimum every candidate is rejected and mu escalates to max_reg
    Zo, Uo, so = solver.fit(z0, U, n_iterations=100, tol=0.0, max_reg=1e2, trace=trace)
    assert so == O.MAX_REG and any(t[0] == O.ACCEPTED for t in trace), "the scenario must accept, then run out of reg"
    ctrl = pddp.controllers.iLQRController(None, model, cost)
    with _maybe_warns():
        Z, Uc, state = ctrl.fit(U.cuda(), encoding=pddp.StateEncoding(enc), n_iterations=100, tol=0.0, max_reg=1e2,
                                z0=z0.cuda(), quiet=True)
    assert int(state) == O.MAX_REG
    # (accept / reject at the optimum is decided at the rounding level, so the two runs may accept a different
    # number of noise-level steps: the gains agree to the size of the last accepted mu, not to 1e-12)
    scale = max(1.0, float(solver.K.abs().max()))
    assert float((ctrl._K.cpu() - solver.K).abs().max()) <= 1e-4 * scale
    assert float((Uc.cpu() - Uo).abs().max()) <= 1e-4 * max(1.0, float(Uo.abs().max()))
    # ... and they are NOT the gains of the last backward pass, which ran with mu ~ max_reg
    last_pass_K = ctrl._solver.matrices("K")[0].cpu()
    assert float((last_pass_K - solver.K).abs().max()) > 1e-2 * scale
    # the feedback law away from the nominal state uses the accepted gains (ilqr.py:339-354)
    dz = torch.tensor([0.05, -0.03], dtype=dtype)
    u = ctrl(Z[4] + dz.cuda(), 4, pddp.StateEncoding(enc))
    assert torch.allclose(u.cpu(), Uc[4].cpu() + ctrl._K[4].cpu() @ dz, atol=1e-10)
    assert torch.allclose(u.cpu(), Uo[4] + solver.K[4] @ dz, atol=1e-4)


def test_scalar_bound_broadcasts_over_action_size_4(pddp):
    """a 1-element u_min / u_max with action_size 4 (the reference broadcasts it in clamp() and in
    u_min - U[i]) must behave like the 4-vector, not read out of bounds."""
    from pddp.controllers.ilqr import backward, forward
    dtype = torch.float64
    torch.manual_seed(9)
    model = pddp.examples.rendezvous.RendezvousDynamicsModel(0.1).to(dtype)
    cost = pddp.costs.QRCost(pddp.examples.rendezvous.RendezvousCost().Q.data.to(dtype),
                             torch.diag(torch.tensor([0.1, 0.2, 0.3, 0.4], dtype=dtype)), state_size=8, angular_indices=())
    enc = pddp.StateEncoding.IGNORE_UNCERTAINTY
    z0 = torch.tensor([-10.0, -10.0, 10.0, 10.0, 0.0, -5.0, 5.0, 0.0], dtype=dtype).cuda()
    U = (3.0 * torch.randn(6, 4, dtype=dtype)).cuda()
    one, four = torch.tensor([0.7], dtype=dtype).cuda(), torch.full((4,), 0.7, dtype=dtype).cuda()
    a = forward(z0, U, model, cost, enc, u_min=-one, u_max=one)
    b = forward(z0, U, model, cost, enc, u_min=-four, u_max=four)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    ka, Ka = backward(*a, reg=1.0, u_min=-one, u_max=one, U=U)
    kb, Kb = backward(*b, reg=1.0, u_min=-four, u_max=four, U=U)
    assert torch.equal(ka, kb) and torch.equal(Ka, Kb)
    with pytest.raises(ValueError):
        forward(z0, U, model, cost, enc, u_min=-torch.ones(3, dtype=dtype).cuda(), u_max=torch.ones(3, dtype=dtype).cuda())
