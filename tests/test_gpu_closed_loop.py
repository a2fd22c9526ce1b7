"""Closed loop on the device: the batched simulator step `pddp_env_step_known` and
`_apply_controller` (ref: pddp/controllers/pddp.py:209-247) -- MPC from the simulator's state at every
step, and open-loop trials -- against runs of the reference's own environments and controller
(fixtures loop_*.npz, oracle/make_golden.py run_closed_loop)."""
import pytest
import torch

import pddp_oracle as O
from golden_util import load_raw, loop_tags, rel_err

pytestmark = pytest.mark.gpu


def build(name, fx):
    """Models / costs / env with fp64 constants, as the fixture's reference objects had them."""
    import pddp_b200 as P
    torch.set_default_dtype(torch.float64)
    try:
        model = {"pendulum": P.examples.pendulum.PendulumDynamicsModel, "cartpole": P.examples.cartpole.CartpoleDynamicsModel,
                 "rendezvous": P.examples.rendezvous.RendezvousDynamicsModel}[name](0.1)
        if name == "rendezvous":
            cost = P.costs.QRCost(fx["Q"], fx["R"], state_size=8, angular_indices=())
        else:
            cost = {"pendulum": P.examples.pendulum.PendulumCost, "cartpole": P.examples.cartpole.CartpoleCost}[name]()
        cost = cost.double()
        env_cls = {"pendulum": P.examples.pendulum.PendulumEnv, "cartpole": P.examples.cartpole.CartpoleEnv,
                   "rendezvous": P.examples.rendezvous.RendezvousEnv}[name]
    finally:
        torch.set_default_dtype(torch.float32)
    return model, cost, env_cls


@pytest.mark.parametrize("name,kind", [("pendulum", "pendulum"), ("cartpole", "cartpole"),
                                       ("double_cartpole", "double_cartpole"), ("rendezvous", "rendezvous")])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_env_step_matches_the_oracle_model(name, kind, dtype):
    import pddp_b200 as P
    env_cls = {"pendulum": P.examples.pendulum.PendulumEnv, "cartpole": P.examples.cartpole.CartpoleEnv,
               "double_cartpole": P.examples.double_cartpole.DoubleCartpoleEnv, "rendezvous": P.examples.rendezvous.RendezvousEnv}[name]
    B = 300
    env = env_cls(dt=0.1, batch_size=B, dtype=dtype, generator=torch.Generator().manual_seed(1))
    assert env.state_size == env.get_state().mean().shape[-1] and env.get_state().mean().shape[0] == B
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, env.state_size, generator=g, dtype=torch.float64).float().double()      # float32-representable
    u = torch.randn(B, env.action_size, generator=g, dtype=torch.float64)
    env.set_state(x)
    env.apply(u.cuda())
    spec = {"pendulum": O.pendulum_spec, "cartpole": O.cartpole_spec, "double_cartpole": O.double_cartpole_spec,
            "rendezvous": O.rendezvous_spec}[kind](**{n: float(getattr(env.model, n).detach())
                                                      for n in env.model._param_order})
    want = O._KNOWN[kind](spec.params, x, u)
    assert rel_err(env.get_state().mean().cpu(), want) <= (1e-12 if dtype == torch.float64 else 1e-5)
    z = env.get_state().encode(P.StateEncoding.DEFAULT)
    assert z.is_cuda and z.shape == (B, O.encoded_size(env.state_size, O.DEFAULT))


@pytest.mark.parametrize("tag", loop_tags())
def test_closed_loop_matches_the_reference(tag):
    import pddp_b200 as P
    from pddp_b200.controllers import _apply_controller
    fx = load_raw(tag)
    name, enc, N, H = fx["name"], int(fx["enc"]), int(fx["N"]), int(fx["H"])
    model, cost, env_cls = build(name, fx)
    lo = fx["u_min"].cuda() if fx["bounded"] else None
    hi = fx["u_max"].cuda() if fx["bounded"] else None
    B = 3                                   # identical instances: every row must reproduce the reference run
    env = env_cls.from_model(model, batch_size=B, dtype=torch.float64)
    env.set_state(fx["x0"].unsqueeze(0).expand(B, -1))
    ctrl = P.iLQRController(env, model, cost)
    U0 = fx["U0"].unsqueeze(0).expand(B, -1, -1).contiguous().cuda()
    Z, U, state = ctrl.fit(U0, encoding=enc, n_iterations=int(fx["fit_iters"]), quiet=True, u_min=lo, u_max=hi,
                           z0=env.get_state().encode(enc))
    assert state.cpu().tolist() == [int(fx["fit_state"])] * B
    assert rel_err(Z.cpu()[1], fx["fit_Z"]) <= 1e-5 and rel_err(U.cpu()[1], fx["fit_U"]) <= 1e-5
    U_fit = U.clone()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        (X, Um, dX), J = _apply_controller(env, cost, ctrl, H, enc, True, True, {}, u_min=lo, u_max=hi)
    assert X.is_cuda and X.shape == (B * H, env.state_size) and J.shape == (B,)
    for b in range(B):
        rows = slice(b * H, (b + 1) * H)
        assert rel_err(X[rows].cpu(), fx["mpc_X"]) <= 1e-5, b
        assert rel_err(Um[rows].cpu(), fx["mpc_U"]) <= 1e-5, b
        assert rel_err(dX[rows].cpu(), fx["mpc_dX"]) <= 1e-5, b
        assert rel_err(J[b].cpu(), fx["mpc_J"]) <= 1e-5, b
    # open-loop trial of the fitted controls from the same start (train_on_start trials, pddp.py:121-141)
    env.set_state(fx["x0"].unsqueeze(0).expand(B, -1))
    (Xo, Uo, dXo), Jo = _apply_controller(env, cost, U_fit, N, enc, False, True, {})
    assert rel_err(Xo[:N].cpu(), fx["ol_X"]) <= 1e-5 and rel_err(dXo[:N].cpu(), fx["ol_dX"]) <= 1e-5
    assert rel_err(Jo[0].cpu(), fx["ol_J"]) <= 1e-5
