"""Generates tests/golden/*.npz by running the UNMODIFIED reference (via oracle/refshim.py) and, in
the same pass, checks oracle/pddp_oracle.py against it (this is the oracle's parity pin).

Run in the build container only (needs /root/reference):   python oracle/make_golden.py
TEST INFRASTRUCTURE ONLY.

Each fixture stores the INPUTS (z0, U, bounds, reg, alphas, model constants, BNN weights / masks /
eps0 exactly as the reference model object drew them) and the REFERENCE's outputs
(Z, F_z, F_u, L, L_z, L_u, L_zz, L_uz, L_uu, k, K, Z_new, U_new, J, and fit traces).
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refshim  # noqa: E402

warnings.filterwarnings("ignore")
ref = refshim.import_reference()
import pddp_oracle as O  # noqa: E402

from pddp.controllers import ilqr as R  # noqa: E402
from pddp.examples.pendulum.model import PendulumDynamicsModel  # noqa: E402
from pddp.examples.pendulum.cost import PendulumCost  # noqa: E402
from pddp.examples.cartpole.model import CartpoleDynamicsModel  # noqa: E402
from pddp.examples.cartpole.cost import CartpoleCost  # noqa: E402
from pddp.examples.double_cartpole.model import DoubleCartpoleDynamicsModel  # noqa: E402
from pddp.examples.double_cartpole.cost import DoubleCartpoleCost  # noqa: E402
from pddp.examples.rendezvous.model import RendezvousDynamicsModel  # noqa: E402
from pddp.examples.rendezvous.cost import RendezvousCost  # noqa: E402
from pddp.models.bnn import bnn_dynamics_model_factory  # noqa: E402
from pddp.models.bnn.modules import BDropout, CDropout  # noqa: E402
from pddp.utils.gaussian_variable import GaussianVariable  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
DT = 0.1
PROBLEMS = {
    "pendulum": (PendulumDynamicsModel, PendulumCost, O.pendulum_spec, [0.0, 0.0], 2.5),
    "cartpole": (CartpoleDynamicsModel, CartpoleCost, O.cartpole_spec, [0.0, 0.0, 0.0, 0.0], 10.0),
    "double_cartpole": (DoubleCartpoleDynamicsModel, DoubleCartpoleCost, O.double_cartpole_spec,
                        [0.0, 0.0, np.pi, 0.0, np.pi, 0.0], 20.0),
    # initial state of RendezvousEnv.reset (ref: examples/rendezvous/env.py:106-108); action_size 4
    "rendezvous": (RendezvousDynamicsModel, lambda: rendezvous_cost_nd(), O.rendezvous_spec,
                   [-10.0, -10.0, 10.0, 10.0, 0.0, -5.0, 5.0, 0.0], 1.0),
}


def rendezvous_cost_nd():
    """The reference's QRCost with RendezvousCost's Q and a NON-DEGENERATE R.  With the stock
    R = 0.1 I the symmetric problem makes Q_uu a multiple of the identity up to rounding; LAPACK's
    geev (Tensor.eig, ilqr.py:631) then returns non-orthogonal eigenvectors for the repeated
    eigenvalue and the reference's (E/e) E^T is a rounding-dependent matrix, not the regularised
    inverse -- nothing can be pinned on it.  Distinct eigenvalues make the reference well defined."""
    from pddp.costs import QRCost
    R = torch.diag(torch.tensor([0.1, 0.15, 0.2, 0.3])) + 0.02 * (torch.ones(4, 4) - torch.eye(4))
    return QRCost(RendezvousCost().Q.data.clone(), R)


def cost_spec(cost, model_cls, dtype):
    """Reference cost object -> oracle QRCostSpec (constants read from the object's buffers)."""
    DA = cost.Q.shape[0]
    return O.QRCostSpec(cost.Q.data.clone(), cost.R.data.clone(), cost.Q_term.data.clone(),
                        cost.x_goal.data.clone().reshape(-1).expand(DA).clone(), cost.u_goal.data.clone().reshape(-1).expand(
                            model_cls.action_size).clone(),
                        model_cls.state_size, tuple(model_cls.angular_indices.tolist()),
                        tuple(model_cls.non_angular_indices.tolist())).to(dtype)


def bnn_spec(model, model_cls, dtype):
    """Reference BNN model object (after its first call) -> oracle BNNSpec."""
    layers, masks = [], []
    for name, mod in model.model._modules.items():
        if isinstance(mod, torch.nn.Linear):
            layers.append((mod.weight.data.clone(), mod.bias.data.clone()))
        elif isinstance(mod, CDropout):
            masks.append(mod.concrete_noise.data.clone())
        elif isinstance(mod, BDropout):
            masks.append(mod.noise.data.clone())
    vec = lambda b: None if b.dim() == 0 else b.data.clone()
    # (sample_input_distribution=False never draws eps_in: zeros of the right shape then)
    eps0 = model.eps_in[0].data.clone() if 0 in model.eps_in else torch.zeros(
        model.n_particles, model_cls.state_size, dtype=layers[0][0].dtype)
    return O.BNNSpec(layers, masks, eps0, model_cls.state_size,
                     model_cls.action_size, tuple(model_cls.angular_indices.tolist()),
                     tuple(model_cls.non_angular_indices.tolist()), vec(model.X_mean),
                     vec(model.X_std_inv), vec(model.dX_mean), vec(model.dX_std)).to(dtype)


def z0_for(name, enc, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    mean = torch.tensor(PROBLEMS[name][3], dtype=dtype) + 1e-2 * torch.randn(
        len(PROBLEMS[name][3]), generator=g, dtype=dtype)
    return GaussianVariable(mean, var=1e-2 * torch.ones_like(mean)).encode(enc).detach()


def close(a, b, tol, what):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    err = (a - b).abs().max().item() if a.numel() else 0.0
    scale = max(b.abs().max().item(), 1e-30) if b.numel() else 1.0
    ok = err <= tol * max(scale, 1.0)
    print("   %-28s max|d|=%.3e scale=%.3e %s" % (what, err, scale, "ok" if ok else "MISMATCH"))
    return ok


NAMES = "Z F_z F_u L L_z L_u L_zz L_uz L_uu".split()
ONLY = sys.argv[1:]


def run_case(tag, name, enc, dtype, N, seed, bnn=None, bounded=False, reg=1.0, n_alpha=10,
             fit_iters=0, input_mode="infer", pstd=None):
    if ONLY and not any(o in tag for o in ONLY):     # python make_golden.py <substring> ...: subset
        return True
    torch.manual_seed(seed)
    model_cls, cost_cls, spec_fn, _, umax = PROBLEMS[name]
    tol = 1e-9 if dtype == torch.float64 else 2e-4
    cost = cost_cls().to(dtype)
    ospec_cost = cost_spec(cost, model_cls, dtype)
    if bnn is None:
        model = model_cls(DT).to(dtype)
        # constants are read back from the model object: they are fp32-rounded Parameters
        odyn = spec_fn(**{n: float(p) for n, p in model.named_parameters()})
    else:
        hidden, P, scale_out = bnn
        model = bnn_dynamics_model_factory(model_cls.state_size, model_cls.action_size, hidden,
                                           model_cls.angular_indices,
                                           model_cls.non_angular_indices)(n_particles=P).to(dtype)
        with torch.no_grad():
            model.model.fc_out.weight *= scale_out
            model.model.fc_out.bias *= scale_out
        model.eval()
    model_opts = {} if bnn is None else {"use_predicted_std": False, "infer_noise_variables": True}
    if pstd is not None:                   # "dependent" | "independent" (ref: modules.py:242-262)
        model_opts["use_predicted_std"] = True
        model_opts["independent_noise"] = pstd == "independent"
    if input_mode == "resample":
        model_opts["infer_noise_variables"] = False
    elif input_mode == "mean":
        model_opts["sample_input_distribution"] = False
    z0 = z0_for(name, enc, dtype, seed)
    U = (0.1 * torch.randn(N, model_cls.action_size)).to(dtype)
    if bounded:
        U = U * 30 * umax / 10          # make some of the nominal controls hit the bounds
    u_min = torch.full((model_cls.action_size,), -umax, dtype=dtype) if bounded else None
    u_max = torch.full((model_cls.action_size,), umax, dtype=dtype) if bounded else None
    if bnn is not None:
        with torch.no_grad():
            model(z0, U[0], 0, enc, **model_opts)      # draws eps_in[0] and the dropout masks
        odyn = bnn_spec(model, model_cls, dtype)

    ok = True
    print("[%s] %s enc=%d %s N=%d%s" % (tag, name, enc, str(dtype)[6:], N,
                                        " bounded" if bounded else ""))
    lin = R.forward(z0, U.clone(), model, cost, enc, True, model_opts, {}, u_min=u_min,
                    u_max=u_max)
    if input_mode != "infer":
        odyn.input_mode = input_mode
        if input_mode == "resample":      # the reference drew eps_in[i] for every step during forward()
            odyn.eps_in = torch.stack([model.eps_in[i].data.clone() for i in range(N)]).to(dtype)
    if pstd is not None:                   # the reference drew eps_out[i] for every step during forward()
        odyn.eps_out = torch.stack([model.eps_out[i].data.clone() for i in range(N)]).to(dtype)
        odyn.independent_noise = pstd == "independent"
    olin = O.linearize(z0, U, odyn, ospec_cost, enc, u_min, u_max)
    for n, a, b in zip(NAMES, olin, lin):
        ok &= close(a, b, tol, n)

    # backward: escalate reg until the reference accepts (as tests/controllers/test_ilqr.py does)
    # and until the gains are sane (the reference happily returns 1e+279 without raising)
    while True:
        try:
            k, K = R.backward(*lin, reg=reg, u_min=u_min, u_max=u_max, U=U, quiet=True)
            if torch.isfinite(K).all() and K.abs().max() < 1e3:
                break
        except RuntimeError:
            pass
        reg *= 10
    ok_k, oK = O.backward_pass(*lin, reg=reg, u_min=u_min, u_max=u_max, U=U)
    ok &= close(ok_k, k, tol * 10, "k (reg=%g)" % reg)
    ok &= close(oK, K, tol * 10, "K")

    alphas = O.fit_alphas(dtype, n_alpha)
    Zb, Ub = R._control_law(model, lin[0], U, k, K, alphas, enc, model_opts, u_min=u_min,
                            u_max=u_max)
    Jb = R._trajectory_cost(cost, Zb, Ub, enc, {})
    oZb, oUb = O.rollout(odyn, lin[0], U, k, K, alphas, enc, u_min, u_max)
    oJb = O.trajectory_cost(ospec_cost, oZb, oUb, enc)
    ok &= close(oZb, Zb, tol * 10, "Z_new")
    ok &= close(oUb, Ub, tol * 10, "U_new")
    ok &= close(oJb, Jb, tol * 10, "J")

    fx = dict(name=name, enc=enc, N=N, dt=DT, reg=reg, bounded=bounded, z0=z0, U=U, alphas=alphas,
              k=k, K=K, Z_new=Zb, U_new=Ub, J=Jb, Q=cost.Q.data, R=cost.R.data,
              Q_term=cost.Q_term.data, x_goal=cost.x_goal.data.reshape(-1).expand(cost.Q.shape[0]))
    if bounded:
        fx.update(u_min=u_min, u_max=u_max)
    fx.update({n: v for n, v in zip(NAMES, lin)})
    if bnn is None:
        fx.update({"p_" + n: v for n, v in odyn.params.items()})
    else:
        fx["P"], fx["hidden"] = odyn.P, np.array(bnn[0])
        for i, (W, b) in enumerate(odyn.weights):
            fx["W%d" % i], fx["b%d" % i] = W, b
        for i, m in enumerate(odyn.masks):
            fx["mask%d" % i] = m
        fx["eps0"] = odyn.eps0
        fx["input_mode"] = {"infer": 0, "resample": 1, "mean": 2}[input_mode]
        if odyn.eps_in is not None:
            fx["eps_in"] = odyn.eps_in
        if odyn.eps_out is not None:
            fx["eps_out"], fx["independent_noise"] = odyn.eps_out, int(odyn.independent_noise)

    if fit_iters:
        class Env:
            def get_state(self):
                D = model_cls.state_size
                return GaussianVariable(z0[:D].clone(), var=1e-2 * torch.ones(D, dtype=dtype))
        ctrl = R.iLQRController(Env(), model, cost, model_opts=model_opts)
        trace = []
        Zf, Uf, st = ctrl.fit(U.clone(), encoding=enc, n_iterations=fit_iters, quiet=True,
                              u_min=u_min, u_max=u_max,
                              on_iteration=lambda i, s, Z, U_, J: trace.append(
                                  (int(s), float(J), ctrl._mu)))
        otrace = []
        solver = O.ILQR(odyn, ospec_cost, enc)
        oZf, oUf, ost = solver.fit(z0, U, n_iterations=fit_iters, u_min=u_min, u_max=u_max,
                                   trace=otrace)
        same_states = [t[0] for t in trace] == [t[0] for t in otrace]
        print("   fit: %d attempts, final state %d, J=%.6g; oracle states match: %s" % (
            len(trace), int(st), trace[-1][1], same_states))
        ok &= same_states and int(st) == ost
        ok &= close(oZf, Zf, tol * 1e3, "fit Z")
        ok &= close(oUf, Uf, tol * 1e3, "fit U")
        fx.update(fit_Z=Zf, fit_U=Uf, fit_state=int(st), fit_iters=fit_iters,
                  fit_trace=np.array(trace, dtype=np.float64))
    np.savez_compressed(os.path.join(OUT, tag + ".npz"),
                        **{k_: (v.detach().numpy() if isinstance(v, torch.Tensor) else v)
                           for k_, v in fx.items()})
    return ok


def run_closed_loop(tag, name, enc, N, H, seed, bounded, fit_iters=5):
    """Closed loop on the reference's own environment: controller.fit, then
    `_apply_controller(env, cost, controller, H, encoding, mpc=True)` (ref: pddp.py:209-247) -- every
    step is one MPC iteration from the simulator's state -- and an open-loop trial with the fitted U.
    Everything is built under a float64 default dtype so that the env's own model is fp64 too."""
    if ONLY and not any(o in tag for o in ONLY):
        return True
    from pddp.controllers.pddp import _apply_controller
    import importlib
    env_mod = importlib.import_module("pddp.examples.%s.env" % name)
    env_cls = getattr(env_mod, "".join(w.capitalize() for w in name.split("_")) + "Env")
    model_cls, cost_cls, _, _, umax = PROBLEMS[name]
    torch.set_default_dtype(torch.float64)
    try:
        torch.manual_seed(seed)
        np.random.seed(seed)
        env = env_cls(dt=DT)
        model, cost = model_cls(DT), (cost_cls() if name != "rendezvous" else rendezvous_cost_nd()).double()
        env.reset()
        x0 = torch.tensor(env._env.state).clone()
        U0 = 0.1 * torch.randn(N, model_cls.action_size)
        u_min = torch.full((model_cls.action_size,), -umax) if bounded else None
        u_max = torch.full((model_cls.action_size,), umax) if bounded else None
        ctrl = R.iLQRController(env, model, cost)
        Zf, Uf, st = ctrl.fit(U0.clone(), encoding=enc, n_iterations=fit_iters, quiet=True, u_min=u_min, u_max=u_max)
        (X, U, dX), J = _apply_controller(env, cost, ctrl, H, enc, True, True, {}, u_min=u_min, u_max=u_max)
        # open-loop trial of the fitted controls from the same start
        env._env.state = x0.numpy().copy()
        env._state = x0.clone()
        (Xo, Uo, dXo), Jo = _apply_controller(env, cost, Uf, N, enc, False, True, {})
    finally:
        torch.set_default_dtype(torch.float32)
    print("[%s] %s closed loop: fit state %d, mpc J=%.6g, open-loop J=%.6g" % (tag, name, int(st), float(J), float(Jo)))
    fx = dict(name=name, enc=enc, N=N, H=H, bounded=bounded, fit_iters=fit_iters, x0=x0, U0=U0, fit_Z=Zf, fit_U=Uf,
              fit_state=int(st), mpc_X=X, mpc_U=U, mpc_dX=dX, mpc_J=J, ol_X=Xo, ol_U=Uo, ol_dX=dXo, ol_J=Jo,
              Q=cost.Q.data, R=cost.R.data, Q_term=cost.Q_term.data)
    if bounded:
        fx.update(u_min=u_min, u_max=u_max)
    np.savez_compressed(os.path.join(OUT, tag + ".npz"),
                        **{k_: (v.detach().numpy() if isinstance(v, torch.Tensor) else v) for k_, v in fx.items()})
    return True


def main():
    os.makedirs(OUT, exist_ok=True)
    f64, f32 = torch.float64, torch.float32
    E = O
    ok = True
    # known dynamics: IGNORE (cfg 1/4), UT-Cholesky (default), full covariance
    for name in PROBLEMS:
        for enc, etag in ((E.IGNORE_UNCERTAINTY, "ign"), (E.UPPER_TRIANGULAR_CHOLESKY, "ut"),
                          (E.FULL_COVARIANCE_MATRIX, "full")):
            # (rendezvous is linear-quadratic: one Newton step lands on the optimum, and whether a second
            # iteration is ACCEPTED or REJECTED there is a rounding coin flip -- one iteration only)
            ok &= run_case("known_%s_%s_f64" % (name, etag), name, enc, f64, 12, 11,
                           fit_iters=0 if enc != E.IGNORE_UNCERTAINTY else 1 if name == "rendezvous" else 4)
    ok &= run_case("known_pendulum_ign_f32", "pendulum", E.IGNORE_UNCERTAINTY, f32, 100, 3,
                   fit_iters=6)
    ok &= run_case("known_pendulum_ign_bounded_f64", "pendulum", E.IGNORE_UNCERTAINTY, f64, 20, 5,
                   bounded=True, fit_iters=5)
    ok &= run_case("known_cartpole_ut_bounded_f64", "cartpole", E.UPPER_TRIANGULAR_CHOLESKY, f64,
                   10, 6, bounded=True)
    # BNN dynamics: cartpole UT-Cholesky (cfg 2/5), double cartpole full covariance (cfg 3)
    ok &= run_case("bnn_cartpole_ut_small_f64", "cartpole", E.UPPER_TRIANGULAR_CHOLESKY, f64, 6, 21,
                   bnn=([32, 32], 12, 0.05), fit_iters=2)
    ok &= run_case("bnn_cartpole_ut_f64", "cartpole", E.UPPER_TRIANGULAR_CHOLESKY, f64, 5, 22,
                   bnn=([200, 200], 50, 0.02))
    ok &= run_case("bnn_cartpole_ut_f32", "cartpole", E.UPPER_TRIANGULAR_CHOLESKY, f32, 5, 23,
                   bnn=([200, 200], 50, 0.02))
    ok &= run_case("bnn_cartpole_ut_bounded_f64", "cartpole", E.UPPER_TRIANGULAR_CHOLESKY, f64, 5,
                   24, bnn=([32, 32], 12, 0.05), bounded=True, n_alpha=16)
    ok &= run_case("bnn_double_cartpole_full_small_f64", "double_cartpole",
                   E.FULL_COVARIANCE_MATRIX, f64, 4, 25, bnn=([32, 32], 16, 0.05))
    ok &= run_case("bnn_double_cartpole_full_f64", "double_cartpole", E.FULL_COVARIANCE_MATRIX, f64,
                   3, 26, bnn=([200, 200], 50, 0.02))
    # the two diagonal encodings
    for name in PROBLEMS:
        for enc, etag in ((E.VARIANCE_ONLY, "var"), (E.STANDARD_DEVIATION_ONLY, "std")):
            ok &= run_case("known_%s_%s_f64" % (name, etag), name, enc, f64, 12, 11)
    ok &= run_case("bnn_cartpole_var_small_f64", "cartpole", E.VARIANCE_ONLY, f64, 5, 31,
                   bnn=([32, 32], 12, 0.05))
    ok &= run_case("bnn_cartpole_std_small_f64", "cartpole", E.STANDARD_DEVIATION_ONLY, f64, 5, 32,
                   bnn=([32, 32], 12, 0.05))
    # action_size 4: eigen-clipped Q_uu and the 4-dimensional box QP (the unbounded
    # rendezvous cases come from the loops over PROBLEMS above)
    ok &= run_case("known_rendezvous_ign_bounded_f64", "rendezvous", E.IGNORE_UNCERTAINTY, f64, 15, 42,
                   bounded=True, fit_iters=5)
    ok &= run_case("known_rendezvous_ut_bounded_f64", "rendezvous", E.UPPER_TRIANGULAR_CHOLESKY, f64, 8, 43,
                   bounded=True)
    ok &= run_case("known_rendezvous_ign_f32", "rendezvous", E.IGNORE_UNCERTAINTY, f32, 30, 44)
    # BNN input-particle options: infer_noise_variables=False, sample_input_distribution=False
    ok &= run_case("bnn_cartpole_ut_resample_small_f64", "cartpole", E.UPPER_TRIANGULAR_CHOLESKY, f64, 5,
                   33, bnn=([32, 32], 12, 0.05), input_mode="resample")
    ok &= run_case("bnn_cartpole_ut_mean_small_f64", "cartpole", E.UPPER_TRIANGULAR_CHOLESKY, f64, 5,
                   34, bnn=([32, 32], 12, 0.05), input_mode="mean")
    ok &= run_case("bnn_double_cartpole_full_resample_small_f64", "double_cartpole",
                   E.FULL_COVARIANCE_MATRIX, f64, 4, 35, bnn=([32, 32], 16, 0.05), input_mode="resample")
    ok &= run_case("bnn_cartpole_ut_resample_f32", "cartpole", E.UPPER_TRIANGULAR_CHOLESKY, f32, 4, 36,
                   bnn=([200, 200], 50, 0.02), input_mode="resample")
    # use_predicted_std=True: exp(log_std) * eps_out[i] added to every particle
    ok &= run_case("bnn_cartpole_ut_pstd_small_f64", "cartpole", E.UPPER_TRIANGULAR_CHOLESKY, f64, 5, 51,
                   bnn=([32, 32], 12, 0.05), pstd="dependent")
    ok &= run_case("bnn_cartpole_ut_pstd_indep_small_f64", "cartpole", E.UPPER_TRIANGULAR_CHOLESKY, f64, 5, 52,
                   bnn=([32, 32], 12, 0.05), pstd="independent")
    ok &= run_case("bnn_double_cartpole_full_pstd_small_f64", "double_cartpole", E.FULL_COVARIANCE_MATRIX, f64,
                   4, 53, bnn=([32, 32], 16, 0.05), pstd="dependent")
    ok &= run_case("bnn_cartpole_ut_pstd_f32", "cartpole", E.UPPER_TRIANGULAR_CHOLESKY, f32, 4, 54,
                   bnn=([200, 200], 50, 0.02), pstd="dependent")
    ok &= run_case("bnn_cartpole_std_pstd_resample_small_f64", "cartpole", E.STANDARD_DEVIATION_ONLY, f64, 4, 55,
                   bnn=([32, 32], 12, 0.05), pstd="dependent", input_mode="resample")
    # closed loop on the simulator
    ok &= run_closed_loop("loop_pendulum_ign_bounded", "pendulum", E.IGNORE_UNCERTAINTY, 15, 8, 61, True)
    ok &= run_closed_loop("loop_cartpole_ut", "cartpole", E.UPPER_TRIANGULAR_CHOLESKY, 12, 6, 62, False)
    ok &= run_closed_loop("loop_rendezvous_ign_bounded", "rendezvous", E.IGNORE_UNCERTAINTY, 10, 5, 63, True)
    print("ALL OK" if ok else "SOME MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
