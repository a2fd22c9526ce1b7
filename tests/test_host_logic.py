"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol the header declares,
struct layouts match, host-side encodings agree with the oracle, unsupported configurations fail
loudly, and the public API keeps the reference's names and argument order."""
import ctypes
import inspect
import os
import re

import pytest
import torch

import pddp_oracle as O
import pddp_b200
from pddp_b200 import _lib, controllers, costs, examples, models
from pddp_b200.utils import encoding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "pddp_b200.h")).read()
    declared = set(re.findall(r"\b(pddp_[a-z0-9_]+)\s*\(", header))
    assert {"pddp_linearize_known", "pddp_backward", "pddp_rollout_known", "pddp_accept_update",
            "pddp_linearize_bnn", "pddp_rollout_bnn", "pddp_cost_derivatives"} <= declared
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert set(_lib.EXPORTS) == declared
    assert b"sm_100a" in _lib.load().pddp_version()


def test_struct_layouts():
    assert ctypes.sizeof(_lib.Shape) == 8 * 4
    assert ctypes.sizeof(_lib.Cost) == (64 + 64 + 16 + 8 + 4) * 8
    assert ctypes.sizeof(_lib.KnownDynamics) == 64
    assert ctypes.sizeof(_lib.BNN) == 16 + 13 * 8 + 8 + 8 + 8 + 8     # + input_mode (padded), eps_in, eps_out, independent_noise (padded)


def test_no_cpu_fallback():
    cost = examples.pendulum.PendulumCost()
    model = examples.pendulum.PendulumDynamicsModel(0.1)
    with pytest.raises(RuntimeError, match="CUDA"):
        controllers.forward(torch.zeros(2), torch.zeros(5, 1), model, cost, encoding.StateEncoding.IGNORE_UNCERTAINTY)
    with pytest.raises(RuntimeError, match="CUDA"):
        model(torch.zeros(2), torch.zeros(1), 0, encoding.StateEncoding.IGNORE_UNCERTAINTY)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pddp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pddp_oracle" not in text and "refshim" not in text, f


@pytest.mark.parametrize("enc", list(encoding.StateEncoding))
def test_encoding_matches_oracle(enc):
    g = torch.Generator().manual_seed(0)
    m = torch.randn(3, 5, generator=g, dtype=torch.float64)
    A = torch.randn(3, 5, 5, generator=g, dtype=torch.float64)
    C = A @ A.mT + 0.1 * torch.eye(5, dtype=torch.float64)
    z = encoding.encode(m, C=C, encoding=enc)
    assert z.shape[-1] == encoding.infer_encoded_state_size(5, enc) == O.encoded_size(5, int(enc))
    assert torch.allclose(z, O.encode(m, C=C, enc=int(enc)), atol=1e-12)
    assert torch.allclose(encoding.decode_covar(z, enc), O.decode_covar(z, int(enc)), atol=1e-12)
    assert encoding.infer_state_size(z.shape[-1], enc) == 5


def test_unsupported_configurations_raise():
    with pytest.raises(NotImplementedError):
        models.geometry_of(5, (1,))                    # no kernels for an arbitrary geometry
    assert models.geometry_of(8, ()) == _lib.GEO_RENDEZVOUS
    with pytest.raises(NotImplementedError):
        models.bnn.bnn_dynamics_model_factory(4, 2, [200, 200], [2], [0, 1, 3])
    Model = models.bnn.bnn_dynamics_model_factory(4, 1, [32, 32], [2], [0, 1, 3])
    with pytest.raises(NotImplementedError):
        controllers.iLQRController(None, Model(n_particles=8), examples.cartpole.CartpoleCost(),
                                   model_opts={"resample": True})
    with pytest.raises(NotImplementedError):
        controllers.backward(*[torch.zeros(1)] * 9, V_zz_reg=True)


def test_cost_constants_match_reference_examples():
    c = examples.cartpole.CartpoleCost()
    assert c.Q[0, 3] == 0.5 and c.Q[3, 3] == 0.25 and c.Q_term.equal(torch.eye(5))
    assert torch.allclose(c.x_goal, torch.tensor([0, 0, 0, -8.742278e-08, -1.0]))   # fp32 sin(pi), cos(pi)
    d = examples.double_cartpole.DoubleCartpoleCost()
    assert torch.allclose(d.Q[0, 4:], torch.tensor([-0.6, 0.0, -0.6, 0.0]))
    assert d.Q_term.equal(100 * torch.eye(8)) and d.x_goal.tolist() == [0, 0, 0, 0, 0, 1, 0, 1]
    p = examples.pendulum.PendulumCost()
    assert p.Q[0, 1] == 0.5 and p.R[0, 0] == pytest.approx(0.1)


def test_api_mirrors_reference_signatures():
    """Argument names/order of the reference (pddp/controllers/ilqr.py:71,183-192,237-248,318-326,
    393-402,490,530-544,678-687,765)."""
    names = lambda f: list(inspect.signature(f).parameters)
    assert names(controllers.iLQRController.__init__)[:6] == ["self", "env", "model", "cost", "model_opts", "cost_opts"]
    assert names(controllers.iLQRController.fit)[:11] == ["self", "U", "encoding", "n_iterations", "tol", "max_reg",
                                                          "batch_rollout", "quiet", "on_iteration", "u_min", "u_max"]
    assert names(controllers.iLQRController.step)[:10] == ["self", "z0", "U", "i", "encoding", "batch_rollout",
                                                           "alphas", "u_min", "u_max", "on_iteration"]
    assert names(controllers.iLQRController.forward)[:8] == ["self", "z", "i", "encoding", "mpc",
                                                             "ignore_uncertainty", "u_min", "u_max"]
    assert names(controllers.forward) == ["z0", "U", "model", "cost", "encoding", "batch_rollout", "model_opts",
                                          "cost_opts", "u_min", "u_max"]
    assert names(controllers.Q) == ["F_z", "F_u", "L_z", "L_u", "L_zz", "L_uz", "L_uu", "V_z", "V_zz"]
    assert names(controllers.backward)[:15] == ["Z", "F_z", "F_u", "L", "L_z", "L_u", "L_zz", "L_uz", "L_uu", "reg",
                                                "V_zz_reg", "u_min", "u_max", "U", "quiet"]
    assert names(controllers._control_law)[:10] == ["model", "Z", "U", "k", "K", "alpha", "encoding", "model_opts",
                                                    "u_min", "u_max"]
    assert names(controllers._trajectory_cost) == ["cost", "Z", "U", "encoding", "cost_opts"]
    assert [s.value for s in controllers.iLQRState] == [0, 1, 2, 3, 4, 5]
    assert names(models.DynamicsModel.forward)[:6] == ["self", "z", "u", "i", "encoding", "identical_inputs"]
    assert names(costs.Cost.forward)[:6] == ["self", "z", "u", "i", "terminal", "encoding"]


def test_q_function_matches_oracle():
    g = torch.Generator().manual_seed(1)
    r = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    args = (r(4, 4), r(4, 1), r(4), r(1), r(4, 4), r(1, 4), r(1, 1), r(4), r(4, 4))
    for a, b in zip(controllers.Q(*args), O.q_terms(*args)):
        assert torch.allclose(a, b)


def test_bnn_model_draws_reference_style_state():
    Model = models.bnn.bnn_dynamics_model_factory(4, 1, [200, 200], [2], [0, 1, 3])
    m = Model(n_particles=50)
    m.resample(torch.Generator().manual_seed(0))
    d = m.descriptor()
    assert (d.P, d.H0, d.H1) == (50, 200, 200)
    eps = m.eps_in[0]
    assert torch.allclose(eps.mean(0), torch.zeros(4), atol=1e-6) and torch.allclose(eps.std(0), torch.ones(4), atol=1e-6)
    for mask in (m.model.drop_0.mask, m.model.drop_1.mask):
        assert mask.shape == (50, 200) and 0 <= float(mask.min()) and float(mask.max()) <= 1
    assert tuple(m.model.fc_0.weight.shape) == (200, 6) and tuple(m.model.fc_out.weight.shape) == (8, 200)


def test_bnn_model_options_map_to_the_descriptor():
    """infer_noise_variables / sample_input_distribution / use_predicted_std / independent_noise
    (ref: pddp/models/bnn/modules.py:242-262, 320-358) -> BNNDynamics.input_mode / eps_in / eps_out;
    noise the reference would draw lazily is drawn per step, standardised, and then kept."""
    Model = models.bnn.bnn_dynamics_model_factory(4, 1, [32, 32], [2], [0, 1, 3])
    m = Model(n_particles=9)
    m.resample(torch.Generator().manual_seed(0))
    d = m.descriptor({}, 6)
    assert d.input_mode == _lib.BNN_INPUT_INFER and d.tensors["eps_in"] is None and d.tensors["eps_out"] is None
    d = m.descriptor({"sample_input_distribution": False}, 6)
    assert d.input_mode == _lib.BNN_INPUT_MEAN
    d = m.descriptor({"infer_noise_variables": False, "use_predicted_std": True, "independent_noise": True}, 6)
    assert d.input_mode == _lib.BNN_INPUT_RESAMPLE and d.independent_noise
    assert tuple(d.tensors["eps_in"].shape) == (6, 9, 4) and tuple(d.tensors["eps_out"].shape) == (6, 9, 4)
    assert torch.equal(d.tensors["eps_in"][0], m.eps_in[0])
    for e in (d.tensors["eps_in"][3], d.tensors["eps_out"][5]):
        assert torch.allclose(e.mean(0), torch.zeros(4), atol=1e-6) and torch.allclose(e.std(0), torch.ones(4), atol=1e-6)
    again = m.descriptor({"infer_noise_variables": False, "use_predicted_std": True}, 4)
    assert torch.equal(again.tensors["eps_in"], d.tensors["eps_in"][:4])        # drawn once, then persistent
    assert torch.equal(again.tensors["eps_out"], d.tensors["eps_out"][:4])
    with pytest.raises(ValueError):
        m.descriptor({"infer_noise_variables": False})                          # horizon needed to lay eps_in out
    with pytest.raises(ValueError):
        from pddp_b200.solver import BNNDynamics
        t = d.tensors
        BNNDynamics(d.geo, [t["W0"], t["W1"], t["W2"]], [t["b0"], t["b1"], t["b2"]], [t["mask0"], t["mask1"]],
                    t["eps0"], input_mode=_lib.BNN_INPUT_RESAMPLE)              # RESAMPLE without eps_in


def test_rendezvous_and_envs_host_side():
    """action_size 4 constants reach the C structs; environments refuse to live on the CPU."""
    c = examples.rendezvous.RendezvousCost()
    assert c.Q[0, 2] == -1 and c.Q[3, 1] == -1 and c.R.shape == (4, 4) and c.Q_term.equal(c.Q)
    s = c.constants().c_struct()
    assert [s.R[i] for i in (0, 1, 5, 15)] == pytest.approx([0.1, 0.0, 0.1, 0.1])
    assert list(s.u_goal) == [0.0] * 4
    m = examples.rendezvous.RendezvousDynamicsModel(0.1)
    assert (m.state_size, m.action_size) == (8, 4) and m.descriptor().geo == _lib.GEO_RENDEZVOUS
    assert m.descriptor().params == pytest.approx([0.1, 1.0, 0.1])
    with pytest.raises(RuntimeError, match="CUDA"):
        examples.pendulum.PendulumEnv(dt=0.1, device="cpu")
    names = list(inspect.signature(controllers._apply_controller).parameters)
    assert names[:8] == ["env", "cost", "controller", "H", "encoding", "mpc", "quiet", "cost_opts"]   # ref: pddp.py:209-217
    assert issubclass(examples.rendezvous.RendezvousEnv, pddp_b200.envs.KnownDynamicsEnv)
