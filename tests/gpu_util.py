"""Helpers for the -m gpu parity tests: golden fixture -> pddp_b200 solver objects."""
import torch

import pddp_b200
from pddp_b200 import _lib
from pddp_b200.solver import BatchedSolver, KnownDynamics, QRCostConstants

GEO = {"pendulum": _lib.GEO_PENDULUM, "cartpole": _lib.GEO_CARTPOLE,
       "double_cartpole": _lib.GEO_DOUBLE_CARTPOLE, "rendezvous": _lib.GEO_RENDEZVOUS}
PARAM_ORDER = {"pendulum": ("dt", "m", "l", "mu", "g"),
               "cartpole": ("dt", "mc", "mp", "l", "mu", "g"),
               "double_cartpole": ("dt", "mc", "mp1", "mp2", "l1", "l2", "mu", "g"),
               "rendezvous": ("dt", "m", "alpha")}


def cost_from_fixture(fx):
    return QRCostConstants(fx.t("Q"), fx.t("R"), fx.t("Q_term"), fx.t("x_goal"))


def dynamics_from_fixture(fx, device="cuda"):
    if not fx.is_bnn:
        p = fx.known_params()
        return KnownDynamics(GEO[fx.name], [p[k] for k in PARAM_ORDER[fx.name]])
    from pddp_b200.solver import BNNDynamics
    return BNNDynamics(GEO[fx.name], [fx.t("W0"), fx.t("W1"), fx.t("W2")],
                       [fx.t("b0"), fx.t("b1"), fx.t("b2")], [fx.t("mask0"), fx.t("mask1")],
                       fx.t("eps0"), input_mode=("infer", "resample", "mean").index(fx.input_mode),
                       eps_in=fx.t("eps_in") if fx.has("eps_in") else None,
                       eps_out=fx.t("eps_out") if fx.has("eps_out") else None,
                       independent_noise=bool(int(fx.raw["independent_noise"])) if fx.has("independent_noise") else False)


def solver_from_fixture(fx, B=1, layout=None):
    s = BatchedSolver(dynamics_from_fixture(fx), cost_from_fixture(fx), fx.enc, B, fx.N,
                      dtype=fx.dtype, layout=layout)
    return s


def tile(t, B):
    return t.unsqueeze(0).expand(B, *t.shape).contiguous().cuda()
