"""Loads tests/golden/*.npz (reference outputs written by oracle/make_golden.py) and rebuilds the
oracle-side problem specs from the stored constants.  Test helper; imports oracle/."""
import glob
import os

import numpy as np
import torch

import pddp_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LIN_NAMES = "Z F_z F_u L L_z L_u L_zz L_uz L_uu".split()
GEOMETRY = {  # name -> (D, nu, angular, non-angular)
    "pendulum": (2, 1, (0,), (1,)),
    "cartpole": (4, 1, (2,), (0, 1, 3)),
    "double_cartpole": (6, 1, (2, 4), (0, 1, 3, 5)),
    "rendezvous": (8, 4, (), tuple(range(8))),
}
SPEC_FN = {"pendulum": O.pendulum_spec, "cartpole": O.cartpole_spec,
           "double_cartpole": O.double_cartpole_spec, "rendezvous": O.rendezvous_spec}


def all_tags():
    """Per-pass fixtures (known_* / bnn_*); the closed-loop fixtures (loop_*) have their own loader."""
    return sorted(t for t in (os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
                  if not t.startswith(("loop_", "train_")))


def loop_tags():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "loop_*.npz")))


def load_raw(tag):
    raw = np.load(os.path.join(GOLDEN_DIR, tag + ".npz"), allow_pickle=False)
    return {k: (torch.from_numpy(np.array(raw[k])) if raw[k].ndim else raw[k].item()) for k in raw.files}


class Fixture:
    def __init__(self, tag):
        self.tag = tag
        raw = np.load(os.path.join(GOLDEN_DIR, tag + ".npz"), allow_pickle=False)
        self.raw = {k: raw[k] for k in raw.files}
        self.name = str(self.raw["name"])
        self.enc = int(self.raw["enc"])
        self.N = int(self.raw["N"])
        self.reg = float(self.raw["reg"])
        self.bounded = bool(self.raw["bounded"])
        self.dtype = torch.float64 if self.raw["z0"].dtype == np.float64 else torch.float32
        self.is_bnn = "W0" in self.raw
        self.D, self.nu, self.ang, self.nonang = GEOMETRY[self.name]

    def t(self, key):
        return torch.from_numpy(np.array(self.raw[key]))

    def has(self, key):
        return key in self.raw

    @property
    def bounds(self):
        if not self.bounded:
            return None, None
        return self.t("u_min"), self.t("u_max")

    def lin(self):
        return tuple(self.t(n) for n in LIN_NAMES)

    def known_params(self):
        return {k[2:]: float(v) for k, v in self.raw.items() if k.startswith("p_")}

    def cost_spec(self):
        return O.QRCostSpec(self.t("Q"), self.t("R"), self.t("Q_term"), self.t("x_goal"),
                            torch.zeros(self.nu, dtype=self.dtype), self.D, self.ang, self.nonang)

    def dyn_spec(self):
        if not self.is_bnn:
            return SPEC_FN[self.name](**self.known_params())
        n_layers = sum(1 for k in self.raw if k.startswith("W"))
        weights = [(self.t("W%d" % i), self.t("b%d" % i)) for i in range(n_layers)]
        masks = [self.t("mask%d" % i) for i in range(n_layers - 1)]
        spec = O.BNNSpec(weights, masks, self.t("eps0"), self.D, self.nu, self.ang, self.nonang)
        spec.input_mode = self.input_mode
        spec.eps_in = self.t("eps_in") if self.has("eps_in") else None
        spec.eps_out = self.t("eps_out") if self.has("eps_out") else None
        spec.independent_noise = bool(int(self.raw["independent_noise"])) if self.has("independent_noise") else False
        return spec

    @property
    def input_mode(self):
        """BNN input-particle option the reference ran with (absent in the older fixtures = default)."""
        return ("infer", "resample", "mean")[int(self.raw["input_mode"])] if self.has("input_mode") else "infer"


def rel_err(a, b):
    """max |a-b| / max(1, max|b|): relative to the tensor's scale (entries that are structurally
    zero would make an element-wise relative error meaningless)."""
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    if b.numel() == 0:
        return 0.0
    if not torch.isfinite(a).all():
        return float("inf")
    return ((a - b).abs().max() / max(1.0, b.abs().max().item())).item()
