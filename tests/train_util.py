"""Loader of the BNN-training fixtures (tests/golden/train_*.npz, written by oracle/make_golden_train.py)."""
import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def train_tags():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "train_*.npz")))


def load(tag):
    raw = np.load(os.path.join(GOLDEN_DIR, tag + ".npz"))
    fx = {k: (torch.from_numpy(np.array(raw[k])) if raw[k].ndim else raw[k].item()) for k in raw.files}
    fx["dtype"] = fx["X"].dtype
    return fx


def augmented_inputs(fx):
    """[x_nonang, sin, cos, u] of the cartpole dataset (angle index 2), as fit() forms it (modules.py:158-163)."""
    X, U = fx["X"], fx["U"]
    return torch.cat([X[:, [0, 1, 3]], X[:, 2:3].sin(), X[:, 2:3].cos(), U], -1)
