"""Compatibility shim that lets the UNMODIFIED reference (anassinator/pddp, written for
torch 0.4.1 / Python 3.6) import and run under torch 2.x / Python 3.12.

TEST INFRASTRUCTURE ONLY.  Used in the build container (where /root/reference exists) to
 * validate oracle/pddp_oracle.py against the real reference, and
 * generate the committed golden fixtures under tests/golden/ (oracle/make_golden.py).
Nothing in the product path (pddp_b200/), bench.py's GPU arm or the -m gpu tests imports this
file; /root/reference does not exist on the GPU box.

What is patched:  removed torch-0.4 linear-algebra spellings
(potrf / potrs / gesv / trtrs / Tensor.eig), uint8 mask indexing, `collections.Iterable`,
and stub `gym` modules (the envs are not on the hot path).
"""
import collections
import collections.abc
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = "/root/reference"
_installed = False


def _potrf(a, upper=True):
    L = torch.linalg.cholesky(a)
    return L.mT if upper else L


def _potrs(b, u, upper=True):
    squeeze = b.dim() == 1
    rhs = b.unsqueeze(-1) if squeeze else b
    out = torch.cholesky_solve(rhs, u, upper=upper)
    return out


def _gesv(B, A):
    return torch.linalg.solve(A, B), None


def _trtrs(b, A, upper=True, transpose=False, unitriangular=False):
    if transpose:
        x = torch.linalg.solve_triangular(A.mT, b, upper=not upper, unitriangular=unitriangular)
    else:
        x = torch.linalg.solve_triangular(A, b, upper=upper, unitriangular=unitriangular)
    return x, None


def _eig(self, eigenvectors=False):
    w, V = torch.linalg.eig(self)
    if (w.imag != 0).any() and False:
        pass
    e = torch.stack([w.real, w.imag], -1)
    return e, V.real


def _fix_index(idx):
    if isinstance(idx, torch.Tensor) and idx.dtype == torch.uint8:
        return idx.bool()
    if isinstance(idx, tuple):
        return tuple(_fix_index(i) for i in idx)
    return idx


def install():
    """Patch torch + stub gym, put the reference on sys.path. Idempotent."""
    global _installed
    if _installed:
        return
    _installed = True
    collections.Iterable = collections.abc.Iterable

    torch.Tensor.potrf = _potrf
    torch.potrf = _potrf
    torch.Tensor.potrs = _potrs
    torch.potrs = _potrs
    torch.Tensor.gesv = _gesv
    torch.gesv = _gesv
    torch.Tensor.trtrs = _trtrs
    torch.trtrs = _trtrs
    torch.Tensor.eig = _eig

    _get, _set = torch.Tensor.__getitem__, torch.Tensor.__setitem__

    def getitem(self, idx):
        return _get(self, _fix_index(idx))

    def setitem(self, idx, val):
        return _set(self, _fix_index(idx), val)

    torch.Tensor.__getitem__ = getitem
    torch.Tensor.__setitem__ = setitem

    # ---- gym stubs (envs are out of scope; only import must succeed) ----
    gym = types.ModuleType("gym")
    spaces = types.ModuleType("gym.spaces")
    utils = types.ModuleType("gym.utils")
    seeding = types.ModuleType("gym.utils.seeding")

    class Box(object):
        def __init__(self, low=None, high=None, shape=None, dtype=np.float32):
            self.low = np.asarray(low, dtype=dtype)
            self.high = np.asarray(high, dtype=dtype)
            self.shape = self.low.shape
            self.dtype = dtype

        def sample(self):
            return np.random.uniform(self.low, self.high).astype(self.dtype)

    class Env(object):
        metadata = {}

        def seed(self, seed=None):
            return [seed]

    def np_random(seed=None):
        return np.random.RandomState(seed), seed

    spaces.Box = Box
    seeding.np_random = np_random
    utils.seeding = seeding
    gym.spaces = spaces
    gym.utils = utils
    gym.Env = Env
    for name, mod in (("gym", gym), ("gym.spaces", spaces), ("gym.utils", utils),
                      ("gym.utils.seeding", seeding)):
        sys.modules.setdefault(name, mod)

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def import_reference():
    """Returns the reference `pddp` package (imported from /root/reference)."""
    install()
    import pddp  # noqa: E402  (the reference, not this repo)
    return pddp
