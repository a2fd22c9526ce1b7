"""State encodings: how a Gaussian (mean, covariance) is packed into the flat vector z.

Host-side mirror of pddp/utils/encoding.py (same enum values, same layouts) used to build z0 and to
read results; plain torch ops on whatever device the inputs live on.  The hot path decodes inside
the CUDA kernels (pddp_b200/csrc/core.cuh)."""
from enum import IntEnum

import torch


class StateEncoding(IntEnum):
    """ref: pddp/utils/encoding.py:25-33"""
    FULL_COVARIANCE_MATRIX = 0
    UPPER_TRIANGULAR_CHOLESKY = DEFAULT = 1
    VARIANCE_ONLY = 2
    STANDARD_DEVIATION_ONLY = 3
    IGNORE_UNCERTAINTY = 4


def infer_encoded_state_size(state_size, encoding=StateEncoding.DEFAULT):
    """ref: pddp/utils/encoding.py:46-67"""
    D = state_size
    if encoding == StateEncoding.FULL_COVARIANCE_MATRIX:
        return D + D * D
    if encoding == StateEncoding.UPPER_TRIANGULAR_CHOLESKY:
        return (3 * D + D * D) // 2
    if encoding in (StateEncoding.VARIANCE_ONLY, StateEncoding.STANDARD_DEVIATION_ONLY):
        return 2 * D
    if encoding == StateEncoding.IGNORE_UNCERTAINTY:
        return D
    raise NotImplementedError("Unknown StateEncoding: {}".format(encoding))


def infer_state_size(encoded_state_size, encoding=StateEncoding.DEFAULT):
    """ref: pddp/utils/encoding.py:70-96"""
    for D in range(1, 65):
        if infer_encoded_state_size(D, encoding) == encoded_state_size:
            return D
    raise ValueError("no state size encodes to %d under %r" % (encoded_state_size, encoding))


def _cholesky(C, jitter=1e-12, max_jitter=10.0):
    """ref: pddp/utils/encoding.py:536-564 (same name as the reference's helper)"""
    return _jittered_cholesky_upper(C, jitter, max_jitter)


def _jittered_cholesky_upper(C, jitter=1e-12, max_jitter=10.0):
    """U with U^T U = C + jitter I, jitter escalating x10 (ref: pddp/utils/encoding.py:536-564)."""
    eye = torch.eye(C.shape[-1], dtype=C.dtype, device=C.device)
    while True:
        L, info = torch.linalg.cholesky_ex(C + jitter * eye)
        if not bool((info != 0).any()):
            return L.mT
        jitter *= 10
        if jitter > max_jitter:
            raise RuntimeError("covariance is not positive definite")


def encode(M, C=None, V=None, S=None, encoding=StateEncoding.DEFAULT):
    """(mean, one of covariance / variance / std) -> z.  ref: pddp/utils/encoding.py:99-141"""
    if encoding == StateEncoding.IGNORE_UNCERTAINTY:
        return M
    D = M.shape[-1]
    if V is None:
        V = S ** 2 if S is not None else torch.diagonal(C, dim1=-2, dim2=-1)
    if encoding == StateEncoding.VARIANCE_ONLY:
        other = V
    elif encoding == StateEncoding.STANDARD_DEVIATION_ONLY:
        other = S if S is not None else V.sqrt()
    else:
        cov = C if C is not None else torch.diag_embed(V)
        if encoding == StateEncoding.FULL_COVARIANCE_MATRIX:
            other = cov.reshape(*M.shape[:-1], D * D)
        else:
            iu = torch.triu_indices(D, D, device=M.device)
            other = _jittered_cholesky_upper(cov)[..., iu[0], iu[1]]
    return torch.cat([M, other], -1)


def decode_mean(Z, encoding=StateEncoding.DEFAULT, state_size=None):
    """ref: pddp/utils/encoding.py:144-156"""
    D = state_size or infer_state_size(Z.shape[-1], encoding)
    return Z[..., :D]


def decode_covar(Z, encoding=StateEncoding.DEFAULT, state_size=None):
    """ref: pddp/utils/encoding.py:159-216"""
    D = state_size or infer_state_size(Z.shape[-1], encoding)
    other = Z[..., D:]
    if encoding == StateEncoding.FULL_COVARIANCE_MATRIX:
        return other.reshape(*Z.shape[:-1], D, D)
    if encoding == StateEncoding.UPPER_TRIANGULAR_CHOLESKY:
        iu = torch.triu_indices(D, D, device=Z.device)
        U = torch.zeros(*Z.shape[:-1], D, D, dtype=Z.dtype, device=Z.device)
        U[..., iu[0], iu[1]] = other
        return U.mT @ U
    if encoding == StateEncoding.VARIANCE_ONLY:
        return torch.diag_embed(other)
    if encoding == StateEncoding.STANDARD_DEVIATION_ONLY:
        return torch.diag_embed(other ** 2)
    return (1e-6 * torch.eye(D, dtype=Z.dtype, device=Z.device)).expand(*Z.shape[:-1], D, D)


def decode_var(Z, encoding=StateEncoding.DEFAULT, state_size=None):
    """ref: pddp/utils/encoding.py:219-258"""
    return torch.diagonal(decode_covar(Z, encoding, state_size), dim1=-2, dim2=-1)


def decode_std(Z, encoding=StateEncoding.DEFAULT, state_size=None):
    """Standard deviations [..., D] (IGNORE_UNCERTAINTY: the constant 1e-3).  ref: pddp/utils/encoding.py:263-301"""
    D = state_size or infer_state_size(Z.shape[-1], encoding)
    if encoding == StateEncoding.STANDARD_DEVIATION_ONLY:
        return Z[..., D:]
    return decode_var(Z, encoding, D).sqrt()


def decode_covar_sqrt(Z, encoding=StateEncoding.DEFAULT, state_size=None):
    """Upper-triangular U with U^T U = covariance, [..., D, D]: the stored factor for UT-Cholesky, the jittered
    Cholesky factor of the stored matrix for FULL_COVARIANCE_MATRIX, diagonal for the variance / std encodings,
    1e-3 I for IGNORE_UNCERTAINTY.  ref: pddp/utils/encoding.py:304-362"""
    D = state_size or infer_state_size(Z.shape[-1], encoding)
    other = Z[..., D:]
    if encoding == StateEncoding.FULL_COVARIANCE_MATRIX:
        return _jittered_cholesky_upper(other.reshape(*Z.shape[:-1], D, D))
    if encoding == StateEncoding.UPPER_TRIANGULAR_CHOLESKY:
        iu = torch.triu_indices(D, D, device=Z.device)
        U = torch.zeros(*Z.shape[:-1], D, D, dtype=Z.dtype, device=Z.device)
        U[..., iu[0], iu[1]] = other
        return U
    if encoding == StateEncoding.VARIANCE_ONLY:
        return torch.diag_embed(other.sqrt())
    if encoding == StateEncoding.STANDARD_DEVIATION_ONLY:
        return torch.diag_embed(other)
    return (1e-3 * torch.eye(D, dtype=Z.dtype, device=Z.device)).expand(*Z.shape[:-1], D, D)
