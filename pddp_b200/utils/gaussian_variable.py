"""Gaussian random variable container: what `env.get_state()` returns and the controller encodes.

Mirror of pddp/utils/gaussian_variable.py (mean + one of covariance / variance / std, the others derived
lazily); plain torch ops on whatever device the tensors live on -- none of this is on the hot path."""
import torch

from .encoding import (StateEncoding, decode_covar, decode_mean, decode_std, decode_var, encode)


class GaussianVariable(object):
    """ref: pddp/utils/gaussian_variable.py:22-275"""

    def __init__(self, mean, covar=None, var=None, std=None):
        self._mean, self._covar, self._var, self._std = mean, covar, var, std
        self._is_from_covar = covar is not None

    def __repr__(self):
        return "GaussianVariable({})".format(self.shape)

    @property
    def device(self):
        return self._mean.device

    @property
    def dtype(self):
        return self._mean.dtype

    @property
    def shape(self):
        return self._mean.shape

    def mean(self):
        return self._mean

    def covar(self):
        if self._covar is None:
            if self._var is not None:
                self._covar = torch.diag_embed(self._var)
            elif self._std is not None:
                self._covar = torch.diag_embed(self._std ** 2)
            else:
                raise NotImplementedError("Cannot compute covariance")
        return self._covar

    def var(self):
        if self._var is None:
            if self._covar is not None:
                self._var = torch.diagonal(self._covar, dim1=-2, dim2=-1)
            elif self._std is not None:
                self._var = self._std ** 2
            else:
                raise NotImplementedError("Cannot compute variance")
        return self._var

    def std(self):
        if self._std is None:
            self._std = self.var().sqrt()
        return self._std

    def sample(self, sample_shape=torch.Size([])):
        """ref: gaussian_variable.py:108-123"""
        if self._is_from_covar:
            return torch.distributions.MultivariateNormal(self.mean(), self.covar()).sample(sample_shape)
        return torch.distributions.Normal(self.mean(), self.std()).sample(sample_shape)

    def encode(self, encoding=StateEncoding.DEFAULT):
        """ref: gaussian_variable.py:125-145 (which moment is handed to `encode` per encoding)"""
        if encoding in (StateEncoding.FULL_COVARIANCE_MATRIX, StateEncoding.UPPER_TRIANGULAR_CHOLESKY):
            return encode(self.mean(), C=self.covar(), encoding=encoding)
        if encoding == StateEncoding.STANDARD_DEVIATION_ONLY:
            return encode(self.mean(), S=self.std(), encoding=encoding)
        return encode(self.mean(), V=self.var(), encoding=encoding)

    @classmethod
    def decode(cls, z, encoding=StateEncoding.DEFAULT, state_size=None):
        """ref: gaussian_variable.py:147-176"""
        mean = decode_mean(z, encoding, state_size)
        if encoding in (StateEncoding.FULL_COVARIANCE_MATRIX, StateEncoding.UPPER_TRIANGULAR_CHOLESKY):
            return cls(mean, covar=decode_covar(z, encoding, state_size))
        if encoding == StateEncoding.STANDARD_DEVIATION_ONLY:
            return cls(mean, std=decode_std(z, encoding, state_size))
        return cls(mean, var=decode_var(z, encoding, state_size))

    def _map(self, fn):
        return GaussianVariable(*[None if t is None else fn(t) for t in (self._mean, self._covar, self._var, self._std)])

    def detach(self):
        return self._map(lambda t: t.detach())

    def clone(self):
        return self._map(lambda t: t.clone())

    def to(self, *args, **kwargs):
        return self._map(lambda t: t.to(*args, **kwargs))

    def cpu(self):
        return self.to("cpu")

    def cuda(self, *args, **kwargs):
        return self._map(lambda t: t.cuda(*args, **kwargs))

    def double(self):
        return self.to(torch.float64)

    def float(self):
        return self.to(torch.float32)

    @classmethod
    def random(cls, n, reg=1e-1, requires_grad=True, **tensor_opts):
        """A random valid Gaussian of size n: covariance L^T L + reg I.  ref: gaussian_variable.py:258-275"""
        mean = torch.randn(n, **tensor_opts)
        L = torch.randn(n, n, **tensor_opts)
        covar = L.t().mm(L) + reg * torch.eye(n, **tensor_opts)
        return cls(mean, covar=covar)
