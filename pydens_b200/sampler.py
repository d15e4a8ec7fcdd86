""" Samplers for collocation points.

The reference re-exports batchflow's samplers (`from batchflow.sampler import *`, pydens/__init__.py:5)
and only ever calls `sampler.sample(batch_size)` (model_torch.py:433).  This module provides the subset
its README / tutorial use — `NumpySampler(name, **kwargs)`, `&` for column concatenation, `dim=` — with the
same host behaviour, plus `device_columns()`: when every column is an independent uniform / normal /
constant, `Solver.fit` lowers the sampler to the in-kernel Philox generator so the training loop never
touches the host (distribution-level equivalent; the RNG streams differ, as they would between numpy
versions).
"""
import numpy as np

__all__ = ['Sampler', 'NumpySampler', 'ConstantSampler']

_ALIASES = {'u': 'uniform', 'n': 'normal'}
COL_UNIFORM, COL_NORMAL, COL_CONST = 0, 1, 2


class Sampler:
    """ Base class: anything with `.sample(size) -> ndarray [size, dim]`. """
    dim = 1

    def sample(self, size):
        raise NotImplementedError

    def __and__(self, other):
        return _Concat(self, other)

    def device_columns(self):
        """ [(kind, a, b)] per column if the sampler can run in-kernel, else None. """
        return None


class _Concat(Sampler):
    def __init__(self, left, right):
        self.left, self.right = left, right
        self.dim = left.dim + right.dim

    def sample(self, size):
        return np.concatenate([self.left.sample(size), self.right.sample(size)], axis=1)

    def device_columns(self):
        l, r = self.left.device_columns(), self.right.device_columns()
        return None if l is None or r is None else l + r


class ConstantSampler(Sampler):
    def __init__(self, constant, **kwargs):
        _ = kwargs
        self.constant = np.atleast_1d(np.asarray(constant, dtype=np.float64)).reshape(1, -1)
        self.dim = self.constant.shape[1]

    def sample(self, size):
        return np.repeat(self.constant, size, axis=0)

    def device_columns(self):
        return [(COL_CONST, float(c), 0.0) for c in self.constant[0]]


class NumpySampler(Sampler):
    """ `NumpySampler('uniform', low=1, high=5)`, `NumpySampler('n', dim=2)`, … — draws from
    `numpy.random.RandomState(seed).<name>(size=(size, dim), **kwargs)`. """

    def __init__(self, name, seed=None, dim=1, **kwargs):
        self.name = _ALIASES.get(name, name)
        self.dim = dim
        self.kwargs = kwargs
        self.state = np.random.RandomState(seed)
        if not hasattr(self.state, self.name):
            raise ValueError('numpy.random has no distribution %r' % name)

    def sample(self, size):
        return getattr(self.state, self.name)(size=(size, self.dim), **self.kwargs)

    def device_columns(self):
        kw = self.kwargs
        if self.name == 'uniform' and set(kw) <= {'low', 'high'}:
            lo, hi = kw.get('low', 0.0), kw.get('high', 1.0)
            if np.isscalar(lo) and np.isscalar(hi):
                return [(COL_UNIFORM, float(lo), float(hi))] * self.dim
        if self.name == 'normal' and set(kw) <= {'loc', 'scale'}:
            loc, scale = kw.get('loc', 0.0), kw.get('scale', 1.0)
            if np.isscalar(loc) and np.isscalar(scale):
                return [(COL_NORMAL, float(loc), float(scale))] * self.dim
        return None
