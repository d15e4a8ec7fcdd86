""" Stand-in for `batchflow.sampler`: NumpySampler with `&` (column concatenation).

Call sites being served: pydens/model_torch.py:433 `sampler.sample(batch_size)`; README.md:82
`NumpySampler('uniform') & NumpySampler('uniform', low=1, high=5)`; tutorial `NS('u', dim=2) & ...`.
"""
import numpy as np

__all__ = ['NumpySampler', 'Sampler']

_ALIASES = {'u': 'uniform', 'n': 'normal'}


class Sampler:
    def sample(self, size):
        raise NotImplementedError

    def __and__(self, other):
        return _Concat(self, other)


class _Concat(Sampler):
    def __init__(self, left, right):
        self.left, self.right = left, right

    def sample(self, size):
        return np.concatenate([self.left.sample(size), self.right.sample(size)], axis=1)


class NumpySampler(Sampler):
    def __init__(self, name, seed=None, dim=1, **kwargs):
        self.name = _ALIASES.get(name, name)
        self.dim = dim
        self.kwargs = kwargs
        self.state = np.random.RandomState(seed)

    def sample(self, size):
        return getattr(self.state, self.name)(size=(size, self.dim), **self.kwargs)
