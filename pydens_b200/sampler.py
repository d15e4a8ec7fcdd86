""" Samplers for collocation points.

The reference re-exports batchflow's samplers (`from batchflow.sampler import *`, pydens/__init__.py:5)
and only ever calls `sampler.sample(batch_size)` (model_torch.py:433).  This module provides the subset
its README / tutorial use — `NumpySampler(name, **kwargs)`, `&` for column concatenation, `dim=` — and the
rest of that sampler algebra on the host: `w & s` (weight), `s1 | s2` (mixture), `+ - * /` with numbers /
arrays / other samplers, `.apply(f)`, `.truncate(high, low, expr)`, `ScipySampler`, `HistoSampler`.
Anything with `.sample(size)` can feed `Solver.fit` (batches are staged through pinned memory); in addition
`device_columns()` says when every column is an independent uniform / normal / constant — possibly shifted
and scaled by numbers — and `Solver.fit` then lowers the sampler to the in-kernel Philox generator so the
training loop never touches the host (distribution-level equivalent; the RNG streams differ, as they would
between numpy versions).
"""
import numpy as np

__all__ = ['Sampler', 'NumpySampler', 'ConstantSampler', 'ScipySampler', 'HistoSampler']

_ALIASES = {'u': 'uniform', 'n': 'normal', 'e': 'exponential', 'g': 'gamma', 'be': 'beta', 'ln': 'lognormal',
            'w': 'weibull', 'p': 'poisson', 'b': 'binomial', 'mvn': 'multivariate_normal', 'c': 'choice'}
COL_UNIFORM, COL_NORMAL, COL_CONST, COL_TNORMAL = 0, 1, 2, 4


def _is_number(x):
    return isinstance(x, (int, float, np.integer, np.floating))


class Sampler:
    """ Base class: anything with `.sample(size) -> ndarray [size, dim]`. """
    dim = 1
    weight = 1.0
    __array_ufunc__ = None                                # `ndarray + sampler` defers to the sampler's reflected ops

    def sample(self, size):
        raise NotImplementedError

    # ---- algebra ----
    def __and__(self, other):
        if _is_number(other):                             # `0.3 & sampler`: weight inside a mixture
            self.weight = self.weight * float(other)
            return self
        return _Concat(self, other)

    __rand__ = __and__

    def __or__(self, other):
        return _Mixture(self, other)

    def __add__(self, other):
        return _Arith(np.add, self, other)

    def __radd__(self, other):
        return _Arith(np.add, other, self)

    def __sub__(self, other):
        return _Arith(np.subtract, self, other)

    def __rsub__(self, other):
        return _Arith(np.subtract, other, self)

    def __mul__(self, other):
        return _Arith(np.multiply, self, other)

    def __rmul__(self, other):
        return _Arith(np.multiply, other, self)

    def __truediv__(self, other):
        return _Arith(np.divide, self, other)

    def __rtruediv__(self, other):
        return _Arith(np.divide, other, self)

    def __neg__(self):
        return _Arith(np.multiply, self, -1.0)

    def apply(self, transform):
        """ Sampler of `transform(points)` (array [size, dim] -> array [size, new_dim]). """
        return _Apply(self, transform)

    def truncate(self, high=None, low=None, expr=None, prob=0.5, max_iters=None, sample_anyway=False):
        """ Rejection sampler: keeps the points with low <= expr(points) <= high (expr = identity by default;
        bounds are numbers or per-column sequences).  `prob` is the guessed acceptance rate that sizes each
        round; after `max_iters` rounds it raises, or pads with unfiltered points if `sample_anyway`. """
        return _Truncated(self, high, low, expr, prob, max_iters, sample_anyway)

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


class _Mixture(Sampler):
    """ `s1 | s2`: every point comes from s1 with probability w1 / (w1 + w2), else from s2. """

    def __init__(self, left, right, seed=None):
        if left.dim != right.dim:
            raise ValueError('mixture of samplers with %d and %d columns' % (left.dim, right.dim))
        self.left, self.right, self.dim = left, right, left.dim
        self.weight = left.weight + right.weight          # so that (a | b) | c keeps the three weights
        self.state = np.random.RandomState(seed)

    def device_columns(self):
        """ In-kernel form: per column ('mix', group key, [(weight, kind, a, b), …]); every column of this
        mixture carries the same key, so one draw per point picks the side for the whole row.  Each side must be
        made of simple columns or be such a mixture itself (`(a | b) | c`), at most 4 components in all. """
        sides = []
        for side in (self.left, self.right):
            cols = side.device_columns()
            if cols is None:
                return None
            keys = {c[1] for c in cols if c[0] == 'mix'}
            if keys and (len(keys) > 1 or any(c[0] != 'mix' for c in cols)):
                return None                               # a side that mixes only some of its columns
            sides.append((side.weight, cols))
        out = []
        for k in range(self.dim):
            comps = []
            for weight, cols in sides:
                col = cols[k]
                if col[0] == 'mix':
                    inner = float(sum(c[0] for c in col[2]))
                    comps += [(weight * c[0] / inner, c[1], c[2], c[3]) for c in col[2]]
                else:
                    comps.append((weight, col[0], col[1], col[2]))
            if len(comps) > 4:
                return None
            out.append(('mix', id(self), comps))
        return out

    def sample(self, size):
        n_left = self.state.binomial(size, self.left.weight / (self.left.weight + self.right.weight))
        pts = np.concatenate([np.asarray(self.left.sample(n_left), dtype=np.float64).reshape(n_left, self.dim),
                              np.asarray(self.right.sample(size - n_left), dtype=np.float64).reshape(size - n_left, self.dim)])
        return pts[self.state.permutation(size)]


class _Arith(Sampler):
    """ Element-wise arithmetic between a sampler and a number / array / another sampler. """

    def __init__(self, op, left, right):
        self.op, self.left, self.right = op, left, right
        dims = [x.dim for x in (left, right) if isinstance(x, Sampler)]
        other = [np.atleast_1d(np.asarray(x, dtype=np.float64)) for x in (left, right) if not isinstance(x, Sampler)]
        self.dim = max(dims + [o.shape[-1] for o in other])
        self.weight = next(x.weight for x in (left, right) if isinstance(x, Sampler))

    @staticmethod
    def _draw(x, size):
        return np.asarray(x.sample(size), dtype=np.float64) if isinstance(x, Sampler) else np.asarray(x, dtype=np.float64)

    def sample(self, size):
        return np.broadcast_to(self.op(self._draw(self.left, size), self._draw(self.right, size)), (size, self.dim)).copy()

    def device_columns(self):
        # (number op columns) stays in the uniform / normal / constant family for + - * and / by a number
        samplers = [x for x in (self.left, self.right) if isinstance(x, Sampler)]
        if len(samplers) != 1 or not _is_number(self.left if samplers[0] is self.right else self.right):
            return None
        cols = samplers[0].device_columns()
        if cols is None or any(col[0] == 'mix' or len(col) != 3 for col in cols):
            return None
        c = float(self.left if samplers[0] is self.right else self.right)
        sampler_first = samplers[0] is self.left
        if self.op is np.add:
            shift, scale = c, 1.0
        elif self.op is np.subtract:
            shift, scale = (-c, 1.0) if sampler_first else (c, -1.0)
        elif self.op is np.multiply:
            shift, scale = 0.0, c
        elif self.op is np.divide and sampler_first and c != 0.0:
            shift, scale = 0.0, 1.0 / c
        else:
            return None
        out = []
        for kind, a, b in cols:
            if kind == COL_UNIFORM:                       # U[a, b) -> scale * U + shift (a > b is fine: a + u (b - a))
                out.append((COL_UNIFORM, scale * a + shift, scale * b + shift))
            elif kind == COL_NORMAL:
                out.append((COL_NORMAL, scale * a + shift, abs(scale) * b))
            else:
                out.append((COL_CONST, scale * a + shift, 0.0))
        return out


class _Apply(Sampler):
    def __init__(self, base, transform):
        self.base, self.transform, self.weight = base, transform, base.weight
        self.dim = np.asarray(transform(np.asarray(base.sample(2), dtype=np.float64))).reshape(2, -1).shape[1]

    def sample(self, size):
        return np.asarray(self.transform(np.asarray(self.base.sample(size), dtype=np.float64))).reshape(size, -1)

    def device_columns(self):
        """ A transform that turns out to be a per-column affine map (y_k = s_k x_k + t_k: unit changes, shifts,
        reflections — the usual use of `.apply`) keeps the columns in the uniform / normal / constant family and runs
        in-kernel.  The callable is opaque, so it is PROBED: affine and column-separable on random points to 1e-12,
        else the sampler stays on the host. """
        cols = self.base.device_columns()
        if cols is None or self.dim != self.base.dim or any(col[0] == 'mix' or len(col) != 3 for col in cols):
            return None
        d = self.dim
        rng = np.random.RandomState(12345)
        try:
            f = lambda pts: np.asarray(self.transform(np.asarray(pts, dtype=np.float64)), dtype=np.float64).reshape(len(pts), -1)   # noqa: E731
            x0 = np.zeros((1, d))                            # probing at the origin keeps exact coefficients exact
            y0 = f(x0)
            if y0.shape != (1, d):
                return None
            scale = np.zeros(d)
            for k in range(d):                               # one-column perturbations: slope, and no cross-talk
                x1 = x0.copy(); x1[0, k] += 1.0
                dy = f(x1) - y0
                scale[k] = dy[0, k]
                if np.abs(np.delete(dy[0], k)).max(initial=0.0) > 1e-12 * max(1.0, np.abs(y0).max()):
                    return None
            shift = y0[0] - scale * x0[0]
            pts = rng.uniform(-3.0, 3.0, size=(64, d))       # affine everywhere we look, batch size does not matter
            if np.abs(f(pts) - (pts * scale + shift)).max() > 1e-12 * max(1.0, np.abs(pts * scale + shift).max()):
                return None
        except Exception:                                    # noqa: BLE001
            return None
        out = []
        for (kind, a, b), sc, sh in zip(cols, scale, shift):
            if kind == COL_UNIFORM:
                out.append((COL_UNIFORM, sc * a + sh, sc * b + sh))
            elif kind == COL_NORMAL:
                if sc == 0.0:
                    out.append((COL_CONST, sh + 0.0 * a, 0.0))
                else:
                    out.append((COL_NORMAL, sc * a + sh, abs(sc) * b))
            else:
                out.append((COL_CONST, sc * a + sh, 0.0))
        return out


class _Truncated(Sampler):
    def __init__(self, base, high, low, expr, prob, max_iters, sample_anyway):
        self.base, self.dim, self.weight = base, base.dim, base.weight
        self.high, self.low, self.expr = high, low, expr
        self.prob, self.max_iters, self.sample_anyway = float(prob), max_iters, sample_anyway

    def _keep(self, pts):
        vals = pts if self.expr is None else np.asarray(self.expr(pts)).reshape(len(pts), -1)
        ok = np.ones(len(pts), dtype=bool)
        if self.high is not None:
            ok &= np.all(vals <= np.asarray(self.high, dtype=np.float64), axis=1)
        if self.low is not None:
            ok &= np.all(vals >= np.asarray(self.low, dtype=np.float64), axis=1)
        return pts[ok]

    def device_columns(self):
        """ Box truncation of independent columns (no `expr`) factorises: a truncated uniform column is the uniform
        on the intersection of the intervals, a truncated normal column is drawn by rejection in the kernel
        (PINN_COL_TNORMAL), a constant either survives or makes the sampler empty. """
        if self.expr is not None:
            return None
        cols = self.base.device_columns()
        if cols is None or any(col[0] == 'mix' or len(col) != 3 for col in cols):
            return None
        try:
            hi = np.broadcast_to(np.asarray(np.inf if self.high is None else self.high, dtype=np.float64), (self.dim,))
            lo = np.broadcast_to(np.asarray(-np.inf if self.low is None else self.low, dtype=np.float64), (self.dim,))
        except ValueError:
            return None
        out = []
        for (kind, a, b), l, h in zip(cols, lo, hi):
            if kind == COL_UNIFORM:
                lo_k, hi_k = max(min(a, b), l), min(max(a, b), h)
                if not lo_k < hi_k:
                    return None                              # empty (or degenerate): let the host path report it
                out.append((COL_UNIFORM, float(lo_k), float(hi_k)))
            elif kind == COL_NORMAL:
                if not l < h:
                    return None
                big = 3.0e38
                out.append((COL_TNORMAL, float(a), float(b), float(max(l, -big)), float(min(h, big))))
            else:
                if not l <= a <= h:
                    return None
                out.append((COL_CONST, float(a), 0.0))
        return out

    def sample(self, size):
        got, have, rounds, pts = [], 0, 0, None
        while have < size:
            if self.max_iters is not None and rounds >= self.max_iters:
                if not self.sample_anyway:
                    raise ValueError('truncate: %d rounds were not enough for %d points (raise prob / max_iters)'
                                     % (rounds, size))
                got.append(pts[:size - have])             # pad with unfiltered points of the last round
                have = size
                break
            pts = np.asarray(self.base.sample(max(int(np.ceil((size - have) / self.prob)), 1)),
                             dtype=np.float64).reshape(-1, self.dim)
            kept = self._keep(pts)
            got.append(kept)
            have += len(kept)
            rounds += 1
        return np.concatenate(got)[:size]


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


class ScipySampler(Sampler):
    """ `ScipySampler('beta', a=2, b=5)`: draws from `scipy.stats.<name>(**kwargs).rvs`. """

    def __init__(self, name, seed=None, dim=1, **kwargs):
        import scipy.stats
        if not hasattr(scipy.stats, name):
            raise ValueError('scipy.stats has no distribution %r' % name)
        self.name, self.dim, self.state = name, dim, np.random.RandomState(seed)
        self.dist = getattr(scipy.stats, name)(**kwargs)

    def sample(self, size):
        return np.asarray(self.dist.rvs(size=(size, self.dim), random_state=self.state), dtype=np.float64).reshape(size, -1)


class HistoSampler(Sampler):
    """ Sampler of a `numpy.histogramdd`-style histogram: picks a bin with probability proportional to its
    count, then a uniform point inside it.  `HistoSampler(histo=(counts, edges))` or `edges=` + `.update(points)`. """

    def __init__(self, histo=None, edges=None, seed=None):
        if histo is not None:
            counts, edges = histo
        elif edges is not None:
            edges = [edges] if np.ndim(edges[0]) == 0 else edges
            counts = np.zeros([len(e) - 1 for e in edges], dtype=np.float64)
        else:
            raise ValueError('HistoSampler needs `histo` or `edges`')
        self.edges = [np.asarray(e, dtype=np.float64) for e in ([edges] if np.ndim(edges[0]) == 0 else edges)]
        self.counts = np.asarray(counts, dtype=np.float64).reshape([len(e) - 1 for e in self.edges])
        self.dim = len(self.edges)
        self.state = np.random.RandomState(seed)

    def update(self, points):
        points = np.asarray(points, dtype=np.float64).reshape(-1, self.dim)
        self.counts += np.histogramdd(points, bins=self.edges)[0]

    def sample(self, size):
        total = self.counts.sum()
        if total <= 0:
            raise ValueError('HistoSampler: the histogram is empty')
        flat = self.state.choice(self.counts.size, size=size, p=self.counts.reshape(-1) / total)
        idx = np.unravel_index(flat, self.counts.shape)
        cols = [e[i] + self.state.uniform(size=size) * (e[i + 1] - e[i]) for e, i in zip(self.edges, idx)]
        return np.stack(cols, axis=1)
