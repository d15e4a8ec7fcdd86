""" Device sampler restatement: Philox4x32-10 known-answer vectors (integer work: bit-exact), host
emulation of the device code == numpy oracle, distribution checks, NumpySampler host behaviour. """
import numpy as np

from oracle import philox as ph
import emul_harness as E
from pydens_b200 import NumpySampler, ConstantSampler


def _kat(c, k):
    r = ph.philox4x32_10(*[np.array([x], dtype=np.uint32) for x in c], *[np.array([x], dtype=np.uint32) for x in k])
    return [int(x[0]) for x in r]


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    assert _kat([0] * 4, [0] * 2) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert _kat([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _kat([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_device_code_matches_numpy_restatement_bit_exact():
    for total, seed, step, off, n in [(2, 123, 7, 1000, 1000), (4, 2 ** 40 + 5, 0, 0, 257),
                                      (7, 9, (3 << 32) | 11, 2 ** 33, 513), (1, 0, 2 ** 47, 5, 64)]:
        a = ph.sample(None, total, seed, step, off, n)
        b = E.emul_sample(None, total, seed, step, off, n)
        assert a.dtype == b.dtype == np.float32
        assert np.array_equal(a, b)
        assert a.min() >= 0.0 and a.max() < 1.0
    cols = [(0, -1.0, 2.0), (1, 0.5, 2.0), (2, 3.0, 0.0), (0, 0.1, 4.0), (0, 0.0, 1.0), (1, 0.0, 1.0)]
    a = ph.sample(cols, 6, 99, 3, 2 ** 33, 4096)
    b = E.emul_sample(cols, 6, 99, 3, 2 ** 33, 4096)
    for k in (0, 2, 3, 4):
        assert np.array_equal(a[:, k], b[:, k])
    for k in (1, 5):                                   # normal columns go through libm
        assert np.abs(a[:, k] - b[:, k]).max() <= 2e-6


def test_stream_is_independent_of_sharding_and_varies_with_step():
    full = ph.sample(None, 3, 42, 5, 0, 1000)
    parts = np.concatenate([ph.sample(None, 3, 42, 5, 0, 250), ph.sample(None, 3, 42, 5, 250, 750)])
    assert np.array_equal(full, parts)
    assert not np.array_equal(full, ph.sample(None, 3, 42, 6, 0, 1000))
    assert not np.array_equal(full, ph.sample(None, 3, 43, 5, 0, 1000))


def test_distribution_moments():
    x = ph.sample([(0, 1.0, 5.0), (1, -2.0, 0.5)], 2, 7, 1, 0, 200000)
    assert abs(x[:, 0].mean() - 3.0) < 0.02 and abs(x[:, 0].var() - 16 / 12) < 0.02
    assert x[:, 0].min() >= 1.0 and x[:, 0].max() < 5.0
    assert abs(x[:, 1].mean() + 2.0) < 0.01 and abs(x[:, 1].std() - 0.5) < 0.01


def test_numpy_sampler_host_api():
    s = NumpySampler('u', seed=1) & NumpySampler('uniform', low=.5, high=5.5, seed=2)
    a = s.sample(700)
    assert a.shape == (700, 2) and a[:, 1].min() >= .5 and a[:, 1].max() <= 5.5
    assert s.device_columns() == [(0, 0.0, 1.0), (0, 0.5, 5.5)]
    s3 = NumpySampler('u', dim=2) & NumpySampler('n', loc=1.0, scale=2.0) & ConstantSampler(3.0)
    assert s3.sample(10).shape == (10, 4)
    assert s3.device_columns() == [(0, 0.0, 1.0), (0, 0.0, 1.0), (1, 1.0, 2.0), (2, 3.0, 0.0)]
    assert NumpySampler('exponential', scale=2.0).device_columns() is None


def test_sampler_algebra_on_the_host():
    """ The rest of the sampler algebra the reference re-exports (pydens/__init__.py:5): weights, mixtures,
    arithmetic, apply, truncate, scipy / histogram samplers. """
    from pydens_b200 import ScipySampler, HistoSampler
    u = NumpySampler('u', seed=3)
    # affine transforms of uniform / normal / constant columns still run in-kernel
    s = 2.0 * NumpySampler('u', low=1, high=2) + 1.0
    assert s.device_columns() == [(0, 3.0, 5.0)]
    x = s.sample(2000)
    assert x.shape == (2000, 1) and x.min() >= 3.0 and x.max() <= 5.0
    s = (1.0 - NumpySampler('n', loc=2.0, scale=0.5, seed=1)) / 2.0
    assert s.device_columns() == [(1, -0.5, 0.25)]
    x = s.sample(50000)
    assert abs(x.mean() + 0.5) < 0.01 and abs(x.std() - 0.25) < 0.01
    assert (-u).device_columns() == [(0, 0.0, -1.0)] and (ConstantSampler(2.0) * 3).device_columns() == [(2, 6.0, 0.0)]
    assert (np.array([1.0, 2.0]) * NumpySampler('u', dim=2)).device_columns() is None     # per-column factors: host
    x = (np.array([1.0, 2.0]) * NumpySampler('u', dim=2, seed=0) + NumpySampler('u', dim=2, seed=1)).sample(100)
    assert x.shape == (100, 2) and x[:, 1].max() <= 3.0
    # mixture with weights: 1/4 of the points from [0, 1), 3/4 from [10, 11)
    mix = 0.25 & NumpySampler('u', seed=5) | 0.75 & NumpySampler('u', low=10, high=11, seed=6)
    x = mix.sample(40000)
    assert x.shape == (40000, 1) and mix.device_columns()[0][0] == 'mix'          # runs in-kernel, too
    assert abs((x < 5).mean() - 0.25) < 0.01
    assert abs((x[:20000] < 5).mean() - 0.25) < 0.02      # shuffled, not blockwise
    tri = (NumpySampler('u', seed=1) | NumpySampler('u', low=2, high=3, seed=2)) | NumpySampler('u', low=4, high=5, seed=3)
    x = tri.sample(30000)
    assert abs((x < 1.5).mean() - 1 / 3) < 0.015 and abs((x > 3.5).mean() - 1 / 3) < 0.015
    # apply / truncate
    disc = NumpySampler('u', low=-1, high=1, dim=2, seed=7).truncate(high=1.0, expr=lambda p: (p ** 2).sum(axis=1), prob=0.7)
    x = disc.sample(5000)
    assert x.shape == (5000, 2) and ((x ** 2).sum(axis=1) <= 1.0).all()
    x = NumpySampler('n', dim=2, seed=8).truncate(low=[-1, 0], high=[1, 2]).sample(1000)
    assert x.shape == (1000, 2) and x[:, 0].min() >= -1 and x[:, 1].min() >= 0 and x[:, 1].max() <= 2
    import pytest
    with pytest.raises(ValueError):
        NumpySampler('u', seed=1).truncate(low=2.0, max_iters=3).sample(10)
    assert NumpySampler('u', seed=1).truncate(low=2.0, max_iters=2, sample_anyway=True).sample(10).shape == (10, 1)
    polar = NumpySampler('u', dim=2, seed=9).apply(lambda p: np.stack([p[:, 0] * np.cos(p[:, 1]), p[:, 0] * np.sin(p[:, 1]), p[:, 0]], axis=1))
    assert polar.dim == 3 and polar.sample(64).shape == (64, 3)
    # scipy / histogram samplers
    x = ScipySampler('beta', a=2.0, b=5.0, seed=1).sample(20000)
    assert x.shape == (20000, 1) and abs(x.mean() - 2 / 7) < 0.01
    h = HistoSampler(edges=[np.linspace(0, 1, 5), np.linspace(0, 2, 3)], seed=2)
    h.update(np.array([[0.1, 0.5], [0.1, 0.6], [0.9, 1.5]]))
    x = h.sample(9000)
    assert x.shape == (9000, 2) and abs((x[:, 0] < 0.25).mean() - 2 / 3) < 0.02
    assert ((x[:, 0] < 0.25) == (x[:, 1] < 1.0)).all()
    h1 = HistoSampler(histo=np.histogram(np.random.RandomState(0).normal(size=5000), bins=30), seed=3)
    assert abs(h1.sample(20000).mean()) < 0.05
    # anything with .sample(size) feeds Solver.fit (CPU: autograd path; on the GPU through pinned staging)
    import torch
    from pydens_b200 import Solver, D
    s = Solver(lambda f, x, e: D(f, x) - e * torch.cos(e * x), ndims=1, nparams=1, initial_condition=1.0, device='cpu')
    s.fit(niters=2, batch_size=16, sampler=NumpySampler('u') & (0.5 & NumpySampler('u', low=1, high=2) | NumpySampler('n', loc=4, scale=.1)))
    assert len(s.losses) == 2


def test_mixture_columns_in_kernel_form():
    """ `s1 | s2` with weights lowers to PINN_COL_MIXTURE columns: the device code (host build) is bit-exact
    with the numpy restatement, the component frequencies follow the weights, and the columns of one mixture
    switch together. """
    key = 'g0'
    cols = [('mix', key, [(1.0, 0, 0.0, 1.0), (3.0, 0, 10.0, 11.0)]),
            ('mix', key, [(1.0, 2, -5.0, 0.0), (3.0, 0, 20.0, 21.0)]),
            (0, 0.0, 1.0),
            ('mix', 'g1', [(0.2, 0, 0.0, 1.0), (0.3, 0, 2.0, 3.0), (0.5, 0, 4.0, 5.0)])]
    a = ph.sample(cols, 4, 77, 5, 2 ** 34, 20000)
    b = E.emul_sample(cols, 4, 77, 5, 2 ** 34, 20000)
    assert np.array_equal(a, b)
    left = a[:, 0] < 5
    assert abs(left.mean() - 0.25) < 0.01
    assert np.array_equal(left, a[:, 1] == -5.0)                   # same group: the whole row switches side
    assert abs((a[:, 3] < 1.5).mean() - 0.2) < 0.01 and abs((a[:, 3] > 3.5).mean() - 0.5) < 0.012
    assert abs(np.corrcoef(left, a[:, 3] > 3.5)[0, 1]) < 0.03     # different groups: independent draws
    normal_mix = [('mix', 0, [(1.0, 1, -3.0, 0.5), (1.0, 1, 3.0, 0.5)])]
    x = ph.sample(normal_mix, 1, 1, 0, 0, 50000)[:, 0]
    y = E.emul_sample(normal_mix, 1, 1, 0, 0, 50000)[:, 0]
    assert np.abs(x - y).max() <= 4e-6 and abs((x < 0).mean() - 0.5) < 0.01 and abs(np.abs(x).mean() - 3.0) < 0.02

    # the sampler algebra produces that form
    u = NumpySampler
    mix = 0.25 & u('u') | 0.75 & u('u', low=10, high=11)
    dc = mix.device_columns()
    assert len(dc) == 1 and dc[0][0] == 'mix' and dc[0][2] == [(0.25, 0, 0.0, 1.0), (0.75, 0, 10.0, 11.0)]
    rows = (u('u') & ConstantSampler(-5.0)) | 3.0 & (u('u', low=10, high=11) & u('n', loc=20, scale=0.1))
    dc = rows.device_columns()
    assert [c[1] for c in dc] == [id(rows)] * 2 and dc[1][2] == [(1.0, 2, -5.0, 0.0), (3.0, 1, 20.0, 0.1)]
    three = (u('u') | u('u', low=2, high=3)) | 2.0 & u('u', low=4, high=5)
    dc = three.device_columns()
    assert [round(c[0], 6) for c in dc[0][2]] == [1.0, 1.0, 2.0] and len({c[1] for c in dc}) == 1
    five = ((u('u') | u('u')) | (u('u') | u('u'))) | u('u')
    assert five.device_columns() is None                           # more than 4 components: host sampling
    both = mix & u('u') & (u('n') | ConstantSampler(1.0))
    dc = both.device_columns()
    assert [c[0] for c in dc] == ['mix', 0, 'mix'] and dc[0][1] != dc[2][1]
    assert (mix + 1.0).device_columns() is None and ((u('u') & mix) | u('u', dim=2)).device_columns() is None
    from pydens_b200 import _native as N
    arr = N.make_columns(both.device_columns(), 3)
    assert (arr[0].kind, arr[0].group, arr[0].n_comp, arr[2].group, arr[1].kind) == (3, 0, 2, 1, 0)
    assert abs(arr[0].cum_w[0] - 0.25) < 1e-7 and arr[0].cum_w[1] == 1.0


def test_random_column_tables_bit_exact_against_numpy_restatement():
    """ Random sampler tables (uniform / normal / constant / mixtures sharing or not sharing their group, random
    seeds, steps, offsets): the device code (host build) equals the numpy restatement bit for bit on every column
    that does not go through libm's log / cos, and within 4e-6 on those that do. """
    rng = np.random.RandomState(11)
    for case in range(60):
        total = int(rng.randint(1, 9))
        keys = ['g%d' % i for i in range(3)]
        cols, has_normal = [], []

        def simple():
            kind = int(rng.randint(3))
            a, b = float(np.round(rng.uniform(-3, 3), 3)), float(np.round(rng.uniform(0.1, 4), 3))
            return (kind, a, b if kind != 2 else 0.0)
        for k in range(total):
            if rng.rand() < 0.35:
                comps = [(float(np.round(rng.uniform(0.1, 2), 2)),) + simple() for _ in range(int(rng.randint(2, 5)))]
                cols.append(('mix', keys[int(rng.randint(len(keys)))], comps))
                has_normal.append(any(c[1] == 1 for c in comps))
            else:
                cols.append(simple())
                has_normal.append(cols[-1][0] == 1)
        seed, step = int(rng.randint(0, 2 ** 62)), int(rng.randint(0, 2 ** 47))
        off, n = int(rng.randint(0, 2 ** 40)), int(rng.choice([1, 33, 1000]))
        a = ph.sample(cols, total, seed, step, off, n)
        b = E.emul_sample(cols, total, seed, step, off, n)
        for k in range(total):
            if has_normal[k]:
                assert np.abs(a[:, k] - b[:, k]).max() <= 4e-6 * max(1.0, np.abs(a[:, k]).max()), (case, k)
            else:
                assert np.array_equal(a[:, k], b[:, k]), (case, k)


def test_truncated_and_transformed_samplers_lower_to_kernel_columns():
    """ batchflow `.truncate(high, low)` / `.apply(affine)` (tutorial cell `NS('u', dim=2) & ...`): box truncation of
    independent columns and per-column affine maps stay in-kernel; anything else stays on the host. """
    from pydens_b200 import NumpySampler as NS
    s = (NS('n', loc=0.5, scale=2.0).truncate(high=1.5, low=-1.0) & NS('u', low=0, high=4).truncate(high=3, low=1)
         & NS('u').apply(lambda x: 2 * x + 1) & NS('u', dim=2).truncate(high=[0.5, 0.9]))
    cols = s.device_columns()
    assert cols == [(4, 0.5, 2.0, -1.0, 1.5), (0, 1.0, 3.0), (0, 1.0, 3.0), (0, 0.0, 0.5), (0, 0.0, 0.9)]
    assert NS('u', dim=2).apply(lambda x: x[:, ::-1]).device_columns() is None           # mixes columns
    assert NS('u').apply(np.sin).device_columns() is None                                # not affine
    assert NS('u', dim=2).truncate(high=1.0, expr=lambda x: x.sum(axis=1)).device_columns() is None
    assert NS('u', low=0, high=1).truncate(low=2.0).device_columns() is None              # empty box
    assert (NS('n').truncate(low=0.0) * 3.0).device_columns() is None                    # arithmetic on a truncated column
    # the device stream of the truncated normal column == its numpy restatement (same rejection sequence) ...
    tn = [(4, 0.5, 2.0, -1.0, 1.5), (0, 0.0, 1.0), (4, 0.0, 1.0, 0.0, 3e38)]
    a = E.emul_sample(tn, 3, 99, 3, 2 ** 33, 30000)
    b = ph.sample(tn, 3, 99, 3, 2 ** 33, 30000)
    assert np.array_equal(a[:, 1], b[:, 1]) and np.abs(a - b).max() <= 4e-6
    assert a[:, 0].min() >= -1.0 and a[:, 0].max() <= 1.5 and a[:, 2].min() >= 0.0
    # ... and follows the truncated distribution the host sampler draws from
    host = NS('n', loc=0.5, scale=2.0, seed=3).truncate(high=1.5, low=-1.0).sample(30000)[:, 0]
    assert abs(a[:, 0].mean() - host.mean()) < 0.03 and abs(a[:, 0].std() - host.std()) < 0.03
    half = a[:, 2]                                            # half-normal: mean sqrt(2/pi)
    assert abs(half.mean() - np.sqrt(2 / np.pi)) < 0.02
