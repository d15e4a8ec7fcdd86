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
