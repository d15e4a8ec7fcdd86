""" Parity of the CUDA fit step (through the C ABI) with the reference: committed goldens written by
the unmodified reference, the oracle port on fresh inputs, bit-exact sampler, size-independent
properties at BASELINE batch sizes.  Tolerances (fp32, stated by SURVEY.md 8c / BASELINE.md):
loss rel <= 1e-5, residual rel-L2 <= 1e-5, gradient rel-L2 <= 1e-4 (whole vector and per tensor). """
import numpy as np
import pytest
import torch

import problems as P
from helpers import load_golden, oracle_problem, rel_l2
from oracle import philox as ph

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from gpu_helpers import make_solver


# (the problems with five / six derivative directions have their own file, test_gpu_zz_directions.py)
@pytest.mark.parametrize('name', [n for n in P.PROBLEMS if n not in P.HI_DIRECTION + P.HI_ORDER])
def test_step_matches_reference_golden(name):
    g = load_golden(name)
    solver = make_solver(name, g['params'])
    eng = solver._get_engine()
    assert eng.n_params == g['params'].size
    loss, grads, residual = solver.loss_and_grads(g['points'])
    grads = grads.cpu().numpy()
    assert abs(loss - float(g['loss'])) <= 1e-5 * abs(float(g['loss']))
    assert rel_l2(residual.cpu().numpy(), g['residual']) <= 1e-5
    assert rel_l2(grads, g['grads']) <= 1e-4
    spec = eng.spec
    for l in range(spec.n_layers):
        w = slice(spec.w_off[l], spec.w_off[l] + spec.widths[l] * spec.widths[l + 1])
        b = slice(spec.b_off[l], spec.b_off[l] + spec.widths[l + 1])
        assert rel_l2(grads[w], g['grads'][w]) <= 1e-4, 'W%d' % l
        assert rel_l2(grads[b], g['grads'][b]) <= 1e-4, 'b%d' % l
    u = solver.predict(*[g['points'][:, i] for i in range(g['points'].shape[1])]).reshape(-1)
    assert rel_l2(u, g['u']) <= 1e-5


@pytest.mark.parametrize('name', ['poisson2d', 'heat2d', 'burgers', 'ode_var', 'wave3d', 'poisson_sin', 'burgers_silu',
                                  'wave1d_gelu', 'heat_softplus'])
def test_gradient_error_vs_fp64_no_worse_than_reference(name):
    g = load_golden(name)
    solver = make_solver(name, g['params'])
    _, grads, _ = solver.loss_and_grads(g['points'])
    prob = oracle_problem(name, torch.float64, g['params'].astype(np.float64))
    _, _, g64 = prob.loss_and_grads(g['points'].astype(np.float64))
    ours, ref = rel_l2(grads.cpu().numpy(), g64.numpy()), rel_l2(g['grads'], g64.numpy())
    assert ours <= max(4 * ref, 5e-6)


@pytest.mark.parametrize('n', [1, 2, 31, 32, 33, 1000, 4097])
def test_ragged_batches_against_oracle(n):
    g = load_golden('burgers')
    solver = make_solver('burgers', g['params'])
    prob = oracle_problem('burgers', torch.float32, g['params'])
    pts = P.make_points('burgers', n, seed=77)
    loss, grads, residual = solver.loss_and_grads(pts)
    l, r, gr = prob.loss_and_grads(pts)
    assert abs(loss - l) <= 1e-5 * abs(l)
    assert rel_l2(residual.cpu().numpy(), r) <= 1e-5
    assert rel_l2(grads.cpu().numpy(), gr.numpy()) <= 1e-4


def test_deterministic_run_to_run():
    g = load_golden('heat2d')
    solver = make_solver('heat2d', g['params'])
    pts = P.make_points('heat2d', 20000, seed=5)
    a = solver.loss_and_grads(pts)
    b = solver.loss_and_grads(pts)
    assert a[0] == b[0] and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])


def test_sampler_bit_exact_against_oracle():
    solver = make_solver('heat_param')
    eng = solver._get_engine()
    cols = [(0, 0.0, 1.0), (0, 0.0, 1.0), (0, 0.0, 0.5), (0, 0.1, 4.0)]
    for step, off, n in [(0, 0, 1000), (7, 12345, 4097), ((3 << 32) | 5, 2 ** 33, 513)]:
        dev = eng.sample(n, cols, step=step, point_offset=off).cpu().numpy()
        ref = ph.sample(cols, 4, eng.seed, step, off, n)
        assert np.array_equal(dev, ref)
    dev = eng.sample(2000, None, step=1).cpu().numpy()
    assert np.array_equal(dev, ph.sample(None, 4, eng.seed, 1, 0, 2000))
    ncols = [(1, 0.5, 2.0), (0, 0.0, 1.0), (2, 3.0, 0.0), (1, 0.0, 1.0)]
    dev = eng.sample(4096, ncols, step=2).cpu().numpy()
    ref = ph.sample(ncols, 4, eng.seed, 2, 0, 4096)
    assert np.array_equal(dev[:, 1:3], ref[:, 1:3]) and np.abs(dev - ref).max() <= 4e-6
    # mixture columns (batchflow `s1 | s2`): the component draw is integer work -> bit-exact
    mcols = [('mix', 'a', [(1.0, 0, 0.0, 1.0), (3.0, 0, 10.0, 11.0)]), ('mix', 'a', [(1.0, 2, -5.0, 0.0), (3.0, 0, 20.0, 21.0)]),
             (0, 0.0, 1.0), ('mix', 'b', [(0.2, 0, 0.0, 1.0), (0.3, 0, 2.0, 3.0), (0.5, 0, 4.0, 5.0)])]
    dev = eng.sample(8192, mcols, step=3, point_offset=2 ** 35).cpu().numpy()
    assert np.array_equal(dev, ph.sample(mcols, 4, eng.seed, 3, 2 ** 35, 8192))
    # and the fused step on a mixture sampler == the same step fed those points explicitly
    from pydens_b200 import _native as N
    n = 5000
    eng._step(None, N.make_columns(mcols, 4), n, 1.0 / n, 0, use_counter=False, step_value=4)
    torch.cuda.synchronize()
    sampled = eng.out.clone()
    eng._step(eng.sample(n, mcols, step=4), None, n, 1.0 / n, 0, use_counter=False, step_value=4)
    torch.cuda.synchronize()
    assert torch.equal(sampled, eng.out)


def test_in_kernel_sampling_equals_explicit_points():
    """ The fused step on in-kernel samples == the same step fed the sampled batch explicitly. """
    g = load_golden('poisson2d')
    solver = make_solver('poisson2d', g['params'])
    eng = solver._get_engine()
    n = 100000
    eng._step(None, None, n, 1.0 / n, 0, use_counter=False, step_value=9)
    torch.cuda.synchronize()
    sampled = eng.out.clone()
    pts = eng.sample(n, None, step=9)
    eng._step(pts, None, n, 1.0 / n, 0, use_counter=False, step_value=9)
    torch.cuda.synchronize()
    assert torch.equal(sampled, eng.out)
    # and the oracle agrees on that batch (full BASELINE cfg2 size)
    prob = oracle_problem('poisson2d', torch.float32, g['params'])
    l, _, gr = prob.loss_and_grads(pts.cpu().numpy())
    assert abs(float(sampled[eng.n_params]) - l) <= 1e-5 * abs(l)
    assert rel_l2(sampled[:eng.n_params].cpu().numpy(), gr.numpy()) <= 1e-4


@pytest.mark.parametrize('name,n', [('poisson2d', 100000), ('ode_param', 1000000), ('heat2d', 1000000)])
def test_full_size_additivity(name, n):
    """ Size-independent property at BASELINE sizes: with the global 1/N scale, the outputs of two
    half batches add up to the output of the whole batch (this is what the multi-GPU path relies on). """
    g = load_golden(name)
    solver = make_solver(name, g['params'])
    eng = solver._get_engine()
    pts = torch.from_numpy(P.make_points(name, n, seed=11)).cuda()
    h = n // 2 + 13
    eng._step(pts, None, n, 1.0 / n, 0, use_counter=False)
    whole = eng.out.clone()
    eng._step(pts[:h].contiguous(), None, h, 1.0 / n, 0, use_counter=False)
    a = eng.out.clone()
    eng._step(pts[h:].contiguous(), None, n - h, 1.0 / n, 0, use_counter=False)
    b = eng.out.clone()
    torch.cuda.synchronize()
    assert torch.isfinite(whole).all()
    np_ = eng.n_params
    assert abs(float(whole[np_] - (a + b)[np_])) <= 1e-5 * abs(float(whole[np_]))
    assert rel_l2((a + b)[:np_].cpu().numpy(), whole[:np_].cpu().numpy()) <= 1e-4


def test_sharded_offsets_reproduce_the_global_batch():
    """ Two 'ranks' with point offsets sample exactly the slices of the global batch. """
    g = load_golden('poisson2d')
    solver = make_solver('poisson2d', g['params'])
    eng = solver._get_engine()
    n = 50001
    eng._step(None, None, n, 1.0 / n, 0, use_counter=False, step_value=4)
    whole = eng.out.clone()
    from pydens_b200.engine import shard_batch
    acc = torch.zeros_like(whole)
    for r in range(2):
        ln, off = shard_batch(n, 2, r)
        eng._step(None, None, ln, 1.0 / n, off, use_counter=False, step_value=4)
        acc += eng.out
    torch.cuda.synchronize()
    np_ = eng.n_params
    assert abs(float(whole[np_] - acc[np_])) <= 1e-5 * abs(float(whole[np_]))
    assert rel_l2(acc[:np_].cpu().numpy(), whole[:np_].cpu().numpy()) <= 1e-4


def test_plan_info_and_errors():
    from pydens_b200 import _native
    solver = make_solver('poisson2d')
    eng = solver._get_engine()
    info = eng.info
    assert info.nf == 2 and info.ns == 2 and info.channels == 5
    assert info.flops_per_point == 6 * 5 * 335 and info.bytes_per_point == 8
    assert info.sm_count >= 100 and info.threads_per_cta % 32 == 0
    with pytest.raises(_native.NativeError):
        eng._step(None, None, 0, 1.0, 0)                                          # n_points must be > 0
