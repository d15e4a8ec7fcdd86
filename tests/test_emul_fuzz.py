""" Randomised cross-check of the per-thread device code (host build, tests/emul) against the fp64 autograd
oracle: random network shapes (widths 1..40, so every output-block / input-block remainder and several blocks
per layer occur), activations, residual layouts, ansatz configurations, domains and equations.  CPU only. """
import numpy as np
import pytest
import torch

import emul_harness as E
import problems as P
from helpers import rel_l2
from oracle import autograd_port as ap
from pydens_b200 import _native as N
from pydens_b200 import tracer as T

ACTS = ['Tanh', 'Sigmoid', P.Sin, 'Softplus', 'SiLU', 'GELU']


def _equations(total, ndims):
    """ (name, callable(u, *xs, D, V)) candidates for a problem with `total` point columns. """
    eqs = [('first', lambda u, *xs, D, V: D(u, xs[0]) - torch.cos(xs[0]) * u),
           ('second', lambda u, *xs, D, V: D(D(u, xs[0]), xs[0]) + 0.5 * D(u, xs[0]) - u ** 2 + 1.0),
           ('var', lambda u, *xs, D, V: D(u, xs[0]) * V('k', 0.7) - torch.sin(xs[0]) + V('k', 0.7) ** 2)]
    if total >= 2:
        j = total - 1
        eqs += [('laplace', lambda u, *xs, D, V: D(D(u, xs[0]), xs[0]) + D(D(u, xs[j]), xs[j]) - xs[j] * u),
                ('mixed', lambda u, *xs, D, V: D(D(u, xs[0]), xs[j]) + D(u, xs[j]) * u - 0.3),
                ('advect', lambda u, *xs, D, V: D(u, xs[j]) + xs[0] * D(u, xs[0]) - torch.exp(-u))]
    if total >= 3:
        eqs += [('three', lambda u, *xs, D, V: D(D(u, xs[0]), xs[0]) + D(D(u, xs[1]), xs[1]) - D(u, xs[2])
                 + 0.1 * D(D(u, xs[0]), xs[1]))]
    return eqs


def _random_problem(seed):
    rng = np.random.RandomState(seed)
    ndims = int(rng.randint(1, 4))
    nparams = int(rng.randint(0, 2)) if ndims < 3 else 0
    total = ndims + nparams
    depth = int(rng.randint(1, 5))                               # hidden layers
    widths = [int(rng.choice([1, 2, 3, 4, 5, 7, 8, 9, 12, 15, 16, 17, 23, 31, 32, 33, 40])) for _ in range(depth)]
    acts = [ACTS[int(rng.randint(len(ACTS)))] for _ in range(depth)]
    layout, skip_done = '', False
    for l in range(depth):
        # a residual block 'R fa+' needs equal widths on both ends of the skip
        if l >= 1 and not skip_done and widths[l] == widths[l - 1] and rng.rand() < 0.7:
            layout += 'R fa+ '
            skip_done = True
        elif l >= 1 and not skip_done and rng.rand() < 0.35:
            widths[l] = widths[l - 1]
            layout += 'R fa+ '
            skip_done = True
        else:
            layout += 'fa '
    layout += 'f'
    has_ic = bool(rng.rand() < 0.5)
    nsp = ndims - 1 if has_ic else ndims
    ic = None
    if has_ic:
        kind = int(rng.randint(3))
        if kind == 0 or nsp == 0:
            ic = float(np.round(rng.uniform(-1, 2), 2))
        elif kind == 1:
            ic = lambda *x: torch.sin(2.0 * x[0]) + 0.5
        else:
            ic = lambda *x: x[0] * (1.0 - x[-1]) + 0.25
    bc = float(np.round(rng.uniform(-1, 1), 2)) if (nsp > 0 and rng.rand() < 0.6) else None
    domain = [(float(np.round(rng.uniform(-1, 0.2), 2)), float(np.round(rng.uniform(0.8, 2.5), 2))) for _ in range(ndims)]
    eqs = _equations(total, ndims)
    name, eq = eqs[int(rng.randint(len(eqs)))]
    ranges = domain + [(0.5, 2.0)] * nparams
    return dict(ndims=ndims, nparams=nparams, total=total, features=widths + [1], acts=acts, layout=layout,
                ic=ic, bc=bc, domain=domain, eq=eq, eq_name=name, ranges=ranges, variables={'k': 0.7} if name == 'var' else None,
                log_scale=float(np.round(rng.uniform(-0.5, 0.5), 2)))


def _layer_plan(cfg):
    acts, skips, stack, i_a = [], [], [], 0
    names = [(a if isinstance(a, str) else a.__name__).lower() for a in cfg['acts']]
    for letter in cfg['layout'].replace(' ', ''):
        if letter == 'f':
            acts.append('none'); skips.append(None)
        elif letter == 'a':
            acts[-1] = names[i_a]; i_a += 1
        elif letter == 'R':
            stack.append(len(acts) - 1)
        elif letter == '+':
            skips[-1] = stack.pop()
    return acts, skips


@pytest.mark.parametrize('seed', list(range(120)))
def test_random_problem_matches_fp64_oracle(seed):
    cfg = _random_problem(seed)
    sym_V = lambda n, init: T.Sym(T.var(n))
    nsp = cfg['ndims'] - 1 if cfg['ic'] is not None else cfg['ndims']
    traced = T.trace(lambda u, *xs: cfg['eq'](u, *xs, D=T.sym_D, V=sym_V), cfg['total'], None,
                     initial_condition=cfg['ic'], ndims_spatial=nsp)
    acts, skips = _layer_plan(cfg)
    spec = N.build_spec([cfg['total']] + cfg['features'], acts, cfg['ndims'], cfg['nparams'], cfg['bc'] is not None,
                        cfg['bc'] if cfg['bc'] is not None else 0.0, cfg['ic'] is not None, cfg['domain'], traced,
                        skips=skips)

    prob = ap.Problem(cfg['eq'], ndims=cfg['ndims'], nparams=cfg['nparams'], initial_condition=cfg['ic'],
                      boundary_condition=cfg['bc'], domain=cfg['domain'], features=cfg['features'],
                      activation=cfg['acts'], dtype=torch.float64, variables=cfg['variables'], seed=seed,
                      layout=cfg['layout'])
    with torch.no_grad():
        prob.log_scale.fill_(cfg['log_scale'])
    params = prob.flat_params().numpy()
    assert spec.n_params == params.size, (spec.n_params, params.size, cfg['layout'], cfg['features'])

    rng = np.random.RandomState(1000 + seed)
    n = int(rng.choice([1, 5, 33, 70]))
    pts = np.concatenate([rng.uniform(lo, hi, size=(n, 1)) for lo, hi in cfg['ranges']], axis=1).astype(np.float32)
    loss, residual, grads = E.emul_step(spec, params.astype(np.float32), pts)
    prob.load_flat(torch.from_numpy(params.astype(np.float32).astype(np.float64)))
    ref_loss, ref_res, ref_grads = prob.loss_and_grads(pts.astype(np.float64))
    tag = '%s %s %s acts=%s' % (cfg['eq_name'], cfg['layout'], cfg['features'], acts)
    # a residual far below its O(1) terms is a cancellation: fp32 rounding of the terms (~1e-7 absolute) then shows
    # up magnified in every RELATIVE error, for torch's fp32 as well — the tolerances are relaxed by that factor
    cond = max(1.0, 0.05 / max(float(np.sqrt(np.mean(np.square(ref_res)))), 1e-30))
    assert abs(loss - ref_loss) <= 2e-5 * cond * max(abs(ref_loss), 1e-6), tag
    assert rel_l2(residual, ref_res) <= 2e-5 * cond, tag
    assert rel_l2(grads, ref_grads.numpy()) <= 1e-4 * cond, tag
    u = E.emul_forward(spec, params.astype(np.float32), pts)
    ref_u = prob.predict(pts.astype(np.float64))
    # u is a sum of O(1) terms that may cancel: the error is measured against that scale, not against |u|
    assert np.abs(u - ref_u).max() <= 1e-5 * max(1.0, np.abs(ref_u).max()), tag


# ---------------------------------------------------------------------------------------------------------------
# five / six derivative directions (kernels step_kernel<5,5> / <6,6>: every direction carries its second derivative)
# ---------------------------------------------------------------------------------------------------------------
def _many_direction_equations(ndims):
    """ (name, callable) candidates whose jet needs 5 or 6 directions in a problem with `ndims` variables. """
    eqs = []
    if ndims == 3:
        eqs += [('hessian3', lambda u, x, y, z, D, V: D(D(u, x), x) + D(D(u, y), y) * 0.5 + D(D(u, z), z) + D(D(u, x), y)
                 - 0.7 * D(D(u, y), z) + 0.2 * D(D(u, x), z) * u - torch.cos(x * z)),
                ('two_diagonals', lambda u, x, y, z, D, V: D(D(u, x), y) + D(D(u, y), z) * V('k', 0.7) + D(u, x) * u - y)]
    if ndims == 4:
        eqs += [('diag_plus_time', lambda u, x, y, z, t, D, V: D(u, t) - D(D(u, x), x) - D(D(u, y), y) - D(D(u, z), z)
                 - 0.4 * D(D(u, x), z) + u ** 2),
                ('wave_cross', lambda u, x, y, z, t, D, V: D(D(u, t), t) - D(D(u, x), x) - D(D(u, y), y) - D(D(u, z), z)
                 + 0.3 * D(D(u, x), y) - torch.sin(t))]
    if ndims == 5:
        eqs += [('heat4', lambda u, a, b, c, d, t, D, V: D(u, t) - 0.2 * (D(D(u, a), a) + D(D(u, b), b) + D(D(u, c), c) + D(D(u, d), d)) + a * u),
                ('lap5', lambda u, a, b, c, d, e, D, V: D(D(u, a), a) + D(D(u, b), b) + D(D(u, c), c) + D(D(u, d), d) + D(D(u, e), e)
                 - torch.exp(-u) * V('k', 0.7)),
                ('first5', lambda u, a, b, c, d, e, D, V: D(u, a) + b * D(u, b) - D(u, c) * D(u, d) + D(u, e) * u - 0.1)]
    if ndims == 6:
        eqs += [('lap6', lambda u, a, b, c, d, e, g, D, V: D(D(u, a), a) + D(D(u, b), b) + D(D(u, c), c) + D(D(u, d), d)
                 + D(D(u, e), e) + D(D(u, g), g) - u * torch.cos(a + g)),
                ('heat5', lambda u, a, b, c, d, e, t, D, V: D(u, t) * (1.0 + a) - D(D(u, a), a) - D(D(u, b), b) - D(D(u, c), c)
                 - D(D(u, d), d) - D(D(u, e), e))]
    return eqs


def _random_many_direction_problem(seed):
    rng = np.random.RandomState(50000 + seed)
    ndims = int(rng.randint(3, 7))
    depth = int(rng.randint(1, 4))
    widths = [int(rng.choice([1, 3, 4, 7, 8, 9, 13, 16, 17, 24])) for _ in range(depth)]
    acts = [ACTS[int(rng.randint(len(ACTS)))] for _ in range(depth)]
    layout = ''
    for l in range(depth):
        if l >= 1 and 'R' not in layout and rng.rand() < 0.4:
            widths[l] = widths[l - 1]
            layout += 'R fa+ '
        else:
            layout += 'fa '
    layout += 'f'
    eqs = _many_direction_equations(ndims)
    name, eq = eqs[int(rng.randint(len(eqs)))]
    has_time = name in ('diag_plus_time', 'wave_cross', 'heat4', 'heat5')
    has_ic = has_time and bool(rng.rand() < 0.7)
    nsp = ndims - 1 if has_ic else ndims
    ic = None
    if has_ic:
        ic = (lambda *x: torch.sin(2.0 * x[0]) * x[1] + 0.5 * x[-1]) if rng.rand() < 0.6 else float(np.round(rng.uniform(-1, 2), 2))
    bc = float(np.round(rng.uniform(-1, 1), 2)) if rng.rand() < 0.6 else None
    domain = [(float(np.round(rng.uniform(-1, 0.2), 2)), float(np.round(rng.uniform(0.8, 2.5), 2))) for _ in range(ndims)]
    return dict(ndims=ndims, nparams=0, total=ndims, features=widths + [1], acts=acts, layout=layout, ic=ic, bc=bc,
                domain=domain, eq=eq, eq_name=name, ranges=domain, variables={'k': 0.7} if name in ('two_diagonals', 'lap5') else None,
                log_scale=float(np.round(rng.uniform(-0.5, 0.5), 2)))


@pytest.mark.parametrize('seed', list(range(40)))
def test_random_many_direction_problem_matches_fp64_oracle(seed):
    cfg = _random_many_direction_problem(seed)
    sym_V = lambda n, init: T.Sym(T.var(n))
    nsp = cfg['ndims'] - 1 if cfg['ic'] is not None else cfg['ndims']
    traced = T.trace(lambda u, *xs: cfg['eq'](u, *xs, D=T.sym_D, V=sym_V), cfg['total'], None,
                     initial_condition=cfg['ic'], ndims_spatial=nsp)
    assert traced.nf in (5, 6) and traced.ns == traced.nf, cfg['eq_name']
    acts, skips = _layer_plan(cfg)
    spec = N.build_spec([cfg['total']] + cfg['features'], acts, cfg['ndims'], 0, cfg['bc'] is not None,
                        cfg['bc'] if cfg['bc'] is not None else 0.0, cfg['ic'] is not None, cfg['domain'], traced,
                        skips=skips)
    prob = ap.Problem(cfg['eq'], ndims=cfg['ndims'], nparams=0, initial_condition=cfg['ic'],
                      boundary_condition=cfg['bc'], domain=cfg['domain'], features=cfg['features'],
                      activation=cfg['acts'], dtype=torch.float64, variables=cfg['variables'], seed=seed,
                      layout=cfg['layout'])
    with torch.no_grad():
        prob.log_scale.fill_(cfg['log_scale'])
    params = prob.flat_params().numpy()
    assert spec.n_params == params.size
    rng = np.random.RandomState(2000 + seed)
    n = int(rng.choice([1, 7, 33, 64]))
    pts = np.concatenate([rng.uniform(lo, hi, size=(n, 1)) for lo, hi in cfg['ranges']], axis=1).astype(np.float32)
    loss, residual, grads = E.emul_step(spec, params.astype(np.float32), pts)
    prob.load_flat(torch.from_numpy(params.astype(np.float32).astype(np.float64)))
    ref_loss, ref_res, ref_grads = prob.loss_and_grads(pts.astype(np.float64))
    tag = '%s %s %s acts=%s' % (cfg['eq_name'], cfg['layout'], cfg['features'], acts)
    cond = max(1.0, 0.05 / max(float(np.sqrt(np.mean(np.square(ref_res)))), 1e-30))
    assert abs(loss - ref_loss) <= 2e-5 * cond * max(abs(ref_loss), 1e-6), tag
    assert rel_l2(residual, ref_res) <= 2e-5 * cond, tag
    assert rel_l2(grads, ref_grads.numpy()) <= 1e-4 * cond, tag
    u = E.emul_forward(spec, params.astype(np.float32), pts)
    ref_u = prob.predict(pts.astype(np.float64))
    assert np.abs(u - ref_u).max() <= 1e-5 * max(1.0, np.abs(ref_u).max()), tag


# ---------------------------------------------------------------------------------------------------------------
# derivatives of order 3 / 4 (kernels hi_step_kernel<NF, K>: every direction carries its whole Taylor jet)
# ---------------------------------------------------------------------------------------------------------------
def _high_order_equations(ndims):
    D3 = lambda D, u, x: D(D(D(u, x), x), x)
    D4 = lambda D, u, x: D(D(D(D(u, x), x), x), x)
    eqs = [('ode3', lambda u, *xs, D, V: D3(D, u, xs[0]) + D(u, xs[0]) * u - torch.cos(xs[0])),
           ('ode4', lambda u, *xs, D, V: D4(D, u, xs[0]) - 2.0 * D(D(u, xs[0]), xs[0]) + u ** 2 - xs[0]),
           ('ode3var', lambda u, *xs, D, V: D3(D, u, xs[0]) * V('k', 0.7) + D(D(u, xs[0]), xs[0]) - V('k', 0.7) * torch.sin(u))]
    if ndims >= 2:
        j = ndims - 1
        eqs += [('kdv', lambda u, *xs, D, V: D(u, xs[j]) + 6.0 * u * D(u, xs[0]) + D3(D, u, xs[0])),
                ('beam', lambda u, *xs, D, V: D(D(u, xs[j]), xs[j]) + 0.5 * D4(D, u, xs[0]) - torch.sin(xs[0])),
                ('ks', lambda u, *xs, D, V: D(u, xs[j]) + u * D(u, xs[0]) + D(D(u, xs[0]), xs[0]) + D4(D, u, xs[0])),
                ('third_in_time', lambda u, *xs, D, V: D3(D, u, xs[j]) + D(D(u, xs[0]), xs[0]) * xs[j] - u)]
        D2 = lambda D, u, x: D(D(u, x), x)
        eqs += [('biharmonic', lambda u, *xs, D, V: D4(D, u, xs[0]) + 2.0 * D2(D, D2(D, u, xs[0]), xs[j]) + D4(D, u, xs[j])
                 - torch.sin(xs[0]) * torch.cos(xs[j])),
                ('mixed3', lambda u, *xs, D, V: D(D2(D, u, xs[0]), xs[j]) - 0.5 * D(D2(D, u, xs[j]), xs[0]) + D(D(u, xs[0]), xs[j]) * u
                 + D3(D, u, xs[0]))]
    if ndims >= 3:
        eqs += [('plate', lambda u, *xs, D, V: D(D(u, xs[2]), xs[2]) + 0.1 * (D4(D, u, xs[0]) + D4(D, u, xs[1])) + D3(D, u, xs[0]) * u),
                ('three3', lambda u, *xs, D, V: D3(D, u, xs[0]) + D3(D, u, xs[1]) - D(u, xs[2]) + xs[1] * D(D(u, xs[0]), xs[0]))]
    return eqs


def _random_high_order_problem(seed):
    rng = np.random.RandomState(90000 + seed)
    ndims = int(rng.randint(1, 4))
    nparams = int(rng.randint(0, 2)) if ndims < 3 else 0
    total = ndims + nparams
    depth = int(rng.randint(1, 4))                               # hidden layers (the oracle's nested autograd needs one)
    widths = [int(rng.choice([1, 2, 3, 4, 5, 7, 8, 9, 12, 16, 17, 23])) for _ in range(depth)]
    acts = [ACTS[int(rng.randint(len(ACTS)))] for _ in range(depth)]
    eqs = _high_order_equations(ndims)
    name, eq = eqs[int(rng.randint(len(eqs)))]
    has_ic = ndims >= 2 and bool(rng.rand() < 0.6)
    if name in ('ode3', 'ode4', 'ode3var') and ndims == 1:
        has_ic = bool(rng.rand() < 0.4)
    nsp = ndims - 1 if has_ic else ndims
    ic = None
    if has_ic:
        if nsp == 0 or rng.rand() < 0.3:
            ic = float(np.round(rng.uniform(-1, 2), 2))
        elif rng.rand() < 0.5:
            ic = lambda *x: torch.sin(2.0 * x[0]) + 0.5
        else:
            ic = lambda *x: x[0] * (1.0 - x[-1]) * x[0] + 0.25 * torch.exp(-x[-1])
    bc = float(np.round(rng.uniform(-1, 1), 2)) if (nsp > 0 and rng.rand() < 0.6) else None
    domain = [(float(np.round(rng.uniform(-1, 0.2), 2)), float(np.round(rng.uniform(0.8, 2.5), 2))) for _ in range(ndims)]
    # residual blocks 'R fa+' (equal widths on both ends of the skip), possibly chained; drawn from a generator of its own
    # so that the problems of the plain-chain seeds stay what they were
    rng2 = np.random.RandomState(170000 + seed)
    layout = 'fa'
    for l in range(1, depth):
        if rng2.rand() < 0.4:
            widths[l] = widths[l - 1]
            layout += ' R fa+'
        else:
            layout += ' fa'
    layout += ' f'
    return dict(ndims=ndims, nparams=nparams, total=total, features=widths + [1], acts=acts, layout=layout, ic=ic, bc=bc,
                domain=domain, eq=eq, eq_name=name, ranges=domain + [(0.5, 2.0)] * nparams,
                variables={'k': 0.7} if name == 'ode3var' else None, log_scale=float(np.round(rng.uniform(-0.5, 0.5), 2)))


@pytest.mark.parametrize('seed', list(range(60)))
def test_random_high_order_problem_matches_fp64_oracle(seed):
    cfg = _random_high_order_problem(seed)
    sym_V = lambda n, init: T.Sym(T.var(n))
    nsp = cfg['ndims'] - 1 if cfg['ic'] is not None else cfg['ndims']
    traced = T.trace(lambda u, *xs: cfg['eq'](u, *xs, D=T.sym_D, V=sym_V), cfg['total'], None,
                     initial_condition=cfg['ic'], ndims_spatial=nsp)
    assert traced.order in (3, 4) and traced.ns == 0 and traced.channels == 1 + traced.nf * traced.order, cfg['eq_name']
    acts, skips = _layer_plan(cfg)
    spec = N.build_spec([cfg['total']] + cfg['features'], acts, cfg['ndims'], cfg['nparams'], cfg['bc'] is not None,
                        cfg['bc'] if cfg['bc'] is not None else 0.0, cfg['ic'] is not None, cfg['domain'], traced,
                        skips=skips)
    assert spec.order == traced.order
    prob = ap.Problem(cfg['eq'], ndims=cfg['ndims'], nparams=cfg['nparams'], initial_condition=cfg['ic'],
                      boundary_condition=cfg['bc'], domain=cfg['domain'], features=cfg['features'],
                      activation=cfg['acts'] or 'Tanh', dtype=torch.float64, variables=cfg['variables'], seed=seed,
                      layout=cfg['layout'])
    with torch.no_grad():
        prob.log_scale.fill_(cfg['log_scale'])
    params = prob.flat_params().numpy()
    assert spec.n_params == params.size
    rng = np.random.RandomState(4000 + seed)
    n = int(rng.choice([1, 7, 33, 64]))
    pts = np.concatenate([rng.uniform(lo, hi, size=(n, 1)) for lo, hi in cfg['ranges']], axis=1).astype(np.float32)
    loss, residual, grads = E.emul_step(spec, params.astype(np.float32), pts)
    prob.load_flat(torch.from_numpy(params.astype(np.float32).astype(np.float64)))
    ref_loss, ref_res, ref_grads = prob.loss_and_grads(pts.astype(np.float64))
    tag = '%s %s %s acts=%s' % (cfg['eq_name'], cfg['layout'], cfg['features'], acts)
    cond = max(1.0, 0.05 / max(float(np.sqrt(np.mean(np.square(ref_res)))), 1e-30))
    if cfg['eq_name'] in ('biharmonic', 'mixed3'):
        cond *= 5.0          # mixed derivatives by polarisation: (P_4 + M_4 - 2 u_xxxx - 2 u_yyyy) / 12 cancels leading digits
    assert abs(loss - ref_loss) <= 2e-5 * cond * max(abs(ref_loss), 1e-6), tag
    assert rel_l2(residual, ref_res) <= 2e-5 * cond, tag
    assert rel_l2(grads, ref_grads.numpy()) <= 1e-4 * cond, tag
    u = E.emul_forward(spec, params.astype(np.float32), pts)
    ref_u = prob.predict(pts.astype(np.float64))
    assert np.abs(u - ref_u).max() <= 1e-5 * max(1.0, np.abs(ref_u).max()), tag


# ---------------------------------------------------------------------------------------------------------------
# criteria other than MSE (tracer.apply_criterion) on random problems of all three families
# ---------------------------------------------------------------------------------------------------------------
def _criterion_problem(seed):
    rng = np.random.RandomState(250000 + seed)
    family = int(rng.randint(3))
    cfg = [_random_problem, _random_many_direction_problem, _random_high_order_problem][family](int(rng.randint(100000)))
    kind = ['l1', 'huber', 'smooth_l1'][int(rng.randint(3))]
    return cfg, kind, bool(rng.rand() < 0.3), int(rng.choice([5, 33, 70]))


@pytest.mark.parametrize('seed', list(range(60)))
def test_random_problem_with_other_criteria_matches_torch_criteria(seed):
    cfg, kind, use_sum, n = _criterion_problem(seed)
    sym_V = lambda n_, init: T.Sym(T.var(n_))
    nsp = cfg['ndims'] - 1 if cfg['ic'] is not None else cfg['ndims']
    nparams = cfg.get('nparams', 0)
    prob = ap.Problem(cfg['eq'], ndims=cfg['ndims'], nparams=nparams, initial_condition=cfg['ic'],
                      boundary_condition=cfg['bc'], domain=cfg['domain'], features=cfg['features'],
                      activation=cfg['acts'] or 'Tanh', dtype=torch.float64, variables=cfg['variables'], seed=seed,
                      layout=cfg['layout'])
    with torch.no_grad():
        prob.log_scale.fill_(cfg['log_scale'])
    params = prob.flat_params().numpy().astype(np.float32)
    rng = np.random.RandomState(260000 + seed)
    # points at least 1e-3 of the width away from the faces: next to a face the ORACLE's nested torch.prod backward
    # loses its digits (whole-jet paragraph)
    pts = np.concatenate([rng.uniform(lo + 1e-3 * (hi - lo), hi - 1e-3 * (hi - lo), size=(n, 1)) for lo, hi in cfg['ranges']],
                         axis=1).astype(np.float32)
    prob.load_flat(torch.from_numpy(params.astype(np.float64)))
    _, r64, _ = prob.loss_and_grads(pts.astype(np.float64))
    thr = float(np.float32(np.median(np.abs(r64))))               # both branches of Huber / SmoothL1 are met
    red = 'sum' if use_sum else 'mean'
    key, crit = {'l1': (('l1',), torch.nn.L1Loss(reduction=red)),
                 'huber': (('huber', thr), torch.nn.HuberLoss(delta=thr, reduction=red)),
                 'smooth_l1': (('smooth_l1', thr), torch.nn.SmoothL1Loss(beta=thr, reduction=red))}[kind]
    traced = T.trace(lambda u, *xs: cfg['eq'](u, *xs, D=T.sym_D, V=sym_V), cfg['total'], None,
                     initial_condition=cfg['ic'], ndims_spatial=nsp, criterion=key)
    acts, skips = _layer_plan(cfg)
    spec = N.build_spec([cfg['total']] + cfg['features'], acts, cfg['ndims'], nparams, cfg['bc'] is not None,
                        cfg['bc'] if cfg['bc'] is not None else 0.0, cfg['ic'] is not None, cfg['domain'], traced,
                        skips=skips)
    loss, _, grads = E.emul_step(spec, params, pts)
    weight = float(n) if use_sum else 1.0
    ref_loss, _, ref_grads = prob.loss_and_grads(pts.astype(np.float64), criterion=crit)
    tag = '%s %s %s %s acts=%s n=%d' % (kind, red, cfg['eq_name'], cfg['layout'], acts, n)
    # the criteria are not smooth: a residual within fp32 rounding of a kink (0 for L1, the threshold for the others)
    # lands on the other branch in fp32 — its weight in the gradient is 1 / n, which bounds what one such point can do
    near_kink = np.abs(np.abs(r64) - (0.0 if kind == 'l1' else thr)) <= 1e-5 * np.maximum(np.abs(r64), thr)
    slack = 1.0 + 1e4 * float(near_kink.mean()) * (1.0 if kind == 'l1' else 1e-5)
    cond = max(1.0, 0.05 / max(float(np.sqrt(np.mean(np.square(r64)))), 1e-30))
    if cfg['eq_name'] in ('biharmonic', 'mixed3'):
        cond *= 5.0
    assert abs(loss * weight - ref_loss) <= 2e-5 * cond * max(abs(ref_loss), 1e-6), tag
    assert rel_l2(grads * np.float32(weight), ref_grads.numpy()) <= 1e-4 * cond * slack, tag
