""" Tracer / lowering: symbolic D, partial derivatives, register programs (CPU only). """
import numpy as np
import pytest
import torch

import problems as P
from pydens_b200 import tracer as T
import emul_harness as E


def eval_expr(e, env, memo=None):
    memo = {} if memo is None else memo
    if e in memo:
        return memo[e]
    k = e.kind
    if k == 'const': r = np.full_like(env['x'][0], e.value)
    elif k == 'coord': r = env['x'][e.value]
    elif k == 'u': r = env['u'][e.value]
    elif k == 'ch': r = env['ch'][e.value]
    elif k == 'var': r = np.full_like(env['x'][0], env['v'][e.value])
    elif k == 'powi': r = eval_expr(e.args[0], env, memo) ** e.value
    else:
        a = [eval_expr(x, env, memo) for x in e.args]
        r = {'add': lambda: a[0] + a[1], 'sub': lambda: a[0] - a[1], 'mul': lambda: a[0] * a[1],
             'div': lambda: a[0] / a[1], 'pow': lambda: a[0] ** a[1], 'neg': lambda: -a[0],
             'sin': lambda: np.sin(a[0]), 'cos': lambda: np.cos(a[0]), 'tan': lambda: np.tan(a[0]),
             'exp': lambda: np.exp(a[0]), 'log': lambda: np.log(a[0]), 'sqrt': lambda: np.sqrt(a[0]),
             'tanh': lambda: np.tanh(a[0]), 'sigmoid': lambda: 1 / (1 + np.exp(-a[0])),
             'abs': lambda: np.abs(a[0]), 'sign': lambda: np.sign(a[0])}[k]()
    memo[e] = r
    return r


@pytest.mark.parametrize('name', list(P.PROBLEMS))
def test_program_equals_expression_and_partials(name):
    tr = E.traced_problem(name)
    cfg = P.PROBLEMS[name]
    total, n = cfg['ndims'] + cfg['nparams'], 50
    rng = np.random.RandomState(0)
    coords = np.stack([rng.uniform(lo, hi, n) for lo, hi in cfg['ranges']])
    C = tr.channels
    ujet = rng.normal(size=(C, n))
    vvals = {nm: 0.7 + i for i, nm in enumerate(tr.var_names)}

    def env_of(uj):
        return {'x': coords, 'ch': uj, 'v': vvals}
    outs = T.run_program(tr.eq_prog, ujet, coords, [vvals[nm] for nm in tr.var_names])
    r0 = eval_expr(tr.residual, env_of(ujet))
    np.testing.assert_allclose(outs[0], r0, rtol=1e-12, atol=1e-12)
    # partials w.r.t. every jet channel: central differences of the residual expression
    for c in range(C):
        h = 1e-6
        up, um = ujet.copy(), ujet.copy()
        up[c] += h; um[c] -= h
        fd = (eval_expr(tr.residual, env_of(up)) - eval_expr(tr.residual, env_of(um))) / (2 * h)
        np.testing.assert_allclose(outs[1 + c], fd, rtol=1e-5, atol=1e-6)
    for i, nm in enumerate(tr.var_names):
        h = 1e-6
        vp, vm = dict(vvals), dict(vvals)
        vp[nm] += h; vm[nm] -= h
        e1 = env_of(ujet); e1['v'] = vp
        e2 = env_of(ujet); e2['v'] = vm
        fd = (eval_expr(tr.residual, e1) - eval_expr(tr.residual, e2)) / (2 * h)
        np.testing.assert_allclose(outs[1 + C + i], fd, rtol=1e-5, atol=1e-6)
    assert tr.n_slots <= T.MAX_SLOTS and len(tr.eq_prog) <= T.MAX_PROG


def test_ic_program_is_the_jet_of_ic():
    tr = E.traced_problem('heat2d')
    n = 40
    rng = np.random.RandomState(1)
    coords = rng.uniform(0, 1, size=(3, n))
    outs = T.run_program(tr.ic_prog, np.zeros((tr.channels, n)), coords, [])
    x, y = coords[0], coords[1]
    ic = 10 * x * y * (1 - x) * (1 - y)
    np.testing.assert_allclose(outs[0], ic, rtol=1e-12)
    # dirs = [x, y, t] with second order on x, y
    assert tr.dirs == [0, 1, 2] and tr.ns == 2 and tr.dir_vecs[2] == [0.0, 0.0, 1.0]
    np.testing.assert_allclose(outs[1], 10 * y * (1 - y) * (1 - 2 * x), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(outs[2], 10 * x * (1 - x) * (1 - 2 * y), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(outs[3], 0 * x, atol=1e-12)
    np.testing.assert_allclose(outs[4], -20 * y * (1 - y), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(outs[5], -20 * x * (1 - x), rtol=1e-10, atol=1e-12)


def test_not_lowerable_cases():
    D = T.sym_D
    with pytest.raises(T.NotLowerable):            # fifth order
        T.trace(lambda f, x: D(D(D(D(D(f, x), x), x), x), x), 1, None)
    with pytest.raises(T.NotLowerable):            # u_xxxy: not carried by the diagonals e_x +- e_y
        T.trace(lambda f, x, y: D(D(D(D(f, x), x), x), y), 2, None)
    with pytest.raises(T.NotLowerable):            # third order next to mixed derivatives of two pairs: 3 + 4 directions
        T.trace(lambda f, x, y, z: D(D(D(f, x), x), x) + D(D(f, x), y) + D(D(f, y), z), 3, None)
    with pytest.raises(T.NotLowerable):            # more than 6 directions (3 axes + 3 diagonals + t)
        T.trace(lambda f, x, y, z, t: D(D(f, x), y) + D(D(f, y), z) + D(D(f, x), z) + D(f, t), 4, None)
    with pytest.raises(T.NotLowerable):            # data-dependent branch
        T.trace(lambda f, x: f if x > 0 else -f, 1, None)
    with pytest.raises(T.NotLowerable):            # unsupported torch function
        T.trace(lambda f, x: torch.cumsum(f, 0), 1, None)
    with pytest.raises(T.NotLowerable):            # initial condition depending on the solution
        T.trace(lambda f, x: D(f, x), 1, None, initial_condition=lambda: T.Sym(T.uleaf()), ndims_spatial=0)


def test_five_and_six_directions_are_promoted_to_second_order():
    """ More than 4 directions: the library holds one kernel per direction count (NS = NF), so every direction
    carries its second-order channel; the residual's partials with respect to the promoted channels are zero. """
    D = T.sym_D
    tr = T.trace(lambda f, x, y, z: D(D(f, x), y) + D(D(f, y), z), 3, None)          # 3 axes + 2 diagonals
    assert (tr.nf, tr.ns) == (5, 5) and tr.dirs == [0, 1, 2, -1, -1]
    tr = T.trace(lambda f, x, y, z, w, t: D(D(f, x), x) + D(D(f, y), y) + D(D(f, z), z) + D(D(f, w), w) - D(f, t) * t, 5, None,
                 initial_condition=lambda x, y, z, w: x * y + z * w, ndims_spatial=4)
    assert (tr.nf, tr.ns) == (5, 5) and tr.dirs == [0, 1, 2, 3, 4] and tr.channels == 11
    # channel layout: 0 value, 1..5 first order, 6..10 second order; t (direction 4) has no second derivative in the
    # equation: d residual / d channel 10 is the constant 0
    rng = np.random.RandomState(0)
    coords, jet = rng.uniform(0.2, 0.9, size=(5, 7)), rng.normal(size=(11, 7))
    outs = T.run_program(tr.eq_prog, jet, coords, [])
    assert np.all(outs[1 + 10] == 0.0) and np.allclose(outs[1 + 5], -coords[4])
    np.testing.assert_allclose(outs[0], jet[6] + jet[7] + jet[8] + jet[9] - jet[5] * coords[4], rtol=1e-12)


def test_third_and_fourth_order_derivatives_carry_whole_jets():
    """ D nested three / four times along one argument: every differentiated axis carries its Taylor jet up to the
    highest order met; channel 1 + d*order + (k-1) = k-th derivative along direction d. """
    D = T.sym_D
    tr = T.trace(lambda u, x, t: D(u, t) + 6.0 * u * D(u, x) + D(D(D(u, x), x), x), 2, None,
                 initial_condition=lambda x: torch.sin(2.0 * x), ndims_spatial=1)
    assert (tr.order, tr.nf, tr.ns, tr.dirs, tr.channels) == (3, 2, 0, [0, 1], 7)
    rng = np.random.RandomState(1)
    coords, jet = rng.uniform(0.2, 0.9, size=(2, 5)), rng.normal(size=(7, 5))
    outs = T.run_program(tr.eq_prog, jet, coords, [])
    # channels: 0 u | 1 u_x, 2 u_xx, 3 u_xxx | 4 u_t, 5 u_tt, 6 u_ttt
    np.testing.assert_allclose(outs[0], jet[4] + 6.0 * jet[0] * jet[1] + jet[3], rtol=1e-12)
    np.testing.assert_allclose(outs[1 + 0], 6.0 * jet[1], rtol=1e-12)
    np.testing.assert_allclose(outs[1 + 1], 6.0 * jet[0], rtol=1e-12)
    assert np.all(outs[1 + 3] == 1.0) and np.all(outs[1 + 4] == 1.0) and np.all(outs[1 + 2] == 0.0) and np.all(outs[1 + 6] == 0.0)
    ic = T.run_program(tr.ic_prog, np.zeros((7, 5)), coords, [])
    x = coords[0]
    np.testing.assert_allclose(ic[0], np.sin(2 * x), rtol=1e-12)
    np.testing.assert_allclose(ic[1], 2 * np.cos(2 * x), rtol=1e-12)
    np.testing.assert_allclose(ic[3], -8 * np.cos(2 * x), rtol=1e-12)
    assert np.all(ic[4] == 0.0) and np.all(ic[6] == 0.0)            # the initial condition does not depend on t
    tr = T.trace(lambda u, x, t: D(D(u, t), t) + D(D(D(D(u, x), x), x), x) * T.Sym(T.var('q')), 2, None)
    assert (tr.order, tr.channels, tr.var_names) == (4, 9, ['q'])
    tr = T.trace(lambda u, x, t: D(D(D(u, x), x), x) + D(u, t), 2, None,              # a variable inside the initial condition
                 initial_condition=lambda x: torch.sin(x) * T.Sym(T.var('amp')), ndims_spatial=1)
    assert tr.ic_has_vars and tr.var_names == ['amp'] and len(tr.ic_prog.outs) == 2 * 7
    ic = T.run_program(tr.ic_prog, np.zeros((7, 5)), coords, [1.5])
    np.testing.assert_allclose(ic[3], -1.5 * np.cos(x), rtol=1e-12)                    # d3/dx3 of amp sin x
    np.testing.assert_allclose(ic[7 + 1], np.cos(x), rtol=1e-12)                       # d/d amp of d/dx


def test_mixed_derivatives_next_to_high_orders_ride_on_two_diagonals():
    """ The biharmonic operator: u_xxxx + 2 u_xxyy + u_yyyy.  Directions x, y, x+y, x-y, each with its jet up to order 4;
    u_xxyy = (P_4 + M_4 - 2 u_xxxx - 2 u_yyyy) / 12. """
    D = T.sym_D
    D2 = lambda u, x: D(D(u, x), x)
    tr = T.trace(lambda u, x, y: D2(D2(u, x), x) + 2.0 * D2(D2(u, x), y) + D2(D2(u, y), y), 2, None)
    assert (tr.order, tr.nf, tr.dirs, tr.channels) == (4, 4, [0, 1, -1, -1], 17)
    assert tr.dir_vecs[2] == [1.0, 1.0] and tr.dir_vecs[3] == [1.0, -1.0]
    rng = np.random.RandomState(2)
    jet = rng.normal(size=(17, 4))
    outs = T.run_program(tr.eq_prog, jet, rng.uniform(size=(2, 4)), [])
    x4, y4, p4, m4 = jet[4], jet[8], jet[12], jet[16]
    np.testing.assert_allclose(outs[0], x4 + y4 + 2.0 * (p4 + m4 - 2.0 * x4 - 2.0 * y4) / 12.0, rtol=1e-12)
    np.testing.assert_allclose(outs[1 + 4], np.full(4, 1.0 - 4.0 / 12.0), rtol=1e-12)
    np.testing.assert_allclose(outs[1 + 12], np.full(4, 2.0 / 12.0), rtol=1e-12)
    # third order: u_xxy = (P_3 - M_3 - 2 u_yyy) / 6
    tr = T.trace(lambda u, x, y: D(D2(u, x), y), 2, None)
    assert (tr.order, tr.nf, tr.channels) == (3, 4, 13)
    jet = rng.normal(size=(13, 3))
    outs = T.run_program(tr.eq_prog, jet, rng.uniform(size=(2, 3)), [])
    np.testing.assert_allclose(outs[0], (jet[9] - jet[12] - 2.0 * jet[6]) / 6.0, rtol=1e-12)


def test_variables_inside_initial_condition():
    D = T.sym_D
    tr = T.trace(lambda f, x, t: D(D(f, x), x) - D(f, t) + T.Sym(T.var('src')), 2, None,
                 initial_condition=lambda x: T.Sym(T.var('amp')) * np.sin(x) + T.Sym(T.var('shift')) ** 2,
                 ndims_spatial=1)
    assert tr.var_names == ['amp', 'shift', 'src'] and tr.ic_has_vars
    C = tr.channels
    assert len(tr.ic_prog.outs) == C * (1 + 3)
    assert min(tr.ic_prog.outs) >= tr.eq_prog.n_slots          # never clobbered by the residual program
    n = 16
    coords = np.random.RandomState(0).uniform(0, 1, size=(2, n))
    vals = [0.7, 0.2, 0.1]
    outs = T.run_program(tr.ic_prog, np.zeros((C, n)), coords, vals)
    x = coords[0]
    np.testing.assert_allclose(outs[0], 0.7 * np.sin(x) + 0.04, rtol=1e-12)
    # dirs: x carries first+second order, t first order -> channels [v, x, t, xx]
    assert tr.dirs == [0, 1] and tr.ns == 1
    np.testing.assert_allclose(outs[1], 0.7 * np.cos(x), rtol=1e-12)
    np.testing.assert_allclose(outs[3], -0.7 * np.sin(x), rtol=1e-12)
    np.testing.assert_allclose(outs[C + 0], np.sin(x), rtol=1e-12)        # d ic / d amp
    np.testing.assert_allclose(outs[2 * C + 0], 0.4 * np.ones(n), rtol=1e-12)  # d ic / d shift
    np.testing.assert_allclose(outs[3 * C + 0], 0 * x, atol=1e-15)       # d ic / d src


def test_numpy_and_torch_entry_points_agree():
    D = T.sym_D
    a = T.trace(lambda f, x: D(f, x) - 2 * np.pi * torch.cos(2 * np.pi * x), 1, None)
    b = T.trace(lambda f, x: D(f, x) - 2 * np.pi * np.cos(2 * np.pi * x), 1, None)
    assert a.residual is b.residual                # hash-consed DAG: identical expression


def test_derivative_with_respect_to_parameter_column_is_a_direction():
    D = T.sym_D
    tr = T.trace(lambda f, x, e: D(f, x) + D(f, e) * e, 2, None)
    assert tr.dirs == [0, 1] and tr.ns == 0


def test_mixed_derivative_is_polarised_onto_a_diagonal_direction():
    D = T.sym_D
    tr = T.trace(lambda f, x, y: D(D(f, x), y) - D(D(f, y), x) + D(D(f, x), y), 2, None)
    assert tr.dirs == [0, 1, -1] and tr.dir_vecs[2] == [1.0, 1.0] and tr.ns == 3 and tr.nf == 3
    # residual == u_xy = (U_vv - U_xx - U_yy) / 2 with channels [u, x, y, v, xx, yy, vv]
    n = 8
    rng = np.random.RandomState(0)
    ujet = rng.normal(size=(tr.channels, n))
    outs = T.run_program(tr.eq_prog, ujet, np.zeros((2, n)), [])
    np.testing.assert_allclose(outs[0], 0.5 * (ujet[6] - ujet[4] - ujet[5]), rtol=1e-12)
    np.testing.assert_allclose(outs[1 + 6], 0.5 * np.ones(n)); np.testing.assert_allclose(outs[1 + 4], -0.5 * np.ones(n))


def test_rewritten_functions_match_torch():
    """ sinh / cosh / exp2 / log2 / log10 / rsqrt / square / reciprocal have no opcode of their own: they lower
    through exp / log / sqrt; value and partials must match torch's. """
    def eq(u, x, y, D=T.sym_D):
        return (torch.sinh(u) * np.cosh(x) + torch.exp2(y) - torch.log2(x + 2.0) + torch.log10(y + 3.0)
                + (x + 1.5).rsqrt() * D(u, x) + u.cosh() - torch.square(D(u, y)) + torch.reciprocal(y + 2.0) + (-u).sinh())
    tr = T.trace(lambda u, x, y: eq(u, x, y), 2, None)
    assert tr.nf == 2 and tr.ns == 0
    rng = np.random.RandomState(3)
    n = 50
    jet = rng.uniform(-1, 1, size=(3, n))
    coords = rng.uniform(0, 1, size=(2, n))
    outs = T.run_program(tr.eq_prog, jet, coords, [])

    ch = [torch.tensor(jet[c], dtype=torch.float64, requires_grad=True) for c in range(3)]
    x, y = (torch.tensor(coords[k], dtype=torch.float64) for k in range(2))
    col = {d: 1 + i for i, d in enumerate(tr.dirs)}            # channel of du/dx_k

    def D_num(v, xx):
        return ch[col[0]] if xx is x else ch[col[1]]
    r = eq(ch[0], x, y, D=D_num)
    assert np.allclose(outs[0], r.detach().numpy(), rtol=1e-12, atol=1e-12)
    grads = torch.autograd.grad(r.sum(), ch)
    for c in range(3):
        assert np.allclose(outs[1 + c], grads[c].numpy(), rtol=1e-10, atol=1e-12)


def test_piecewise_linear_functions_lower_through_abs():
    """ relu / clamp / maximum / minimum / hypot / F.softplus / F.silu have no opcode of their own: they lower through
    |.|, sigmoid, exp and log; values and partials must match torch's (away from the kinks, which the random points are). """
    import torch.nn.functional as F

    def eq(u, x, y, D=T.sym_D):
        return (torch.relu(u - 0.2) * torch.clamp(x, min=0.3) + torch.clamp(D(u, x), -0.4, 0.5) - torch.maximum(u, y)
                + torch.minimum(D(u, y), x - 0.5) * F.softplus(3.0 * u) + F.silu(D(u, x)) + torch.hypot(u, y + 0.1)
                + (u * 2.0).clamp(max=0.7) - (x - u).relu() + np.maximum(u, 0.25) + torch.clamp_min(y - u, 0.1))
    tr = T.trace(lambda u, x, y: eq(u, x, y), 2, None)
    assert tr.nf == 2 and tr.ns == 0
    rng = np.random.RandomState(4)
    n = 200
    jet = rng.uniform(-1, 1, size=(3, n))
    coords = rng.uniform(0, 1, size=(2, n))
    outs = T.run_program(tr.eq_prog, jet, coords, [])
    ch = [torch.tensor(jet[c], dtype=torch.float64, requires_grad=True) for c in range(3)]
    x, y = (torch.tensor(coords[k], dtype=torch.float64) for k in range(2))
    col = {d: 1 + i for i, d in enumerate(tr.dirs)}

    def D_num(v, xx):
        return ch[col[0]] if xx is x else ch[col[1]]

    def eq_t(u, x, y, D):                                    # the same expression with torch-only spellings
        return (torch.relu(u - 0.2) * torch.clamp(x, min=0.3) + torch.clamp(D(u, x), -0.4, 0.5) - torch.maximum(u, y)
                + torch.minimum(D(u, y), x - 0.5) * F.softplus(3.0 * u) + F.silu(D(u, x)) + torch.hypot(u, y + 0.1)
                + (u * 2.0).clamp(max=0.7) - (x - u).relu() + torch.clamp(u, min=0.25) + torch.clamp_min(y - u, 0.1))
    r = eq_t(ch[0], x, y, D_num)
    assert np.allclose(outs[0], r.detach().numpy(), rtol=1e-12, atol=1e-12)
    grads = torch.autograd.grad(r.sum(), ch)
    for c in range(3):
        assert np.allclose(outs[1 + c], grads[c].numpy(), rtol=1e-10, atol=1e-12)
    with pytest.raises(T.NotLowerable):
        T.trace(lambda u, x: F.softplus(u, beta=2.0) + x, 1, None)


def test_comparisons_and_where_lower_to_indicators():
    """ Piecewise coefficients and sources — torch.where(x < 0.5, 1.0, 10.0), (x > 0.2) & (y <= 0.7), ~mask, masks used as
    numbers — lower to 0 / 1 indicators (1 + sign(.)) / 2: values and partials must match torch's on points off the
    thresholds.  What stays off the fused path: `if` on a comparison, == / !=, torch.where guarding a branch that may
    be NaN / inf where it is not selected, boolean algebra on things that are not comparisons. """
    def eq(u, x, y, D):
        a = torch.where(x < 0.5, 1.0, 10.0)
        band = ((x > 0.2) & (y <= 0.7)) | ~(y > 0.1)
        ring = torch.logical_and(x >= 0.3, torch.logical_not(y < 0.4))
        return (a * D(u, x) + D(u, y) * torch.where(band, torch.sin(3.0 * x), u * u) + (x > y) * 2.0 + (0.3 < x) * u
                + torch.where(ring, u, -u * x) - torch.gt(u, 0.1) * y + (u * y).where(x <= 0.6, y ** 2)
                + torch.where(torch.tensor(True), x, y) + ((x > 0.5) ^ (y > 0.5)) * 0.25)
    tr = T.trace(lambda u, x, y: eq(u, x, y, T.sym_D), 2, None)
    assert tr.nf == 2 and tr.ns == 0
    rng = np.random.RandomState(7)
    n = 300
    jet = rng.uniform(-1, 1, size=(3, n))
    coords = rng.uniform(0, 1, size=(2, n))
    outs = T.run_program(tr.eq_prog, jet, coords, [])
    ch = [torch.tensor(jet[c], dtype=torch.float64, requires_grad=True) for c in range(3)]
    x, y = (torch.tensor(coords[k], dtype=torch.float64) for k in range(2))
    col = {d: 1 + i for i, d in enumerate(tr.dirs)}
    r = eq(ch[0], x, y, lambda v, xx: ch[col[0]] if xx is x else ch[col[1]])
    assert np.allclose(outs[0], r.detach().numpy(), rtol=1e-12, atol=1e-12)
    grads = torch.autograd.grad(r.sum(), ch)
    for c in range(3):
        assert np.allclose(outs[1 + c], grads[c].numpy(), rtol=1e-10, atol=1e-12)

    def raises(f, total=1):
        with pytest.raises(T.NotLowerable):
            T.trace(f, total, None)
    raises(lambda u, x: torch.where(x > 0.5, torch.sqrt(x - 0.5), 0.0 * x) + u)        # guard around a NaN branch
    raises(lambda u, x: torch.where(x > 0.5, 1.0 / x, x) + u)
    raises(lambda u, x: torch.where(x > 0.5, torch.exp(x), x) + u)
    raises(lambda u, x: u if x > 0.5 else -u)                                            # control flow
    raises(lambda u, x: (x == 0.5) * u)
    raises(lambda u, x: (x & u))                                                         # not comparisons
    raises(lambda u, x: torch.where(x, u, -u))
