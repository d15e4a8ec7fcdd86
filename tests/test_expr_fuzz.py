""" Randomised residual expressions: the tracer (symbolic D, partial derivatives, constant folding, hash-consing,
slot reuse in the register programs) + the device interpreter (host build) against torch autograd in fp64.
Each case draws an expression tree over u, its first / second derivatives, the point columns, constants and a
trainable variable, with the operators the fused path supports, on a small fixed network.  CPU only. """
import numpy as np
import pytest
import torch

import emul_harness as E
from helpers import rel_l2
from oracle import autograd_port as ap
from pydens_b200 import _native as N
from pydens_b200 import tracer as T

UNARY = [
    ('sin', lambda t: torch.sin(t)), ('cos', lambda t: torch.cos(t)), ('tanh', lambda t: torch.tanh(t)),
    ('exp', lambda t: torch.exp(-t * t)), ('sqrt', lambda t: torch.sqrt(t * t + 1.0)),
    ('log', lambda t: torch.log(t * t + 0.5)), ('sigmoid', lambda t: torch.sigmoid(t)),
    ('abs', lambda t: abs(t)), ('neg', lambda t: -t), ('sq', lambda t: t ** 2), ('cube', lambda t: t ** 3),
    ('recip', lambda t: 1.0 / (t * t + 1.0)), ('sinh', lambda t: torch.sinh(0.5 * t)), ('cos2', lambda t: torch.cos(2.0 * t)),
    ('powf', lambda t: (t * t + 1.0) ** 1.5), ('scale', lambda t: 0.37 * t - 0.2),
]
BINARY = [
    ('add', lambda a, b: a + b), ('sub', lambda a, b: a - b), ('mul', lambda a, b: a * b),
    ('div', lambda a, b: a / (b * b + 1.0)), ('mix', lambda a, b: 0.5 * a - 1.5 * b + 0.1),
]


def random_equation(rng, total, high=False):
    """ -> (callable eq(u, *xs, D, V), description).  Leaves are created lazily so D() is called inside the trace.
    `high`: also derivatives of order 3 / 4 (and mixed ones carried by the diagonals) among the leaves. """
    second = rng.rand() < 0.6
    mixed = total >= 2 and rng.rand() < 0.3
    use_var = rng.rand() < 0.4
    extra = []
    if high:
        extra = ['uxxx'] + (['uxxxx'] if rng.rand() < 0.6 else []) + (['uyyy'] if total >= 2 and rng.rand() < 0.4 else [])
        if total >= 2 and rng.rand() < 0.4:
            extra += ['uxxy'] + (['uxxyy'] if 'uxxxx' in extra else [])

    def leaf():
        kinds = ['u', 'u', 'ux', 'x', 'const'] + (['uxx'] if second else []) + (['uxy'] if mixed else []) \
            + (['y', 'uy'] if total >= 2 else []) + (['var'] if use_var else []) + extra
        k = kinds[int(rng.randint(len(kinds)))]
        c = float(np.round(rng.uniform(-2, 2), 2))
        return (k, c)

    def tree(depth):
        if depth == 0 or rng.rand() < 0.2:
            return ('leaf', leaf())
        if rng.rand() < 0.45:
            return ('un', int(rng.randint(len(UNARY))), tree(depth - 1))
        return ('bin', int(rng.randint(len(BINARY))), tree(depth - 1), tree(depth - 1))

    # make sure the residual depends on u and on a derivative
    root = ('bin', 0, ('bin', 0, tree(int(rng.randint(2, 5))), ('leaf', ('ux', 0.0))), ('leaf', ('u', 0.0)))
    if high:                                               # ... and on a derivative of order 3 at least
        root = ('bin', 0, root, ('leaf', (extra[int(rng.randint(len(extra)))], 0.0)))

    def describe(t):
        if t[0] == 'leaf':
            return t[1][0] if t[1][0] != 'const' else str(t[1][1])
        if t[0] == 'un':
            return '%s(%s)' % (UNARY[t[1]][0], describe(t[2]))
        return '%s(%s, %s)' % (BINARY[t[1]][0], describe(t[2]), describe(t[3]))

    def eq(u, *xs, D, V):
        x, y = xs[0], xs[-1]
        cache = {}

        def val(kind, c):
            if kind == 'const':
                return torch.tensor(c, dtype=torch.float64)        # a 0-d tensor: torch.sin(0.62) is not valid torch
            if kind not in cache:
                cache[kind] = {'u': lambda: u, 'x': lambda: x, 'y': lambda: y, 'ux': lambda: D(u, x),
                               'uy': lambda: D(u, y), 'uxx': lambda: D(D(u, x), x), 'uxy': lambda: D(D(u, x), y),
                               'uxxx': lambda: D(D(D(u, x), x), x), 'uxxxx': lambda: D(D(D(D(u, x), x), x), x),
                               'uyyy': lambda: D(D(D(u, y), y), y), 'uxxy': lambda: D(D(D(u, x), x), y),
                               'uxxyy': lambda: D(D(D(D(u, x), x), y), y),
                               'var': lambda: V('k', 0.7)}[kind]()
            return cache[kind]

        def ev(t):
            if t[0] == 'leaf':
                return val(*t[1])
            if t[0] == 'un':
                return UNARY[t[1]][1](ev(t[2]))
            return BINARY[t[1]][1](ev(t[2]), ev(t[3]))
        return ev(root)
    return eq, describe(root), use_var


@pytest.mark.parametrize('seed', list(range(150)))
def test_random_expression_matches_autograd(seed):
    rng = np.random.RandomState(5000 + seed)
    total = int(rng.randint(1, 3))
    eq, text, use_var = random_equation(rng, total)
    features, acts = [6, 5, 1], ['Tanh', 'Sigmoid']
    sym_V = lambda n, init: T.Sym(T.var(n))
    try:
        traced = T.trace(lambda u, *xs: eq(u, *xs, D=T.sym_D, V=sym_V), total, None)
    except T.NotLowerable as exc:                          # e.g. program too long: the tracer must say so, not mis-lower
        assert 'slots' in str(exc) or 'instructions' in str(exc) or 'directions' in str(exc), (text, exc)
        return
    spec = N.build_spec([total] + features, ['tanh', 'sigmoid', 'none'], total, 0, False, 0.0, False,
                        [(0.0, 1.0)] * total, traced)
    uses_var = 'k' in traced.var_names
    if 'var' in text and not uses_var:
        return                                             # the variable cancelled symbolically (k - k, 0 * k, …)
    prob = ap.Problem(eq, ndims=total, features=features, activation=acts, dtype=torch.float64,
                      variables={'k': 0.7} if uses_var else None, seed=seed, layout='fafaf')
    params = prob.flat_params().numpy().astype(np.float32)
    assert spec.n_params == params.size, text
    pts = rng.uniform(0.05, 0.95, size=(40, total)).astype(np.float32)
    loss, residual, grads = E.emul_step(spec, params, pts)
    prob.load_flat(torch.from_numpy(params.astype(np.float64)))
    ref_loss, ref_res, ref_grads = prob.loss_and_grads(pts.astype(np.float64))
    cond = max(1.0, 0.05 / max(float(np.sqrt(np.mean(np.square(ref_res)))), 1e-30))
    assert abs(loss - ref_loss) <= 3e-5 * cond * max(abs(ref_loss), 1e-6), text
    assert rel_l2(residual, ref_res) <= 3e-5 * cond, text
    assert rel_l2(grads, ref_grads.numpy()) <= 1e-4 * cond, text


@pytest.mark.parametrize('seed', list(range(100)))
def test_random_high_order_expression_matches_autograd(seed):
    """ The same with derivatives of order 3 / 4 among the leaves: the whole-jet path (tracer polarisation included). """
    rng = np.random.RandomState(7000 + seed)
    total = int(rng.randint(1, 3))
    eq, text, use_var = random_equation(rng, total, high=True)
    features, acts = [6, 5, 1], ['Tanh', 'Sigmoid']
    sym_V = lambda n, init: T.Sym(T.var(n))
    try:
        traced = T.trace(lambda u, *xs: eq(u, *xs, D=T.sym_D, V=sym_V), total, None)
    except T.NotLowerable as exc:
        assert any(w in str(exc) for w in ('slots', 'instructions', 'directions', 'outputs', 'diagonals')), (text, exc)
        return
    assert traced.order in (3, 4), text
    spec = N.build_spec([total] + features, ['tanh', 'sigmoid', 'none'], total, 0, False, 0.0, False,
                        [(0.0, 1.0)] * total, traced)
    uses_var = 'k' in traced.var_names
    if 'var' in text and not uses_var:
        return
    prob = ap.Problem(eq, ndims=total, features=features, activation=acts, dtype=torch.float64,
                      variables={'k': 0.7} if uses_var else None, seed=seed, layout='fafaf')
    params = prob.flat_params().numpy().astype(np.float32)
    assert spec.n_params == params.size, text
    pts = rng.uniform(0.05, 0.95, size=(40, total)).astype(np.float32)
    loss, residual, grads = E.emul_step(spec, params, pts)
    prob.load_flat(torch.from_numpy(params.astype(np.float64)))
    ref_loss, ref_res, ref_grads = prob.loss_and_grads(pts.astype(np.float64))
    cond = max(1.0, 0.05 / max(float(np.sqrt(np.mean(np.square(ref_res)))), 1e-30))
    if 'uxxy' in text:
        cond *= 5.0                                        # polarisation cancels leading digits
    assert abs(loss - ref_loss) <= 3e-5 * cond * max(abs(ref_loss), 1e-6), text
    assert rel_l2(residual, ref_res) <= 3e-5 * cond, text
    assert rel_l2(grads, ref_grads.numpy()) <= 1e-4 * cond, text


@pytest.mark.parametrize('seed', list(range(80)))
def test_random_initial_condition_matches_autograd(seed):
    """ Random `initial_condition` callables (over the spatial columns and trainable variables) in the ansatz
    u = S(t) * net + ic(x): the kernel needs the jet of ic along every derivative direction and, with variables,
    its partials — all produced by the tracer. """
    rng = np.random.RandomState(9000 + seed)
    nsp = int(rng.randint(1, 3))                           # spatial columns; the last column is time
    total = nsp + 1
    use_var = rng.rand() < 0.5

    def tree(depth):
        if depth == 0 or rng.rand() < 0.25:
            kinds = ['x', 'x', 'const'] + (['y'] if nsp == 2 else []) + (['var'] if use_var else [])
            return ('leaf', (kinds[int(rng.randint(len(kinds)))], float(np.round(rng.uniform(-2, 2), 2))))
        if rng.rand() < 0.45:
            return ('un', int(rng.randint(len(UNARY))), tree(depth - 1))
        return ('bin', int(rng.randint(len(BINARY))), tree(depth - 1), tree(depth - 1))
    root = ('bin', 0, tree(int(rng.randint(1, 4))), ('leaf', ('x', 0.0)))

    def make_ic(V):
        def ic(*xs):
            def ev(t):
                if t[0] == 'leaf':
                    kind, c = t[1]
                    return {'x': lambda: xs[0], 'y': lambda: xs[-1], 'const': lambda: torch.tensor(c, dtype=torch.float64),
                            'var': lambda: V('amp', 0.6)}[kind]()
                if t[0] == 'un':
                    return UNARY[t[1]][1](ev(t[2]))
                return BINARY[t[1]][1](ev(t[2]), ev(t[3]))
            return ev(root)
        return ic

    def eq(u, *xs, D, V):
        x, t = xs[0], xs[-1]
        r = D(u, t) - 0.3 * D(D(u, x), x) + 0.1 * u * D(u, x)
        if nsp == 2:
            r = r - 0.2 * D(D(u, xs[1]), xs[1]) + 0.05 * D(D(u, x), xs[1])
        return r
    sym_V = lambda n, init: T.Sym(T.var(n))
    try:
        traced = T.trace(lambda u, *xs: eq(u, *xs, D=T.sym_D, V=sym_V), total, None,
                         initial_condition=make_ic(sym_V), ndims_spatial=nsp)
    except T.NotLowerable as exc:                          # a jet too long for the program memory: must say so
        assert 'slots' in str(exc) or 'instructions' in str(exc), exc
        return
    has_var = 'amp' in traced.var_names
    features = [7, 6, 1]
    bc = float(np.round(rng.uniform(-1, 1), 2)) if rng.rand() < 0.5 else None
    domain = [(float(np.round(rng.uniform(-0.5, 0.1), 2)), float(np.round(rng.uniform(0.9, 1.8), 2))) for _ in range(total)]
    spec = N.build_spec([total] + features, ['tanh', 'tanh', 'none'], total, 0, bc is not None, bc or 0.0, True, domain, traced)
    holder = {}
    prob = ap.Problem(eq, ndims=total, features=features, activation='Tanh', dtype=torch.float64,
                      initial_condition=make_ic(lambda n, init: holder['p'].V(n, init)), boundary_condition=bc, domain=domain,
                      variables={'amp': 0.6} if has_var else None, seed=seed, layout='fafaf')
    holder['p'] = prob
    if use_var and not has_var:
        return                                             # the variable cancelled symbolically
    with torch.no_grad():
        prob.log_scale.fill_(float(np.round(rng.uniform(-0.5, 0.5), 2)))
    params = prob.flat_params().numpy().astype(np.float32)
    assert spec.n_params == params.size
    pts = np.concatenate([rng.uniform(lo, hi, size=(40, 1)) for lo, hi in domain], axis=1).astype(np.float32)
    loss, residual, grads = E.emul_step(spec, params, pts)
    prob.load_flat(torch.from_numpy(params.astype(np.float64)))
    ref_loss, ref_res, ref_grads = prob.loss_and_grads(pts.astype(np.float64))
    cond = max(1.0, 0.05 / max(float(np.sqrt(np.mean(np.square(ref_res)))), 1e-30))
    assert abs(loss - ref_loss) <= 3e-5 * cond * max(abs(ref_loss), 1e-6)
    assert rel_l2(residual, ref_res) <= 3e-5 * cond
    assert rel_l2(grads, ref_grads.numpy()) <= 1e-4 * cond
    u = E.emul_forward(spec, params, pts)
    ref_u = prob.predict(pts.astype(np.float64))
    assert np.abs(u - ref_u).max() <= 1e-5 * max(1.0, np.abs(ref_u).max())


# functions that lower through |.| (relu, clamp, maximum / minimum, hypot, F.softplus, F.silu): tables of their own so
# that the problems of the seeds above stay what they were
PIECEWISE_UNARY = [
    ('relu', lambda t: torch.relu(t)), ('clamp', lambda t: torch.clamp(t, -0.5, 0.8)), ('clamp_min', lambda t: torch.clamp(t, min=0.1)),
    ('softplus', lambda t: torch.nn.functional.softplus(t)), ('silu', lambda t: torch.nn.functional.silu(t)),
    ('relu_m', lambda t: (t - 0.3).relu()),
]
PIECEWISE_BINARY = [
    ('maximum', lambda a, b: torch.maximum(a, b)), ('minimum', lambda a, b: torch.minimum(a, b)),
    ('hypot', lambda a, b: torch.hypot(a, b + 0.5)),
]


@pytest.mark.parametrize('seed', list(range(60)))
def test_random_expression_with_piecewise_functions_matches_autograd(seed, monkeypatch):
    import sys
    mod = sys.modules[__name__]
    monkeypatch.setattr(mod, 'UNARY', UNARY + PIECEWISE_UNARY * 2)
    monkeypatch.setattr(mod, 'BINARY', BINARY + PIECEWISE_BINARY)
    rng = np.random.RandomState(15000 + seed)
    total = int(rng.randint(1, 3))
    eq, text, use_var = random_equation(rng, total)
    features, acts = [6, 5, 1], ['Tanh', 'Sigmoid']
    sym_V = lambda n, init: T.Sym(T.var(n))
    try:
        traced = T.trace(lambda u, *xs: eq(u, *xs, D=T.sym_D, V=sym_V), total, None)
    except T.NotLowerable as exc:
        assert 'slots' in str(exc) or 'instructions' in str(exc) or 'directions' in str(exc), (text, exc)
        return
    spec = N.build_spec([total] + features, ['tanh', 'sigmoid', 'none'], total, 0, False, 0.0, False,
                        [(0.0, 1.0)] * total, traced)
    uses_var = 'k' in traced.var_names
    if 'var' in text and not uses_var:
        return
    prob = ap.Problem(eq, ndims=total, features=features, activation=acts, dtype=torch.float64,
                      variables={'k': 0.7} if uses_var else None, seed=seed, layout='fafaf')
    params = prob.flat_params().numpy().astype(np.float32)
    pts = rng.uniform(0.05, 0.95, size=(40, total)).astype(np.float32)
    loss, residual, grads = E.emul_step(spec, params, pts)
    prob.load_flat(torch.from_numpy(params.astype(np.float64)))
    ref_loss, ref_res, ref_grads = prob.loss_and_grads(pts.astype(np.float64))
    cond = max(1.0, 0.05 / max(float(np.sqrt(np.mean(np.square(ref_res)))), 1e-30))
    assert abs(loss - ref_loss) <= 3e-5 * cond * max(abs(ref_loss), 1e-6), text
    assert rel_l2(residual, ref_res) <= 3e-5 * cond, text
    assert rel_l2(grads, ref_grads.numpy()) <= 1e-4 * cond, text


def _interface_equations():
    """ piecewise coefficients / sources whose conditions are on the coordinates (exact fp32 inputs: both precisions take
    the same branch) """
    def diffusivity(u, x, y, D, V):
        a = torch.where(x < 0.5, 1.0, 10.0)
        return a * D(D(u, x), x) + D(D(u, y), y) - torch.where((x > 0.2) & (y <= 0.7), torch.sin(3.0 * x), 0.0 * x)

    def reaction(u, x, y, D, V):
        k = torch.where(y > x, 2.0, 0.5) * V('k', 0.7)
        return D(u, y) - D(D(u, x), x) + k * u * (1.0 - u) + (x > 0.6) * 0.3 - (~(y > 0.25)) * u

    def third(u, x, y, D, V):
        return D(u, y) + torch.where(x >= 0.4, 6.0, 1.5) * u * D(u, x) + D(D(D(u, x), x), x) * ((y < 0.5) | (x < 0.3))
    return [('diffusivity', diffusivity, None), ('reaction', reaction, {'k': 0.7}), ('third', third, None)]


@pytest.mark.parametrize('which', [0, 1, 2])
def test_interface_problems_match_autograd(which):
    """ torch.where / comparisons in the equation run as indicator arithmetic in the device interpreter (host build):
    loss, residual and gradients against the fp64 oracle executing the user's torch.where. """
    name, eq, variables = _interface_equations()[which]
    features, acts = [8, 6, 1], ['Tanh', 'Sigmoid']
    sym_V = lambda n, init: T.Sym(T.var(n))
    traced = T.trace(lambda u, *xs: eq(u, *xs, D=T.sym_D, V=sym_V), 2, None, initial_condition=0.3, ndims_spatial=1)
    spec = N.build_spec([2] + features, ['tanh', 'sigmoid', 'none'], 2, 0, True, 0.1, True, [(0.0, 1.0)] * 2, traced)
    prob = ap.Problem(eq, ndims=2, features=features, activation=acts, dtype=torch.float64, variables=variables, seed=which,
                      layout='fafaf', initial_condition=0.3, boundary_condition=0.1)
    params = prob.flat_params().numpy().astype(np.float32)
    assert spec.n_params == params.size
    rng = np.random.RandomState(77 + which)
    pts = rng.uniform(0.02, 0.98, size=(96, 2)).astype(np.float32)
    loss, residual, grads = E.emul_step(spec, params, pts)
    prob.load_flat(torch.from_numpy(params.astype(np.float64)))
    ref_loss, ref_res, ref_grads = prob.loss_and_grads(pts.astype(np.float64))
    assert abs(loss - ref_loss) <= 2e-5 * abs(ref_loss), name
    assert rel_l2(residual, ref_res) <= 2e-5, name
    assert rel_l2(grads, ref_grads.numpy()) <= 1e-4, name
