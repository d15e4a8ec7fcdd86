""" Host-side logic of the pydens API on CPU (autograd path): signatures, reshape_and_concat contract,
V / constraints / freeze semantics, lowering decisions.  The fused path itself is exercised with
-m gpu (test_gpu_parity.py, test_gpu_fit.py). """
import numpy as np
import pytest
import torch

import pydens
from pydens import Solver, D, V, NumpySampler, ConvBlockModel, TorchModel


def pde(f, x, y):
    return D(D(f, x), x) + D(D(f, y), y) - 5 * torch.sin(np.pi * (x + y))


def test_import_surface_matches_reference():
    for name in ('Solver', 'D', 'V', 'TorchModel', 'ConvBlockModel', 'NumpySampler'):
        assert hasattr(pydens, name)
    assert pydens.__version__.startswith('1.0.2')


def test_readme_poisson_on_cpu_autograd_path():
    torch.manual_seed(0)
    solver = Solver(equation=pde, ndims=2, boundary_condition=1, layout='fa fa fa f', activation='Tanh',
                    units=[10, 12, 15, 1], device='cpu')
    assert solver._traced is not None and solver._traced.dirs == [0, 1] and solver._traced.ns == 2
    assert sum(p.numel() for p in solver.model.conv_block.parameters()) == 373
    solver.fit(batch_size=100, niters=30)
    assert len(solver.losses) == 30 and isinstance(solver.losses[0], np.ndarray) and solver.losses[0].shape == ()
    solver.fit(batch_size=50, niters=5, optimizer=None)          # re-uses the optimizer, list keeps growing
    assert len(solver.losses) == 35
    grid = np.linspace(0, 1, 7)
    out = solver.predict(grid, 0.0)                               # boundary y=0 -> bc value exactly
    assert out.shape == (7, 1) and np.allclose(out, 1.0)


def test_fused_backend_raises_without_gpu_or_lowering():
    if torch.cuda.is_available():
        pytest.skip('CPU-only check')
    solver = Solver(pde, ndims=2, boundary_condition=1, device='cpu', backend='fused')
    with pytest.raises(RuntimeError):
        solver.fit(niters=1, batch_size=10)
    with pytest.raises(RuntimeError):
        Solver(lambda f, x: D(D(D(D(D(f, x), x), x), x), x), ndims=1, device='cpu', backend='fused')     # fifth order


def test_reshape_and_concat_contract():
    rc = Solver.reshape_and_concat
    out = rc([np.linspace(0, 1, 5), 2, [1., 2., 3., 4., 5.], torch.arange(5.)])
    assert out.shape == (5, 4) and out.dtype == torch.float32
    assert torch.all(out[:, 1] == 2)
    assert rc([0.5, 1]).shape == (1, 2)
    assert torch.equal(rc([np.array([[3.0], [4.0]]), np.array([7.0, 8.0, 9.0])[:1].repeat(2)])[:, 0],
                       torch.tensor([3.0, 4.0]))
    short = rc([np.arange(6.0), np.array([9.0, 10.0])])           # wrong-sized array: first element tiled
    assert torch.all(short[:, 1] == 9.0)


def test_variables_constraints_and_freezing():
    def odevar(f, x):
        return D(f, x) - 2 * np.pi * torch.cos(2 * np.pi * x) + V('new_var', data=torch.Tensor([1.0]))
    torch.manual_seed(1)
    solver = Solver(odevar, ndims=1, initial_condition=1, constraints=lambda f, x: f(torch.tensor([0.5])),
                    device='cpu')
    assert isinstance(solver.model.new_var, torch.nn.Parameter)
    assert solver._traced.var_names == ['new_var']
    solver.model.freeze_trainable(variables=('new_var',))
    solver.fit(niters=5, batch_size=20, lr=0.1)
    assert float(solver.model.new_var) == 1.0
    solver.model.unfreeze_trainable(variables=['new_var'])
    solver.fit(niters=5, batch_size=20, lr=0.1, loss_terms=['equation', 'constraint_0'])
    assert float(solver.model.new_var) != 1.0
    solver.model.freeze_layers(['fc1'], ['log_scale'])            # README spelling
    assert not solver.model.conv_block.linears[0].weight.requires_grad
    assert not solver.model.log_scale.requires_grad


def test_lowering_decisions():
    s = Solver(lambda f, x: D(D(D(f, x), x), x), ndims=1, device='cpu')
    assert s._traced is not None and s._traced.order == 3                         # whole-jet kernels (orders 3 and 4)
    assert Solver(lambda f, x: D(D(D(D(D(f, x), x), x), x), x), ndims=1, device='cpu')._traced is None
    s = Solver(lambda f, x, y: D(D(f, x), y), ndims=2, device='cpu')
    assert s._traced is not None and s._traced.dirs == [0, 1, -1]                 # mixed derivative: polarised
    s = Solver(lambda f, x, y, z: D(D(f, x), y) + D(D(f, y), z) + D(D(f, x), z), ndims=3, device='cpu')
    assert s._traced is not None and s._traced.nf == 6 and s._traced.ns == 6      # x, y, z and the three diagonals
    s = Solver(lambda f, x, y, z, t: D(D(f, x), y) + D(D(f, y), z) + D(D(f, x), z) - D(f, t), ndims=4, device='cpu')
    assert s._traced is None and 'directions' in s._lower_error                   # seven: more than the kernels carry
    s = Solver(lambda f, x, y, z, w, t: D(D(f, x), x) + D(D(f, y), y) + D(D(f, z), z) + D(D(f, w), w) - D(f, t), ndims=5,
               device='cpu')
    assert s._traced.nf == 5 and s._traced.ns == 5                                # t promoted to second order
    s = Solver(lambda f, t: D(f, t), ndims=1, initial_condition=lambda: V('init', data=torch.Tensor([3.0])),
               device='cpu')
    assert s._traced is not None and s._traced.var_names == ['init'] and s._traced.ic_has_vars
    s.fit(niters=2, batch_size=8)                                 # (CPU here: autograd path)
    s = Solver(pde, ndims=2, layout='faR fa+ f', features=[6, 6, 1], device='cpu')
    assert s._traced is not None and [c[2] for c in s._chain] == [None, 0, None]     # residual layouts lower
    s.fit(niters=2, batch_size=8)
    s = Solver(pde, ndims=2, layout='fa R fa+ R fa+ f', features=[5, 5, 5, 1], device='cpu')
    assert [c[2] for c in s._chain] == [None, 0, 1, None]                           # chained residual blocks
    s = Solver(pde, ndims=2, layout='R fa fa + f', features=[2, 2, 1], device='cpu')   # skip from the raw input
    assert s._traced is None and 'dense chain' in s._lower_error
    s.fit(niters=2, batch_size=8)
    s = Solver(pde, ndims=2, layout='fa fa f', features=[4, 4, 1], activation='Sin', device='cpu')
    assert s._traced is not None and [c[1] for c in s._chain] == ['sin', 'sin', 'none']
    s = Solver(pde, ndims=2, layout='fa fa fa f', features=[4, 4, 4, 1], activation=['Softplus', torch.nn.SiLU, 'GELU'],
               device='cpu')
    assert [c[1] for c in s._chain] == ['softplus', 'silu', 'gelu', 'none']
    s = Solver(pde, ndims=2, layout='fa fa f', features=[4, 4, 1], activation='ELU', device='cpu')
    assert s._traced is None                                      # not among the fused activations
    s = Solver(pde, ndims=2, layout='fa f', features=[4, 1], activation=torch.nn.Softplus(beta=2.0), device='cpu')
    assert s._traced is None                                      # right name, different function
    s = Solver(pde, ndims=2, layout='fa f', features=[4, 1], activation=torch.nn.GELU(approximate='tanh'), device='cpu')
    assert s._traced is None

    class MyModel(ConvBlockModel):
        def forward(self, xs):
            return super().forward(xs) * 2
    assert Solver(pde, ndims=2, model=MyModel, device='cpu')._traced is None


def test_sampler_and_domain_handling():
    s = Solver(lambda f, x, e: D(f, x) - e * np.pi * torch.cos(e * np.pi * x), ndims=1, nparams=1,
               initial_condition=2.0, device='cpu')
    s.fit(niters=3, batch_size=16, sampler=NumpySampler('u') & NumpySampler('u', low=.5, high=5.5), lr=0.01)
    assert np.allclose(s.predict(0.0, 3.0), 2.0)                  # t = t0 -> initial condition
    with pytest.raises(ValueError):
        Solver(pde, ndims=2, domain=3, device='cpu')
    s2 = Solver(pde, ndims=2, domain=[(-1, 1), (0, 2)], boundary_condition=0.5, device='cpu')
    assert np.allclose(s2.predict(-1.0, 1.0), 0.5)


def test_constraints_lower_to_fused_launches_when_pointwise():
    def odevar(u, t):
        return D(u, t) - 2 * np.pi * torch.cos(2 * np.pi * t)

    def initial(*args):
        return V('init', data=torch.Tensor([3.0]))
    s = Solver(odevar, ndims=1, initial_condition=initial, device='cpu',
               constraints=[lambda u, t: u(torch.tensor([0.5])),                 # README.md:118
                            lambda u, t: u(np.array([0.25, 0.75])) ** 2 - 1.0,
                            lambda u, t: u(0.1) - u(0.9),                          # two evaluations: autograd adds it
                            lambda u, t: (u(0.5) * t).mean()])                     # uses the batch points
    tr, pts = s._lower_constraint(0)
    assert tr.channels == 1 and tr.var_names == ['init'] and pts.shape == (1, 1) and float(pts[0, 0]) == 0.5
    tr, pts = s._lower_constraint(1)
    assert pts.shape == (2, 1) and len(tr.eq_prog) >= 2
    assert s._lower_constraint(2) is None and s._lower_constraint(3) is None
    s.fit(niters=2, batch_size=8, loss_terms=['equation', 'constraint_0', 'constraint_2'])   # CPU: autograd path
    assert len(s.losses) == 2


def test_criteria_other_than_mse():
    """ reference :365 / :448 `criterion`: MSELoss, L1Loss, HuberLoss, SmoothL1Loss (mean reduction) have a fused form —
    the Solver re-traces equation and constraints for the criterion of the fit; anything else stays on autograd.  On the
    CPU the autograd path applies the criterion as the reference does. """
    from torch import nn
    key = Solver._criterion_key
    assert key(nn.MSELoss()) == ('mse',) and key(nn.L1Loss()) == ('l1',)
    assert key(nn.HuberLoss(delta=0.25)) == ('huber', 0.25) and key(nn.SmoothL1Loss(beta=0.5)) == ('smooth_l1', 0.5)
    assert key(nn.SmoothL1Loss(beta=0.0)) == ('l1',)
    assert key(nn.functional.l1_loss) == ('l1',) and key(nn.functional.huber_loss) == ('huber', 1.0) and key(nn.functional.mse_loss) == ('mse',)
    assert key(nn.MSELoss(reduction='sum')) == ('mse', 'sum') and key(nn.HuberLoss(reduction='sum', delta=2.0)) == ('huber', 2.0, 'sum')
    assert key(nn.MSELoss(reduction='none')) is None and key(nn.BCELoss()) is None and key(lambda a, b: (a - b).abs().mean()) is None

    torch.manual_seed(0)
    solver = Solver(lambda u, t: D(u, t) - 2 * np.pi * torch.cos(2 * np.pi * t), ndims=1, initial_condition=1,
                    constraints=lambda u, t: u(torch.tensor([0.5])) - 1.0, layout='fa fa f', units=[8, 8, 1], activation='Tanh',
                    device='cpu')
    mse_prog = solver._traced.eq_prog
    assert solver._lower_constraint(0) is not None
    solver._switch_criterion(('huber', 0.5))
    assert solver._crit_key == ('huber', 0.5) and solver._traced is not None and solver._traced_constraints == {}
    assert len(solver._traced.eq_prog) > len(mse_prog)          # the residual transform is part of the program
    assert solver._lower_constraint(0) is not None                     # constraints follow the criterion
    from pydens_b200 import _native as N
    tr = solver._traced
    spec = N.build_spec([1, 8, 8, 1], ['tanh', 'tanh', 'none'], 1, 0, False, 0.0, True, [(0.0, 1.0)], tr)
    assert spec.n_eq == len(tr.eq_prog)
    solver._switch_criterion(('mse',))
    assert len(solver._traced.eq_prog) == len(mse_prog)
    for crit in (nn.L1Loss(), nn.HuberLoss(delta=0.3)):
        before = len(solver.losses)
        solver.fit(niters=3, batch_size=32, criterion=crit, loss_terms=('equation', 'constraint_0'))
        assert len(solver.losses) == before + 3 and np.isfinite(solver.losses[-1])
