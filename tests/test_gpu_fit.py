""" Solver.fit on the fused path (GPU): trajectories against the reference's own fit (goldens),
CUDA-graph replay == plain launches, in-kernel samplers, constraints hybrid, tutorial flows. """
import os

import numpy as np
import pytest
import torch

import problems as P
from helpers import load_golden

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from gpu_helpers import make_solver, Replay
    from pydens_b200 import Solver, D, V, NumpySampler


@pytest.mark.parametrize('adam', ['kernel', 'torch'])
@pytest.mark.parametrize('name', [n for n in P.GOLDEN_TRAJ if n not in P.HI_DIRECTION + P.HI_ORDER])
def test_fit_trajectory_matches_reference_fit(name, adam, monkeypatch):
    """ Same init, same point stream, same Adam: the loss curve of the fused fit follows the curve of the
    reference's `Solver.fit` (BASELINE: residual MSE within 1e-5 of reference on identical points).  Both forms of
    optimizer.step(): in the tail of the step kernel (pinn_step_adam, the default) and torch's fused Adam kernels. """
    monkeypatch.setenv('PYDENS_B200_FUSED_ADAM', '1' if adam == 'kernel' else '0')
    # 'torch' also keeps tiny batches on one launch per step; 'kernel' is the default configuration, where they go to the
    # persistent cluster kernel by themselves
    monkeypatch.setenv('PYDENS_B200_AUTO_PERSISTENT', '1' if adam == 'kernel' else '0')
    g = load_golden(name)
    niters, batch, lr = int(g['traj_meta'][0]), int(g['traj_meta'][1]), float(g['traj_meta'][2])
    solver = make_solver(name, g['params'])
    batches = [P.make_points(name, batch, seed=1000 + i) for i in range(niters)]
    solver.fit(niters=niters, batch_size=batch, sampler=Replay(batches), lr=lr)
    losses = np.asarray(solver.losses, dtype=np.float64)
    ref = g['traj_losses'].astype(np.float64)
    assert losses.shape == ref.shape
    assert np.max(np.abs(losses - ref) / np.maximum(np.abs(ref), 1e-6)) <= 2e-3
    assert abs(losses[-1] - ref[-1]) <= 1e-5 * max(1.0, abs(ref[-1]))
    final = solver.flat_params().cpu().numpy()
    assert np.linalg.norm(final - g['traj_params']) / np.linalg.norm(g['traj_params']) <= 1e-3


def test_graph_replay_equals_plain_launches():
    g = load_golden('poisson2d')
    curves = []
    for no_graph in ('0', '1'):
        os.environ['PYDENS_B200_NO_GRAPH'] = no_graph
        try:
            solver = make_solver('poisson2d', g['params'])
            solver.fit(niters=40, batch_size=5000, lr=0.005)
        finally:
            os.environ.pop('PYDENS_B200_NO_GRAPH', None)
        curves.append(np.asarray(solver.losses, dtype=np.float64))
    assert len(curves[0]) == 40 and np.isfinite(curves[0]).all()
    np.testing.assert_allclose(curves[0], curves[1], rtol=1e-6)
    assert curves[0][-1] < curves[0][0]


def test_readme_examples_train():
    def pde(f, x, y):
        return D(D(f, x), x) + D(D(f, y), y) - 5 * torch.sin(np.pi * (x + y))
    torch.manual_seed(0)
    solver = Solver(equation=pde, ndims=2, boundary_condition=1, layout='fa fa fa f', activation='Tanh',
                    units=[10, 12, 15, 1])
    solver.fit(batch_size=100, niters=1500)
    assert len(solver.losses) == 1500 and solver._engine is not None
    assert np.mean(solver.losses[-50:]) < 0.1 * np.mean(solver.losses[:10])

    def odeparam(f, x, e):
        return D(f, x) - e * np.pi * torch.cos(e * np.pi * x)
    s = NumpySampler('uniform') & NumpySampler('uniform', low=1, high=5)
    solver = Solver(equation=odeparam, ndims=1, nparams=1, initial_condition=1)
    solver.fit(batch_size=1000, sampler=s, niters=2000, lr=0.01)
    assert np.mean(solver.losses[-50:]) < 0.2 * np.mean(solver.losses[:10])
    xs = np.linspace(0, 1, 50)
    approx = solver.predict(xs, 2.0).reshape(-1)
    assert np.abs(approx - (np.sin(2.0 * np.pi * xs) + 1)).max() < 0.5


def test_tutorial_ode_accuracy():
    """ tutorial problem 1: u' = 2 pi cos(2 pi x), u(0) = .5 -> sin(2 pi x) + .5 """
    torch.manual_seed(0)
    solver = Solver(lambda f, x: D(f, x) - 2 * np.pi * torch.cos(2 * np.pi * x), ndims=1, initial_condition=.5,
                    activation='Tanh', layout='fafaf', features=[12, 10, 1])
    solver.fit(niters=1500, batch_size=400, lr=0.02)
    xs = np.linspace(0, 1, 100)
    err = np.abs(solver.predict(xs).reshape(-1) - (np.sin(2 * np.pi * xs) + .5)).max()
    assert err < 0.1


def test_variable_constraint_hybrid():
    def odevar(f, x):
        return D(f, x) - 2 * np.pi * torch.cos(2 * np.pi * x) + V('new_var', data=torch.Tensor([1.0]))
    torch.manual_seed(0)
    solver = Solver(odevar, ndims=1, initial_condition=1, constraints=lambda f, x: f(torch.tensor([0.5])))
    solver.model.freeze_trainable(variables=('new_var',))
    solver.fit(niters=200, batch_size=500, lr=0.1)
    assert float(solver.model.new_var.detach()) == 1.0 and solver._engine is not None
    solver.model.unfreeze_trainable(variables=['new_var'])
    solver.fit(niters=100, batch_size=100, lr=0.1, loss_terms=['equation', 'constraint_0'])
    assert float(solver.model.new_var.detach()) != 1.0
    assert len(solver.losses) == 300 and np.isfinite(solver.losses).all()


def test_fused_constraint_matches_autograd_constraint(monkeypatch):
    """ README.md:112-126 flow: V in the initial condition + a constraint at t = 0.5.  The constraint runs as one
    more fused launch (CUDA-graph capturable); the autograd-added constraint is the yardstick. """
    def odevar(u, t):
        return D(u, t) - 2 * np.pi * torch.cos(2 * np.pi * t)

    def initial(*args):
        return V('init', data=torch.Tensor([3.0]))

    def make():
        torch.manual_seed(0)
        return Solver(odevar, ndims=1, initial_condition=initial, layout='fafaf', features=[12, 10, 1], activation='Tanh',
                      constraints=lambda u, t: u(torch.tensor([0.5])) - 0.25)
    rng = np.random.RandomState(5)
    batches = [rng.uniform(size=(150, 1)).astype(np.float32) for _ in range(40)]
    fused = make()
    fused.fit(niters=40, batch_size=150, lr=0.05, sampler=Replay(batches), loss_terms=['equation', 'constraint_0'])
    assert fused._engine is not None and fused._engine._constraint_plans[0] is not None
    monkeypatch.setenv('PYDENS_B200_FUSED_CONSTRAINTS', '0')
    hybrid = make()
    hybrid.fit(niters=40, batch_size=150, lr=0.05, sampler=Replay(batches), loss_terms=['equation', 'constraint_0'])
    assert hybrid._engine._constraint_plans[0] is None
    a, b = np.asarray(fused.losses, dtype=np.float64), np.asarray(hybrid.losses, dtype=np.float64)
    assert np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-6)) <= 2e-3
    assert abs(float(fused.model.init.detach()) - float(hybrid.model.init.detach())) <= 1e-4
    assert float(fused.model.init.detach()) != 3.0
    monkeypatch.delenv('PYDENS_B200_FUSED_CONSTRAINTS')
    # in-kernel sampling + graph replay, then the README's second stage: only the variable trains
    solver = make()
    solver.fit(niters=200, batch_size=150, lr=0.05, loss_terms=['equation', 'constraint_0'])
    solver.model.freeze_layers(['fc1', 'fc2', 'fc3'], ['log_scale'])
    w0 = solver.model.conv_block.linears[0].weight.detach().clone()
    solver.fit(niters=100, batch_size=150, lr=0.05, loss_terms=['equation', 'constraint_0'])
    assert torch.equal(w0, solver.model.conv_block.linears[0].weight.detach())
    assert np.isfinite(solver.losses).all() and np.mean(solver.losses[-20:]) < np.mean(solver.losses[:20])
    assert abs(float(solver.predict(0.5).reshape(-1)[0]) - 0.25) < 0.2          # the constraint pulls u(0.5) to 0.25


def test_plans_sharing_a_kernel_instantiation_coexist():
    """ The dynamic shared-memory limit belongs to the kernel, not to the plan: a later, smaller plan of the same
    instantiation must not lower it under an earlier plan that is still in use. """
    g = load_golden('poisson2d')
    big = make_solver('poisson2d', g['params'])
    ref_loss, ref_grads, _ = big.loss_and_grads(g['points'])
    ref_u = big.predict(g['points'][:, 0], g['points'][:, 1])

    def pde(f, x, y):
        return D(D(f, x), x) + D(D(f, y), y) - 5 * torch.sin(np.pi * (x + y))
    small = Solver(pde, ndims=2, boundary_condition=1, layout='faf', features=[3, 1], activation='Tanh')
    small.loss_and_grads(g['points'])
    small.predict(0.5, 0.5)
    loss, grads, _ = big.loss_and_grads(g['points'])
    assert loss == ref_loss and torch.equal(grads, ref_grads)
    assert np.array_equal(big.predict(g['points'][:, 0], g['points'][:, 1]), ref_u)
    big.fit(niters=3, batch_size=1000)
    small.fit(niters=3, batch_size=1000)


def test_mixture_sampler_runs_in_kernel():
    s = Solver(lambda f, x, e: D(f, x) - e * np.pi * torch.cos(e * np.pi * x), ndims=1, nparams=1, initial_condition=1.0)
    sampler = NumpySampler('u') & (0.5 & NumpySampler('u', low=1, high=2) | NumpySampler('n', loc=4, scale=.1))
    assert sampler.device_columns() is not None
    s.fit(niters=80, batch_size=2000, sampler=sampler, lr=0.01)
    assert len(s.losses) == 80 and np.isfinite(s.losses).all()
    pts = s._engine.sample(30000, sampler.device_columns(), step=3).cpu().numpy()
    assert abs((pts[:, 1] < 3).mean() - 1 / 3) < 0.015 and pts[:, 0].min() >= 0 and pts[:, 0].max() < 1


def test_autograd_path_on_gpu_matches_fused_step():
    """ backend='torch' (device-aware restatement of the reference loop) and the fused kernel agree. """
    g = load_golden('heat_param')
    fused = make_solver('heat_param', g['params'])
    loss, grads, _ = fused.loss_and_grads(g['points'])
    ref = make_solver('heat_param', backend='torch')
    # copy the weights into the autograd solver
    eng = fused._get_engine()
    with torch.no_grad():
        for (p_f, _), p_t in zip(eng.entries, [q for q in ref.model.conv_block.parameters()]):
            pass
    lin_f = fused.model.conv_block.linears
    lin_t = ref.model.conv_block.linears
    with torch.no_grad():
        for a, b in zip(lin_f, lin_t):
            b.weight.copy_(a.weight); b.bias.copy_(a.bias)
    pts = torch.from_numpy(g['points']).cuda()
    xs = [pts[:, i:i + 1].clone().requires_grad_() for i in range(pts.shape[1])]
    u = ref.ctx.run(ref.model, ref.reshape_and_concat(xs))
    r = ref.ctx.run(ref.equation, u, *xs)
    l = torch.nn.functional.mse_loss(r, torch.zeros_like(r))
    assert abs(float(l) - loss) <= 1e-5 * abs(loss)


def test_network_too_large_for_the_kernel_falls_back_loudly():
    """ 300-wide layers: the weight matrices do not fit shared memory -> pinn_plan_create says
    PINN_E_UNSUPPORTED; backend='auto' warns and trains on the autograd path, backend='fused' raises. """
    from pydens_b200 import _native

    def ode(f, x):
        return D(f, x) - 2 * np.pi * torch.cos(2 * np.pi * x)
    solver = Solver(ode, ndims=1, initial_condition=.5, layout='fafaf', features=[300, 300, 1], activation='Tanh')
    with pytest.warns(UserWarning):
        solver.fit(niters=3, batch_size=64)
    assert len(solver.losses) == 3 and solver._engine is None
    assert solver.predict(np.linspace(0, 1, 5)).shape == (5, 1)
    strict = Solver(ode, ndims=1, initial_condition=.5, layout='fafaf', features=[300, 300, 1], activation='Tanh',
                    backend='fused')
    with pytest.raises(_native.NativeError):
        strict.fit(niters=1, batch_size=8)


def test_limits_eight_columns_and_deep_network():
    """ PINN_MAX_DIMS point columns (bias of the first layer takes the separate path), 10 dense layers. """
    def eq(f, x, y, z, t, a, b, c, d):
        return D(D(f, x), x) + D(f, t) * a - b * c + d * D(f, y)
    torch.manual_seed(3)
    solver = Solver(eq, ndims=4, nparams=4, initial_condition=lambda x, y, z: x * y + z, boundary_condition=0.25,
                    layout='fa' * 9 + 'f', features=[6, 7, 8, 9, 10, 9, 8, 7, 6, 1], activation='Tanh', backend='fused')
    pts = torch.rand(1000, 8)
    loss, grads, _ = solver.loss_and_grads(pts)
    ref = Solver(eq, ndims=4, nparams=4, initial_condition=lambda x, y, z: x * y + z, boundary_condition=0.25,
                 layout='fa' * 9 + 'f', features=[6, 7, 8, 9, 10, 9, 8, 7, 6, 1], activation='Tanh', backend='torch')
    with torch.no_grad():
        for a, b in zip(solver.model.conv_block.linears, ref.model.conv_block.linears):
            b.weight.copy_(a.weight); b.bias.copy_(a.bias)
    xs = [pts[:, i:i + 1].cuda().clone().requires_grad_() for i in range(8)]
    u = ref.ctx.run(ref.model, ref.reshape_and_concat(xs))
    r = ref.ctx.run(ref.equation, u, *xs)
    l = torch.nn.functional.mse_loss(r, torch.zeros_like(r))
    l.backward()
    assert abs(float(l.detach()) - loss) <= 1e-5 * abs(loss)
    g_ref = torch.cat([p.grad.reshape(-1) for lin in ref.model.conv_block.linears for p in (lin.weight, lin.bias)])
    g_our = grads[:g_ref.numel()]
    assert float((g_our - g_ref).norm() / g_ref.norm()) <= 1e-4


# ---------------------------------------------------------------------------------------------------------------
# persistent multi-step kernel: k whole optimizer steps (Adam included) per launch — Solver.fit(steps_per_launch=k)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('kernel', ['small', 'tile'])
@pytest.mark.parametrize('name,k', [('poisson2d', 8), ('poisson2d', 40), ('ode_param', 5), ('burgers', 7), ('heat1d_icvar', 4),
                                    ('heat2d', 6), ('wave3d', 4), ('mixed_ic', 5), ('poisson_sin', 10)])
def test_persistent_kernel_follows_the_reference_fit(name, k, kernel, monkeypatch):
    """ README regime (batch 100): the loss curve of the in-kernel Adam loop against the reference's own Solver.fit
    on identical points (golden trajectory), same tolerances as the one-launch-per-step path.  Two kernels serve
    `steps_per_launch`: the (point, unit)-parallel one for batches <= 128 ('small') and the thread-per-point one. """
    monkeypatch.setenv('PINN_MULTI_KERNEL', kernel)
    monkeypatch.setenv('PINN_FORCE_KERNEL', 'thread')       # the wide-net problems too (the tile kernel has no multi-step form)
    g = load_golden(name)
    niters, batch, lr = int(g['traj_meta'][0]), int(g['traj_meta'][1]), float(g['traj_meta'][2])
    solver = make_solver(name, g['params'])
    batches = [P.make_points(name, batch, seed=1000 + i) for i in range(niters)]
    solver.fit(niters=niters, batch_size=batch, sampler=Replay(batches), lr=lr, steps_per_launch=k)
    losses = np.asarray(solver.losses, dtype=np.float64)
    ref = g['traj_losses'].astype(np.float64)
    assert losses.shape == ref.shape
    assert np.max(np.abs(losses - ref) / np.maximum(np.abs(ref), 1e-6)) <= 2e-3
    assert abs(losses[-1] - ref[-1]) <= 1e-5 * max(1.0, abs(ref[-1]))
    final = solver.flat_params().cpu().numpy()
    assert np.linalg.norm(final - g['traj_params']) / np.linalg.norm(g['traj_params']) <= 1e-3


def test_persistent_kernel_equals_stepwise_fit_and_keeps_the_optimizer_state(monkeypatch):
    monkeypatch.setenv('PYDENS_B200_AUTO_PERSISTENT', '0')             # only an explicit steps_per_launch goes persistent here
    g = load_golden('poisson2d')
    a = make_solver('poisson2d', g['params'])
    a.fit(niters=24, batch_size=100, lr=0.005)                           # in-kernel sampler, one launch per step
    a.fit(niters=12, batch_size=100, optimizer=None)
    b = make_solver('poisson2d', g['params'])
    b.fit(niters=24, batch_size=100, lr=0.005, steps_per_launch=10)      # 10 + 10 + 4 steps in three launches
    b.fit(niters=12, batch_size=100, optimizer=None)                     # torch's Adam continues on the same state
    la, lb = np.asarray(a.losses, dtype=np.float64), np.asarray(b.losses, dtype=np.float64)
    assert la.shape == lb.shape == (36,)
    assert np.max(np.abs(la - lb) / np.maximum(np.abs(la), 1e-6)) <= 1e-3
    c = make_solver('poisson2d', g['params'])
    c.fit(niters=24, batch_size=100, lr=0.005)
    c.fit(niters=12, batch_size=100, optimizer=None, steps_per_launch=6)  # and the other way round
    lc = np.asarray(c.losses, dtype=np.float64)
    assert np.max(np.abs(la - lc) / np.maximum(np.abs(la), 1e-6)) <= 1e-3


def test_adam_in_the_step_kernel_equals_torch_adam_and_shares_its_state(monkeypatch):
    """ pinn_step_adam against pinn_step + torch's fused Adam + pinn_record_loss: same loss curve, same parameters, and
    the torch optimizer object sees the state the kernel wrote (step counters, moments) — a later fit that has to call
    optimizer.step() itself (an autograd constraint) continues from it. """
    g = load_golden('poisson2d')
    runs = {}
    for mode in ('1', '0'):
        monkeypatch.setenv('PYDENS_B200_FUSED_ADAM', mode)
        s = make_solver('poisson2d', g['params'])
        s.fit(niters=30, batch_size=3000, lr=0.005)                      # in-kernel sampler, graph replay
        s.fit(niters=13, batch_size=3000, optimizer=None)                # continues on the same optimizer
        st = [s.optimizer.state[q] for q in s.optimizer.param_groups[0]['params']]     # one fixed order for both modes
        assert all(abs(float(x['step']) - 43.0) < 0.5 for x in st)
        runs[mode] = (np.asarray(s.losses, dtype=np.float64), s.flat_params().cpu().numpy(),
                      torch.cat([x['exp_avg'].reshape(-1) for x in st]).cpu().numpy())
    la, lb = runs['1'][0], runs['0'][0]
    assert la.shape == lb.shape == (43,)
    assert np.max(np.abs(la - lb) / np.maximum(np.abs(lb), 1e-6)) <= 1e-4
    assert np.linalg.norm(runs['1'][1] - runs['0'][1]) / np.linalg.norm(runs['0'][1]) <= 1e-5
    assert np.linalg.norm(runs['1'][2] - runs['0'][2]) / np.linalg.norm(runs['0'][2]) <= 1e-3
    # frozen parameters stay put under the in-kernel update
    monkeypatch.setenv('PYDENS_B200_FUSED_ADAM', '1')
    s = make_solver('poisson2d', g['params'])
    s.model.freeze_trainable(variables=['log_scale'])
    before = float(s.model.log_scale.detach())
    w_before = s.flat_params().cpu().numpy().copy()
    s.fit(niters=10, batch_size=500, lr=0.005)
    assert float(s.model.log_scale.detach()) == before
    assert np.abs(s.flat_params().cpu().numpy() - w_before).max() > 0
    assert len(s.losses) == 10 and np.isfinite(np.asarray(s.losses)).all()


def test_tiny_batches_go_to_the_persistent_kernel_by_themselves(monkeypatch):
    """ The README call as written (batch_size=100, no steps_per_launch): same curve as one launch per step, and the
    engine really took the persistent path (no per-step graphs were captured). """
    g = load_golden('poisson2d')
    monkeypatch.setenv('PYDENS_B200_AUTO_PERSISTENT', '0')
    a = make_solver('poisson2d', g['params'])
    a.fit(niters=120, batch_size=100, lr=0.005)
    assert a._engine._graphs
    monkeypatch.setenv('PYDENS_B200_AUTO_PERSISTENT', '1')
    b = make_solver('poisson2d', g['params'])
    b.fit(niters=120, batch_size=100, lr=0.005)
    assert not b._engine._graphs
    la, lb = np.asarray(a.losses, dtype=np.float64), np.asarray(b.losses, dtype=np.float64)
    assert la.shape == lb.shape == (120,)
    assert np.max(np.abs(la - lb) / np.maximum(np.abs(la), 1e-6)) <= 2e-3
    b.fit(niters=10, batch_size=100, optimizer=None)                     # too few iterations: stepwise, same optimizer state
    assert len(b.losses) == 130 and np.isfinite(np.asarray(b.losses)).all()


def test_persistent_kernel_respects_frozen_parameters():
    g = load_golden('heat2d')
    solver = make_solver('heat2d', g['params'])
    solver.model.freeze_trainable(variables=['log_scale'])
    before = float(solver.model.log_scale)
    w_before = solver.flat_params().cpu().numpy().copy()
    solver.fit(niters=6, batch_size=64, lr=0.001, steps_per_launch=3)
    assert float(solver.model.log_scale) == before
    assert np.abs(solver.flat_params().cpu().numpy() - w_before).max() > 0
