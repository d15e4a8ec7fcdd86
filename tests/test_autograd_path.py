""" The package's autograd path (`backend='torch'`: the device-aware restatement of the reference loop that runs
whatever the fused kernel does not cover) against the goldens of the unmodified reference — on CPU.  It is the
host-side mirror of the reference interface (model_torch.py:19-178, 426-464): same modules, same ansatz, same `D`. """
import numpy as np
import pytest
import torch

import problems as P
from helpers import load_golden, rel_l2
from pydens_b200 import Solver, D, V


def pkg_V(name, init):
    return V(name, data=torch.Tensor([init]))


def cpu_solver(name, params=None):
    cfg = P.PROBLEMS[name]
    torch.manual_seed(0)
    solver = Solver(P.bind(name, D, pkg_V), ndims=cfg['ndims'], nparams=cfg['nparams'],
                    initial_condition=P.make_ic(name, pkg_V), boundary_condition=cfg['bc'], domain=cfg['domain'],
                    layout=cfg['layout'], features=cfg['features'], activation=cfg['activation'],
                    device='cpu', backend='torch')
    if params is not None:
        with torch.no_grad():
            off = 0
            for p in flat_parameters(solver, name):
                n = p.numel()
                p.copy_(torch.from_numpy(params[off:off + n]).reshape(p.shape))
                off += n
    return solver


def flat_parameters(solver, name):
    """ parameters in the golden layout: W_0, b_0, …, log_scale, variables in registry order """
    model = solver.model
    out = []
    for lin in model.conv_block.linears:
        out += [lin.weight, lin.bias]
    out.append(model.log_scale)
    out += [getattr(model, v) for v in P.PROBLEMS[name].get('variables', {})]
    return out


@pytest.mark.parametrize('name', list(P.PROBLEMS))
def test_autograd_path_matches_reference_on_explicit_points(name):
    g = load_golden(name)
    solver = cpu_solver(name, g['params'])
    pts = torch.from_numpy(g['points'])
    xs = [pts[:, i:i + 1].clone().requires_grad_() for i in range(pts.shape[1])]
    u = solver.ctx.run(solver.model, solver.reshape_and_concat(xs))
    residual = solver.ctx.run(solver.equation, u, *xs)
    loss = torch.nn.MSELoss()(residual, torch.zeros_like(xs[0]))
    loss.backward()
    params = flat_parameters(solver, name)
    grads = np.concatenate([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).numpy() for p in params])
    n = grads.size
    assert abs(float(loss.detach()) - float(g["loss"])) <= 2e-6 * abs(float(g['loss']))
    assert rel_l2(residual.detach().numpy().reshape(-1), g['residual']) <= 2e-6
    assert rel_l2(grads, g['grads'][:n]) <= 2e-5
    assert rel_l2(solver.predict(*[g['points'][:, i] for i in range(pts.shape[1])]).reshape(-1), g['u']) <= 2e-6


class Replay:
    def __init__(self, batches):
        self.batches, self.i = batches, 0

    def sample(self, size):
        b = self.batches[self.i]
        self.i += 1
        return b


@pytest.mark.parametrize('name', list(P.GOLDEN_TRAJ))
def test_autograd_path_fit_follows_reference_fit(name):
    g = load_golden(name)
    niters, batch, lr = int(g['traj_meta'][0]), int(g['traj_meta'][1]), float(g['traj_meta'][2])
    solver = cpu_solver(name, g['params'])
    batches = [P.make_points(name, batch, seed=1000 + i) for i in range(niters)]
    solver.fit(niters=niters, batch_size=batch, sampler=Replay(batches), lr=lr)
    losses = np.asarray(solver.losses, dtype=np.float64)
    ref = g['traj_losses'].astype(np.float64)
    assert losses.shape == ref.shape
    assert np.max(np.abs(losses - ref) / np.maximum(np.abs(ref), 1e-6)) <= 1e-3
    final = np.concatenate([p.detach().reshape(-1).numpy() for p in flat_parameters(solver, name)])
    assert rel_l2(final, g['traj_params'][:final.size]) <= 1e-3
    assert all(isinstance(l, np.ndarray) and l.ndim == 0 for l in solver.losses)      # 0-d arrays, like :464
