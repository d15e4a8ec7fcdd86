""" The oracle (oracle/autograd_port.py) pinned against outputs of the UNMODIFIED reference
(tests/golden/*.npz, written by oracle/make_golden.py in the build container). CPU only. """
import numpy as np
import pytest
import torch

import problems as P
from helpers import load_golden, oracle_problem, rel_l2
from oracle import autograd_port as ap


@pytest.mark.parametrize('name', list(P.PROBLEMS))
def test_port_matches_reference_fp32(name):
    g = load_golden(name)
    prob = oracle_problem(name, torch.float32, g['params'])
    loss, residual, grads = prob.loss_and_grads(g['points'])
    # same ATen ops in the same order as the reference: agreement to fp32 rounding noise
    assert abs(loss - float(g['loss'])) <= 2e-6 * abs(float(g['loss']))
    assert rel_l2(residual, g['residual']) <= 2e-6
    assert rel_l2(grads.numpy(), g['grads']) <= 2e-5
    assert rel_l2(prob.predict(g['points']), g['u']) <= 2e-6


@pytest.mark.parametrize('name', list(P.PROBLEMS))
def test_fp64_port_brackets_reference(name):
    """ fp64 evaluation of the same weights: the reference's fp32 result is within fp32 noise of it. """
    g = load_golden(name)
    prob = oracle_problem(name, torch.float64, g['params'].astype(np.float64))
    loss, residual, grads = prob.loss_and_grads(g['points'].astype(np.float64))
    assert abs(loss - float(g['loss'])) <= 1e-4 * abs(float(g['loss']))
    assert rel_l2(g['grads'], grads.numpy()) <= 2e-3


@pytest.mark.parametrize('name', list(P.GOLDEN_TRAJ))
def test_port_trajectory_matches_reference_fit(name):
    g = load_golden(name)
    niters, batch, lr = g['traj_meta']
    niters, batch = int(niters), int(batch)
    prob = oracle_problem(name, torch.float32, g['params'])
    losses = ap.fit(prob, niters, batch, lr=float(lr),
                    point_stream=lambda i: torch.from_numpy(P.make_points(name, batch, seed=1000 + i)))
    ref = g['traj_losses']
    assert np.max(np.abs(losses - ref) / np.maximum(np.abs(ref), 1e-6)) <= 1e-3
    assert abs(losses[-1] - ref[-1]) <= 1e-5 * max(1.0, abs(ref[-1]))


@pytest.mark.parametrize('weight_decay', [0.0, 0.01])
def test_adam_restatement_matches_torch_adam(weight_decay):
    """ oracle/adam.py (the arithmetic of the step kernel's optimizer tail) against torch.optim.Adam itself — the
    optimizer the reference constructs (model_torch.py:419-422) and steps (:461) — over 200 steps of random gradients,
    with a frozen slice. """
    import torch
    from oracle import adam as oadam
    rng = np.random.default_rng(5)
    n = 373
    p0 = rng.standard_normal(n).astype(np.float32)
    w = torch.nn.Parameter(torch.from_numpy(p0.copy()))
    frozen = torch.nn.Parameter(torch.from_numpy(p0[:7].copy()), requires_grad=False)
    opt = torch.optim.Adam([w], lr=0.005, weight_decay=weight_decay)
    p, m, v = p0.copy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
    pf, mf, vf = p0[:7].copy(), np.zeros(7, np.float32), np.zeros(7, np.float32)
    for t in range(1, 201):
        g = (rng.standard_normal(n) * (10.0 ** rng.uniform(-4, 1))).astype(np.float32)
        w.grad = torch.from_numpy(g.copy())
        opt.step()
        oadam.adam_step(p, g, m, v, t, lr=0.005, weight_decay=weight_decay)
        oadam.adam_step(pf, g[:7], mf, vf, t, mask=np.zeros(7))
    ref = w.detach().numpy()
    assert np.max(np.abs(p - ref)) <= 5e-6                     # parameters are O(1): a few fp32 ulps accumulated over 200 steps
    st = opt.state[w]
    rm, rv = st['exp_avg'].numpy(), st['exp_avg_sq'].numpy()
    assert np.max(np.abs(m - rm)) <= 1e-5 * np.max(np.abs(rm))        # a signed running mean: absolute, not per element
    assert np.max(np.abs(v - rv)) <= 1e-5 * np.max(np.abs(rv))
    assert np.array_equal(pf, frozen.detach().numpy()) and not mf.any() and not vf.any()
