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
