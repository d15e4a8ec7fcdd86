""" Shared helpers of the test-suite (oracle side). """
import os

import numpy as np
import torch

import problems as P
from oracle import autograd_port as ap

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + '.npz')))


def oracle_problem(name, dtype=torch.float32, params=None):
    cfg = P.PROBLEMS[name]
    holder = {}
    ic = P.make_ic(name, lambda n, init: holder['prob'].V(n, init))
    prob = ap.Problem(lambda u, *xs, D, V: cfg['equation'](u, *xs, D=D, V=V),
                      ndims=cfg['ndims'], nparams=cfg['nparams'], initial_condition=ic,
                      boundary_condition=cfg['bc'], domain=cfg['domain'], features=cfg['features'],
                      activation=cfg['activation'], dtype=dtype, variables=cfg.get('variables'), layout=cfg['layout'])
    holder['prob'] = prob
    if params is not None:
        prob.load_flat(torch.as_tensor(params))
    return prob


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def criterion_case(kind, residual):
    """ -> (tracer criterion key, the torch criterion it stands for); thresholds at the median |residual| of the
    problem's golden so that both branches of Huber / SmoothL1 are met """
    med = float(np.float32(np.median(np.abs(residual))))
    return {'l1': (('l1',), torch.nn.L1Loss()),
            'huber': (('huber', med), torch.nn.HuberLoss(delta=med)),
            'huber_default': (('huber', 1.0), torch.nn.HuberLoss()),
            'smooth_l1': (('smooth_l1', med), torch.nn.SmoothL1Loss(beta=med)),
            'smooth_l1_zero': (('smooth_l1', 0.0), torch.nn.SmoothL1Loss(beta=0.0))}[kind]
