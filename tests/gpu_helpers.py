""" Helpers of the -m gpu tests: build a pydens_b200.Solver for a registry problem. """
import numpy as np
import torch

import problems as P
from pydens_b200 import Solver, D, V


def pkg_V(name, init):
    return V(name, data=torch.Tensor([init]))


def make_solver(name, params=None, backend='fused', seed=0, **kw):
    cfg = P.PROBLEMS[name]
    torch.manual_seed(seed)
    solver = Solver(P.bind(name, D, pkg_V), ndims=cfg['ndims'], nparams=cfg['nparams'],
                    initial_condition=P.make_ic(name, pkg_V), boundary_condition=cfg['bc'], domain=cfg['domain'],
                    layout=cfg['layout'], features=cfg['features'], activation=cfg['activation'],
                    device='cuda', backend=backend, seed=1234, **kw)
    if params is not None:
        solver.load_flat_params(params)
    elif 'log_scale' in cfg:
        with torch.no_grad():
            solver.model.log_scale.fill_(cfg['log_scale'])
    return solver


class Replay:
    """ Host sampler replaying recorded batches (numpy), like oracle/make_golden.py's. """

    def __init__(self, batches):
        self.batches, self.i = batches, 0

    def sample(self, size):
        b = self.batches[self.i]
        self.i += 1
        assert b.shape[0] == size
        return b
