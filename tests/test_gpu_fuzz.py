""" The randomised problems of test_emul_fuzz.py through the bare C ABI on the GPU, against the fp64 autograd
oracle: random widths / activations / residual layouts / ansatz configurations / equations, batch sizes that
leave ragged tiles.  Covers every kernel variant family (plain and general, shared- and global-memory state). """
import numpy as np
import pytest
import torch

from helpers import rel_l2
from test_emul_fuzz import _random_problem, _layer_plan

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from gpu_helpers import abi_step
    from oracle import autograd_port as ap
    from pydens_b200 import _native as N, tracer as T


@pytest.mark.parametrize('seed', list(range(60)))
def test_random_problem_on_gpu_matches_fp64_oracle(seed):
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
    params = prob.flat_params().numpy().astype(np.float32)
    rng = np.random.RandomState(2000 + seed)
    n = int(rng.choice([1, 31, 257, 3000]))
    pts = np.concatenate([rng.uniform(lo, hi, size=(n, 1)) for lo, hi in cfg['ranges']], axis=1).astype(np.float32)
    loss, residual, grads, u = abi_step(spec, params, pts)
    prob.load_flat(torch.from_numpy(params.astype(np.float64)))
    ref_loss, ref_res, ref_grads = prob.loss_and_grads(pts.astype(np.float64))
    tag = '%s %s %s acts=%s n=%d' % (cfg['eq_name'], cfg['layout'], cfg['features'], acts, n)
    # a residual far below its O(1) terms is a cancellation: fp32 rounding of the terms (~1e-7 absolute) then shows
    # up magnified in every RELATIVE error, for torch's fp32 as well — the tolerances are relaxed by that factor
    cond = max(1.0, 0.05 / max(float(np.sqrt(np.mean(np.square(ref_res)))), 1e-30))
    assert abs(loss - ref_loss) <= 2e-5 * cond * max(abs(ref_loss), 1e-6), tag
    assert rel_l2(residual, ref_res) <= 2e-5 * cond, tag
    assert rel_l2(grads, ref_grads.numpy()) <= 1e-4 * cond, tag
    ref_u = prob.predict(pts.astype(np.float64))
    # u is a sum of O(1) terms that may cancel: the error is measured against that scale, not against |u|
    assert np.abs(u - ref_u).max() <= 1e-5 * max(1.0, np.abs(ref_u).max()), tag
