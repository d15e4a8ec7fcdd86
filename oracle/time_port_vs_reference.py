""" TEST INFRASTRUCTURE ONLY — is the port (`oracle/autograd_port.py`, the CPU baseline `bench.py` times) as fast
as the UNMODIFIED reference on the same host?

Run in the build container only (needs /root/reference):

    python oracle/time_port_vs_reference.py > profiles/r2_port_vs_reference_cpu.json

Both arms run the same workload on the same cores in the same process: the reference's own `Solver.fit`
(`/root/reference/pydens/model_torch.py:364-464`, imported as is through oracle/batchflow_standin, progress bar
silenced) and `autograd_port.fit` exactly as `bench.py: cpu_reference` calls it.  The reference cannot travel to the
GPU box (it is not part of this repository and depends on the un-vendored batchflow), so `bench.py --impl reference`
times the port there; this script pins how the two relate where both exist.
"""
import json
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, 'batchflow_standin'))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, ROOT)

import pydens as ref                                     # noqa: E402  (the unmodified reference)
from pydens import model_torch as ref_mt                 # noqa: E402
import problems as P                                     # noqa: E402
from oracle import autograd_port as ap                   # noqa: E402

ref_mt.tqdm = lambda x, *a, **k: x                       # silence the progress bar only

# (label, problem, batch, lr, warm-up iterations, timed iterations)
CASES = [('cfg1 (README: batch 100)', 'poisson2d', 100, 0.005, 20, 300),
         ('cfg2 (batch 100 000)', 'poisson2d', 100000, 0.005, 2, 10),
         ('cfg3 sample (batch 100 000 of 1 M)', 'ode_param', 100000, 0.01, 2, 10),
         ('cfg4 sample (batch 100 000 of 1 M)', 'heat2d', 100000, 0.001, 1, 4),
         ('cfg5 sample (batch 20 000 of 500 k)', 'wave3d', 20000, 0.001, 1, 3)]


class UniformRanges:
    """ a host sampler in the reference's sense: `.sample(size)` -> float array [size, total] """

    def __init__(self, ranges):
        self.ranges = ranges

    def sample(self, size):
        return torch.cat([torch.rand((size, 1)) * (hi - lo) + lo for lo, hi in self.ranges], dim=1).numpy()


def time_reference(name, batch, lr, warm, iters):
    cfg = P.PROBLEMS[name]
    torch.manual_seed(0)
    ref_V = lambda n, init: ref.V(n, data=torch.Tensor([init]))         # noqa: E731
    solver = ref.Solver(P.bind(name, ref.D, ref_V), ndims=cfg['ndims'], nparams=cfg['nparams'],
                        initial_condition=P.make_ic(name, ref_V), boundary_condition=cfg['bc'], domain=cfg['domain'],
                        layout=cfg['layout'], features=cfg['features'], activation=cfg['activation'])
    sampler = UniformRanges(cfg['ranges'])
    solver.fit(niters=warm, batch_size=batch, sampler=sampler, lr=lr)
    t0 = time.perf_counter()
    solver.fit(niters=iters, batch_size=batch, sampler=sampler, lr=lr)
    return time.perf_counter() - t0


def time_port(name, batch, lr, warm, iters):
    cfg = P.PROBLEMS[name]
    torch.manual_seed(0)
    prob = ap.Problem(lambda u, *xs, D, V: cfg['equation'](u, *xs, D=D, V=V), ndims=cfg['ndims'],
                      nparams=cfg['nparams'], initial_condition=cfg['ic'], boundary_condition=cfg['bc'],
                      domain=cfg['domain'], features=cfg['features'], activation=cfg['activation'],
                      variables=cfg.get('variables'))
    ranges = cfg['ranges']

    def stream(i):
        return torch.cat([torch.rand((batch, 1)) * (hi - lo) + lo for lo, hi in ranges], dim=1)
    ap.fit(prob, warm, batch, lr=lr, point_stream=stream)
    t0 = time.perf_counter()
    ap.fit(prob, iters, batch, lr=lr, point_stream=stream)
    return time.perf_counter() - t0


def main():
    cores = os.cpu_count() or 1
    out = {'host_cores': cores, 'torch_threads': torch.get_num_threads(), 'torch': torch.__version__, 'cases': []}
    for label, name, batch, lr, warm, iters in CASES:
        best = {}
        for arm, fn in (('reference', time_reference), ('port', time_port)):
            best[arm] = min(fn(name, batch, lr, warm, iters) for _ in range(2))
        row = {'workload': label, 'problem': name, 'batch': batch, 'iterations': iters,
               'reference_points_per_s': batch * iters / best['reference'], 'port_points_per_s': batch * iters / best['port'],
               'reference_ms_per_step': 1e3 * best['reference'] / iters, 'port_ms_per_step': 1e3 * best['port'] / iters,
               'port_over_reference': best['reference'] / best['port']}
        out['cases'].append(row)
        sys.stderr.write('%-40s reference %9.3f ms/step   port %9.3f ms/step   port/reference speed %.3f\n'
                         % (label, row['reference_ms_per_step'], row['port_ms_per_step'], row['port_over_reference']))
    out['note'] = ('both arms on this container\'s CPU cores, PyTorch default threading, best of two runs; '
                   'port_over_reference > 1 means the port (the baseline bench.py reports) is FASTER than the unmodified '
                   'reference, i.e. the reported GPU/CPU ratio is conservative')
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
