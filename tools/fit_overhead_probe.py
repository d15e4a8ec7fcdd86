""" Fixed host cost of one `Solver.fit` call on the host-batch path (what the K = 20 e2e figure of bench.py pays 1/20 of
per step): wall time of fits of several lengths -> slope (per step) and intercept (per call), then a cProfile of one
K = 20 fit.

    python tools/fit_overhead_probe.py
"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
os.environ.setdefault('PYDENS_B200_PROGRESS', '0')

import numpy as np                 # noqa: E402
import torch                       # noqa: E402
from bench import make_solver      # noqa: E402


def main():
    solver, cfg, lr = make_solver('cfg2', torch.device('cuda'))
    B, total = 100000, 2
    pool = [torch.rand(B, total).pin_memory() for _ in range(32)]

    class HostBatches:
        i = 0

        def sample(self, size):
            self.i += 1
            return pool[self.i % 32]
    hb = HostBatches()
    solver.fit(niters=16, batch_size=B, sampler=hb, lr=lr)
    xs, ys = [], []
    for K in (8, 20, 50, 200):
        ts = []
        for _ in range(7):
            torch.cuda.synchronize(); t = time.perf_counter()
            solver.fit(niters=K, batch_size=B, sampler=hb, lr=lr)
            torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
        med = sorted(ts)[3]
        xs.append(K); ys.append(med)
        print('K = %3d: %.1f us per fit, %.2f us per step' % (K, med * 1e6, med / K * 1e6), flush=True)
    slope, icpt = np.polyfit(xs, ys, 1)
    print('per step %.2f us, per call %.1f us' % (slope * 1e6, icpt * 1e6))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(20):
        solver.fit(niters=20, batch_size=B, sampler=hb, lr=lr)
    torch.cuda.synchronize()
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats('cumulative').print_stats(28)


if __name__ == '__main__':
    main()
