""" The reference README's parametric example (README.md:76-90): one network for a whole family of ODEs,
u' = eps * pi * cos(eps * pi * x), u(0) = 1, eps ~ U[1, 5).  The sampler is lowered to the in-kernel Philox
generator, so the training loop never touches the host.

    python examples/parametric_family.py
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pydens import Solver, NumpySampler, D                     # noqa: E402


def odeparam(f, x, e):
    return D(f, x) - e * np.pi * torch.cos(e * np.pi * x)


s = NumpySampler('uniform') & NumpySampler('uniform', low=1, high=5)
solver = Solver(equation=odeparam, ndims=1, nparams=1, initial_condition=1)
start = time.time()
solver.fit(batch_size=1000, sampler=s, niters=5000, lr=0.01)
print('fit: %.3f s, final loss %.4f' % (time.time() - start, float(solver.losses[-1])))
xs = np.linspace(0, 1, 200)
for eps in (1.5, 3.0, 4.5):
    err = np.abs(solver.predict(xs, eps).reshape(-1) - (np.sin(eps * np.pi * xs) + 1)).max()
    print('eps = %.1f: max |u - exact| = %.3f' % (eps, err))
