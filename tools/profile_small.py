""" One launch of the tiny-batch kernel (README problem, batch 100, 50 steps per launch) for ncu. """
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault('PYDENS_B200_PROGRESS', '0')
import numpy as np, torch
from pydens_b200 import Solver, D

def pde(f, x, y):
    return D(D(f, x), x) + D(D(f, y), y) - 5 * torch.sin(np.pi * (x + y))

torch.manual_seed(0)
solver = Solver(pde, ndims=2, boundary_condition=1, layout='fa fa fa f', activation='Tanh', units=[10, 12, 15, 1])
for _ in range(4):
    solver.fit(batch_size=100, niters=50, steps_per_launch=50)
torch.cuda.synchronize()
print('ok', float(solver.losses[-1]))
