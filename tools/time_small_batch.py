""" Wall time of the README example (batch 100 x 1500 iterations) and of the other tutorial-sized fits. """
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault('PYDENS_B200_PROGRESS', '0')
import numpy as np, torch
from pydens_b200 import Solver, D

def pde(f, x, y):
    return D(D(f, x), x) + D(D(f, y), y) - 5 * torch.sin(np.pi * (x + y))

for per_graph in (() if 'multi' in sys.argv[1:] else ('1', '4', '16', '64')):
    os.environ['PYDENS_B200_GRAPH_STEPS'] = per_graph
    os.environ['PYDENS_B200_AUTO_PERSISTENT'] = '0'          # one launch per step
    torch.manual_seed(0)
    solver = Solver(pde, ndims=2, boundary_condition=1, layout='fa fa fa f', activation='Tanh', units=[10, 12, 15, 1])
    solver.fit(batch_size=100, niters=100)
    torch.cuda.synchronize(); t = time.perf_counter()
    solver.fit(batch_size=100, niters=1500)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print('steps/graph %3s: 1500 x batch 100 in %.1f ms  (%.1f us/step, %.3g points/s), final loss %.4f'
          % (per_graph, dt * 1e3, dt / 1500 * 1e6, 150000 / dt, float(np.mean(solver.losses[-20:]))))

# persistent multi-step kernel: k whole optimizer steps (Adam included) per launch
os.environ.pop('PYDENS_B200_GRAPH_STEPS', None)
os.environ.pop('PYDENS_B200_AUTO_PERSISTENT', None)
torch.manual_seed(0)
solver = Solver(pde, ndims=2, boundary_condition=1, layout='fa fa fa f', activation='Tanh', units=[10, 12, 15, 1])
solver.fit(batch_size=100, niters=100)
torch.cuda.synchronize(); t = time.perf_counter()
solver.fit(batch_size=100, niters=1500)                      # the README call as written
torch.cuda.synchronize(); dt = time.perf_counter() - t
print('README call as written: 1500 x batch 100 in %.1f ms  (%.1f us/step), final loss %.4f'
      % (dt * 1e3, dt / 1500 * 1e6, float(np.mean(solver.losses[-20:]))))
for k in (10, 50, 250, 1500):
    torch.manual_seed(0)
    solver = Solver(pde, ndims=2, boundary_condition=1, layout='fa fa fa f', activation='Tanh', units=[10, 12, 15, 1])
    solver.fit(batch_size=100, niters=100, steps_per_launch=k)
    torch.cuda.synchronize(); t = time.perf_counter()
    solver.fit(batch_size=100, niters=1500, steps_per_launch=k)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print('steps_per_launch %4d: 1500 x batch 100 in %.1f ms  (%.1f us/step, %.3g points/s), final loss %.4f'
          % (k, dt * 1e3, dt / 1500 * 1e6, 150000 / dt, float(np.mean(solver.losses[-20:]))))
