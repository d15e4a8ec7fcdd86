import os, sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
os.environ.setdefault('PYDENS_B200_PROGRESS', '0')
import numpy as np, torch
import problems as P
from pydens_b200 import Solver, D, V
cfg = P.PROBLEMS['poisson2d']
def mk():
    torch.manual_seed(0)
    return Solver(P.bind('poisson2d', D, None), ndims=2, boundary_condition=1, layout=cfg['layout'], features=cfg['features'], activation='Tanh', backend='fused', seed=1)
B, K = 100000, 400
host_pool = [torch.rand(B, 2).pin_memory() for _ in range(32)]
class HB:
    i = 0
    def sample(self, n):
        self.i += 1
        return host_pool[self.i % 32]
for label, env in [('default', {}), ('no loss copy', {'PYDENS_B200_EXP_NOLOSSCOPY': '1'}), ('no graph', {'PYDENS_B200_NO_GRAPH': '1'})]:
    for k, v in env.items(): os.environ[k] = v
    s = mk(); hb = HB()
    s.fit(niters=70, batch_size=B, sampler=hb)
    torch.cuda.synchronize(); t = time.perf_counter()
    s.fit(niters=K, batch_size=B, sampler=hb, optimizer=None)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print('%-14s %.1f us/step' % (label, dt / K * 1e6))
    for k in env: os.environ.pop(k)
# pure device sampler for comparison
s = mk(); s.fit(niters=70, batch_size=B)
torch.cuda.synchronize(); t = time.perf_counter(); s.fit(niters=K, batch_size=B, optimizer=None); torch.cuda.synchronize()
print('device sampler %.1f us/step' % ((time.perf_counter() - t) / K * 1e6))
