""" Time the tensor-core tile kernel (cfg5 by default) for several grid sizes (PINN_WIDE_CTAS, read at every launch):
fewer CTAs = a smaller global slab (0.9 MB per CTA for cfg5) = more of it resident in L2.

    python tools/sweep_wide_ctas.py cfg5 148 136 128 120 112 96 74
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
os.environ.setdefault('PYDENS_B200_PROGRESS', '0')

import torch                      # noqa: E402
import problems as P              # noqa: E402
from bench import WORKLOADS       # noqa: E402
from pydens_b200 import Solver, D, V   # noqa: E402


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else 'cfg5'
    grids = [int(v) for v in sys.argv[2:]] or [148, 128, 112, 96, 74]
    name, batch, lr = WORKLOADS[workload]
    cfg = P.PROBLEMS[name]
    torch.manual_seed(0)
    solver = Solver(P.bind(name, D, lambda n, init: V(n, data=torch.Tensor([init]))), ndims=cfg['ndims'],
                    nparams=cfg['nparams'], initial_condition=cfg['ic'], boundary_condition=cfg['bc'],
                    domain=cfg['domain'], layout=cfg['layout'], features=cfg['features'],
                    activation=cfg['activation'], device='cuda', backend='fused', seed=123)
    eng = solver._get_engine()
    pts = torch.from_numpy(P.make_points(name, batch, seed=3)).cuda()
    ref = None
    for g in grids:
        os.environ['PINN_WIDE_CTAS'] = str(g)
        n = 5
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        eng._step(pts, None, batch, 1.0 / batch, 0, use_counter=False, step_value=0)      # warm
        ev[0].record()
        for i in range(n):
            eng._step(pts, None, batch, 1.0 / batch, 0, use_counter=False, step_value=0)
            ev[i + 1].record()
        torch.cuda.synchronize()
        out = eng.out.clone()
        if ref is None:
            ref = out
        rel = float((out - ref).norm() / ref.norm())
        ms = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(n))
        print('ctas %4d: median %.3f ms  min %.3f  (grad+loss vs first grid: rel %.2e)' % (g, ms[n // 2], ms[0], rel), flush=True)


if __name__ == '__main__':
    main()
