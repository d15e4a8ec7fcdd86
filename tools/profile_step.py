""" Launch the fused step kernel a few times with plain launches (for ncu).

    ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 3 -c 1 \
        -o gpurun_out/prof python tools/profile_step.py cfg2
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
    workload = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
    n_launch = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    sampled = len(sys.argv) > 3 and sys.argv[3] == 'sampled'
    name, batch, lr = WORKLOADS[workload]
    cfg = P.PROBLEMS[name]
    torch.manual_seed(0)
    solver = Solver(P.bind(name, D, lambda n, init: V(n, data=torch.Tensor([init]))), ndims=cfg['ndims'],
                    nparams=cfg['nparams'], initial_condition=cfg['ic'], boundary_condition=cfg['bc'],
                    domain=cfg['domain'], layout=cfg['layout'], features=cfg['features'],
                    activation=cfg['activation'], device='cuda', backend='fused', seed=123)
    eng = solver._get_engine()
    pts = torch.from_numpy(P.make_points(name, batch, seed=3)).cuda()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n_launch + 1)]
    ev[0].record()
    for i in range(n_launch):
        eng._step(None if sampled else pts, None, batch, 1.0 / batch, 0, use_counter=False, step_value=i)
        ev[i + 1].record()
    torch.cuda.synchronize()
    i = eng.info
    print('%s: %d threads/CTA, %d B smem, %d regs, smem-resident=%d' % (workload, i.threads_per_cta, i.smem_bytes,
                                                                         i.regs_per_thread, i.activations_in_smem))
    print('launch ms:', ['%.3f' % ev[k].elapsed_time(ev[k + 1]) for k in range(n_launch)])


if __name__ == '__main__':
    main()
