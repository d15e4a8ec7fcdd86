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


def abi_step(spec, params, points):
    """ One fused step through the bare C ABI (include/pinn_b200.h): plan from `spec`, explicit points.
    -> (loss, residual [n], grads [n_params], u [n]) as numpy. """
    import ctypes as C
    from pydens_b200 import _native as N
    lib = N.load()
    dev = torch.device('cuda:0')
    plan = C.c_void_p()
    N.check(lib.pinn_plan_create(C.byref(spec), 0, C.byref(plan)))
    try:
        n = points.shape[0]
        flat = torch.from_numpy(np.ascontiguousarray(params, dtype=np.float32)).to(dev)
        pts = torch.from_numpy(np.ascontiguousarray(points, dtype=np.float32)).to(dev)
        out = torch.zeros(spec.n_params + 4, dtype=torch.float32, device=dev)
        res = torch.zeros(n, dtype=torch.float32, device=dev)
        u = torch.zeros(n, dtype=torch.float32, device=dev)
        ws = torch.zeros(lib.pinn_workspace_bytes(plan, n), dtype=torch.uint8, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        N.check(lib.pinn_step(plan, C.c_void_p(flat.data_ptr()), C.c_void_p(pts.data_ptr()), None, C.c_uint64(0), None,
                              C.c_uint64(0), C.c_uint64(0), C.c_int64(n), C.c_float(1.0 / n),
                              C.c_void_p(out.data_ptr()), C.c_void_p(res.data_ptr()),
                              C.c_void_p(ws.data_ptr()), C.c_size_t(ws.numel()), stream))
        N.check(lib.pinn_forward(plan, C.c_void_p(flat.data_ptr()), C.c_void_p(pts.data_ptr()), C.c_int64(n),
                                 C.c_void_p(u.data_ptr()), C.c_void_p(ws.data_ptr()), C.c_size_t(ws.numel()), stream))
        torch.cuda.synchronize(dev)
        o = out.cpu().numpy()
        return float(o[spec.n_params]), res.cpu().numpy(), o[:spec.n_params].copy(), u.cpu().numpy()
    finally:
        lib.pinn_plan_destroy(plan)
