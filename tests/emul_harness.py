""" TEST INFRASTRUCTURE: builds and drives tests/emul/libpinn_emul.so — a g++ build of the
per-thread device code (pydens_b200/csrc/pinn_device.cuh) — so the kernel math can be checked on
CPU.  Not part of the product. """
import ctypes as C
import os
import subprocess

import numpy as np

from pydens_b200 import _native as N
from pydens_b200 import tracer as T
import problems as P

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'emul', 'pinn_emul.cpp')
LIB = os.path.join(HERE, 'emul', 'libpinn_emul.so')
DEPS = [SRC] + [os.path.join(HERE, '..', 'pydens_b200', 'csrc', f) for f in ('pinn_device.cuh', 'pinn_device_hi.cuh', 'pinn_host_plan.h')] \
    + [os.path.join(HERE, '..', 'include', 'pinn_b200.h')]
_lib = None


def lib():
    global _lib
    if _lib is None:
        stale = (not os.path.exists(LIB)) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in DEPS)
        if stale:
            subprocess.check_call(['g++', '-O2', '-mfma', '-std=c++17', '-shared', '-fPIC', '-x', 'c++',
                                   '-Wno-unknown-pragmas', '-o', LIB, SRC])
        _lib = C.CDLL(LIB)
    return _lib


def traced_problem(name, criterion=None):
    cfg = P.PROBLEMS[name]
    total = cfg['ndims'] + cfg['nparams']
    nsp = cfg['ndims'] - 1 if P.has_ic(name) else cfg['ndims']
    sym_V = lambda n, init: T.Sym(T.var(n))
    eq = P.bind(name, T.sym_D, sym_V)
    return T.trace(eq, total, None, initial_condition=P.make_ic(name, sym_V), ndims_spatial=nsp, criterion=criterion)


def spec_for(name, criterion=None):
    cfg = P.PROBLEMS[name]
    total = cfg['ndims'] + cfg['nparams']
    tr = traced_problem(name, criterion)
    widths = [total] + list(cfg['features'])
    acts, skips = P.layer_plan(name)
    dom = cfg['domain']
    if isinstance(dom[0], (int, float)):
        dom = [tuple(dom)] * cfg['ndims']
    return N.build_spec(widths, acts, cfg['ndims'], cfg['nparams'], cfg['bc'] is not None,
                        cfg['bc'] if cfg['bc'] is not None else 0.0, P.has_ic(name), dom, tr, skips=skips)


def emul_step(spec, params, points):
    params = np.ascontiguousarray(params, dtype=np.float32)
    points = np.ascontiguousarray(points, dtype=np.float32)
    n = points.shape[0]
    out = np.zeros(spec.n_params + 4, dtype=np.float32)
    res = np.zeros(n, dtype=np.float32)
    msg = C.create_string_buffer(256)
    rc = lib().emul_step(C.byref(spec), params.ctypes.data_as(C.c_void_p), points.ctypes.data_as(C.c_void_p),
                         C.c_longlong(n), C.c_float(1.0 / n), out.ctypes.data_as(C.c_void_p),
                         res.ctypes.data_as(C.c_void_p), msg, 256)
    if rc:
        raise RuntimeError(msg.value.decode())
    return float(out[spec.n_params]), res, out[:spec.n_params]


def emul_forward(spec, params, points):
    params = np.ascontiguousarray(params, dtype=np.float32)
    points = np.ascontiguousarray(points, dtype=np.float32)
    n = points.shape[0]
    u = np.zeros(n, dtype=np.float32)
    msg = C.create_string_buffer(256)
    rc = lib().emul_forward(C.byref(spec), params.ctypes.data_as(C.c_void_p), points.ctypes.data_as(C.c_void_p),
                            C.c_longlong(n), u.ctypes.data_as(C.c_void_p), msg, 256)
    if rc:
        raise RuntimeError(msg.value.decode())
    return u


def emul_sample(cols, total, seed, step, offset, n):
    arr = N.make_columns(cols if cols is not None else [(0, 0.0, 1.0)] * total, total)
    out = np.zeros((n, total), dtype=np.float32)
    lib().emul_sample(arr, C.c_int(total), C.c_ulonglong(seed), C.c_ulonglong(step), C.c_ulonglong(offset),
                      C.c_longlong(n), out.ctypes.data_as(C.c_void_p))
    return out
