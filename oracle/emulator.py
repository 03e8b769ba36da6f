"""Build and call the host emulation of the augmented-Lagrangian vector phase  --  TEST INFRASTRUCTURE ONLY.

``oracle/al_emulate.cpp`` replays the launch sequence of the CUDA solver on the CPU with the arithmetic header the
kernel itself includes (``optiml_b200/csrc/al_math.cuh``).  Only ``tests/`` import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(HERE, '_build')
LIB = os.path.join(BUILD_DIR, 'libal_emulate.so')
RULES = {'adagrad': 0, 'sgd': 1, 'rmsprop': 2, 'adadelta': 3, 'adam': 4, 'amsgrad': 5, 'adamax': 6}
MOMENTUM = {'none': 0, 'polyak': 1, 'nesterov': 2}
STATUS = {0: 'unknown', 1: 'optimal', 2: 'stopped'}


def build():
    src = os.path.join(HERE, 'al_emulate.cpp')
    hdr = os.path.join(os.path.dirname(HERE), 'optiml_b200', 'csrc', 'al_math.cuh')
    os.makedirs(BUILD_DIR, exist_ok=True)
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(['g++', '-O2', '-ffp-contract=off', '-std=c++17', '-shared', '-fPIC', '-o', LIB, src], check=True)
    return LIB


class Result:
    pass


def al_emulate(M, q, lb, ub, x0, A=None, b=0., rho=1., rule='adagrad', momentum_type='none', step_size=1.,
               momentum=0.9, decay=0.9, beta1=0.9, beta2=0.999, offset=None, tol=1e-8, epochs=1000, svr=False,
               finalise_every=0):
    lib = C.CDLL(build())
    f64 = lambda v: np.ascontiguousarray(v, dtype=np.float64)
    M, q, lb, ub, x0 = f64(M), f64(q), f64(lb), f64(ub), f64(x0)
    n, nv = M.shape[0], len(q)
    assert nv == (2 * n if svr else n)
    if offset is None:
        offset = 1e-6 if rule == 'adadelta' else 1e-8
    lr = f64(np.broadcast_to(step_size, (epochs,)))
    mom = f64(np.broadcast_to(momentum, (epochs + 1,)))
    A = None if A is None else f64(A)
    out = Result()
    out.x, out.g_x, out.lam_lb, out.lam_ub = (np.empty(nv) for _ in range(4))
    fh, ph = np.full(epochs + 1, np.nan), np.full(epochs + 1, np.nan)
    mu, it, st = C.c_double(0), C.c_int64(0), C.c_int(0)
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    lib.al_emulate.argtypes = [C.c_int64, C.c_int] + [C.c_void_p] * 6 + [C.c_double, C.c_double, C.c_int, C.c_int,
                                                                          C.c_void_p, C.c_void_p] + \
        [C.c_double] * 5 + [C.c_int64, C.c_int] + [C.c_void_p] * 4 + [C.POINTER(C.c_double), C.c_void_p, C.c_void_p,
                                                                      C.POINTER(C.c_int64), C.POINTER(C.c_int)]
    lib.al_emulate(n, int(svr), p(M), p(q), p(lb), p(ub), p(x0), p(A), float(b), float(rho), RULES[rule],
                   MOMENTUM[momentum_type], p(lr), p(mom), float(decay), float(beta1), float(beta2), float(offset),
                   float(tol), int(epochs), int(finalise_every), p(out.x), p(out.g_x), p(out.lam_lb), p(out.lam_ub),
                   C.byref(mu), p(fh), p(ph), C.byref(it), C.byref(st))
    out.mu, out.iter, out.status = mu.value, it.value, STATUS[st.value]
    out.f_hist, out.pf_hist = fh[:out.iter + 1], ph[:out.iter + 1]
    out.dual_x = np.concatenate(([out.mu] if A is not None else [], out.lam_lb, out.lam_ub))
    return out
