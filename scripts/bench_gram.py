"""K1 timing + parity spot-check on the BASELINE shapes (1 GPU)."""
import ctypes as C, sys, time
sys.path.insert(0, '.')
import numpy as np
from optiml_b200 import _native as N
from optiml_b200.configs import make_config
from optiml_b200.runtime import default_context
from oracle import svm_oracle as O

ctx = default_context()
for cfg, n, kid, kind in (('C4', 50000, 2, 'gaussian'), ('C3', 20000, 0, 'linear'), ('C2', 10000, 1, 'poly'), ('C1', 2000, 2, 'gaussian'),
                          ('C4', 50000, 4, 'laplacian'), ('C4', 50000, 3, 'sigmoid')):
    spec, X, y = make_config(cfg, n=n)
    d = X.shape[1]
    gamma = 1. / (d * X.var())
    dX = ctx.upload_matrix(X)
    ld = N.padded_ld(n)
    dQ = ctx.malloc(n * ld * 8)
    args = (ctx.handle, C.c_void_p(dX.dptr), n, dX.ld, C.c_void_p(dX.dptr), n, dX.ld, d, 1, kid, gamma, 0., 3., None, None, 0.0, 0, n, C.c_void_p(dQ), ld)
    for _ in range(2):
        N.call('svmb200_gram', *args)
    ctx.sync()
    ts = []
    for _ in range(5):
        ctx.timer_start(); N.call('svmb200_gram', *args); ts.append(ctx.timer_stop_ms())
    ms = min(ts)
    rows = np.empty((128, ld))
    ctx.d2h(rows, dQ + 4096 * ld * 8 if n > 8192 else dQ)
    r0 = 4096 if n > 8192 else 0
    want = O.kernel_matrix(kind, X[r0:r0 + 128], X, gamma=gamma, degree=3)
    if kind == 'gaussian':
        want[np.arange(128), np.arange(r0, r0 + 128)] = 1.0
    flops = 2.0 * n * n * d
    err = np.abs(rows[:, :n] - want) / (np.abs(want) + 1e-3 * np.abs(want).max())
    print(f'{cfg} n={n} d={d} {kind}: {ms:.3f} ms  {2.0*n*n*d/ms/1e9:.2f} TFLOP/s  write {8.0*n*ld/ms/1e6:.0f} GB/s  max rel err {err.max():.2e}', flush=True)
    ctx.free(dQ); dX.release()
