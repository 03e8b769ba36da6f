"""Checks of the opt-in symmetric pass (K2s, optiml_b200/csrc/k2_symv.cuh: the product Q u from the upper triangle of the
matrix alone), written once and run twice: on the host emulation of the kernels in the CPU suite
(tests/test_emulated_kernels.py, small tile shapes built with -DSVMB200_SYMV_*) and on the B200 in the `-m gpu` suite
(tests/test_gpu_symmetric.py, the shipped shape).

``device`` is a callable returning a context manager that yields a probe with ``launches()`` and ``assert_clean()``
(tests/shared_gram_checks.py).
"""
import contextlib
import ctypes as C

import numpy as np

from oracle import svm_oracle as O
from optiml_b200 import _native as N
from shared_gram_checks import psd


@contextlib.contextmanager
def symmetric_pass(on=True):
    """``runtime.use_symmetric_pass(on)`` for the duration of a test"""
    from optiml_b200 import runtime
    saved = runtime._symmetric
    runtime.use_symmetric_pass(on)
    try:
        yield
    finally:
        runtime._symmetric = saved
        if runtime._default_ctx is not None:
            N.call('svmb200_ctx_set_symmetric', runtime._default_ctx.handle, int(bool(saved)))


def _symv(ctx, dQ, n, ld, u):
    du, dw = ctx.malloc(8 * ld), ctx.malloc(8 * n)
    ctx.h2d(du, u)
    ctx.memset(dw, 0xFF, 8 * n)
    N.call('svmb200_symv', ctx.handle, C.c_void_p(dQ), n, ld, C.c_void_p(du), C.c_void_p(dw))
    w = np.empty(n)
    ctx.d2h(w, dw)
    ctx.free(du)
    ctx.free(dw)
    return w


def check_symmetric_product(device, n, data_seed=0, poison=True, **dev_kw):
    """svmb200_symv against NumPy on a symmetric matrix whose LOWER triangle (below the diagonal blocks' reach) holds
    NaN: a single read of it would poison the result.  Returns the product (for bitwise comparisons between runs)."""
    from optiml_b200.runtime import default_context
    rng = np.random.default_rng(data_seed + n)
    A = rng.standard_normal((n, n))
    Qs = (A + A.T) / 2
    u = np.zeros(N.padded_ld(n))
    u[:n] = rng.standard_normal(n)
    want = Qs @ u[:n]
    with device(**dev_kw) as probe:
        ctx = default_context()
        ld = N.padded_ld(n)
        Q = np.zeros((n, ld))
        Q[:, :n] = Qs
        Q[:, n:] = 3.0   # padding columns meet u = 0: any finite value must do
        dq = _upload(ctx, Q)
        w_clean = _symv(ctx, dq, n, ld, u)
        ctx.free(dq, Q.nbytes)
        assert np.abs(w_clean - want).max() <= 1e-13 * n * max(1.0, np.abs(want).max())
        if poison:
            # everything strictly below the band structure's diagonal blocks: for ANY band height the kernel may read
            # (r, c) with c < r only inside its diagonal block, i.e. c >= r - (BH - 1); poison c < r - 512 (BH <= 512)
            Qp = Q.copy()
            r, c = np.tril_indices(n, -513)
            Qp[r, c] = np.nan
            if r.size:
                dq = _upload(ctx, Qp)
                w_p = _symv(ctx, dq, n, ld, u)
                ctx.free(dq, Qp.nbytes)
                assert np.array_equal(w_p, w_clean)
        probe.assert_clean()
    return w_clean


def _upload(ctx, M):
    d = ctx.malloc(M.nbytes)
    ctx.h2d(d, M)
    return d


def check_lower_triangle_is_never_read(device, n, bh, **dev_kw):
    """with the band height known (``bh``), everything below the diagonal BLOCKS is NaN"""
    from optiml_b200.runtime import default_context
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n))
    Qs = (A + A.T) / 2
    ld = N.padded_ld(n)
    u = np.zeros(ld)
    u[:n] = rng.standard_normal(n)
    Q = np.zeros((n, ld))
    Q[:, :n] = Qs
    rows, cols = np.indices((n, n))
    below = cols < (rows // bh) * bh
    Q[:, :n][below] = np.nan
    with device(**dev_kw) as probe:
        ctx = default_context()
        dq = _upload(ctx, Q)
        w = _symv(ctx, dq, n, ld, u)
        ctx.free(dq, Q.nbytes)
        probe.assert_clean()
    want = Qs @ u[:n]
    assert np.all(np.isfinite(w))
    assert np.abs(w - want).max() <= 1e-13 * n * max(1.0, np.abs(want).max())
    return w


def check_symmetric_solves(device, n=150, max_iter=30, **dev_kw):
    """PG (plain, label-sign view and SVR block layout), Frank-Wolfe and an augmented-Lagrangian rule with the symmetric
    pass against the default pass and the oracle: same iterate to rounding, `symmetric_pass` reported, and the number
    of launches per iteration one higher (tile pass + combine instead of one K2)."""
    from optiml_b200.opti import Quadratic
    from optiml_b200.opti.constrained import FrankWolfe, ProjectedGradient
    from optiml_b200.runtime import DeviceHessian, default_context
    rng = np.random.default_rng(n)
    Q, q, ub = psd(rng, n), rng.standard_normal(n), rng.uniform(0.5, 2.0, n)
    M = psd(rng, n)
    y = rng.standard_normal(n)
    q2 = np.hstack((-y, y)) + 0.1
    res = {}
    for sym in (False, True):
        with device(**dev_kw) as probe, symmetric_pass(sym):
            pg = ProjectedGradient(quad=Quadratic(Q, q), ub=ub, max_iter=max_iter)
            pg.profile = True   # keeps a small problem on the two-kernel loop (no persistent kernel)
            pg.minimize()
            assert pg.symmetric_pass is sym
            fw = FrankWolfe(quad=Quadratic(Q, q), ub=ub, max_iter=max_iter, t=0.2).minimize()
            assert fw.symmetric_pass is sym
            ctx = default_context()
            H = DeviceHessian(ctx, n, 'svr')
            block = np.zeros((n, H.ld))
            block[:, :n] = M
            ctx.h2d(H.matrix.dptr, block)
            svr = ProjectedGradient(quad=Quadratic(H, q2), ub=np.ones(2 * n), max_iter=max_iter).minimize()
            assert svr.symmetric_pass is sym
            res[sym] = (pg.x.copy(), np.asarray(pg.f_hist).copy(), fw.x.copy(), svr.x.copy())
            for s in (pg, fw, svr):
                s.f.release()   # device buffers go back while their context is alive (the emulated one ends with the block)
            del pg, fw, svr, H
            probe.assert_clean()
    for a, b in zip(res[False], res[True]):
        assert np.abs(a - b).max() <= 1e-10 * max(1.0, np.abs(a).max())
    want = O.projected_gradient(Q, q, ub, max_iter=max_iter)
    assert np.abs(res[True][0] - want.x).max() <= 1e-9
    Qfull = np.vstack((np.hstack((M, -M)), np.hstack((-M, M))))
    want = O.projected_gradient(Qfull, q2, np.ones(2 * n), max_iter=max_iter)
    assert np.abs(res[True][3] - want.x).max() <= 1e-9


def check_symmetric_fit(device, n=120, **dev_kw):
    """the estimator path: DualSVC / DualSVR with the symmetric pass == the default pass to rounding, same support set"""
    from optiml_b200.ml.svm import DualSVC, DualSVR
    from optiml_b200.ml.svm.kernels import GaussianKernel, PolyKernel
    from optiml_b200.opti.constrained import FrankWolfe
    from sklearn.datasets import make_classification, make_regression
    X, y = make_classification(n_samples=n, n_features=6, random_state=1)
    Xr, yr = make_regression(n_samples=n, n_features=5, noise=0.1, random_state=2)
    yr = (yr - yr.mean()) / yr.std()
    out = {}
    for sym in (False, True):
        with device(**dev_kw) as probe, symmetric_pass(sym):
            svc = DualSVC(kernel=GaussianKernel(), optimizer=FrankWolfe, max_iter=40).fit(X, y)
            svr = DualSVR(kernel=PolyKernel(degree=2), optimizer=FrankWolfe, max_iter=40, epsilon=0.1).fit(Xr, yr)
            assert svc.optimizer.symmetric_pass is sym and svr.optimizer.symmetric_pass is sym
            out[sym] = (svc.alphas_.copy(), svc.support_.copy(), svc.decision_function(X[:20]).copy(),
                        svr.alphas_.copy(), svr.predict(Xr[:20]).copy())
            svc.obj.release()
            svr.obj.release()
            del svc, svr
            probe.assert_clean()
    a, b = out[False], out[True]
    assert np.abs(a[0] - b[0]).max() <= 1e-10
    assert np.array_equal(a[1], b[1])
    assert np.abs(a[2] - b[2]).max() <= 1e-9
    assert np.abs(a[3] - b[3]).max() <= 1e-9
    assert np.abs(a[4] - b[4]).max() <= 1e-8
