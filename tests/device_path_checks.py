"""Checks of the device-side replacements of the O(n d) host steps of a fit (csrc/devmath.cu), shared by the CPU suite
(host emulation of the kernels, tests/test_emulated_kernels.py) and the GPU suite (tests/test_gpu_kernels.py)."""
import ctypes as C
import pickle

import numpy as np
import pytest

from optiml_b200 import _native as N


def device_var(ctx, X, want=True):
    dX = ctx.upload_matrix(X)
    v, bad = C.c_double(0), C.c_int(0)
    N.call('svmb200_device_variance', ctx.handle, C.c_void_p(dX.dptr), X.shape[0], X.shape[1], dX.ld, int(want), C.byref(v),
           C.byref(bad))
    dX.release()
    return v.value, bad.value


def check_device_variance(shapes, seed=0):
    """kernels.py:93, 127: gamma='scale' = 1 / (d * X.var()) -- the device pass returns NumPy's bits on every shape (odd d:
    padded leading dimension; tiny; one leaf; many leaves) and sees every non-finite element"""
    from optiml_b200.runtime import default_context
    ctx = default_context()
    rng = np.random.default_rng(seed)
    for n, d in shapes:
        X = rng.standard_normal((n, d)) * rng.uniform(0.1, 100) + rng.uniform(-50, 50)
        v, bad = device_var(ctx, X)
        assert v == X.var() and bad == 0, (n, d, v, X.var())
        host = C.c_double(0)
        N.call('svmb200_host_variance', N.ptr(X), X.size, 3, C.byref(host))
        assert host.value == v
    X = rng.standard_normal((301, 7))
    for val in (np.nan, np.inf, -np.inf):
        Xb = X.copy()
        Xb[int(rng.integers(301)), int(rng.integers(7))] = val
        assert device_var(ctx, Xb, want=False)[1] == 1
    assert device_var(ctx, X, want=False)[1] == 0


def check_fit_keeps_host_passes_off_the_path(n=400, d=5):
    """fit validates X and takes its variance on the device, leaves the support vectors in HBM; the host attribute, the
    pickled estimator and both decision paths agree with the reference semantics X[support_]"""
    from optiml_b200.ml.svm import DualSVC, DualSVR
    from optiml_b200.ml.svm.kernels import GaussianKernel, LinearKernel, PolyKernel
    rng = np.random.default_rng(4)
    X = rng.standard_normal((n, d))
    y = (X[:, 0] + 0.3 * rng.standard_normal(n) > 0).astype(int)
    m = DualSVC(kernel=GaussianKernel(), C=1, max_iter=40).fit(X, y)
    assert m._sv_host is None and m._sv_device is not None            # no host gather inside fit
    dec_dev = m.decision_function(X[:50])
    sv = m.support_vectors_                                            # materialised now, from the device snapshot
    assert np.array_equal(sv, X[m.support_]) and m._sv_host is sv
    twin = pickle.loads(pickle.dumps(m))
    assert twin._sv_device is None and np.array_equal(twin.support_vectors_, sv)
    assert np.array_equal(twin.decision_function(X[:50]), dec_dev)     # host-SV path: same bits (same gamma, same kernels)
    assert np.array_equal(twin.predict(X[:50]), m.predict(X[:50]))
    m.support_vectors_ = sv.copy()                                     # assigning drops the device copy
    assert m._sv_device is None and np.array_equal(m.decision_function(X[:50]), dec_dev)
    # the training matrix may change after fit: the device snapshot was taken at fit time (like the reference's copy)
    Xc = X.copy()
    m2 = DualSVC(kernel=GaussianKernel(), C=1, max_iter=40).fit(Xc, y)
    Xc[:] = 0.0
    assert np.array_equal(m2.support_vectors_, X[m2.support_])
    # SVR, and the linear kernel (coef_ needs the support vectors on the host)
    t = X @ rng.standard_normal(d)
    r = DualSVR(kernel=PolyKernel(degree=2), C=1, max_iter=30).fit(X, t)
    assert r._sv_host is None and np.array_equal(r.support_vectors_, X[r.support_])
    lin = DualSVC(kernel=LinearKernel(), C=1, max_iter=30).fit(X, y)
    assert lin._sv_device is None and np.array_equal(lin.support_vectors_, X[lin.support_])
    assert np.allclose(lin.coef_, lin.dual_coef_ @ X[lin.support_], rtol=0, atol=0)
    # non-finite input: sklearn's error, raised from the device-side test
    Xn = X.copy()
    Xn[7, 2] = np.nan
    for est in (DualSVC(kernel=GaussianKernel(gamma=0.5)), DualSVC(kernel=GaussianKernel()), DualSVC(kernel=LinearKernel())):
        with pytest.raises(ValueError, match='NaN'):
            est.fit(Xn, y)
    for e in (m, m2, r, lin):
        e.obj.release()
