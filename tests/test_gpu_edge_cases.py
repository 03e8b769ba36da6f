"""Edge cases of the path, through the public API and directly through the C ABI: tiny and degenerate problems,
error codes, input conversion, the row-block chunking of the host-pointer Gram entry point."""
import ctypes as C

import numpy as np
import pytest

from oracle import svm_oracle as O

pytestmark = pytest.mark.gpu


def test_tiny_and_degenerate_fits():
    from optiml_b200.ml.svm import DualSVC, DualSVR
    from optiml_b200.ml.svm.kernels import GaussianKernel, LinearKernel
    X2 = np.array([[0., 1.], [1., 0.]])
    m = DualSVC(kernel=GaussianKernel(gamma=1.), C=1., max_iter=50).fit(X2, [0, 1])
    ref = O.svc_dual_fit(X2, [0, 1], kind='gaussian', gamma=1., C=1., max_iter=50)
    assert m.optimizer.iter == ref.pg.iter and m.optimizer.status == ref.pg.status
    assert np.abs(m.alphas_ - ref.alphas_).max() <= 1e-12 and abs(m.intercept_ - ref.intercept_) <= 1e-12
    assert np.array_equal(m.predict(X2), [0, 1])
    assert m.decision_function(X2[:1]).shape == (1,)  # a single test row
    with pytest.raises(ValueError):
        DualSVC().fit(np.array([[1., 2.]]), [1])      # one sample: 'Q is too small' (opti/_base.py:249-250)
    with pytest.raises(ValueError):
        DualSVC().fit(np.zeros((0, 3)), [])           # no samples (sklearn validation, as in the reference)
    # max_iter = 1 and a regression target that is constant
    rng = np.random.default_rng(0)
    X = rng.standard_normal((40, 3))
    m = DualSVR(kernel=LinearKernel(), epsilon=0.1, C=1., max_iter=1).fit(X, np.full(40, 2.5))
    ref = O.svr_dual_fit(X, np.full(40, 2.5), kind='linear', epsilon=0.1, C=1., max_iter=1)
    assert m.optimizer.iter == 1 and np.abs(m.alphas_ - ref.alphas_).max() <= 1e-12
    assert np.abs(m.predict(X) - O.decision_function(ref, X)).max() <= 1e-10


def test_input_conversion_float32_int_sparse():
    import scipy.sparse as sp
    from optiml_b200.ml.svm.kernels import GaussianKernel, PolyKernel
    rng = np.random.default_rng(1)
    X = rng.standard_normal((37, 5))
    K = GaussianKernel(gamma=0.3)(X)
    assert np.array_equal(GaussianKernel(gamma=0.3)(sp.csr_matrix(X)), K)          # CSR input is densified
    K32 = GaussianKernel(gamma=0.3)(X.astype(np.float32))                           # float32 is computed in FP64
    assert K32.dtype == np.float64 and np.abs(K32 - K).max() < 1e-6
    Xi = rng.integers(-3, 4, size=(20, 4))
    # integer inputs are promoted to FP64; the contraction is exact, the device pow() is within 2 ulp of NumPy's
    assert np.allclose(PolyKernel(degree=2, gamma=1., coef0=1.)(Xi), (Xi @ Xi.T + 1.) ** 2, rtol=1e-14, atol=0)


def test_kernel_matrix_host_row_block_chunking():
    """nx * ld * 8 > 1 GiB: svmb200_kernel_matrix_host walks the rows in blocks"""
    from optiml_b200.ml.svm.kernels import LinearKernel
    rng = np.random.default_rng(2)
    X = rng.standard_normal((12000, 8))
    K = LinearKernel()(X)
    idx = rng.integers(0, 12000, size=(2000, 2))
    want = np.einsum('ij,ij->i', X[idx[:, 0]], X[idx[:, 1]])
    assert np.abs(K[idx[:, 0], idx[:, 1]] - want).max() <= 1e-12 * np.abs(want).max() + 1e-14
    assert np.abs(K[-1] - X @ X[-1]).max() <= 1e-13


def test_c_abi_error_codes_and_messages():
    from optiml_b200 import _native as N
    from optiml_b200.runtime import default_context
    ctx = default_context()
    lib = N.load_library()
    X = np.ascontiguousarray(np.random.default_rng(3).standard_normal((8, 4)))
    out = np.empty((8, 8))
    rc = lib.svmb200_kernel_matrix_host(ctx.handle, N.ptr(X), 8, None, 8, 4, 99, 1., 0., 1., N.ptr(out))
    assert rc == 1 and b'unknown kernel id' in lib.svmb200_last_error()             # SVMB200_ERR_ARG
    rc = lib.svmb200_kernel_matrix_host(ctx.handle, N.ptr(X), 8, None, 8, 4, N.KERNEL_GAUSSIAN, -1., 0., 1., N.ptr(out))
    assert rc == 1 and b'gamma' in lib.svmb200_last_error()
    with pytest.raises(N.NativeError, match='ld must be'):
        N.call('svmb200_matvec', ctx.handle, C.c_void_p(256), 4, 3, C.c_void_p(256), C.c_void_p(256))
    h = C.c_void_p()
    q = np.zeros(4)
    rc = lib.svmb200_pg_create(ctx.handle, C.c_void_p(256), 4, 16, 0, 4, 7, N.ptr(q), None, N.ptr(q), None, 1e-6, 10, C.byref(h))
    assert rc == 1 and b'hessian' in lib.svmb200_last_error()
    rc = lib.svmb200_fw_create(ctx.handle, C.c_void_p(256), 4, 16, 0, 4, 0, N.ptr(q), None, N.ptr(q), None, 1e-6, 10, 1.5,
                               C.byref(h))
    assert rc == 1 and b't has to lie' in lib.svmb200_last_error()
    r0, nr = C.c_int64(0), C.c_int64(0)
    assert lib.svmb200_shard_rows(50000, 7, 8, C.byref(r0), C.byref(nr)) == 0 and (r0.value, nr.value) == (43904, 6096)


def test_bcqp_host_entry_point_matches_oracle():
    """svmb200_bcqp_pg_host: the one-call form a reference-side binding would use"""
    from optiml_b200 import _native as N
    from optiml_b200.runtime import default_context
    rng = np.random.default_rng(4)
    n = 257
    G = rng.standard_normal((n + 9, n))
    Q = np.ascontiguousarray(G.T @ G / n)
    q, ub, lb = rng.standard_normal(n), rng.uniform(0.5, 2., n), -rng.uniform(0., 1., n)
    x, g = np.empty(n), np.empty(n)
    fh, ngh = np.empty(31), np.empty(31)
    it, st = C.c_int64(0), C.c_int(0)
    N.call('svmb200_bcqp_pg_host', default_context().handle, N.ptr(Q), N.ptr(q), N.ptr(lb), N.ptr(ub), None, n, 1e-6, 30,
           N.ptr(x), N.ptr(g), N.ptr(fh), N.ptr(ngh), C.byref(it), C.byref(st))
    want = O.projected_gradient(Q, q, ub, lb=lb, max_iter=30)
    assert it.value == want.iter and N.STATUS[st.value] == want.status
    assert np.abs(x - want.x).max() <= 1e-11 and np.abs(g - want.g_x).max() <= 1e-10
    assert np.allclose(fh[:it.value + 1], want.f_hist, rtol=1e-11, atol=1e-12)
    assert np.allclose(ngh[:it.value + 1], want.ng_hist, rtol=1e-10, atol=1e-12)
