"""Kernel functors (reference: optiml/ml/svm/kernels.py).  ``kernel(X, Y=None)`` returns the Gram
matrix as a host ndarray, computed by the CUDA Gram kernel (csrc/gram.cu); the estimators use
``gram_spec`` instead and keep the training Gram matrix in HBM."""
import ctypes as C

import numpy as np
from sklearn.base import BaseEstimator
from sklearn.metrics.pairwise import check_pairwise_arrays

from ... import _native as N
from ...runtime import default_context


def _check_gamma(gamma):
    if isinstance(gamma, str):
        if gamma not in ('scale', 'auto'):
            raise ValueError(f'unknown gamma type {gamma}')
    elif not gamma > 0:
        raise ValueError('gamma must be > 0')


def _dense_f64(X, Y, finite=True):
    """sklearn's pairwise validation (2-D, finite, matching feature counts), then dense float64:
    sparse / float32 inputs are converted (the reference keeps float32 Grams in float32).
    ``finite=False``: the caller runs the all-finite test on the device copy (``device_variance``) -- sklearn's is a full
    single-threaded pass over X on the host (6.6 ms at n = 50 000, d = 128, on every rank of a sharded fit)."""
    same = Y is None or Y is X
    X, Y = check_pairwise_arrays(X, None if same else Y, accept_sparse='csr', ensure_all_finite=bool(finite))
    dense = [np.ascontiguousarray(A.toarray() if hasattr(A, 'toarray') else A, dtype=np.float64) for A in (X, Y)]
    return dense[0], (None if same else dense[1])


def host_threads():
    """threads for the host-side helpers (variance, support-vector gather): a few per rank, never more than the
    cores this rank can claim; SVMB200_HOST_THREADS overrides"""
    import os
    env = os.environ.get('SVMB200_HOST_THREADS')
    if env:
        return max(1, int(env))
    ranks = max(1, int(os.environ.get('LOCAL_WORLD_SIZE', '1')))
    return max(1, min(4, (os.cpu_count() or 1) // ranks))


def variance(X):
    """``X.var()`` bit for bit (svmb200_host_variance walks NumPy's pairwise-summation tree, fused and threaded):
    40 ms -> 4 ms at n = 50 000, d = 128, paid by every rank of a sharded fit before its Gram build can start"""
    if isinstance(X, np.ndarray) and X.dtype == np.float64 and X.flags['C_CONTIGUOUS'] and X.size >= 1 << 16:
        v = C.c_double(0)
        N.call('svmb200_host_variance', N.ptr(X), X.size, host_threads(), C.byref(v))
        return v.value
    return X.var()


def device_variance(dX, X_host=None, want_variance=True):
    """``X.var()`` bit for bit from the copy of X in HBM (svmb200_device_variance: NumPy's pairwise-summation tree, leaves on
    the device) and, in the same pass, the all-finite test of sklearn's input validation -- a non-finite element raises
    sklearn's own ValueError (from ``assert_all_finite`` on the host copy when there is one)."""
    v, bad = C.c_double(0), C.c_int(0)
    N.call('svmb200_device_variance', dX.ctx.handle, C.c_void_p(dX.dptr), dX.rows, dX.cols, dX.ld, int(bool(want_variance)),
           C.byref(v), C.byref(bad))
    if bad.value:
        if X_host is not None:
            from sklearn.utils import assert_all_finite
            assert_all_finite(X_host)
        raise ValueError('Input contains NaN or infinity.')
    return v.value


def gather_rows(X, idx):
    """``X[idx]`` for a C-contiguous float64 matrix and int64 row indices (threaded copy into fresh memory)"""
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    if isinstance(X, np.ndarray) and X.ndim == 2 and X.dtype == np.float64 and X.flags['C_CONTIGUOUS'] \
            and idx.size * X.shape[1] >= 1 << 16:
        out = np.empty((idx.size, X.shape[1]))
        N.call('svmb200_host_gather_rows', N.ptr(X), X.shape[1], idx.ctypes.data_as(C.c_void_p), idx.size, N.ptr(out),
               host_threads())
        return out
    return X[idx]


class Kernel(BaseEstimator):
    """Base class: a kernel maps two sample sets to their Gram matrix (kernels.py:9-37)."""

    kernel_id = None

    def resolve_gamma(self, X, device=None):
        """'scale' -> 1/(d * X.var()), 'auto' -> 1/d, evaluated on the FIRST argument of the call
        exactly as kernels.py:93-94 / 127-128 do (bit-identical to NumPy's X.var()).  ``device``: the DeviceMatrix that
        holds X in HBM -- the variance and the all-finite test of the input then run there (``X`` may be None)."""
        gamma = getattr(self, 'gamma', None)
        d = device.cols if device is not None else X.shape[1]
        if device is not None:
            var = device_variance(device, X, want_variance=(gamma == 'scale'))
        elif gamma == 'scale':
            var = variance(X)
        if gamma is None:
            return 1.
        return (1. / (d * var) if gamma == 'scale' else 1. / d if gamma == 'auto' else gamma)

    def gram_spec(self, X, device=None):
        """(kernel_id, gamma, coef0, degree) for a call whose first argument is X."""
        return (self.kernel_id, float(self.resolve_gamma(X, device)), float(getattr(self, 'coef0', 0.)),
                float(getattr(self, 'degree', 1)))

    def __call__(self, X, Y=None):
        if self.kernel_id is None:
            raise NotImplementedError
        X, Y = _dense_f64(X, Y)
        kid, gamma, coef0, degree = self.gram_spec(X)
        ny = X.shape[0] if Y is None else Y.shape[0]
        out = np.empty((X.shape[0], ny))
        N.call('svmb200_kernel_matrix_host', default_context().handle, N.ptr(X), X.shape[0], N.ptr(Y), ny, X.shape[1],
               kid, gamma, coef0, degree, N.ptr(out))
        return out


class LinearKernel(Kernel):
    """K(X, Y) = <X, Y>  (kernels.py:40-51)."""
    kernel_id = N.KERNEL_LINEAR


class PolyKernel(Kernel):
    """K(X, Y) = (gamma <X, Y> + coef0) ** degree  (kernels.py:54-95)."""
    kernel_id = N.KERNEL_POLY

    def __init__(self, degree=3, gamma='scale', coef0=0.):
        if not degree > 0:
            raise ValueError('degree must be > 0')
        self.degree = degree
        _check_gamma(gamma)
        self.gamma = gamma
        self.coef0 = coef0


class GaussianKernel(Kernel):
    """K(X, Y) = exp(-gamma ||X - Y||^2)  (kernels.py:98-129)."""
    kernel_id = N.KERNEL_GAUSSIAN

    def __init__(self, gamma='scale'):
        _check_gamma(gamma)
        self.gamma = gamma


class LaplacianKernel(Kernel):
    """K(X, Y) = exp(-gamma ||X - Y||_1)  (kernels.py:132-163; widening)."""
    kernel_id = N.KERNEL_LAPLACIAN

    def __init__(self, gamma='scale'):
        _check_gamma(gamma)
        self.gamma = gamma


class SigmoidKernel(Kernel):
    """K(X, Y) = tanh(gamma <X, Y> + coef0)  (kernels.py:166-201; widening)."""
    kernel_id = N.KERNEL_SIGMOID

    def __init__(self, gamma='scale', coef0=0.):
        _check_gamma(gamma)
        self.gamma = gamma
        self.coef0 = coef0


linear = LinearKernel()
poly = PolyKernel()
gaussian = GaussianKernel()
laplacian = LaplacianKernel()
sigmoid = SigmoidKernel()
