"""K1 parity: Gram matrices from the CUDA kernel vs golden vectors of the reference's kernels.py and
vs the CPU oracle on ragged shapes.  Bar (north_star): entries within 1e-12 relative."""
import numpy as np
import pytest

from oracle import svm_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-12


def assert_gram_close(got, want):
    scale = np.abs(want).max()
    err = np.abs(got - want)
    bound = RTOL * np.abs(want) + 1e-15 * scale  # pure relative + a floor of ~4 ulp of the largest entry
    assert got.shape == want.shape
    assert (err <= bound).all(), f'max rel err {np.max(err / np.maximum(np.abs(want), 1e-300)):.3e}'


def kernels():
    from optiml_b200.ml.svm.kernels import LinearKernel, PolyKernel, GaussianKernel
    return {
        'linear': (LinearKernel(), dict(kind='linear')),
        'poly_d3_scale': (PolyKernel(), dict(kind='poly', degree=3, gamma='scale', coef0=0.)),
        'poly_d2_auto_c1': (PolyKernel(degree=2, gamma='auto', coef0=1.), dict(kind='poly', degree=2, gamma='auto', coef0=1.)),
        'poly_d4_g05_c05': (PolyKernel(degree=4, gamma=0.5, coef0=0.5), dict(kind='poly', degree=4, gamma=0.5, coef0=0.5)),
        'gauss_scale': (GaussianKernel(), dict(kind='gaussian', gamma='scale')),
        'gauss_auto': (GaussianKernel(gamma='auto'), dict(kind='gaussian', gamma='auto')),
        'gauss_g03': (GaussianKernel(gamma=0.3), dict(kind='gaussian', gamma=0.3)),
    }


@pytest.mark.parametrize('name', ['linear', 'poly_d3_scale', 'poly_d2_auto_c1', 'poly_d4_g05_c05', 'gauss_scale',
                                  'gauss_auto', 'gauss_g03'])
def test_kernel_golden(golden, name):
    g = golden('kernels')
    k, _ = kernels()[name]
    assert_gram_close(k(g['X']), g[name + '_XX'])
    assert_gram_close(k(g['X'], g['Y']), g[name + '_XY'])
    if name.startswith('gauss'):
        assert np.all(np.diag(k(g['X'])) == 1.0)  # sklearn forces the self-distance to exactly 0


@pytest.mark.parametrize('shape', [(1, 1, 1), (2, 3, 1), (129, 127, 17), (257, 1, 33), (130, 300, 784), (1000, 1000, 20),
                                   (383, 129, 16), (128, 128, 15)])
@pytest.mark.parametrize('name', ['linear', 'poly_d3_scale', 'gauss_scale'])
def test_kernel_ragged_shapes_vs_oracle(shape, name):
    nx, ny, d = shape
    rng = np.random.default_rng(nx * 1000 + ny + d)
    X = rng.standard_normal((nx, d)) + 0.5
    Y = rng.standard_normal((ny, d)) - 0.25
    k, kw = kernels()[name]
    kw = dict(kw)
    kind = kw.pop('kind')
    if nx * d > 1:  # 'scale' needs a non-zero variance
        assert_gram_close(k(X, Y), O.kernel_matrix(kind, X, Y, **kw))
    if nx > 1:
        assert_gram_close(k(X), O.kernel_matrix(kind, X, None, **kw))


def test_kernel_c1_full_golden(golden):
    from optiml_b200.configs import make_config
    from optiml_b200.ml.svm.kernels import GaussianKernel
    g = golden('c1_svc_gaussian')
    spec, X, y = make_config('C1')
    K = GaussianKernel()(X)
    assert_gram_close(K[::97, ::89], g['K_sub'])
    assert_gram_close(K[0], g['K_row0'])
    # (-2<a,b> + |a|^2) + |b|^2 is rounded in that order (as in sklearn), so K is symmetric only to 1 ulp
    assert np.abs(K - K.T).max() <= 4e-16


def test_kernel_input_validation():
    from optiml_b200.ml.svm.kernels import GaussianKernel, PolyKernel
    with pytest.raises(ValueError):
        GaussianKernel()(np.ones((4, 3)), np.ones((4, 2)))  # feature-count mismatch (sklearn check_pairwise_arrays)
    with pytest.raises(ValueError):
        PolyKernel()(np.array([[np.nan, 1.], [0., 1.]]))
    # ints are promoted to float64, float32 is computed in FP64
    Ki = PolyKernel(gamma=1.)(np.arange(12).reshape(4, 3))
    assert Ki.dtype == np.float64 and Ki[1, 2] == float(np.dot([3, 4, 5], [6, 7, 8])) ** 3


# ---- widening: Laplacian (CUDA-core pairwise kernel) and Sigmoid (tanh epilogue)
def extra_kernels():
    from optiml_b200.ml.svm.kernels import LaplacianKernel, SigmoidKernel
    return {'lap_scale': (LaplacianKernel(), dict(kind='laplacian')),
            'lap_g03': (LaplacianKernel(gamma=0.3), dict(kind='laplacian', gamma=0.3)),
            'sig_scale': (SigmoidKernel(), dict(kind='sigmoid')),
            'sig_g01_c05': (SigmoidKernel(gamma=0.01, coef0=0.5), dict(kind='sigmoid', gamma=0.01, coef0=0.5))}


@pytest.mark.parametrize('name', ['lap_scale', 'lap_g03', 'sig_scale', 'sig_g01_c05'])
def test_extra_kernel_golden(golden, name):
    g, ex = golden('kernels'), golden('kernels_extra')
    k, _ = extra_kernels()[name]
    assert_gram_close(k(g['X']), ex[name + '_XX'])
    assert_gram_close(k(g['X'], g['Y']), ex[name + '_XY'])


@pytest.mark.parametrize('shape', [(2, 3, 1), (129, 127, 17), (257, 1, 33), (130, 300, 100), (383, 129, 16)])
@pytest.mark.parametrize('name', ['lap_scale', 'sig_scale'])
def test_extra_kernel_ragged_shapes_vs_oracle(shape, name):
    nx, ny, d = shape
    rng = np.random.default_rng(nx * 1000 + ny + d)
    X = rng.standard_normal((nx, d)) + 0.5
    Y = rng.standard_normal((ny, d)) - 0.25
    k, kw = extra_kernels()[name]
    kw = dict(kw)
    kind = kw.pop('kind')
    assert_gram_close(k(X, Y), O.kernel_matrix(kind, X, Y, **kw))
    assert_gram_close(k(X), O.kernel_matrix(kind, X, None, **kw))


@pytest.mark.parametrize('name,kern', [('lap', 'LaplacianKernel'), ('sig', 'SigmoidKernel')])
def test_extra_kernel_svc_fit(golden, name, kern):
    from optiml_b200.ml.svm import SVC
    from optiml_b200.ml.svm import kernels as K
    from optiml_b200.ml.svm.losses import hinge
    from optiml_b200.opti.constrained import FrankWolfe
    iris, ex = golden('iris_ovr'), golden('kernels_extra')
    kernel = K.LaplacianKernel() if name == 'lap' else K.SigmoidKernel(gamma=0.05, coef0=0.)
    m = SVC(loss=hinge, kernel=kernel, reg_intercept=True, dual=True, optimizer=FrankWolfe, max_iter=300).fit(
        iris['X_train'], (iris['y_train'] == 0).astype(int))
    assert np.abs(m.alphas_ - ex[f'iris_{name}_alphas']).max() <= 1e-8
    assert np.array_equal(m.support_, ex[f'iris_{name}_support'])
    assert abs(m.intercept_ - float(ex[f'iris_{name}_intercept'])) <= 1e-8
    assert np.abs(m.decision_function(iris['X_test']) - ex[f'iris_{name}_decision']).max() <= 1e-8


def test_device_variance_is_bit_identical_to_numpy():
    """svmb200_device_variance (gamma='scale', kernels.py:93, 127) incl. every BASELINE shape"""
    import device_path_checks as D
    from optiml_b200.configs import make_config
    from optiml_b200.runtime import default_context
    D.check_device_variance([(1, 1), (3, 2), (129, 1), (300, 7), (2049, 5), (70001, 3), (20000, 131), (123457, 9)])
    for cfg, n in (('C1', None), ('C2', None), ('C3', 6000), ('C4', None), ('C5', 30000)):
        spec, X, y = make_config(cfg, n=n)
        assert D.device_var(default_context(), X) == (X.var(), 0)


def test_fit_keeps_the_host_passes_off_the_path():
    import device_path_checks as D
    D.check_fit_keeps_host_passes_off_the_path(n=3000, d=20)
