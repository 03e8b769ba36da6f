"""When /root/reference is present (the build container; never the GPU box) run the REAL reference side by side
with the oracle on freshly generated problems -- more coverage than the committed golden vectors.  CPU only."""
import numpy as np
import pytest

from oracle import ref_shim
from oracle import svm_oracle as O

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason='reference checkout not present')


@pytest.fixture(scope='module')
def ref():
    return ref_shim.load_reference()


@pytest.mark.parametrize('ndim,seed', [(3, 1), (17, 2), (40, 3), (90, 4)])
def test_pg_and_fw_bit_exact_on_generated_bcqp(ref, ndim, seed):
    Q, q, ub = ref.generate_box_constrained_quadratic(ndim=ndim, seed=seed)
    lb = ub / (3 + seed)
    for lbv in (None, lb):
        r = ref.ProjectedGradient(quad=ref.Quadratic(Q, q), ub=ub, lb=lbv, max_iter=300).minimize()
        o = O.projected_gradient(Q, q, ub, lb=lbv, max_iter=300)
        assert (r.iter, r.status) == (o.iter, o.status) and np.array_equal(r.x, o.x) and r.f_x == o.f_x
        for t in (0., 0.3):
            r = ref.FrankWolfe(quad=ref.Quadratic(Q, q), ub=ub, lb=lbv, t=t, max_iter=200).minimize()
            o = O.frank_wolfe(Q, q, ub, lb=lbv, t=t, max_iter=200)
            assert (r.iter, r.status) == (o.iter, o.status) and np.array_equal(r.x, o.x) and r.f_x == o.f_x


def test_kernels_bit_exact_on_random_shapes(ref):
    rng = np.random.default_rng(11)
    for nx, ny, d in ((1, 1, 3), (5, 9, 1), (33, 20, 12), (64, 64, 40)):
        X, Y = rng.standard_normal((nx, d)) * 2 + 1, rng.standard_normal((ny, d))
        pairs = [(ref.LinearKernel(), dict(kind='linear')),
                 (ref.PolyKernel(degree=2, gamma=0.7, coef0=1.5), dict(kind='poly', degree=2, gamma=0.7, coef0=1.5)),
                 (ref.GaussianKernel(gamma=0.4), dict(kind='gaussian', gamma=0.4)),
                 (ref.LaplacianKernel(gamma=0.4), dict(kind='laplacian', gamma=0.4)),
                 (ref.SigmoidKernel(gamma=0.1, coef0=0.2), dict(kind='sigmoid', gamma=0.1, coef0=0.2))]
        if nx > 1:
            pairs += [(ref.GaussianKernel(), dict(kind='gaussian', gamma='scale')), (ref.PolyKernel(), dict(kind='poly'))]
        for kern, kw in pairs:
            kw = dict(kw)
            kind = kw.pop('kind')
            assert np.array_equal(kern(X, Y), O.kernel_matrix(kind, X, Y, **kw))
            assert np.array_equal(kern(X), O.kernel_matrix(kind, X, None, **kw))


def test_estimators_bit_exact_on_small_problems(ref):
    from sklearn.datasets import make_classification, make_regression
    X, y = make_classification(n_samples=150, n_features=6, random_state=3)
    for kern, kind in ((ref.GaussianKernel(), 'gaussian'), (ref.LinearKernel(), 'linear'), (ref.PolyKernel(degree=2), 'poly')):
        m = ref.SVC(loss=ref.hinge, kernel=kern, C=0.7, reg_intercept=True, dual=True, optimizer=ref.ProjectedGradient,
                    max_iter=120).fit(X, y)
        o = O.svc_dual_fit(X, y, kind=kind, C=0.7, degree=2, max_iter=120)
        assert np.array_equal(m.alphas_, o.alphas_) and np.array_equal(m.support_, o.support_)
        assert m.intercept_ == o.intercept_ and np.array_equal(m.decision_function(X[:20]), O.decision_function(o, X[:20]))
        assert np.array_equal(m.predict(X[:20]), O.svc_predict(o, X[:20]))
    X, y = make_regression(n_samples=120, n_features=5, noise=0.2, random_state=1)
    y = (y - y.mean()) / y.std()
    m = ref.SVR(loss=ref.epsilon_insensitive, epsilon=0.05, kernel=ref.GaussianKernel(), C=2., reg_intercept=True, dual=True,
                optimizer=ref.ProjectedGradient, max_iter=150).fit(X, y)
    o = O.svr_dual_fit(X, y, kind='gaussian', C=2., epsilon=0.05, max_iter=150)
    assert np.array_equal(m.alphas_, o.alphas_) and m.intercept_ == o.intercept_
    assert np.array_equal(m.predict(X[:20]), O.decision_function(o, X[:20]))
