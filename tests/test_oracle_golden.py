"""Pin the CPU oracle (oracle/svm_oracle.py) against golden vectors produced by the REAL reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import svm_oracle as O
from optiml_b200.configs import make_config

KERNEL_CASES = {
    'linear': dict(kind='linear'),
    'poly_d3_scale': dict(kind='poly', degree=3, gamma='scale', coef0=0.),
    'poly_d2_auto_c1': dict(kind='poly', degree=2, gamma='auto', coef0=1.),
    'poly_d4_g05_c05': dict(kind='poly', degree=4, gamma=0.5, coef0=0.5),
    'gauss_scale': dict(kind='gaussian', gamma='scale'),
    'gauss_auto': dict(kind='gaussian', gamma='auto'),
    'gauss_g03': dict(kind='gaussian', gamma=0.3),
}


@pytest.mark.parametrize('name', sorted(KERNEL_CASES))
def test_kernels_bit_exact(golden, name):
    g = golden('kernels')
    kw = dict(KERNEL_CASES[name])
    kind = kw.pop('kind')
    assert np.array_equal(O.kernel_matrix(kind, g['X'], None, **kw), g[name + '_XX'])
    assert np.array_equal(O.kernel_matrix(kind, g['X'], g['Y'], **kw), g[name + '_XY'])


@pytest.mark.parametrize('p', ['p2', 'p5', 'p64', 'p200'])
def test_pg_three_pass_bit_exact(golden, p):
    g = golden('bcqp')
    lb = g[p + '_lb'] if p + '_lb' in g else None
    r = O.projected_gradient(g[p + '_Q'], g[p + '_q'], g[p + '_ub'], lb=lb, passes=3)
    assert r.iter == int(g[p + '_iter']) and r.status == str(g[p + '_status'])
    assert np.array_equal(r.x, g[p + '_x'])
    assert np.array_equal(r.f_hist, g[p + '_f_hist'])
    assert np.array_equal(r.g_x, g[p + '_g'])


def test_pg_known_answers(golden):
    # reference tests test_projected_gradient.py:9-12 / test_lower_bound.py:9-25
    g = golden('bcqp')
    assert int(g['p2_iter']) == 2 and str(g['p2_status']) == 'optimal' and np.array_equal(g['p2_x'], [0., 0.])
    assert int(g['p5_iter']) == 23 and str(g['p5_status']) == 'optimal'
    assert np.allclose(g['p5_x'], [2.0763082893739573, 2.7799187922401147, 9.753636925763574,
                                   8.950232049077629, 2.9779895119966024], rtol=0, atol=1e-12)


@pytest.mark.parametrize('p', ['p5', 'p64'])
def test_pg_one_pass_matches_three_pass(golden, p):
    g = golden('bcqp')
    lb = g[p + '_lb'] if p + '_lb' in g else None
    r = O.projected_gradient(g[p + '_Q'], g[p + '_q'], g[p + '_ub'], lb=lb, passes=1)
    assert r.iter == int(g[p + '_iter']) and r.status == str(g[p + '_status'])
    scale = max(1., np.abs(g[p + '_x']).max())
    assert np.abs(r.x - g[p + '_x']).max() <= 1e-9 * scale
    assert np.allclose(r.f_hist, g[p + '_f_hist'], rtol=1e-10, atol=1e-10)


def test_pg_one_pass_p200_same_optimum(golden):
    # ill-conditioned (ecc 0.99) and mostly free steps: the two forms reach the same optimum through
    # trajectories that separate by ~1e-7 (see test_oracle_sensitivity.py)
    g = golden('bcqp')
    r = O.projected_gradient(g['p200_Q'], g['p200_q'], g['p200_ub'], passes=1)
    assert r.status == 'optimal' and abs(r.iter - int(g['p200_iter'])) <= 5
    assert np.abs(r.x - g['p200_x']).max() <= 5e-6
    assert abs(r.f_x - g['p200_f_hist'][-1]) <= 1e-9 * abs(g['p200_f_hist'][-1])


def _check_fit(fit, g, prefix='', exact=True):
    tol = 0. if exact else 1e-10
    assert fit.pg.iter == int(g[prefix + 'iter']) and fit.pg.status == str(g[prefix + 'status'])
    assert np.abs(fit.alphas_ - g[prefix + 'alphas']).max() <= tol
    assert np.array_equal(fit.support_, g[prefix + 'support'])
    assert np.abs(fit.dual_coef_ - g[prefix + 'dual_coef']).max() <= tol
    assert abs(fit.intercept_ - float(g[prefix + 'intercept'])) <= tol
    assert np.abs(fit.pg.f_hist - g[prefix + 'f_hist']).max() <= tol * max(1., np.abs(g[prefix + 'f_hist']).max())


@pytest.mark.parametrize('c', [0, 1, 2])
def test_svc_iris_ovr(golden, c):
    g = golden('iris_ovr')
    fit = O.svc_dual_fit(g['X_train'], (g['y_train'] == c).astype(int), kind='gaussian')
    _check_fit(fit, g, f'c{c}_')
    assert np.array_equal(O.decision_function(fit, g['X_test']), g[f'c{c}_decision'])
    assert np.array_equal(O.svc_predict(fit, g['X_test']), g[f'c{c}_predict'])


def test_svc_iris_appendix_b(golden):
    g = golden('iris_ovr')
    got = [(int(g[f'c{c}_iter']), str(g[f'c{c}_status']), len(g[f'c{c}_support'])) for c in range(3)]
    assert got == [(1000, 'stopped', 14), (406, 'optimal', 35), (541, 'optimal', 34)]
    assert abs(float(g['c0_intercept']) - (-0.2403970197973344)) < 1e-9


@pytest.mark.parametrize('name,kind', [('linear', 'linear'), ('poly', 'poly'), ('gauss', 'gaussian')])
def test_svr_diabetes(golden, name, kind):
    g = golden('diabetes_svr')
    fit = O.svr_dual_fit(g['X_train'], g['y_train'], kind=kind, epsilon=0.1, C=1)
    _check_fit(fit, g, name + '_')
    assert np.array_equal(O.decision_function(fit, g['X_test']), g[name + '_decision'])
    if kind == 'linear':
        assert np.array_equal(fit.coef_, g['linear_coef'])


def test_c1_full_size(golden):
    g = golden('c1_svc_gaussian')
    spec, X, y = make_config('C1')
    assert np.array_equal([X.sum(), (X * X).sum()], g['X_checksum'])
    fit = O.svc_dual_fit(X, y, kind='gaussian', C=1)
    _check_fit(fit, g)
    assert float(g['f_x']) == -113.50663083792071 and len(g['support']) == 1015
    assert np.array_equal(fit.K[::97, ::89], g['K_sub']) and np.array_equal(fit.K[0], g['K_row0'])
    assert np.array_equal(O.decision_function(fit, X[:256]), g['decision'])
    # single-pass restatement tracks the same non-converged 1000-iteration trajectory
    fit1 = O.svc_dual_fit(X, y, kind='gaussian', C=1, passes=1)
    assert np.abs(fit1.alphas_ - g['alphas']).max() <= 1e-11
    assert np.array_equal(fit1.support_, g['support'])


def test_c2_c3_c4_small(golden):
    g = golden('c2small_svr_poly')
    spec, X, y = make_config('C2', n=600)
    _check_fit(O.svr_dual_fit(X, y, kind='poly', degree=3, epsilon=0.1, C=1), g)
    g = golden('c3small_svc_linear')
    spec, X, y = make_config('C3', n=500)
    fit = O.svc_dual_fit(X, y, kind='linear', C=1)
    _check_fit(fit, g)
    assert np.array_equal(fit.coef_, g['coef'])
    g = golden('c4small_svc_gaussian')
    spec, X, y = make_config('C4', n=1200)
    _check_fit(O.svc_dual_fit(X, y, kind='gaussian', C=2.5, max_iter=300), g)


@pytest.mark.parametrize('key,p,t', [('p2', 'p2', 0.), ('p5', 'p5', 0.), ('p64', 'p64', 0.), ('p200', 'p200', 0.),
                                     ('p64_t05', 'p64', 0.5)])
def test_frank_wolfe_three_pass_bit_exact(golden, key, p, t):
    """widening: the oracle's Frank-Wolfe vs the reference's FrankWolfe.minimize"""
    g, fw = golden('bcqp'), golden('frank_wolfe')
    lb = g[p + '_lb'] if p + '_lb' in g else None
    r = O.frank_wolfe(g[p + '_Q'], g[p + '_q'], g[p + '_ub'], lb=lb, t=t, passes=3)
    assert r.iter == int(fw[key + '_iter']) and r.status == str(fw[key + '_status'])
    assert np.array_equal(r.x, fw[key + '_x'])
    assert np.array_equal(r.f_hist, fw[key + '_f_hist'])
    assert np.array_equal(r.g_x, fw[key + '_g'])


@pytest.mark.parametrize('name,kw', [('lap_scale', dict(kind='laplacian')), ('lap_g03', dict(kind='laplacian', gamma=0.3)),
                                     ('sig_scale', dict(kind='sigmoid')), ('sig_g01_c05', dict(kind='sigmoid', gamma=0.01, coef0=0.5))])
def test_extra_kernels_bit_exact(golden, name, kw):
    """widening: Laplacian / Sigmoid kernels vs the reference"""
    g, ex = golden('kernels'), golden('kernels_extra')
    kw = dict(kw)
    kind = kw.pop('kind')
    assert np.array_equal(O.kernel_matrix(kind, g['X'], None, **kw), ex[name + '_XX'])
    assert np.array_equal(O.kernel_matrix(kind, g['X'], g['Y'], **kw), ex[name + '_XY'])
