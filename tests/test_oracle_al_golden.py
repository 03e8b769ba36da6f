"""Augmented-Lagrangian dual path on the CPU:

* the NumPy oracle (oracle/al_oracle.py) against golden vectors produced by the REAL reference
  (tests/golden/make_golden_al.py): every update rule, momentum variant, reg_intercept in {True, False}, SVC and SVR,
  and runs that end through the optimality test of the multiplier update;
* the host emulation of the CUDA vector phase (oracle/al_emulate.cpp, built from the arithmetic header the kernel
  includes) against the same goldens -- the per-variable arithmetic, the order of the phases and the double
  buffering by state parity are checked here without a GPU.
"""
import numpy as np
import pytest

from al_cases import CASES, SVC_RUNS, SVR_RUNS, TOL_RUNS, svc_key, svr_key
from oracle import al_oracle as AL, svm_oracle as O


def check_fit(g, key, fit, X_test, rtol=1e-9):
    al = fit.al
    assert al.iter == int(g[key + '_iter']) and al.status == str(g[key + '_status'])
    scale = max(1., np.abs(g[key + '_alphas']).max())
    assert np.abs(fit.alphas_ - g[key + '_alphas']).max() <= rtol * scale
    dscale = max(1., np.abs(g[key + '_dual_x']).max())
    assert np.abs(al.dual_x - g[key + '_dual_x']).max() <= rtol * dscale
    pf = g[key + '_pf_hist']
    assert len(al.pf_hist) == len(pf) == al.iter + 1
    assert np.abs(al.pf_hist - pf).max() <= rtol * max(1., np.abs(pf).max())
    assert abs(al.f_x - float(g[key + '_f_x'])) <= 1e-8 * max(1., abs(float(g[key + '_f_x'])))
    assert np.abs(al.g_x - g[key + '_g_x']).max() <= 1e-8 * max(1., np.abs(g[key + '_g_x']).max())
    assert np.array_equal(fit.support_, g[key + '_support'])
    assert abs(fit.intercept_ - float(g[key + '_intercept'])) <= 1e-8 * max(1., abs(float(g[key + '_intercept'])))
    dec = O.decision_function(fit, X_test)
    assert np.abs(dec - g[key + '_decision']).max() <= 1e-8 * max(1., np.abs(g[key + '_decision']).max())


@pytest.mark.parametrize('name,ri,c', SVC_RUNS)
def test_oracle_svc_matches_reference(golden, name, ri, c):
    g, iris = golden('al_stochastic'), golden('iris_ovr')
    rule, lr, kw, iters = CASES[name]
    fit = AL.svc_dual_al_fit(iris['X_train'], (iris['y_train'] == c).astype(int), reg_intercept=ri, learning_rate=lr,
                             max_iter=iters, random_state=c + 1, rule=rule, **kw)
    check_fit(g, svc_key(name, ri, c), fit, iris['X_test'])


@pytest.mark.parametrize('name,tol', TOL_RUNS)
def test_oracle_optimality_exit_matches_reference(golden, name, tol):
    g, iris = golden('al_stochastic'), golden('iris_ovr')
    rule, lr, kw, _ = CASES[name]
    fit = AL.svc_dual_al_fit(iris['X_train'], (iris['y_train'] == 2).astype(int), reg_intercept=False, learning_rate=lr,
                             tol=tol, max_iter=1000, random_state=7, rule=rule, **kw)
    assert fit.al.status == 'optimal' and fit.al.iter < 999
    check_fit(g, f'svc_{name}_tol_c2', fit, iris['X_test'])


@pytest.mark.parametrize('kernel,name,ri', SVR_RUNS)
def test_oracle_svr_matches_reference(golden, kernel, name, ri):
    g = golden('al_stochastic')
    rule, lr, kw, iters = CASES[name]
    fit = AL.svr_dual_al_fit(g['svr_X'], g['svr_y'], kind=kernel, epsilon=0.1, reg_intercept=ri, learning_rate=lr,
                             max_iter=min(iters, 400), random_state=3, rule=rule, **kw)
    check_fit(g, svr_key(kernel, name, ri), fit, g['svr_X_test'])


# ------------------------------------------------------------------------------- host emulation of the CUDA vector phase
def emulate_svc(X, yb, ri, rule, lr, kw, iters, seed, tol=1e-4, finalise_every=0):
    from oracle.emulator import al_emulate
    _, ys = O.binarize_labels(yb)
    n = len(ys)
    yy = np.outer(ys, ys)
    Q = O.gaussian_kernel(X) * yy + (yy if ri else 0)
    return al_emulate(Q, -np.ones(n), np.zeros(n), np.ones(n), AL.start_point(n, seed), A=None if ri else ys.astype(float),
                      rho=1., rule=rule, step_size=lr, tol=tol, epochs=iters, finalise_every=finalise_every, **kw)


def check_emulation(g, key, e, rtol=1e-9):
    assert e.iter == int(g[key + '_iter']) and e.status == str(g[key + '_status'])
    assert np.abs(e.x - g[key + '_alphas']).max() <= rtol * max(1., np.abs(g[key + '_alphas']).max())
    assert np.abs(e.dual_x - g[key + '_dual_x']).max() <= rtol * max(1., np.abs(g[key + '_dual_x']).max())
    pf = g[key + '_pf_hist']
    assert np.abs(e.pf_hist - pf).max() <= rtol * max(1., np.abs(pf).max())
    assert abs(e.f_hist[-1] - float(g[key + '_f_x'])) <= 1e-8 * max(1., abs(float(g[key + '_f_x'])))
    assert np.abs(e.g_x - g[key + '_g_x']).max() <= 1e-8 * max(1., np.abs(g[key + '_g_x']).max())


@pytest.mark.parametrize('name,ri,c', [r for r in SVC_RUNS if r[2] == 1])
def test_emulated_vector_phase_svc(golden, name, ri, c):
    g, iris = golden('al_stochastic'), golden('iris_ovr')
    rule, lr, kw, iters = CASES[name]
    for fe in (0, 5):  # 5: a FINALISE launch (step-wise host loop) before every fifth iteration changes nothing
        e = emulate_svc(iris['X_train'], (iris['y_train'] == c).astype(int), ri, rule, lr, kw, iters, c + 1, finalise_every=fe)
        check_emulation(g, svc_key(name, ri, c), e)


@pytest.mark.parametrize('name,tol', TOL_RUNS)
def test_emulated_vector_phase_optimality_exit(golden, name, tol):
    g, iris = golden('al_stochastic'), golden('iris_ovr')
    rule, lr, kw, _ = CASES[name]
    for fe in (0, 1):
        e = emulate_svc(iris['X_train'], (iris['y_train'] == 2).astype(int), False, rule, lr, kw, 1000, 7, tol=tol,
                        finalise_every=fe)
        assert e.status == 'optimal'
        check_emulation(g, f'svc_{name}_tol_c2', e)


@pytest.mark.parametrize('kernel,name,ri', SVR_RUNS)
def test_emulated_vector_phase_svr_block_layout(golden, kernel, name, ri):
    """the 2n-variable SVR problem on the n x n resident matrix M = K + bias: Q = [[M, -M], [-M, M]]"""
    from oracle.emulator import al_emulate
    g = golden('al_stochastic')
    rule, lr, kw, iters = CASES[name]
    X, y = g['svr_X'], g['svr_y']
    n = len(y)
    M = O.kernel_matrix(kernel, X) + (1. if ri else 0.)
    e_row = np.hstack((np.ones(n), -np.ones(n)))
    e = al_emulate(M, np.hstack((-y, y)) + 0.1, np.zeros(2 * n), np.ones(2 * n), AL.start_point(2 * n, 3),
                   A=None if ri else e_row, rho=1., rule=rule, step_size=lr, tol=1e-4, epochs=min(iters, 400), svr=True, **kw)
    check_emulation(g, svr_key(kernel, name, ri), e)


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_reference_own_lagrangian_quadratic_test(golden, seed):
    """opti/constrained/tests/test_lagrangian_quadratic.py:18-22: a general equality row (A = [2, 7]) and the
    optimality exit after ~200 iterations; oracle and emulated vector phase against the reference's run"""
    from oracle.emulator import al_emulate
    g, bc = golden('al_stochastic'), golden('bcqp')
    Q, q, ub, A = bc['p2_Q'], bc['p2_q'], bc['p2_ub'], np.array([2., 7.])
    key = f'alq2d_s{seed}'
    kw = dict(A=A, b=0., rho=1., rule='adagrad', step_size=1., tol=1e-8, epochs=15000)
    a = AL.al_stochastic(lambda v: Q @ v, q, np.zeros(2), ub, AL.start_point(2, seed), **kw)
    e = al_emulate(Q, q, np.zeros(2), ub, AL.start_point(2, seed), finalise_every=1, **kw)
    for r in (a, e):
        assert r.iter == int(g[key + '_iter']) and r.status == str(g[key + '_status']) == 'optimal'
        assert np.abs(r.x - g[key + '_x']).max() <= 1e-14 and np.abs(r.dual_x - g[key + '_dual_x']).max() <= 1e-12
        assert np.abs(r.pf_hist - g[key + '_pf_hist']).max() <= 1e-10 * np.abs(g[key + '_pf_hist']).max()
        assert np.allclose(r.x, np.zeros(2))  # the reference's assertion: x_star is the origin
