"""Widening step 3: the augmented-Lagrangian dual driven by the reference's full-batch stochastic
optimisers, through the public API (``SVC/SVR(dual=True, optimizer=AdaGrad, ...)`` and the optimiser classes
directly) against golden vectors of the REAL reference (tests/golden/make_golden_al.py) and against the NumPy oracle.
The trajectories are stable (every case agrees with the reference to ~1e-12), so north_star's bar -- alpha within
1e-8, identical support set and predictions -- is asserted on every case."""
import numpy as np
import pytest

from al_cases import CASES, OPTIMIZER_CLASS, SVC_RUNS, SVR_RUNS, TOL_RUNS, svc_key, svr_key
from oracle import al_oracle as AL, svm_oracle as O
from optiml_b200.configs import make_config

pytestmark = pytest.mark.gpu


def optimizer_class(rule):
    import optiml_b200.opti.unconstrained.stochastic as S
    return getattr(S, OPTIMIZER_CLASS[rule])


def check_estimator(g, key, m, X_test):
    o = m.optimizer
    assert o.iter == int(g[key + '_iter']) and o.status == str(g[key + '_status'])
    assert np.abs(m.alphas_ - g[key + '_alphas']).max() <= 1e-8 * max(1., np.abs(g[key + '_alphas']).max())
    assert np.abs(o.f.dual_x - g[key + '_dual_x']).max() <= 1e-8 * max(1., np.abs(g[key + '_dual_x']).max())
    pf, hist = g[key + '_pf_hist'], np.array(m.train_loss_history)
    assert len(hist) == len(pf) == o.iter + 1
    assert np.abs(hist - pf).max() <= 1e-9 * max(1., np.abs(pf).max())
    assert abs(o.f_x - float(g[key + '_f_x'])) <= 1e-8 * max(1., abs(float(g[key + '_f_x'])))
    assert np.abs(o.g_x - g[key + '_g_x']).max() <= 1e-8 * max(1., np.abs(g[key + '_g_x']).max())
    assert np.array_equal(m.support_, g[key + '_support'])
    assert abs(m.intercept_ - float(g[key + '_intercept'])) <= 1e-8 * max(1., abs(float(g[key + '_intercept'])))
    assert np.abs(m.decision_function(X_test) - g[key + '_decision']).max() <= 1e-8 * max(1., np.abs(g[key + '_decision']).max())
    assert o.epoch == o.iter + 1 and np.all(o.f.dual_x[o.f.n_eq:] >= 0)


@pytest.mark.filterwarnings('ignore::sklearn.exceptions.ConvergenceWarning')
@pytest.mark.parametrize('name,ri,c', SVC_RUNS)
def test_svc_matches_reference(golden, name, ri, c):
    from optiml_b200.ml.svm import SVC
    from optiml_b200.ml.svm.kernels import gaussian
    from optiml_b200.ml.svm.losses import hinge
    g, iris = golden('al_stochastic'), golden('iris_ovr')
    rule, lr, kw, iters = CASES[name]
    m = SVC(loss=hinge, kernel=gaussian, reg_intercept=ri, dual=True, optimizer=optimizer_class(rule), learning_rate=lr,
            max_iter=iters, random_state=c + 1, **kw).fit(iris['X_train'], (iris['y_train'] == c).astype(int))
    check_estimator(g, svc_key(name, ri, c), m, iris['X_test'])


@pytest.mark.parametrize('name,tol', TOL_RUNS)
def test_optimality_exit_matches_reference(golden, name, tol):
    """a loose tolerance ends the run through the test of the multiplier update (opti/_base.py:143-147): x has moved,
    f_x / g_x stay those of the last evaluation, iter is not advanced, no ConvergenceWarning"""
    import warnings
    from optiml_b200.ml.svm import SVC
    from optiml_b200.ml.svm.kernels import gaussian
    from optiml_b200.ml.svm.losses import hinge
    g, iris = golden('al_stochastic'), golden('iris_ovr')
    rule, lr, kw, _ = CASES[name]
    with warnings.catch_warnings():
        warnings.simplefilter('error')
        m = SVC(loss=hinge, kernel=gaussian, reg_intercept=False, dual=True, optimizer=optimizer_class(rule),
                learning_rate=lr, tol=tol, max_iter=1000, random_state=7, **kw).fit(iris['X_train'],
                                                                                    (iris['y_train'] == 2).astype(int))
    assert m.optimizer.status == 'optimal'
    check_estimator(g, f'svc_{name}_tol_c2', m, iris['X_test'])


@pytest.mark.filterwarnings('ignore::sklearn.exceptions.ConvergenceWarning')
@pytest.mark.parametrize('kernel,name,ri', SVR_RUNS)
def test_svr_matches_reference(golden, kernel, name, ri):
    from optiml_b200.ml.svm import SVR
    from optiml_b200.ml.svm import kernels as Kn
    from optiml_b200.ml.svm.losses import epsilon_insensitive
    g = golden('al_stochastic')
    rule, lr, kw, iters = CASES[name]
    m = SVR(loss=epsilon_insensitive, epsilon=0.1, kernel=getattr(Kn, kernel), reg_intercept=ri, dual=True,
            optimizer=optimizer_class(rule), learning_rate=lr, max_iter=min(iters, 400), random_state=3, **kw)
    m = m.fit(g['svr_X'], g['svr_y'])
    check_estimator(g, svr_key(kernel, name, ri), m, g['svr_X_test'])
    if kernel == 'linear':
        assert np.allclose(m.coef_, m.dual_coef_ @ m.support_vectors_)


def test_convergence_warning_and_refit():
    """ml/svm/_base.py:715-717: a run that hits max_iter warns; the fitted optimiser instance can be refitted"""
    from sklearn.exceptions import ConvergenceWarning
    from optiml_b200.ml.svm import SVC
    from optiml_b200.ml.svm.kernels import gaussian
    from optiml_b200.ml.svm.losses import hinge
    from optiml_b200.opti.unconstrained.stochastic import AdaGrad
    rng = np.random.default_rng(0)
    X = rng.standard_normal((60, 3))
    y = (X[:, 0] > 0).astype(int)
    m = SVC(loss=hinge, kernel=gaussian, reg_intercept=False, dual=True, optimizer=AdaGrad, learning_rate=1., max_iter=20,
            random_state=0)
    with pytest.warns(ConvergenceWarning, match='max_iter reached'):
        m.fit(X, y)
    a = m.alphas_.copy()
    with pytest.warns(ConvergenceWarning):
        m.fit(X, y)
    assert np.array_equal(a, m.alphas_) and isinstance(m.optimizer, AdaGrad)


# ------------------------------------------------------------------------------- the optimiser classes used directly
def random_problem(n, seed, eq=True):
    rng = np.random.default_rng(seed)
    G = rng.standard_normal((n + 5, n))
    Q = np.ascontiguousarray(G.T @ G / n)
    q = rng.standard_normal(n)
    ub = rng.uniform(0.5, 2., n)
    lb = -rng.uniform(0., 1., n)
    A = np.where(rng.random(n) < 0.5, 1., -1.) if eq else None
    return Q, q, lb, ub, A


@pytest.mark.parametrize('n,eq', [(2, True), (3, False), (64, True), (257, True), (1000, False)])
def test_adagrad_on_host_resident_quadratic_matches_oracle(n, eq):
    """AdaGrad(f=AugmentedLagrangianQuadratic(primal=Quadratic(Q, q), A, b, lb, ub, rho)) with general bounds, a general
    right-hand side b and rho != 1, on ragged sizes (ndim <= 3 takes the step-wise loop with its histories)"""
    from optiml_b200.opti import Quadratic
    from optiml_b200.opti.constrained import AugmentedLagrangianQuadratic
    from optiml_b200.opti.unconstrained.stochastic import AdaGrad
    Q, q, lb, ub, A = random_problem(n, n, eq)
    b = np.array([0.3]) if eq else None
    f = AugmentedLagrangianQuadratic(primal=Quadratic(Q, q), A=A, b=b, lb=lb, ub=ub, rho=2.5)
    opt = AdaGrad(f=f, step_size=0.5, epochs=120, tol=1e-10, random_state=n).minimize()
    want = AL.al_stochastic(lambda v: Q @ v, q, lb, ub, AL.start_point(n, n), A=A, b=0.3 if eq else 0., rho=2.5,
                            rule='adagrad', step_size=0.5, tol=1e-10, epochs=120)
    assert opt.iter == want.iter and opt.status == want.status
    assert np.abs(opt.x - want.x).max() <= 1e-10 and np.abs(f.dual_x - want.dual_x).max() <= 1e-9
    assert abs(opt.f_x - want.f_x) <= 1e-9 * max(1., abs(want.f_x))
    assert np.abs(opt.g_x - want.g_x).max() <= 1e-9 * max(1., np.abs(want.g_x).max())
    assert abs(opt.primal_f_x - want.primal_f_x) <= 1e-9 * max(1., abs(want.primal_f_x))
    assert abs(opt.dgap - want.dgap) <= 1e-9
    if n <= 3:
        assert len(opt.f_x_history) == (want.iter + 1 if n == 2 else 0)  # opti/_base.py:107-110: only for ndim == 2
    # the objective evaluated on the host side of the mirror (one device pass per call) agrees with the oracle
    c = AL._constraints(want.x, A, 0.3 if eq else 0., lb, ub)
    fo, go, _ = AL.al_function_jacobian(Q @ want.x, want.x, q, want.dual_x, c, A, 0.3 if eq else 0., lb, ub, 2.5)
    f.dual_x = want.dual_x.copy()
    fv, gv = f.function_jacobian(want.x)
    assert abs(fv - fo) <= 1e-10 * max(1., abs(fo)) and np.abs(gv - go).max() <= 1e-10 * max(1., np.abs(go).max())


def test_stepwise_loop_equals_resident_loop_and_callback_protocol(capsys):
    """a generic callback forces one synchronisation per iteration: same iterates as the device-resident loop;
    StopIteration ends the run with status 'unknown'; verbose prints the reference's columns"""
    from optiml_b200.opti import Quadratic
    from optiml_b200.opti.constrained import AugmentedLagrangianQuadratic
    from optiml_b200.opti.unconstrained.stochastic import Adam
    Q, q, lb, ub, A = random_problem(96, 5)
    mk = lambda: AugmentedLagrangianQuadratic(primal=Quadratic(Q, q), A=A, b=np.zeros(1), lb=lb, ub=ub, rho=1.)
    kw = dict(step_size=0.01, momentum_type='nesterov', momentum=0.5, epochs=60, tol=1e-12, random_state=1)
    f0 = mk()
    base = Adam(f=f0, **kw).minimize()
    seen = []
    f1 = mk()
    step = Adam(f=f1, callback=lambda o: seen.append((o.iter, o.f_x, o.primal_f_x, o.x.copy(), o.g_x.copy())), **kw).minimize()
    assert step.iter == base.iter == 59 and step.status == base.status == 'stopped'
    assert np.array_equal(step.x, base.x) and np.array_equal(f1.dual_x, f0.dual_x) and step.f_x == base.f_x
    assert [s[0] for s in seen] == list(range(60))
    assert np.array_equal(np.array([s[1] for s in seen]), base.f_hist)
    assert np.array_equal(np.array([s[2] for s in seen]), base.pf_hist)
    want = AL.al_stochastic(lambda v: Q @ v, q, lb, ub, AL.start_point(96, 1), A=A, rho=1., rule='adam', step_size=0.01,
                            momentum_type='nesterov', momentum=0.5, tol=1e-12, epochs=60)
    assert np.abs(seen[-1][4] - want.g_x).max() <= 1e-9 * max(1., np.abs(want.g_x).max())  # g_x at the callback point

    def stop_at_7(o):
        if o.iter == 7:
            raise StopIteration

    early = Adam(f=mk(), callback=stop_at_7, **kw).minimize()
    assert early.iter == 7 and early.status == 'unknown' and np.array_equal(early.x, seen[7][3])
    capsys.readouterr()
    Adam(f=mk(), verbose=20, **kw).minimize()
    out = capsys.readouterr().out
    assert out.startswith('epoch\titer\t cost\t')
    rows = [ln for ln in out.split('\n') if ln.strip() and not ln.startswith('epoch')]
    assert len(rows) == 3 and rows[1].startswith('  20\t  20\t') and '\tpcost: ' in rows[1] and '\tdgap: ' in rows[1]


@pytest.mark.filterwarnings('ignore::sklearn.exceptions.ConvergenceWarning')
def test_c1_size_adagrad_both_formulations():
    """BASELINE config C1 (n = 2000) for 150 iterations against the oracle, plus size-independent properties of the
    multipliers: complementary signs, lambda >= 0, the primal cost equals x'Qx/2 + q'x recomputed from alpha"""
    from optiml_b200.ml.svm import SVC
    from optiml_b200.ml.svm.kernels import GaussianKernel
    from optiml_b200.ml.svm.losses import hinge
    from optiml_b200.opti.unconstrained.stochastic import AdaGrad
    spec, X, y = make_config('C1')
    for ri in (True, False):
        m = SVC(loss=hinge, kernel=GaussianKernel(), C=1, reg_intercept=ri, dual=True, optimizer=AdaGrad, learning_rate=1.,
                max_iter=150, random_state=11).fit(X, y)
        want = AL.svc_dual_al_fit(X, y, kind='gaussian', C=1., reg_intercept=ri, learning_rate=1., max_iter=150,
                                  random_state=11)
        assert m.optimizer.iter == want.al.iter == 149
        assert np.abs(m.alphas_ - want.alphas_).max() <= 1e-8
        assert np.abs(m.optimizer.f.dual_x - want.al.dual_x).max() <= 1e-8 * max(1., np.abs(want.al.dual_x).max())
        assert np.array_equal(m.support_, want.support_) and abs(m.intercept_ - want.intercept_) <= 1e-8
        hist = np.array(m.train_loss_history)
        assert np.abs(hist - want.al.pf_hist).max() <= 1e-9 * np.abs(want.al.pf_hist).max()
        x = m.alphas_
        assert abs(hist[-1] - (0.5 * x @ (want.Q @ x) - x.sum())) <= 1e-9 * abs(hist[-1])
        n_eq = m.optimizer.f.n_eq
        lam_lb, lam_ub = m.optimizer.f.dual_x[n_eq:n_eq + len(x)], m.optimizer.f.dual_x[n_eq + len(x):]
        assert lam_lb.min() >= 0 and lam_ub.min() >= 0 and not np.any((lam_lb > 0) & (lam_ub > 0))


def test_c_abi_argument_checks():
    import ctypes as C
    from optiml_b200 import _native as N
    from optiml_b200.runtime import default_context
    ctx, lib = default_context(), N.load_library()
    h = C.c_void_p()
    v = np.zeros(4)
    lr = np.ones(10)
    args = lambda **o: [ctx.handle, C.c_void_p(256), 4, 4, 0, 4, 0, N.ptr(v), N.ptr(v), N.ptr(v), o.get('x0', N.ptr(v)), None, 0.,
                        o.get('rho', 1.), o.get('rule', 0), o.get('mom', 0), N.ptr(lr), None, 0.9, 0.9, 0.999,
                        o.get('offset', 1e-8), 1e-4, 10, C.byref(h)]
    for bad, text in ((dict(rho=0.), b'rho'), (dict(rule=9), b'update rule'), (dict(mom=1), b'no momentum'),
                      (dict(offset=0.), b'offset'), (dict(x0=None), b'start point')):
        assert lib.svmb200_al_create(*args(**bad)) == 1 and text in lib.svmb200_last_error()
    assert lib.svmb200_al_multipliers(None, None, None, None) == 1


def test_reference_acceptance_recipe_ovr_adagrad():
    """the reference's own test (ml/tests/test_svc.py:134-147): OneVsRest over iris, unseeded AdaGrad with
    learning_rate=1., both intercept formulations, accuracy >= 0.97"""
    import warnings
    from sklearn.datasets import load_iris
    from sklearn.model_selection import train_test_split
    from sklearn.multiclass import OneVsRestClassifier as OVR
    from sklearn.preprocessing import MinMaxScaler
    from optiml_b200.ml.svm import SVC
    from optiml_b200.ml.svm.kernels import gaussian
    from optiml_b200.ml.svm.losses import hinge
    from optiml_b200.opti.unconstrained.stochastic import AdaGrad
    X, y = load_iris(return_X_y=True)
    X_scaled = MinMaxScaler().fit_transform(X)
    X_train, X_test, y_train, y_test = train_test_split(X_scaled, y, train_size=0.75, random_state=123456)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for ri in (True, False):
            svc = OVR(SVC(loss=hinge, kernel=gaussian, reg_intercept=ri, dual=True, optimizer=AdaGrad, learning_rate=1.))
            svc = svc.fit(X_train, y_train)
            assert svc.score(X_test, y_test) >= 0.97


def test_fitted_estimators_pickle_without_device_handles():
    """joblib.dump / sklearn meta-estimators copy fitted models: the copy predicts identically and carries no device
    state (the Hessian stays on the GPU of the original; asking the copy for Q says so)"""
    import copy
    import pickle
    import warnings
    from sklearn.multiclass import OneVsRestClassifier as OVR
    from optiml_b200.ml.svm import DualSVC, SVC
    from optiml_b200.ml.svm.kernels import gaussian
    from optiml_b200.ml.svm.losses import hinge
    from optiml_b200.opti.constrained import FrankWolfe
    from optiml_b200.opti.unconstrained.stochastic import Adam
    rng = np.random.default_rng(3)
    X = rng.standard_normal((150, 4))
    y = (X[:, 0] + 0.3 * X[:, 1] > 0).astype(int)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        models = [DualSVC(kernel=gaussian, max_iter=100).fit(X, y),
                  SVC(loss=hinge, kernel=gaussian, dual=True, reg_intercept=False, optimizer=Adam, learning_rate=0.01,
                      momentum_type='polyak', max_iter=60, random_state=0).fit(X, y),
                  OVR(DualSVC(kernel=gaussian, optimizer=FrankWolfe, max_iter=80)).fit(X, rng.integers(0, 3, 150))]
    for m in models:
        for clone in (pickle.loads(pickle.dumps(m)), copy.deepcopy(m)):
            assert np.array_equal(clone.predict(X[:40]), m.predict(X[:40]))
    solo = pickle.loads(pickle.dumps(models[0]))
    assert np.array_equal(solo.alphas_, models[0].alphas_) and solo.optimizer.status == models[0].optimizer.status
    assert len(solo.train_loss_history) == len(models[0].train_loss_history)
    with pytest.raises(RuntimeError, match='device copy of Q was released'):
        solo.obj.Q
    assert models[0].obj.Q.shape == (150, 150)  # the original still owns its device matrix


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_reference_own_lagrangian_quadratic_test(golden, seed):
    """opti/constrained/tests/test_lagrangian_quadratic.py:18-22 as written there (a general equality row, two
    variables, so the step-wise loop with its histories; the run ends through the optimality test)"""
    from optiml_b200.opti import Quadratic
    from optiml_b200.opti.constrained import AugmentedLagrangianQuadratic
    from optiml_b200.opti.unconstrained.stochastic import AdaGrad
    g, bc = golden('al_stochastic'), golden('bcqp')
    Q, q, ub = bc['p2_Q'], bc['p2_q'], bc['p2_ub']
    A, b, lb = [2, 7], np.zeros(1), np.zeros_like(q)
    ld = AugmentedLagrangianQuadratic(primal=Quadratic(Q, q), A=A, b=b, lb=lb, ub=ub, rho=1)
    opt = AdaGrad(ld, step_size=1, epochs=15000, random_state=seed).minimize()
    assert np.allclose(opt.x, np.zeros(2))  # ld.x_star() of the reference's assertion: the origin
    key = f'alq2d_s{seed}'
    assert opt.iter == int(g[key + '_iter']) and opt.status == str(g[key + '_status']) == 'optimal'
    assert np.abs(opt.x - g[key + '_x']).max() <= 1e-12 and np.abs(ld.dual_x - g[key + '_dual_x']).max() <= 1e-10
    assert abs(opt.f_x - float(g[key + '_f_x'])) <= 1e-10 and np.abs(opt.g_x - g[key + '_g_x']).max() <= 1e-9
    assert np.abs(np.array(opt.f_x_history) - g[key + '_pf_hist']).max() <= 1e-10 * np.abs(g[key + '_pf_hist']).max()
    assert len(opt.x0_history) == opt.iter + 1 and opt.epoch == opt.iter + 1
