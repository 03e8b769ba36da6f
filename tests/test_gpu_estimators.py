"""End-to-end parity of the estimators (fit -> alphas_, support_, intercept_, decision_function, predict)
against golden vectors produced by the reference, through the public API (which calls the C ABI)."""
import numpy as np
import pytest

from oracle import svm_oracle as O
from optiml_b200.configs import make_config

pytestmark = pytest.mark.gpu


def api():
    from optiml_b200.ml.svm import SVC, SVR, DualSVC, DualSVR
    from optiml_b200.ml.svm.kernels import GaussianKernel, PolyKernel, LinearKernel, gaussian, linear
    from optiml_b200.ml.svm.losses import hinge, epsilon_insensitive
    from optiml_b200.opti.constrained import ProjectedGradient
    return locals()


def check_stable_fit(m, g, prefix='', atol=1e-8):
    """north_star bar: alpha within 1e-8, identical support set and predictions."""
    assert m.optimizer.iter == int(g[prefix + 'iter']) and m.optimizer.status == str(g[prefix + 'status'])
    assert np.abs(m.alphas_ - g[prefix + 'alphas']).max() <= atol
    assert np.array_equal(m.support_, g[prefix + 'support'])
    assert abs(m.intercept_ - float(g[prefix + 'intercept'])) <= 1e-8 * max(1., abs(float(g[prefix + 'intercept'])))
    fh, gh = np.array(m.train_loss_history), g[prefix + 'f_hist']
    assert len(fh) == len(gh) == m.optimizer.iter + 1
    assert np.abs(fh - gh).max() <= 1e-9 * np.abs(gh).max()


def test_c1_full_parity(golden):
    """BASELINE config C1 at full size: the reference runs this on CPU (1.3 s); all steps but two are
    bound-clipped, the trajectory is stable and alpha matches to ~1e-14."""
    A = api()
    g = golden('c1_svc_gaussian')
    spec, X, y = make_config('C1')
    m = A['SVC'](loss=A['hinge'], kernel=A['GaussianKernel'](), C=1, dual=True, reg_intercept=True,
                 optimizer=A['ProjectedGradient']).fit(X, y)
    check_stable_fit(m, g)
    assert len(m.support_) == 1015 and abs(m.optimizer.f_x - (-113.50663083792071)) < 1e-9
    dec = m.decision_function(X[:256])
    assert np.abs(dec - g['decision']).max() <= 1e-9
    assert np.array_equal(m.predict(X[:256]), g['predict'])
    assert m.score(X, y) == pytest.approx(0.992, abs=1e-12)
    assert np.array_equal(m.support_vectors_, X[m.support_])
    assert np.allclose(m.dual_coef_, m.alphas_[m.support_] * np.where(y[m.support_] == 1, 1, -1), rtol=0, atol=0)


def test_c4_headline_full_parity(golden):
    """BASELINE headline config (n=50 000, d=128, 20 GB Hessian): alpha / support set / intercept / loss history
    vs the unmodified reference solver run on the host (954.6 s there)."""
    A = api()
    g = golden('c4_full_svc_gaussian')
    spec, X, y = make_config('C4')
    assert np.array_equal([X.sum(), (X * X).sum()], g['X_checksum'])
    m = A['DualSVC'](kernel=A['GaussianKernel'](), C=1).fit(X, y)
    check_stable_fit(m, g)
    assert len(m.support_) == 49020
    # size-independent properties: feasibility, monotone descent, gradient consistency
    assert m.alphas_.min() >= -1e-12 and m.alphas_.max() <= 1 + 1e-12
    fh = np.array(m.train_loss_history)
    assert np.all(np.diff(fh) <= 1e-9 * np.abs(fh[:-1]))
    g_fresh = m.obj.jacobian(m.alphas_)  # one more pass over the resident Q
    assert np.abs(g_fresh - m.optimizer.g_x).max() <= 1e-9 * np.abs(g_fresh).max()
    assert abs(0.5 * m.alphas_ @ (g_fresh - 1.) - m.optimizer.f_x) <= 1e-10 * abs(m.optimizer.f_x)
    m.obj.release()


@pytest.mark.parametrize('c', [0, 1, 2])
def test_iris_ovr_binary_problems(golden, c):
    """ml/tests/test_svc.py:96-103 recipe, one binary problem at a time.  These trajectories are chaotic
    (free steps dominate): parity is asserted on the prefix of the loss history that precedes the
    amplification, on the optimum reached, and on the predictions."""
    A = api()
    g = golden('iris_ovr')
    yb = (g['y_train'] == c).astype(int)
    m = A['SVC'](loss=A['hinge'], kernel=A['gaussian'], reg_intercept=True, dual=True,
                 optimizer=A['ProjectedGradient']).fit(g['X_train'], yb)
    p = f'c{c}_'
    fh, gh = np.array(m.train_loss_history), g[p + 'f_hist']
    assert np.abs(fh[:100] - gh[:100]).max() <= 1e-9
    assert m.optimizer.status == str(g[p + 'status'])
    assert abs(m.optimizer.f_x - float(g[p + 'f_x'])) <= (1e-3 if c == 0 else 1e-9)  # c0 is stopped, not optimal
    assert np.array_equal(m.support_, g[p + 'support'])
    assert np.array_equal(m.predict(g['X_test']), g[p + 'predict'])
    ys = np.where(yb == 1, 1., -1.)
    r0 = O.svc_dual_fit(g['X_train'], yb, kind='gaussian')
    rng = np.random.default_rng(c)
    E = rng.integers(-1, 2, size=r0.Q.shape)
    E = np.triu(E) + np.triu(E, 1).T
    r1 = O.projected_gradient(r0.Q * (1 + E * 2.2e-16), -np.ones(len(ys)), np.ones(len(ys)))
    env = np.abs(r1.x - r0.pg.x).max()
    assert np.abs(m.alphas_ - g[p + 'alphas']).max() <= max(1e-8, 20 * env)


def test_iris_ovr_accuracy_like_reference_test(golden):
    """the reference's own acceptance test: OneVsRest accuracy >= 0.97 (it reaches 1.0)."""
    from sklearn.multiclass import OneVsRestClassifier as OVR
    A = api()
    g = golden('iris_ovr')
    svc = OVR(A['SVC'](loss=A['hinge'], kernel=A['gaussian'], reg_intercept=True, dual=True,
                       optimizer=A['ProjectedGradient']))
    svc = svc.fit(g['X_train'], g['y_train'])
    assert svc.score(g['X_test'], g['y_test']) >= 0.97
    assert svc.score(g['X_test'], g['y_test']) == 1.0


@pytest.mark.parametrize('name,kernel', [('linear', 'LinearKernel'), ('poly', 'PolyKernel'), ('gauss', 'GaussianKernel')])
def test_svr_diabetes(golden, name, kernel):
    """stand-in for ml/tests/test_svr.py:112-119 (Boston needs a download).  Chaotic trajectories: early
    history, objective level and fit quality are compared; alpha only against the sensitivity envelope."""
    A = api()
    g = golden('diabetes_svr')
    m = A['SVR'](loss=A['epsilon_insensitive'], kernel=A[kernel](), reg_intercept=True, dual=True,
                 optimizer=A['ProjectedGradient'], epsilon=0.1, C=1).fit(g['X_train'], g['y_train'])
    p = name + '_'
    fh, gh = np.array(m.train_loss_history), g[p + 'f_hist']
    assert m.optimizer.iter == 1000 and m.optimizer.status == 'stopped' and len(fh) == 1001
    assert np.abs(fh[:60] - gh[:60]).max() <= 1e-8 * np.abs(gh[:60]).max()
    assert abs(fh[-1] - gh[-1]) <= 2e-2 * abs(gh[-1])  # stopped, not converged: same level, not same point
    assert m.alphas_.min() >= -1e-12 and m.alphas_.max() <= 1 + 1e-12
    if name == 'linear':
        assert m.coef_.shape == (10,)
    from sklearn.metrics import r2_score
    r2_ref = r2_score(g['y_test'], g[p + 'decision'])
    # the stopped (non-converged) iterates differ, the model must not be worse than the reference's
    assert m.score(g['X_test'], g['y_test']) >= r2_ref - 0.1


@pytest.mark.parametrize('cfg,n,gold,builder', [
    ('C2', 600, 'c2small_svr_poly', lambda A: A['DualSVR'](kernel=A['PolyKernel'](degree=3), epsilon=0.1, C=1)),
    ('C3', 500, 'c3small_svc_linear', lambda A: A['DualSVC'](kernel=A['LinearKernel'](), C=1)),
    ('C4', 1200, 'c4small_svc_gaussian', lambda A: A['DualSVC'](kernel=A['GaussianKernel'](), C=2.5, max_iter=300)),
])
def test_reduced_configs_iteration_map(golden, cfg, n, gold, builder):
    """reduced C2/C3/C4 recipes: the first iterations (before chaotic amplification) match the reference's
    loss history to 1e-9 relative; a max_iter=20 fit matches the oracle's alpha to 1e-10."""
    A = api()
    g = golden(gold)
    spec, X, y = make_config(cfg, n=n)
    m = builder(A).fit(X, y)
    fh, gh = np.array(m.train_loss_history), g['f_hist']
    assert len(fh) == len(gh) and m.optimizer.status == str(g['status'])
    assert np.abs(fh[:20] - gh[:20]).max() <= 1e-9 * np.abs(gh[:20]).max()
    short = builder(A).set_params(max_iter=20).fit(X, y)
    if spec['task'] == 'svr':
        ref = O.svr_dual_fit(X, y, kind='poly', degree=3, epsilon=0.1, C=1, max_iter=20)
    else:
        ref = O.svc_dual_fit(X, y, kind=spec['kernel'], C=short.C, max_iter=20)
    assert np.abs(short.alphas_ - ref.alphas_).max() <= 1e-10
    assert np.array_equal(short.support_, ref.support_)
    assert abs(short.intercept_ - ref.intercept_) <= 1e-9 * max(1., abs(ref.intercept_))
    dec = short.decision_function(X[:128])
    assert np.abs(dec - O.decision_function(ref, X[:128])).max() <= 1e-8 * max(1., np.abs(dec).max())
    if spec['kernel'] == 'linear':
        assert np.abs(short.coef_ - ref.coef_).max() <= 1e-10


def test_full_size_c2_c3_properties(golden):
    """C2 (SVR poly, n=10 000 -> 20 000 variables, only K+1 resident) and C3 (linear, n=20 000, d=784) at full
    size: early loss history vs the reference's full-size run when its golden file is present, plus
    size-independent properties."""
    import os
    GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    A = api()
    for cfg, gold, mk in (('C2', 'c2_full_svr_poly', lambda: A['DualSVR'](kernel=A['PolyKernel'](degree=3), epsilon=0.1, C=1)),
                          ('C3', 'c3_full_svc_linear', lambda: A['DualSVC'](kernel=A['LinearKernel'](), C=1))):
        spec, X, y = make_config(cfg)
        m = mk().fit(X, y)
        fh = np.array(m.train_loss_history)
        assert m.optimizer.iter == 1000 and m.optimizer.status == 'stopped'
        assert np.all(np.diff(fh) <= 1e-9 * np.abs(fh[:-1]))
        assert m.alphas_.min() >= -1e-12 and m.alphas_.max() <= 1 + 1e-12
        g_fresh = m.obj.jacobian(m.alphas_)
        assert np.abs(g_fresh - m.optimizer.g_x).max() <= 1e-8 * max(1., np.abs(g_fresh).max())
        if os.path.exists(os.path.join(GOLDEN, gold + '.npz')):
            g = golden(gold)  # the reference's own full-size run (198 s / 270 s on 8 host cores)
            assert np.array_equal(m.support_, g['support'])
            if cfg == 'C3':
                # stable trajectory: north_star's bar at full size
                assert np.abs(m.alphas_ - g['alphas']).max() <= 1e-8
                assert np.abs(fh - g['f_hist']).max() <= 1e-9 * np.abs(g['f_hist']).max()
                assert abs(m.intercept_ - float(g['intercept'])) <= 1e-8
                assert np.abs(m.coef_ - g['coef']).max() <= 1e-8 * np.abs(g['coef']).max()
                assert np.abs(m.decision_function(X[:256]) - g['decision']).max() <= 1e-7
            else:
                # C2 leaves the reference's trajectory after ~100 iterations (chaotic regime, DESIGN.md section 2;
                # pinned by test_c2_full_size_iteration_map_from_reference_states_and_envelope)
                assert np.abs(fh[:60] - g['f_hist'][:60]).max() <= 1e-9 * np.abs(g['f_hist'][:60]).max()
                assert fh[-1] <= g['f_hist'][-1] * 1.1
        m.obj.release()


def test_c2_full_size_iteration_map_from_reference_states_and_envelope(golden):
    """BASELINE config C2 at full size (DualSVR, PolyKernel(3), n = 10 000 -> 20 000 variables).  Almost every step of
    this run is a FREE exact line-search step (48 of 1000 hit a bound), the regime in which the reference's iteration is a
    chaotic map: the reference algorithm itself, fed a Gram matrix that differs by +-1 ulp, ends 1.9e-2 away in alpha
    and 5 % away in f (``env_*`` in tests/golden/c2_full_states.npz, made by make_golden_c2_states.py; the same is shown
    on a reduced problem in tests/test_oracle_sensitivity.py and, on the device, by profiles/r2_c2_isolate.json).  No
    implementation other than the bit-identical one can hold 1e-8 after 1000 iterations, so parity is pinned where it
    can be:
      (a) the ITERATION MAP: started from seven states of the REAL reference's own trajectory -- before, at and long
          after the point where the loss histories part (k = 98) -- five device iterations land on the reference's state
          five iterations later to 1e-11 (5e-10 from the degenerate start point, where the NumPy oracle itself is
          4.9e-11 away: the first denominator d'Qd cancels five digits);
      (b) the end of the full run stays inside 20x the reference algorithm's own 1-ulp envelope, with the reference's
          status, iteration count and support set."""
    A = api()
    from optiml_b200.opti import Quadratic
    from optiml_b200.opti.constrained import ProjectedGradient
    g, ref = golden('c2_full_states'), golden('c2_full_svr_poly')
    spec, X, y = make_config('C2')
    n = len(y)
    m = A['DualSVR'](kernel=A['PolyKernel'](degree=3), epsilon=0.1, C=1).fit(X, y)
    # (b)
    assert m.optimizer.iter == int(ref['iter']) == 1000 and m.optimizer.status == str(ref['status']) == 'stopped'
    assert np.array_equal(m.support_, ref['support'])
    env = float(g['env_dalpha'])
    assert env > 1e-3                                                    # the envelope is 6 orders above north_star's bar
    assert np.abs(m.alphas_ - ref['alphas']).max() <= 20 * env
    f_spread = abs(float(g['env_f_base'][-1]) - float(g['env_f_pert'][-1]))
    assert abs(m.optimizer.f_x - float(ref['f_hist'][-1])) <= 20 * f_spread
    fh = np.array(m.train_loss_history)
    assert np.abs(fh[:90] - ref['f_hist'][:90]).max() <= 1e-9 * np.abs(ref['f_hist'][:90]).max()
    # (a) on the Hessian the fit left in HBM (only M = K + 1 is resident, block signs applied by the solver)
    H = m.obj.device_hessian()
    q, ub = np.hstack((-y, y)) + 0.1, np.ones(2 * n)
    steps = int(g['steps'])
    k_at, f_at = list(g['k_at']), g['f_at']
    for k in g['ks']:
        k = int(k)
        opt = ProjectedGradient(quad=Quadratic(H, q), ub=ub, x=g[f'x_{k}'].copy(), max_iter=steps).minimize()
        want_x, want_f = g[f'x_{k + steps}'], float(f_at[k_at.index(k + steps)])
        assert opt.iter == steps and opt.status == 'stopped'
        tol = 5e-10 if k == 0 else 1e-11
        assert np.abs(opt.x - want_x).max() <= tol, (k, np.abs(opt.x - want_x).max())
        assert abs(opt.f_x - want_f) <= 1e-9 * max(1., abs(want_f))
        assert np.array_equal(opt.x > 1e-6, want_x > 1e-6)
    m.obj.release()


def test_sklearn_protocol_and_errors():
    from sklearn.base import clone
    A = api()
    m = A['DualSVC'](kernel=A['GaussianKernel'](gamma=0.5), C=3.)
    c = clone(m)
    assert c.get_params()['C'] == 3. and c.kernel.gamma == 0.5 and c.optimizer is A['ProjectedGradient']
    X = np.array([[0., 0.], [1., 1.], [0., 1.], [1., 0.], [2., 2.], [2., 0.]])
    with pytest.raises(ValueError):
        A['DualSVC']().fit(X, [0, 1, 2, 0, 1, 2])  # > 2 classes (ml/svm/_base.py:437-439)
    with pytest.raises(ValueError):
        A['DualSVR']().fit(X, np.ones((6, 2)))  # multi-target (:981-983)
    with pytest.raises(NotImplementedError):
        A['SVC'](loss=A['hinge'], dual=True, reg_intercept=False, optimizer=A['ProjectedGradient']).fit(X, [0, 1, 0, 1, 0, 1])
    # a second fit on the same estimator works (the reference fails here)
    m = A['DualSVC'](kernel=A['GaussianKernel'](), C=1., max_iter=50)
    a1 = m.fit(X, [0, 1, 0, 1, 1, 0]).alphas_.copy()
    a2 = m.fit(X, [0, 1, 0, 1, 1, 0]).alphas_
    assert np.array_equal(a1, a2) and isinstance(m.optimizer, A['ProjectedGradient'])
    assert m.loss is A['hinge']
