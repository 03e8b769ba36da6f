"""Checks of the shared-Gram path (multi-vector pass, signed Hessian views, lockstep batches, the
one-vs-rest / multi-target meta-estimators), written once and run twice: on the host emulation of the kernels in the
CPU suite (tests/test_emulated_kernels.py) and on the B200 in the `-m gpu` suite (tests/test_gpu_shared_gram.py).

``device`` is a callable returning a context manager that yields a probe with ``launches()`` (kernels launched so far)
and ``assert_clean()``.
"""
import contextlib
import ctypes as C
import warnings

import numpy as np
import pytest

from oracle import svm_oracle as O
from optiml_b200 import _native as N


def psd(rng, n, rank=None, shift=1.0):
    G = rng.standard_normal((n, rank or max(3, n // 4)))
    return G @ G.T / G.shape[1] + shift


@contextlib.contextmanager
def real_device(**_):
    """The B200 behind the default context."""
    from optiml_b200.runtime import default_context

    class Probe:
        def launches(self):
            return default_context().launch_count()

        def assert_clean(self):
            default_context().sync()

    yield Probe()


def upload_shared(M):
    """An unsigned resident matrix the problems of a batch share."""
    from optiml_b200.runtime import DeviceHessian, default_context
    ctx = default_context()
    n = M.shape[0]
    shared = DeviceHessian(ctx, n, 'plain')
    block = np.zeros((max(shared.nrows, 1), shared.ld))
    block[:shared.nrows, :n] = M[shared.row0:shared.row0 + shared.nrows]
    ctx.h2d(shared.matrix.dptr, block)
    return shared


# --------------------------------------------------------------------------------------------- K2 x NB
def check_multi_vector_pass(device, n, count, **dev_kw):
    """svmb200_matvec_multi == svmb200_matvec, bit for bit, in ceil(count / 4) passes"""
    from optiml_b200.runtime import default_context
    rng = np.random.default_rng(n + count)
    with device(**dev_kw) as probe:
        ctx = default_context()
        ld = N.padded_ld(n)
        Q = np.zeros((n, ld))
        Q[:, :n] = rng.standard_normal((n, n))
        dQ = ctx.malloc(Q.nbytes)
        ctx.h2d(dQ, Q)
        us, dus, dws = [], [], []
        for b in range(count):
            u = np.zeros(ld)
            u[:n] = rng.standard_normal(n)
            us.append(u)
            dus.append(ctx.malloc(u.nbytes))
            ctx.h2d(dus[-1], u)
            dws.append(ctx.malloc(8 * n))
        single = []
        for b in range(count):
            N.call('svmb200_matvec', ctx.handle, C.c_void_p(dQ), n, ld, C.c_void_p(dus[b]), C.c_void_p(dws[b]))
            w = np.empty(n)
            ctx.d2h(w, dws[b])
            single.append(w)
            ctx.memset(dws[b], 0xFF, 8 * n)
        before = probe.launches()
        N.call('svmb200_matvec_multi', ctx.handle, C.c_void_p(dQ), n, ld, (C.c_void_p * count)(*dus),
               (C.c_void_p * count)(*dws), count)
        assert probe.launches() - before == (count + 3) // 4   # passes over the matrix
        for b in range(count):
            w = np.empty(n)
            ctx.d2h(w, dws[b])
            assert np.array_equal(w, single[b])
            assert np.abs(w - Q[:, :n] @ us[b][:n]).max() <= 1e-12 * n
        for p in [dQ] + dus + dws:
            ctx.free(p)
        probe.assert_clean()


# --------------------------------------------------------------------------------------------- signed views, batches
def make_solvers(kind, quad_for, signs, ub, max_iter, eps_list):
    from optiml_b200.opti.constrained import AugmentedLagrangianQuadratic, FrankWolfe, ProjectedGradient
    from optiml_b200.opti.unconstrained.stochastic import AdaGrad, Adam
    out = []
    for c, s in enumerate(signs):
        quad = quad_for(c)
        if kind == 'pg':
            out.append(ProjectedGradient(quad=quad, ub=ub, max_iter=max_iter, eps=eps_list[c % len(eps_list)]))
        elif kind == 'fw':
            out.append(FrankWolfe(quad=quad, ub=ub, max_iter=max_iter, t=0.2, eps=eps_list[c % len(eps_list)]))
        else:
            # ml/svm/_base.py:638-655: equality row y'alpha = 0 when the intercept is not regularised
            f = AugmentedLagrangianQuadratic(primal=quad, A=s, b=np.zeros(1), lb=np.zeros_like(ub), ub=ub, rho=1.)
            if kind == 'adagrad':
                out.append(AdaGrad(f=f, step_size=1., epochs=max_iter, random_state=c, tol=1e-4))
            else:
                out.append(Adam(f=f, step_size=0.05, epochs=max_iter, random_state=c, tol=1e-4, momentum_type='nesterov',
                                momentum=0.5))
    return out


def solver_state(s):
    hist = [s.f_hist] + ([s.pf_hist, s.f.dual_x] if hasattr(s, 'pf_hist') else [s.ng_hist])
    return [np.asarray(s.x), np.asarray(s.g_x), np.array([s.iter, s.f_x]), np.array([s.status == 'optimal'])] + hist


def check_signed_views_and_batches(device, kind, count, n, max_iter, **dev_kw):
    """(1) a solver on the signed view (s s') o M == the same solver on the materialised Q;  (2) the lockstep batch ==
    the solvers run one after the other -- including a problem that stops early (its threshold is crossed mid-run
    while the others keep going).  All comparisons are bitwise."""
    from optiml_b200.opti import Quadratic
    from optiml_b200.opti.batch import batchable, minimize_batch
    rng = np.random.default_rng(count * 17 + len(kind))
    M = psd(rng, n, shift=0.0) + 1.0
    signs = [np.where(rng.random(n) < 0.3 + 0.1 * c, 1.0, -1.0) for c in range(count)]
    q, ub = -np.ones(n), np.ones(n)
    eps_list = [1e-6, 1e-6, 1e-6]
    with device(**dev_kw) as probe:
        shared = upload_shared(M)
        if kind in ('pg', 'fw'):
            # a stopping threshold for problem 1 that its own criterion (|d| or the gap) crosses mid-run
            first = make_solvers(kind, lambda c: Quadratic(shared.with_signs(signs[c]), q), signs, ub, max_iter, eps_list)[1]
            crit = first.minimize().ng_hist
            eps_list[1] = float(np.min(crit[:max_iter // 2])) * (1 + 1e-12)

        def run(quad_for, batch):
            solvers = make_solvers(kind, quad_for, signs, ub, max_iter, eps_list)
            launches = None
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                if batch:
                    assert batchable(solvers)
                    launches = probe.launches()
                    minimize_batch(solvers)
                    launches = probe.launches() - launches
                    assert all(s.batch_size_ == count for s in solvers)
                else:
                    for s in solvers:
                        s.minimize()
            return [solver_state(s) for s in solvers], solvers, launches

        materialised, _, _ = run(lambda c: Quadratic(signs[c][:, None] * M * signs[c][None, :], q), False)
        views, _, _ = run(lambda c: Quadratic(shared.with_signs(signs[c]), q), False)
        batched, lock, launches = run(lambda c: Quadratic(shared.with_signs(signs[c]), q), True)
        # the signed view answers Quadratic's own queries like the materialised matrix
        x = rng.random(n)
        quad = Quadratic(shared.with_signs(signs[0]), q)
        Qs = signs[0][:, None] * M * signs[0][None, :]
        assert np.array_equal(quad.Q, Qs)
        assert np.abs(quad.jacobian(x) - (Qs @ x + q)).max() <= 1e-12 * n
        probe.assert_clean()
    for a, b, c in zip(materialised, views, batched):
        for va, vb, vc in zip(a, b, c):
            assert np.array_equal(va, vb), 'signed view differs from the materialised Hessian'
            assert np.array_equal(vb, vc), 'lockstep batch differs from the sequential solves'
    iters = [s.iter for s in lock]
    if kind in ('pg', 'fw'):
        assert iters[1] <= max_iter // 2 and iters[1] < max(iters)   # problem 1 met its stopping test early, others ran on
    # ceil(count / 4) passes over M per iteration + one vector launch for all problems (+ set-up and the final state)
    per_iter = (count + 3) // 4 + 1
    assert launches <= count * 2 + (max_iter + 2) * per_iter


def check_batch_argument_checks(device, **dev_kw):
    from optiml_b200.opti import Quadratic
    from optiml_b200.opti.batch import batchable, minimize_batch
    from optiml_b200.opti.constrained import FrankWolfe, ProjectedGradient
    from optiml_b200.runtime import DeviceHessian, default_context
    rng = np.random.default_rng(2)
    n = 40
    Q, q, ub = psd(rng, n), -np.ones(n), np.ones(n)
    with device(**dev_kw) as probe:
        a, b = Quadratic(Q, q), Quadratic(Q, q)   # two uploads: different resident matrices
        mixed = [ProjectedGradient(quad=a, ub=ub, max_iter=5), ProjectedGradient(quad=b, ub=ub, max_iter=5)]
        assert not batchable(mixed)
        assert not batchable([ProjectedGradient(quad=a, ub=ub, max_iter=5), FrankWolfe(quad=a, ub=ub, max_iter=5)])
        assert not batchable([ProjectedGradient(quad=a, ub=ub, max_iter=5), ProjectedGradient(quad=a, ub=ub, max_iter=6)])
        assert not batchable([ProjectedGradient(quad=a, ub=ub, max_iter=5, verbose=True)] * 2)
        minimize_batch(mixed)   # falls back to one solve after the other
        assert all(s.batch_size_ == 1 and s.iter == 5 for s in mixed)
        # the C entry point refuses what the host check refuses
        ok = [ProjectedGradient(quad=a, ub=ub, max_iter=5), ProjectedGradient(quad=a, ub=ub, max_iter=6)]
        created = [s._create(False) for s in ok]
        try:
            handles = (C.c_void_p * 2)(*[h.value for h, _ in created])
            with pytest.raises(N.NativeError, match='iteration limit'):
                N.call('svmb200_pg_run_batch', handles, 2, None, None)
            twice = (C.c_void_p * 2)(created[0][0].value, created[0][0].value)
            with pytest.raises(N.NativeError, match='twice'):
                N.call('svmb200_pg_run_batch', twice, 2, None, None)
        finally:
            for h, _ in created:
                N.load_library().svmb200_pg_destroy(h)
        # a batch of one is a plain solve
        solo = ProjectedGradient(quad=a, ub=ub, max_iter=7).minimize()
        single = ProjectedGradient(quad=a, ub=ub, max_iter=7)
        h, nvars = single._create(False)
        try:
            it, st = (C.c_int64 * 1)(), (C.c_int * 1)()
            N.call('svmb200_pg_run_batch', (C.c_void_p * 1)(h.value), 1, it, st)
            single._finish_resident(h, nvars, int(it[0]), N.STATUS[int(st[0])])
        finally:
            N.load_library().svmb200_pg_destroy(h)
        assert np.array_equal(single.x, solo.x) and single.iter == solo.iter and np.array_equal(single.f_hist, solo.f_hist)
        with pytest.raises(N.NativeError, match=r'\+-1'):
            H = a.device_hessian()
            bad = DeviceHessian(default_context(), n, 'plain', matrix=H.matrix, row0=H.row0, nrows=H.nrows)
            bad.signs = np.full(n, 0.5)
            ProjectedGradient(quad=Quadratic(bad, q), ub=ub, max_iter=3).minimize()
        with pytest.raises(ValueError, match='plain layout'):
            DeviceHessian(default_context(), n, 'svr', matrix=a.device_hessian().matrix, signs=np.ones(n))
        probe.assert_clean()


# --------------------------------------------------------------------------------------------- meta-estimators
def check_one_vs_rest(device, X, y, Xt, optimizer, max_iter, oracle_tol=None, yt=None, inside=None, **dev_kw):
    """The shared-Gram lockstep fit reproduces sklearn's clone-per-class fit bit for bit (recipe of
    ml/tests/test_svc.py:96-147: one-vs-rest, Gaussian kernel); returns the fitted meta-estimator."""
    from sklearn.multiclass import OneVsRestClassifier as SklearnOVR
    from optiml_b200.ml.multiclass import OneVsRestClassifier
    from optiml_b200.ml.svm import SVC
    from optiml_b200.ml.svm.kernels import gaussian
    from optiml_b200.ml.svm.losses import hinge
    from optiml_b200.opti.constrained import ProjectedGradient
    # the stochastic optimisers draw their start point: seeded, so that both fits start from the same one
    kw = dict(learning_rate=1., random_state=0) if optimizer.__name__ == 'AdaGrad' else {}
    est = SVC(loss=hinge, kernel=gaussian, reg_intercept=True, dual=True, optimizer=optimizer, max_iter=max_iter, **kw)
    classes = np.unique(y)
    with device(**dev_kw) as probe, warnings.catch_warnings():
        warnings.simplefilter('ignore')
        shared = OneVsRestClassifier(est).fit(X, y)
        assert [e.fit_times_['batch'] for e in shared.estimators_] == [len(classes)] * len(classes)
        assert len({id(e.obj.device_hessian().matrix) for e in shared.estimators_}) == 1   # one M for all classes
        cloned = SklearnOVR(est).fit(X, y)
        for a, b in zip(shared.estimators_, cloned.estimators_):
            assert np.array_equal(a.alphas_, b.alphas_) and a.intercept_ == b.intercept_
            assert np.array_equal(a.support_, b.support_) and np.array_equal(a.dual_coef_, b.dual_coef_)
            assert a.train_loss_history == b.train_loss_history and a.optimizer.iter == b.optimizer.iter
            assert a.optimizer.status == b.optimizer.status
        assert np.array_equal(shared.predict(Xt), cloned.predict(Xt))
        assert np.array_equal(shared.decision_function(Xt), cloned.decision_function(Xt))
        assert np.array_equal(shared.classes_, cloned.classes_)
        if oracle_tol is not None and optimizer is ProjectedGradient:
            # against the oracle (= the reference's algorithm) on the same binary problems
            for c, e in zip(classes, shared.estimators_):
                want = O.svc_dual_fit(X, (y == c).astype(int), kind='gaussian', max_iter=max_iter)
                assert np.abs(e.alphas_ - want.alphas_).max() <= oracle_tol
                assert np.array_equal(e.support_, want.support_)
                assert abs(e.intercept_ - want.intercept_) <= oracle_tol
        if yt is not None:
            shared.test_score_ = shared.score(Xt, yt)
        if inside is not None:
            inside(shared)   # further assertions that need the device (decision values, ...)
        probe.assert_clean()
    return shared


def check_multi_output(device, n, max_iter, **dev_kw):
    from sklearn.multioutput import MultiOutputRegressor as SklearnMOR
    from optiml_b200.ml.multiclass import MultiOutputRegressor
    from optiml_b200.ml.svm import SVR
    from optiml_b200.ml.svm.kernels import GaussianKernel
    from optiml_b200.ml.svm.losses import epsilon_insensitive
    from optiml_b200.opti.constrained import FrankWolfe
    rng = np.random.default_rng(4)
    X = rng.standard_normal((n, 3))
    Y = np.stack((np.sin(X[:, 0]) + 0.1 * X[:, 1], X[:, 2] ** 2 - X[:, 0], np.cos(X[:, 1])), axis=1)
    est = SVR(loss=epsilon_insensitive, kernel=GaussianKernel(gamma=0.5), reg_intercept=True, dual=True,
              optimizer=FrankWolfe, max_iter=max_iter, epsilon=0.05)
    with device(**dev_kw) as probe:
        shared = MultiOutputRegressor(est).fit(X, Y)
        cloned = SklearnMOR(est).fit(X, Y)
        for a, b in zip(shared.estimators_, cloned.estimators_):
            assert np.array_equal(a.alphas_, b.alphas_) and a.intercept_ == b.intercept_
            assert a.obj.device_hessian() is shared.estimators_[0].obj.device_hessian()
        assert np.array_equal(shared.predict(X[:7]), cloned.predict(X[:7]))
        # a single target / a foreign estimator go through sklearn's own fit
        one = MultiOutputRegressor(est).fit(X, Y[:, :1])
        assert np.array_equal(one.estimators_[0].alphas_, shared.estimators_[0].alphas_)
        probe.assert_clean()


# --------------------------------------------------------------------------------------------- reference goldens
def check_against_reference_meta_estimators(device, g, max_iter, **dev_kw):
    """tests/golden/shared_gram.npz: the REAL reference wrapped in sklearn's OneVsRestClassifier / MultiOutputRegressor
    (FrankWolfe, 400 iterations).  With max_iter = 400 everything is compared; with fewer iterations the loss
    histories (Frank-Wolfe iterates do not depend on the iteration limit)."""
    from optiml_b200.ml.multiclass import MultiOutputRegressor, OneVsRestClassifier
    from optiml_b200.ml.svm import SVC, SVR
    from optiml_b200.ml.svm.kernels import GaussianKernel
    from optiml_b200.ml.svm.losses import epsilon_insensitive, hinge
    from optiml_b200.opti.constrained import FrankWolfe
    full = max_iter == 400
    with device(**dev_kw) as probe, warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ovr = OneVsRestClassifier(SVC(loss=hinge, kernel=GaussianKernel(), C=1, reg_intercept=True, dual=True,
                                      optimizer=FrankWolfe, max_iter=max_iter)).fit(g['ovr_X_train'], g['ovr_y_train'])
        mor = MultiOutputRegressor(SVR(loss=epsilon_insensitive, epsilon=0.1, kernel=GaussianKernel(), C=1, reg_intercept=True,
                                       dual=True, optimizer=FrankWolfe, max_iter=max_iter)).fit(g['mor_X_train'], g['mor_Y_train'])
        for prefix, model in (('ovr_c', ovr), ('mor_t', mor)):
            for i, e in enumerate(model.estimators_):
                p = f'{prefix}{i}_'
                assert e.fit_times_['batch'] == len(model.estimators_)
                want = g[p + 'f_hist'][:max_iter + 1]
                assert np.abs(np.array(e.train_loss_history) - want).max() <= 1e-9 * max(1., np.abs(want).max())
                if full:
                    assert e.optimizer.iter == int(g[p + 'iter']) and e.optimizer.status == str(g[p + 'status'])
                    assert np.abs(e.alphas_ - g[p + 'alphas']).max() <= 1e-8
                    assert np.array_equal(e.support_, g[p + 'support'])
                    assert abs(e.intercept_ - float(g[p + 'intercept'])) <= 1e-8
        if full:
            assert np.array_equal(ovr.predict(g['ovr_X_test']), g['ovr_predict'])
            assert np.abs(ovr.decision_function(g['ovr_X_test']) - g['ovr_decision']).max() <= 1e-8
            assert np.abs(mor.predict(g['mor_X_test']) - g['mor_predict']).max() <= 1e-8
            assert abs(ovr.score(g['ovr_X_test'], g['ovr_y_test']) - float(g['ovr_score'])) <= 1e-12
        probe.assert_clean()
