"""The solver kernels of csrc/pg.cu executed on the CPU, thread by thread (tests/cuda_emu), in the `not gpu` suite.

The emulated library is compiled from the PRODUCT sources (only the launch syntax is rewritten), so these tests run the
real K2 / K3 code -- indexing, barriers, launch sequences, double-buffering, reduction order -- against the reference's
golden vectors and the oracle without a GPU, and pin the properties the batched one-vs-rest path
rests on:  the multi-vector pass is bit-identical to the single-vector pass;  a solver on the signed view
(s s') o M is bit-identical to a solver on the materialised Q;  a lockstep batch is bit-identical to the solvers run
one after the other.  Timing, PTX and inter-CTA memory ordering are out of the emulation's reach: the `-m gpu` tests
remain the parity tests proper.
"""
import ctypes as C
import warnings

import numpy as np
import pytest

from emu import emulated_device
from oracle import svm_oracle as O
from optiml_b200 import _native as N


def _psd(rng, n, rank=None, shift=1.0):
    G = rng.standard_normal((n, rank or max(3, n // 4)))
    return G @ G.T / G.shape[1] + shift


# --------------------------------------------------------------------------------------------- K2 / K3 vs references
@pytest.mark.parametrize('p', ['p2', 'p5', 'p64'])
def test_projected_gradient_reference_goldens(golden, p):
    """the reference's own BCQP problems (opti/constrained/tests/test_projected_gradient.py:9-12, test_lower_bound.py)"""
    import test_gpu_pg as T
    with emulated_device() as lib:
        T.test_bcqp_golden_stable(golden, p)
        assert lib.emu_sticky_error() == 0


@pytest.mark.parametrize('key,p,t', [('p5', 'p5', 0.), ('p64_t05', 'p64', 0.5)])
def test_frank_wolfe_reference_goldens(golden, key, p, t):
    import test_gpu_frank_wolfe as T
    with emulated_device() as lib:
        T.test_bcqp_golden(golden, key, p, t)
        assert lib.emu_sticky_error() == 0


@pytest.mark.parametrize('order', [1, 2])
def test_thread_schedule_does_not_change_a_bit(order):
    """Resuming the threads of a block in reversed / shuffled order between barriers must not change the result: a
    missing __syncthreads() or a read of a half-written partial would."""
    from optiml_b200.opti import Quadratic
    from optiml_b200.opti.constrained import ProjectedGradient, FrankWolfe
    rng = np.random.default_rng(3)
    n = 131  # three 64-row groups, the last one ragged
    Q, q, ub = _psd(rng, n), rng.standard_normal(n), rng.uniform(0.5, 2., n)
    out = {}
    for o in (0, order):
        with emulated_device(order=o, seed=7) as lib:
            pg = ProjectedGradient(quad=Quadratic(Q, q), ub=ub, max_iter=12).minimize()
            fw = FrankWolfe(quad=Quadratic(Q, q), ub=ub, max_iter=8, t=0.3).minimize()
            assert lib.emu_sticky_error() == 0
            out[o] = (pg.x.copy(), pg.f_hist.copy(), fw.x.copy(), fw.f_hist.copy())
    for a, b in zip(out[0], out[order]):
        assert np.array_equal(a, b)
    want = O.projected_gradient(Q, q, ub, max_iter=12)
    assert np.abs(out[0][0] - want.x).max() <= 1e-12


def test_svr_block_layout_vs_oracle():
    """Q = [[M,-M],[-M,M]] with only M resident (ml/svm/_base.py:1098-1099)"""
    from optiml_b200.opti import Quadratic
    from optiml_b200.opti.constrained import ProjectedGradient
    from optiml_b200.runtime import DeviceHessian, default_context
    rng = np.random.default_rng(11)
    n = 70
    M = _psd(rng, n)
    y = rng.standard_normal(n)
    q = np.hstack((-y, y)) + 0.1
    ub = np.ones(2 * n)
    Qfull = np.vstack((np.hstack((M, -M)), np.hstack((-M, M))))
    want = O.projected_gradient(Qfull, q, ub, max_iter=15)
    with emulated_device() as lib:
        ctx = default_context()
        H = DeviceHessian(ctx, n, 'svr')
        block = np.zeros((n, H.ld))
        block[:, :n] = M
        ctx.h2d(H.matrix.dptr, block)
        got = ProjectedGradient(quad=Quadratic(H, q), ub=ub, max_iter=15).minimize()
        assert lib.emu_sticky_error() == 0
    assert got.iter == want.iter and got.status == want.status
    assert np.abs(got.x - want.x).max() <= 1e-11
    assert np.abs(got.f_hist - want.f_hist).max() <= 1e-10 * np.abs(want.f_hist).max()


# --------------------------------------------------------------------------------------------- K2 x NB
@pytest.mark.parametrize('n,count', [(64, 2), (131, 3), (200, 4), (131, 7)])
def test_multi_vector_pass_is_bit_identical_to_single(n, count):
    from optiml_b200.runtime import default_context
    rng = np.random.default_rng(n + count)
    with emulated_device(order=2, seed=n) as lib:
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
        before = lib.emu_launches()
        N.call('svmb200_matvec_multi', ctx.handle, C.c_void_p(dQ), n, ld, (C.c_void_p * count)(*dus),
               (C.c_void_p * count)(*dws), count)
        assert lib.emu_launches() - before == (count + 3) // 4   # passes over the matrix
        for b in range(count):
            w = np.empty(n)
            ctx.d2h(w, dws[b])
            assert np.array_equal(w, single[b])
            assert np.abs(w - Q[:, :n] @ us[b][:n]).max() <= 1e-12 * n
        for p in [dQ] + dus + dws:
            ctx.free(p)
        assert lib.emu_sticky_error() == 0


# --------------------------------------------------------------------------------------------- signed views, batches
def _solvers(kind, quad_for, signs, ub, max_iter, eps_list):
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


def _state(s):
    hist = [s.f_hist] + ([s.pf_hist, s.f.dual_x] if hasattr(s, 'pf_hist') else [s.ng_hist])
    return [np.asarray(s.x), np.asarray(s.g_x), np.array([s.iter, s.f_x]), np.array([s.status == 'optimal'])] + hist


@pytest.mark.parametrize('kind,count', [('pg', 3), ('pg', 5), ('fw', 2), ('adagrad', 3), ('adam', 2)])
def test_signed_views_and_lockstep_batches_are_bit_identical(kind, count):
    """(1) a solver on the signed view (s s') o M == the same solver on the materialised Q;  (2) the lockstep batch ==
    the solvers run one after the other -- including problems that stop early (a loose eps ends problem 1 long before
    the iteration limit while the others keep going)."""
    from optiml_b200.opti import Quadratic
    from optiml_b200.opti.batch import batchable, minimize_batch
    from optiml_b200.runtime import DeviceHessian, default_context
    rng = np.random.default_rng(count * 17 + len(kind))
    n = 96
    K = _psd(rng, n, shift=0.0)
    M = K + 1.0
    signs = [np.where(rng.random(n) < 0.3 + 0.1 * c, 1.0, -1.0) for c in range(count)]
    q, ub = -np.ones(n), np.ones(n)
    max_iter = 14
    eps_list = [1e-6, 1e-6, 1e-6]
    with emulated_device(order=2, seed=count) as lib:
        ctx = default_context()
        shared = DeviceHessian(ctx, n, 'plain')
        block = np.zeros((n, shared.ld))
        block[:, :n] = M
        ctx.h2d(shared.matrix.dptr, block)
        if kind in ('pg', 'fw'):
            # a stopping threshold for problem 1 that its own criterion (|d| or the gap) crosses mid-run
            probe = _solvers(kind, lambda c: Quadratic(shared.with_signs(signs[c]), q), signs, ub, max_iter, eps_list)[1]
            crit = probe.minimize().ng_hist
            eps_list[1] = float(np.min(crit[:7])) * (1 + 1e-12)

        def run(quad_for, batch):
            solvers = _solvers(kind, quad_for, signs, ub, max_iter, eps_list)
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                if batch:
                    assert batchable(solvers)
                    launches = lib.emu_launches()
                    minimize_batch(solvers)
                    launches = lib.emu_launches() - launches
                    assert all(s.batch_size_ == count for s in solvers)
                else:
                    launches = None
                    for s in solvers:
                        s.minimize()
            return [_state(s) for s in solvers], solvers, launches

        materialised, _, _ = run(lambda c: Quadratic(signs[c][:, None] * M * signs[c][None, :], q), False)
        views, solo, _ = run(lambda c: Quadratic(shared.with_signs(signs[c]), q), False)
        batched, lock, launches = run(lambda c: Quadratic(shared.with_signs(signs[c]), q), True)
        assert lib.emu_sticky_error() == 0
    for a, b, c in zip(materialised, views, batched):
        for va, vb, vc in zip(a, b, c):
            assert np.array_equal(va, vb), 'signed view differs from the materialised Hessian'
            assert np.array_equal(vb, vc), 'lockstep batch differs from the sequential solves'
    iters = [s.iter for s in lock]
    if kind in ('pg', 'fw'):
        assert min(iters) < max(iters) == max_iter   # one problem met its stopping test early, the others ran on
    # passes over M: ceil(count / 4) multi-vector launches per iteration (+ one vector launch for all problems)
    per_iter = (count + 3) // 4 + 1
    assert launches <= count * 2 + (max_iter + 2) * per_iter
    # the signed view answers Quadratic's own queries like the materialised matrix
    x = rng.random(n)
    with emulated_device():
        ctx = default_context()
        shared = DeviceHessian(ctx, n, 'plain')
        block = np.zeros((n, shared.ld))
        block[:, :n] = M
        ctx.h2d(shared.matrix.dptr, block)
        quad = Quadratic(shared.with_signs(signs[0]), q)
        Qs = signs[0][:, None] * M * signs[0][None, :]
        assert np.array_equal(quad.Q, Qs)
        assert np.abs(quad.jacobian(x) - (Qs @ x + q)).max() <= 1e-12 * n


def test_batch_argument_checks():
    from optiml_b200.opti import Quadratic
    from optiml_b200.opti.batch import batchable, minimize_batch
    from optiml_b200.opti.constrained import FrankWolfe, ProjectedGradient
    rng = np.random.default_rng(2)
    n = 40
    Q, q, ub = _psd(rng, n), -np.ones(n), np.ones(n)
    with emulated_device() as lib:
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
                lib.svmb200_pg_destroy(h)
        with pytest.raises(N.NativeError, match=r'\+-1'):
            from optiml_b200.runtime import DeviceHessian, default_context
            H = a.device_hessian()
            bad = DeviceHessian(default_context(), n, 'plain', matrix=H.matrix, row0=H.row0, nrows=H.nrows)
            bad.signs = np.full(n, 0.5)
            ProjectedGradient(quad=Quadratic(bad, q), ub=ub, max_iter=3).minimize()
        assert lib.emu_sticky_error() == 0


# --------------------------------------------------------------------------------------------- meta-estimators
def test_one_vs_rest_shares_the_gram_matrix_and_matches_sklearn_clones(golden):
    """ml/tests/test_svc.py:96-103 recipe (iris, one-vs-rest, Gaussian kernel, ProjectedGradient): the shared-Gram
    lockstep fit reproduces sklearn's clone-per-class fit bit for bit, and the per-class problems match the
    reference's own alphas."""
    from sklearn.multiclass import OneVsRestClassifier as SklearnOVR
    from optiml_b200.ml.multiclass import OneVsRestClassifier
    from optiml_b200.ml.svm import SVC
    from optiml_b200.ml.svm.kernels import gaussian
    from optiml_b200.ml.svm.losses import hinge
    from optiml_b200.opti.constrained import ProjectedGradient
    iris = golden('iris_ovr')
    X, y, Xt, yt = iris['X_train'], iris['y_train'], iris['X_test'], iris['y_test']
    est = SVC(loss=hinge, kernel=gaussian, reg_intercept=True, dual=True, optimizer=ProjectedGradient, max_iter=40)
    with emulated_device() as lib:
        grams = lib.emu_launches()
        shared = OneVsRestClassifier(est).fit(X, y)
        assert [e.fit_times_['batch'] for e in shared.estimators_] == [3, 3, 3]
        hessians = {id(e.obj.device_hessian().matrix) for e in shared.estimators_}
        assert len(hessians) == 1                                  # one M in "HBM" for the three classes
        cloned = SklearnOVR(est).fit(X, y)
        for a, b in zip(shared.estimators_, cloned.estimators_):
            assert np.array_equal(a.alphas_, b.alphas_) and a.intercept_ == b.intercept_
            assert np.array_equal(a.support_, b.support_) and np.array_equal(a.dual_coef_, b.dual_coef_)
            assert a.train_loss_history == b.train_loss_history and a.optimizer.iter == b.optimizer.iter == 40
        assert np.array_equal(shared.predict(Xt), cloned.predict(Xt))
        assert np.array_equal(shared.classes_, cloned.classes_)
        # against the oracle (= the reference's algorithm) on the same binary problems
        for c, e in enumerate(shared.estimators_):
            want = O.svc_dual_fit(X, (y == c).astype(int), kind='gaussian', max_iter=40)
            assert np.abs(e.alphas_ - want.alphas_).max() <= 1e-9
            assert np.array_equal(e.support_, want.support_)
            assert abs(e.intercept_ - want.intercept_) <= 1e-9
        assert lib.emu_sticky_error() == 0
    # two classes / foreign estimators go through sklearn's own fit
    from sklearn.linear_model import LogisticRegression
    assert OneVsRestClassifier(LogisticRegression()).fit(X, y).score(Xt, yt) > 0.8


def test_multi_output_regressor_shares_the_gram_matrix():
    from sklearn.multioutput import MultiOutputRegressor as SklearnMOR
    from optiml_b200.ml.multiclass import MultiOutputRegressor
    from optiml_b200.ml.svm import SVR
    from optiml_b200.ml.svm.kernels import GaussianKernel
    from optiml_b200.ml.svm.losses import epsilon_insensitive
    from optiml_b200.opti.constrained import FrankWolfe
    rng = np.random.default_rng(4)
    X = rng.standard_normal((60, 3))
    Y = np.stack((np.sin(X[:, 0]) + 0.1 * X[:, 1], X[:, 2] ** 2 - X[:, 0]), axis=1)
    est = SVR(loss=epsilon_insensitive, kernel=GaussianKernel(gamma=0.5), reg_intercept=True, dual=True,
              optimizer=FrankWolfe, max_iter=25, epsilon=0.05)
    with emulated_device() as lib:
        shared = MultiOutputRegressor(est).fit(X, Y)
        cloned = SklearnMOR(est).fit(X, Y)
        for a, b in zip(shared.estimators_, cloned.estimators_):
            assert np.array_equal(a.alphas_, b.alphas_) and a.intercept_ == b.intercept_
        assert a.obj.device_hessian() is shared.estimators_[0].obj.device_hessian()
        assert np.array_equal(shared.predict(X[:7]), cloned.predict(X[:7]))
        assert lib.emu_sticky_error() == 0


def test_emulated_device_leaves_no_allocation_behind():
    from optiml_b200.opti import Quadratic
    from optiml_b200.opti.constrained import ProjectedGradient
    import gc
    rng = np.random.default_rng(9)
    Q = _psd(rng, 30)
    with emulated_device() as lib:
        base = lib.emu_live_allocations()
        s = ProjectedGradient(quad=Quadratic(Q, -np.ones(30)), ub=np.ones(30), max_iter=3).minimize()
        s.f.release()
        del s
        gc.collect()
    assert lib.emu_live_allocations() <= base   # context, workspace, scratch, matrices: all returned, canaries intact
    assert lib.emu_sticky_error() == 0
