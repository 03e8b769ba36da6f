"""The kernels of csrc/gram.cu and csrc/pg.cu executed on the CPU, thread by thread (tests/cuda_emu), in the `not gpu` suite.

The emulated library is compiled from the PRODUCT sources (only the launch syntax is rewritten), so these tests run the
real K1 / K2 / K3 code -- indexing, barriers, launch sequences, double-buffering, reduction order -- against the reference's
golden vectors and the oracle without a GPU, and pin the properties the batched one-vs-rest path
rests on:  the multi-vector pass is bit-identical to the single-vector pass;  a solver on the signed view
(s s') o M is bit-identical to a solver on the materialised Q;  a lockstep batch is bit-identical to the solvers run
one after the other.  Timing, PTX and inter-CTA memory ordering are out of the emulation's reach: the `-m gpu` tests
remain the parity tests proper.
"""
import contextlib
import ctypes as C

import numpy as np
import pytest

import shared_gram_checks as S
from emu import emulated_device
from oracle import svm_oracle as O
from optiml_b200 import _native as N


_psd = S.psd


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


@pytest.mark.filterwarnings('ignore::sklearn.exceptions.ConvergenceWarning')
@pytest.mark.parametrize('which', ['svc_adam_nesterov_equality_row', 'svr_adadelta_blocks', 'optimality_exit'])
def test_augmented_lagrangian_reference_goldens(golden, which):
    """Whole estimator fits through the real kernels (K1 + K2 + al_vector_kernel) against the REAL reference's runs
    (tests/golden/al_stochastic.npz): alphas, multipliers, histories, support set, intercept, decision values to 1e-8"""
    import test_gpu_al as T
    with emulated_device() as lib:
        if which == 'svc_adam_nesterov_equality_row':
            T.test_svc_matches_reference(golden, 'adam_nesterov', False, 1)   # reg_intercept=False: y'alpha = 0 relaxed
        elif which == 'svr_adadelta_blocks':
            T.test_svr_matches_reference(golden, 'gaussian', 'adadelta', True)
        else:
            T.test_optimality_exit_matches_reference(golden, 'sgd_nesterov', 0.1)
        assert lib.emu_sticky_error() == 0


@pytest.mark.parametrize('order', [1, 2])
def test_thread_schedule_does_not_change_a_bit(order):
    """Resuming the threads of a block in reversed / shuffled order between barriers, and running the blocks of a
    launch in reversed / shuffled order, must not change the result: a missing __syncthreads(), a read of a
    half-written partial, or a reduction whose shape follows the arrival order at a ticket would."""
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


# --------------------------------------------------------------------------------------------- K1 (csrc/gram.cu)
# The Gram kernel itself -- TMA producers, mbarrier rings, the m8n8k4 FP64 tensor-core contraction, the fused
# epilogues -- runs on the emulation from the product source (its PTX wrappers have emulation twins).
@pytest.mark.parametrize('name', ['linear', 'poly_d3_scale', 'poly_d4_g05_c05', 'gauss_scale', 'gauss_g03'])
def test_gram_kernel_reference_goldens(golden, name):
    """kernels.py:49-51, 91-95, 125-129 of the REAL reference (tests/golden/kernels.npz): 1e-12 relative"""
    import test_gpu_kernels as T
    with emulated_device() as lib:
        T.test_kernel_golden(golden, name)
        assert lib.emu_sticky_error() == 0


@pytest.mark.parametrize('name', ['lap_scale', 'sig_g01_c05'])
def test_extra_kernels_reference_goldens(golden, name):
    import test_gpu_kernels as T
    with emulated_device() as lib:
        T.test_extra_kernel_golden(golden, name)
        assert lib.emu_sticky_error() == 0


@pytest.mark.parametrize('shape', [(1, 1, 1), (2, 3, 1), (129, 127, 17), (257, 1, 33), (128, 128, 15)])
def test_gram_kernel_ragged_shapes_vs_oracle(shape):
    import test_gpu_kernels as T
    with emulated_device(order=2, seed=shape[0]) as lib:
        for name in ('linear', 'poly_d3_scale', 'gauss_scale'):
            T.test_kernel_ragged_shapes_vs_oracle(shape, name)
        T.test_kernel_input_validation()
        assert lib.emu_sticky_error() == 0


def test_gram_kernel_schedules_and_sharding_give_the_same_bits(monkeypatch):
    """The Hessian Q = s_i s_j (k(x_i, x_j) + 1) built (a) with the two consumer groups in lockstep, ping-pong and
    free-running (SVMB200_GRAM_EXCLUSIVE = 2 / 1 / 0), (b) under shuffled thread and block schedules, (c) in row shards:
    always the same bits, and the oracle's values."""
    from optiml_b200.ml.svm.kernels import GaussianKernel
    from optiml_b200.runtime import DeviceHessian, default_context
    rng = np.random.default_rng(8)
    n, d = 200, 37
    X = rng.standard_normal((n, d))
    signs = np.where(rng.random(n) < 0.5, 1.0, -1.0)
    kid, gamma, coef0, degree = GaussianKernel().gram_spec(X)

    def build(row0, nrows, order):
        with emulated_device(order=order, seed=3) as lib:
            ctx = default_context()
            dX, dS = ctx.upload_matrix(X), ctx.upload_vector(signs)
            H = DeviceHessian(ctx, n, 'plain', row0=row0, nrows=nrows)
            N.call('svmb200_gram', ctx.handle, C.c_void_p(dX.dptr), n, dX.ld, C.c_void_p(dX.dptr), n, dX.ld, d, 1, kid, gamma,
                   coef0, degree, C.c_void_p(dS.dptr), C.c_void_p(dS.dptr), 1.0, row0, nrows, C.c_void_p(H.matrix.dptr), H.ld)
            out = H.shard_to_host()
            assert lib.emu_sticky_error() == 0
        return out

    full = build(0, n, 0)
    want = signs[:, None] * (O.gaussian_kernel(X) + 1.0) * signs[None, :]
    assert np.abs(full - want).max() <= 1e-12
    for mode, order in (('2', 2), ('1', 0), ('1', 2), ('0', 1)):
        monkeypatch.setenv('SVMB200_GRAM_EXCLUSIVE', mode)
        assert np.array_equal(build(0, n, order), full), (mode, order)
    monkeypatch.delenv('SVMB200_GRAM_EXCLUSIVE')
    assert np.array_equal(np.vstack((build(0, 64, 2), build(64, 128, 0), build(192, 8, 1))), full)


def test_gram_kernel_random_shapes_signs_shards_and_modes(monkeypatch):
    """Seeded fuzz of svmb200_gram on the emulation: sizes around the 128 x 64 tile and the 16-wide k chunk, all five
    kernels, label signs on either side, bias, self / cross Gram, arbitrary row ranges (empty ones too), the three
    consumer-group modes, shuffled schedules.  Against the oracle to 1e-12; pad columns must be exact zeros."""
    from optiml_b200.runtime import default_context
    kinds = {0: 'linear', 1: 'poly', 2: 'gaussian', 3: 'sigmoid', 4: 'laplacian'}
    for seed in range(80):
        rng = np.random.default_rng(9000 + seed)
        na = int(rng.choice([1, 2, 7, 63, 64, 65, 127, 128, 129, 200, 260]))
        same = bool(rng.random() < 0.5)
        nb = na if same else int(rng.choice([1, 3, 31, 64, 65, 130]))
        d = int(rng.choice([1, 2, 3, 15, 16, 17, 32, 33, 50]))
        kid = int(rng.integers(0, 5))
        gamma, coef0 = float(rng.uniform(0.01, 0.5)), float(rng.choice([0., 0.5, 1.]))
        degree, bias = float(rng.choice([1, 2, 3, 4])), float(rng.choice([0., 1.]))
        use_sa, use_sb = bool(rng.random() < 0.5), bool(rng.random() < 0.5)
        row0 = int(rng.integers(0, na))
        nrows = int(rng.integers(0, na - row0 + 1))
        monkeypatch.setenv('SVMB200_GRAM_EXCLUSIVE', str(rng.choice(['0', '1', '2'])))
        A = rng.standard_normal((na, d))
        B = A if same else rng.standard_normal((nb, d))
        sa = np.where(rng.random(na) < 0.5, 1., -1.)
        sb = sa if same else np.where(rng.random(nb) < 0.5, 1., -1.)
        label = dict(seed=seed, na=na, nb=nb, d=d, kind=kinds[kid], same=same, row0=row0, nrows=nrows)
        with emulated_device(order=2, seed=seed) as lib:
            ctx = default_context()
            dA = ctx.upload_matrix(A)
            dB = dA if same else ctx.upload_matrix(B)
            dsa = ctx.upload_vector(sa) if use_sa else None
            dsb = ctx.upload_vector(sb) if use_sb else None
            ldo = N.padded_ld(nb)
            nbytes = 8 * max(nrows, 1) * ldo
            dout = ctx.malloc(nbytes)
            ctx.memset(dout, 0xFF, nbytes)
            N.call('svmb200_gram', ctx.handle, C.c_void_p(dA.dptr), na, dA.ld, C.c_void_p(dB.dptr), nb, dB.ld, d, int(same), kid,
                   gamma, coef0, degree, C.c_void_p(dsa.dptr) if dsa else None, C.c_void_p(dsb.dptr) if dsb else None, bias,
                   row0, nrows, C.c_void_p(dout), ldo)
            out = np.empty((max(nrows, 1), ldo))
            ctx.d2h(out, dout)
            ctx.free(dout)
            assert lib.emu_sticky_error() == 0, label
        if nrows == 0:
            continue
        kw = {} if kid == 0 else dict(gamma=gamma)
        if kid in (1, 3):
            kw['coef0'] = coef0
        if kid == 1:
            kw['degree'] = degree
        K = O.kernel_matrix(kinds[kid], A, None if same else B, **kw)
        want = (K + bias) * (sa[:, None] if use_sa else 1.0) * (sb[None, :] if use_sb else 1.0)
        assert np.abs(out[:nrows, :nb] - want[row0:row0 + nrows]).max() <= 1e-12 * max(1., np.abs(want).max()), label
        assert np.all(out[:nrows, nb:] == 0.0), label


# --------------------------------------------------------------------------------------------- shared-Gram path
# (the same checks run on the B200 in tests/test_gpu_shared_gram.py)
@contextlib.contextmanager
def emu_probe(order=0, seed=1, defines=()):
    with emulated_device(order=order, seed=seed, defines=defines) as lib:
        class Probe:
            def launches(self):
                return lib.emu_launches()

            def assert_clean(self):
                assert lib.emu_sticky_error() == 0

        yield Probe()


@pytest.mark.parametrize('n,count', [(64, 2), (131, 3), (200, 4), (131, 7)])
def test_multi_vector_pass_is_bit_identical_to_single(n, count):
    S.check_multi_vector_pass(emu_probe, n, count, order=2, seed=n)


@pytest.mark.parametrize('kind,count', [('pg', 3), ('pg', 5), ('fw', 2), ('adagrad', 3), ('adam', 2)])
def test_signed_views_and_lockstep_batches_are_bit_identical(kind, count):
    S.check_signed_views_and_batches(emu_probe, kind, count, n=96, max_iter=14, order=2, seed=count)


@pytest.mark.parametrize('defines', [('SVMB200_MULTI_R=2', 'SVMB200_MULTI_U=4'), ('SVMB200_MULTI_H=2',),
                                     ('SVMB200_MULTI_H=2', 'SVMB200_MULTI_R=8', 'SVMB200_MULTI_U=1')])
def test_every_shape_of_the_multi_vector_pass_gives_the_same_bits(defines):
    """rows per work item (R), loads in flight (U) and thread groups per CTA (H) are tuning knobs (scripts/sweep_multi.py):
    none of them may change a bit"""
    S.check_multi_vector_pass(emu_probe, 131, 4, order=2, seed=5, defines=defines)
    S.check_multi_vector_pass(emu_probe, 200, 3, order=1, defines=defines)
    S.check_signed_views_and_batches(emu_probe, 'pg', 4, n=70, max_iter=6, defines=defines)


def test_batch_argument_checks():
    S.check_batch_argument_checks(emu_probe)


def test_one_vs_rest_shares_the_gram_matrix_and_matches_sklearn_clones(golden):
    """ml/tests/test_svc.py:96-103 recipe on iris; the per-class problems also match the oracle"""
    from optiml_b200.opti.constrained import ProjectedGradient
    iris = golden('iris_ovr')
    ovr = S.check_one_vs_rest(emu_probe, iris['X_train'], iris['y_train'], iris['X_test'], ProjectedGradient, max_iter=40,
                              oracle_tol=1e-9, yt=iris['y_test'])
    assert ovr.test_score_ >= 0.9   # 40 of the recipe's 1000 iterations
    # foreign estimators go through sklearn's own fit
    from sklearn.linear_model import LogisticRegression
    from optiml_b200.ml.multiclass import OneVsRestClassifier
    assert OneVsRestClassifier(LogisticRegression()).fit(iris['X_train'], iris['y_train']).score(
        iris['X_test'], iris['y_test']) > 0.8


def test_one_vs_rest_augmented_lagrangian_batch(golden):
    """ml/tests/test_svc.py:134-147 recipe (AdaGrad on the augmented-Lagrangian dual), shortened"""
    from optiml_b200.opti.unconstrained.stochastic import AdaGrad
    iris = golden('iris_ovr')
    S.check_one_vs_rest(emu_probe, iris['X_train'][:80], iris['y_train'][:80], iris['X_test'], AdaGrad, max_iter=12)


def test_meta_estimators_follow_the_reference_wrapped_in_sklearn(golden):
    """the first 30 Frank-Wolfe iterations of every class / target against the real reference's loss histories"""
    S.check_against_reference_meta_estimators(emu_probe, golden('shared_gram'), max_iter=30)


def test_multi_output_regressor_shares_the_gram_matrix():
    S.check_multi_output(emu_probe, n=60, max_iter=25)


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


def test_one_vs_rest_host_contract_labels_clone_pickle():
    """sklearn's contract around the shared-Gram fit: string labels, a multilabel indicator target, two classes (a single
    binary problem: sklearn's own path), clone / get_params, pickling the fitted meta-estimator without device handles"""
    import pickle
    from sklearn.base import clone
    from optiml_b200.ml.multiclass import OneVsRestClassifier
    from optiml_b200.ml.svm import SVC
    from optiml_b200.ml.svm.kernels import GaussianKernel
    from optiml_b200.ml.svm.losses import hinge
    from optiml_b200.opti.constrained import FrankWolfe
    rng = np.random.default_rng(21)
    X = rng.standard_normal((90, 4))
    centres = rng.standard_normal((3, 4)) * 2
    idx = rng.integers(0, 3, 90)
    X += centres[idx]
    names = np.array(['setosa', 'versicolor', 'virginica'])[idx]
    est = SVC(loss=hinge, kernel=GaussianKernel(gamma=0.5), reg_intercept=True, dual=True, optimizer=FrankWolfe, max_iter=15)
    with emulated_device() as lib:
        ovr = OneVsRestClassifier(est).fit(X, names)
        assert list(ovr.classes_) == ['setosa', 'versicolor', 'virginica'] and len(ovr.estimators_) == 3
        assert [e.fit_times_['batch'] for e in ovr.estimators_] == [3, 3, 3]
        pred = ovr.predict(X)
        assert pred.dtype.kind == 'U' and (pred == names).mean() > 0.8
        # multilabel indicator: one binary problem per column, same lockstep fit
        Y = np.stack((idx == 0, idx != 1, idx == 2), axis=1).astype(int)
        multi = OneVsRestClassifier(est).fit(X, Y)
        assert multi.multilabel_ and multi.predict(X).shape == Y.shape
        assert np.array_equal(multi.estimators_[0].alphas_, ovr.estimators_[0].alphas_)   # same column, same problem
        # two classes: sklearn fits ONE estimator through its own path (nothing to share)
        two = OneVsRestClassifier(est).fit(X, (idx == 0).astype(int))
        assert len(two.estimators_) == 1 and 'batch' not in two.estimators_[0].fit_times_
        assert np.array_equal(two.estimators_[0].alphas_, ovr.estimators_[0].alphas_)
        # clone / params / pickle
        again = clone(ovr)
        assert again.get_params()['estimator__max_iter'] == 15 and not hasattr(again, 'estimators_')
        blob = pickle.dumps(ovr)
        assert len(blob) < 200_000   # no n x n matrix, no device handle
        back = pickle.loads(blob)
        assert np.array_equal(back.predict(X), pred)
        assert lib.emu_sticky_error() == 0


def test_device_variance_is_bit_identical_to_numpy():
    import device_path_checks as D
    with emulated_device():
        D.check_device_variance([(1, 1), (3, 2), (7, 1), (1, 9), (129, 1), (300, 7), (257, 16), (1000, 33), (2049, 5), (40, 40)])


def test_fit_keeps_the_host_passes_off_the_path():
    import device_path_checks as D
    with emulated_device():
        D.check_fit_keeps_host_passes_off_the_path(n=150, d=5)


def test_gram_tile_bands_cover_every_tile_once():
    """K1 enumerates its tiles in bands of 32 tile rows (L2 reuse of X); more than one band, a ragged last band and a row
    shard that starts inside a band must still produce every output exactly once"""
    from optiml_b200.runtime import default_context
    with emulated_device():
        ctx = default_context()
        rng = np.random.default_rng(0)
        n, d = 4230, 3                      # 34 tile rows: one full band + a band of two
        X = rng.standard_normal((n, d))
        want = X @ X.T
        dX = ctx.upload_matrix(X)
        ld = N.padded_ld(n)
        for r0, nr in ((0, n), (100, 4120)):
            dQ = ctx.malloc(nr * ld * 8)
            ctx.memset(dQ, 0xFF, nr * ld * 8)
            N.call('svmb200_gram', ctx.handle, C.c_void_p(dX.dptr), n, dX.ld, C.c_void_p(dX.dptr), n, dX.ld, d, 1,
                   N.KERNEL_LINEAR, 1.0, 0., 1., None, None, 0.0, r0, nr, C.c_void_p(dQ), ld)
            out = np.empty((nr, ld))
            ctx.d2h(out, dQ)
            assert np.abs(out[:, :n] - want[r0:r0 + nr]).max() <= 1e-14 and np.all(out[:, n:] == 0)
            ctx.free(dQ)
        dX.release()


@pytest.mark.parametrize('n,grid,layout,order', [(70, 5, 'plain', 0), (600, 43, 'plain', 2), (1100, 79, 'plain', 1), (150, 11, 'svr', 1),
                                                 (200, 15, 'plain', 0), (45, 4, 'plain', 0)])
def test_persistent_loop_is_bit_identical_to_the_two_kernel_loop(monkeypatch, n, grid, layout, order):
    """k_persistent.cuh: the whole projected-gradient solve as ONE cooperative launch with the matrix in shared memory
    (what BASELINE config C1 runs on a B200) against the K2 + K3 launch pairs -- same iterates, histories, stopping
    iteration, with two and three virtual vector CTAs (n = 600, 1100), a run that reaches optimality (n = 200), under
    reversed / shuffled thread schedules; the SVR block layout is outside its scope and must take the two-kernel loop"""
    from optiml_b200.opti import Quadratic
    from optiml_b200.opti.constrained import ProjectedGradient
    from optiml_b200.runtime import DeviceHessian, default_context
    rng = np.random.default_rng(n)
    G = rng.standard_normal((n + 5, n))
    M = G.T @ G / n
    nv = 2 * n if layout == 'svr' else n
    q, ub = rng.standard_normal(nv), np.full(nv, 1.5)
    iters = 300 if n == 200 else 30

    def solve(persistent_grid):
        monkeypatch.setenv('SVMB200_PERSISTENT_GRID', str(persistent_grid))
        with emulated_device(order=order, seed=3) as lib:
            ctx = default_context()
            H = DeviceHessian(ctx, n, layout)
            block = np.zeros((n, H.ld))
            block[:, :n] = M
            ctx.h2d(H.matrix.dptr, block)
            before = lib.emu_launches()
            o = ProjectedGradient(quad=Quadratic(H, q), ub=ub, max_iter=iters).minimize()
            launches = lib.emu_launches() - before
            H.release()
            return o.x, o.g_x, o.f_hist, o.ng_hist, o.iter, o.status, launches

    two_kernel, persistent = solve(0), solve(grid)
    if layout == 'svr':
        assert persistent[6] == two_kernel[6]                    # fell back: same launches
    else:
        assert persistent[6] <= 4 < two_kernel[6]                # product + INIT, the loop, (FINALISE)
    assert persistent[4:6] == two_kernel[4:6]
    if n == 200:
        assert persistent[5] == 'optimal' and persistent[4] < iters
    for a, b in zip(two_kernel[:4], persistent[:4]):
        assert np.array_equal(a, b)


# --------------------------------------------------------------------------------------------- symmetric pass (K2s)
# (the same checks run on the B200 in tests/test_gpu_symmetric.py with the shipped tile shape)
import symv_checks as SY   # noqa: E402

# small tile shapes so that a few hundred rows cover several bands, ragged last bands and narrow last panels:
#   BH = TR * NRB rows per band, BW = 512 * NCH columns per panel
_SYMV_SMALL = ('SVMB200_SYMV_TR=4', 'SVMB200_SYMV_NRB=2', 'SVMB200_SYMV_NCH=1', 'SVMB200_SYMV_LB=2',
               'SVMB200_SYMV_STAGES=2')                                                                    # 8 x 512
_SYMV_TALL = ('SVMB200_SYMV_TR=16', 'SVMB200_SYMV_NRB=4', 'SVMB200_SYMV_NCH=2', 'SVMB200_SYMV_LB=8')      # 64 x 1024
_SYMV_WIDE32 = ('SVMB200_SYMV_TR=32', 'SVMB200_SYMV_NRB=2', 'SVMB200_SYMV_NCH=1', 'SVMB200_SYMV_LB=16',
                'SVMB200_SYMV_STAGES=2')                                                                   # 64 x 512


@pytest.mark.parametrize('n,defines', [(70, _SYMV_SMALL), (600, _SYMV_SMALL), (1100, _SYMV_TALL), (200, _SYMV_WIDE32),
                                       (643, _SYMV_WIDE32), (130, ())])
def test_symmetric_pass_product(n, defines):
    w0 = SY.check_symmetric_product(emu_probe, n, defines=defines)
    w1 = SY.check_symmetric_product(emu_probe, n, defines=defines, order=2, seed=n)   # other thread / block schedule
    assert np.array_equal(w0, w1)


@pytest.mark.parametrize('n,bh,defines', [(333, 8, _SYMV_SMALL), (700, 64, _SYMV_TALL), (300, 128, ())])
def test_symmetric_pass_never_reads_below_the_diagonal_blocks(n, bh, defines):
    SY.check_lower_triangle_is_never_read(emu_probe, n, bh, defines=defines, order=1)


def test_symmetric_pass_solvers_follow_the_default_pass():
    SY.check_symmetric_solves(emu_probe, n=150, max_iter=25, defines=_SYMV_SMALL)
    SY.check_symmetric_solves(emu_probe, n=90, max_iter=12, defines=_SYMV_TALL, order=2, seed=4)


def test_symmetric_pass_estimators():
    SY.check_symmetric_fit(emu_probe, n=120, defines=_SYMV_SMALL)


_SYMV_FEW_SLOTS = _SYMV_TALL + ('SVMB200_SYMV_PLAN_SLOTS=5', 'SVMB200_SYMV_PLAN_OVERHEAD=64.0')   # the planner balances for 5
# resident CTAs and a start-up cost scaled down with the tiles: tail grading pays at test sizes


def test_symmetric_pass_with_short_bands_and_cut_panels():
    """With few resident CTAs the plan ends a shard on short bands (a quarter of the height) and cuts panels into halves:
    same product, bit for bit reproducible under another schedule, lower triangle still unread"""
    import ctypes as C
    from optiml_b200 import _native as N
    n = 1100
    with emulated_device(defines=_SYMV_FEW_SLOTS):
        bands, short, items, ratio = C.c_int64(), C.c_int64(), C.c_int64(), C.c_double()
        N.call('svmb200_symv_plan_info', n, N.padded_ld(n), 0, 1, 148, C.byref(bands), C.byref(short), C.byref(items), C.byref(ratio))
        assert short.value >= 3 and bands.value > -(-n // 64) and ratio.value < 1.15
    w0 = SY.check_symmetric_product(emu_probe, n, defines=_SYMV_FEW_SLOTS)
    w1 = SY.check_symmetric_product(emu_probe, n, defines=_SYMV_FEW_SLOTS, order=2, seed=3)
    assert np.array_equal(w0, w1)
    SY.check_lower_triangle_is_never_read(emu_probe, 700, 64, defines=_SYMV_FEW_SLOTS)
    SY.check_symmetric_solves(emu_probe, n=300, max_iter=10, defines=_SYMV_FEW_SLOTS)


def test_symmetric_switches(monkeypatch):
    """the three ways to opt in: runtime.use_symmetric_pass (contexts that exist and contexts created later), the
    environment variable read when a context is created, the C entry point; lockstep batches keep the full pass"""
    import ctypes as C
    from optiml_b200 import _native as N, runtime
    from optiml_b200.opti import Quadratic
    from optiml_b200.opti.batch import minimize_batch
    from optiml_b200.opti.constrained import ProjectedGradient

    def state(ctx):
        on = C.c_int(-1)
        N.call('svmb200_ctx_get_symmetric', ctx.handle, C.byref(on))
        return on.value

    saved = runtime._symmetric
    try:
        with emulated_device():
            runtime._symmetric = None
            monkeypatch.delenv('SVMB200_SYMMETRIC', raising=False)
            assert state(runtime.Context(device=0)) == 0                 # off by default
            monkeypatch.setenv('SVMB200_SYMMETRIC', '1')
            assert state(runtime.Context(device=0)) == 1                 # environment variable, read at creation
            monkeypatch.delenv('SVMB200_SYMMETRIC')
            ctx = runtime.default_context()
            assert state(ctx) == 0
            runtime.use_symmetric_pass()                                 # the context that exists ...
            assert state(ctx) == 1 and state(runtime.Context(device=0)) == 1   # ... and those created later
            rng = np.random.default_rng(2)
            Q, q, ub = _psd(rng, 80), rng.standard_normal(80), np.ones(80)
            shared = S.upload_shared(Q)
            signs = [np.where(rng.random(80) < 0.5, 1.0, -1.0) for _ in range(3)]
            solvers = [ProjectedGradient(quad=Quadratic(shared.with_signs(s), q), ub=ub, max_iter=5) for s in signs]
            minimize_batch(solvers)
            assert all(s.batch_size_ == 3 and not s.symmetric_pass for s in solvers)   # batches: the multi-vector full pass
            assert state(ctx) == 1                                       # ... and the switch is back afterwards
            single = ProjectedGradient(quad=Quadratic(Q, q), ub=ub, max_iter=5)
            single.profile = True
            assert single.minimize().symmetric_pass
            runtime.use_symmetric_pass(False)
            assert state(ctx) == 0
            for s in solvers + [single]:
                s.f.release()
            shared.release()
    finally:
        runtime._symmetric = saved
