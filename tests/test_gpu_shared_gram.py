"""Widening step 4 on the B200: problems that share one resident Gram matrix -- the multi-vector
streaming pass, signed Hessian views, lockstep batches and the one-vs-rest / multi-target meta-estimators.

The checks are the ones the CPU suite runs on the host emulation of the kernels (tests/shared_gram_checks.py), at
sizes that need several column segments and row groups, plus the reference's one-vs-rest recipes against its goldens.

First hardware runs: the driver's round-1 GPU test pass and profiles/r2_s1_pytest_gpu.log (all 15 green).
"""
import warnings

import numpy as np
import pytest

import shared_gram_checks as S

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('n,count', [(1000, 2), (8200, 3), (9001, 4), (2500, 7)])
def test_multi_vector_pass_is_bit_identical_to_single(n, count):
    S.check_multi_vector_pass(S.real_device, n, count)   # 9001 columns: two segments, ragged last row group


@pytest.mark.parametrize('kind,count,n,max_iter', [('pg', 3, 1500, 200), ('pg', 6, 700, 120), ('fw', 2, 1500, 150),
                                                   ('adagrad', 3, 900, 150), ('adam', 4, 900, 100)])
def test_signed_views_and_lockstep_batches_are_bit_identical(kind, count, n, max_iter):
    S.check_signed_views_and_batches(S.real_device, kind, count, n=n, max_iter=max_iter)


def test_batch_argument_checks():
    S.check_batch_argument_checks(S.real_device)


def test_one_vs_rest_iris_against_reference_goldens(golden):
    """ml/tests/test_svc.py:96-103: OVR(SVC(hinge, gaussian, reg_intercept=True, dual=True, ProjectedGradient)) on
    iris; every binary problem against the reference's own run (goldens), with the criteria these chaotic trajectories
    allow; bitwise against the clone-per-class fit on the same GPU."""
    from optiml_b200.opti.constrained import ProjectedGradient
    iris = golden('iris_ovr')

    def against_goldens(ovr):
        # these PG trajectories are chaotic: same criteria as test_gpu_estimators.py::
        # test_iris_ovr_binary_problems -- loss-history prefix, status, optimum level, support set, predictions
        for c, e in enumerate(ovr.estimators_):
            p = f'c{c}_'
            fh, gh = np.array(e.train_loss_history), iris[p + 'f_hist']
            assert np.abs(fh[:100] - gh[:100]).max() <= 1e-9
            assert e.optimizer.status == str(iris[p + 'status'])
            assert abs(e.optimizer.f_x - float(iris[p + 'f_x'])) <= (1e-3 if c == 0 else 1e-9)
            assert np.array_equal(e.support_, iris[p + 'support'])
            assert np.array_equal(e.predict(iris['X_test']), iris[p + 'predict'])

    ovr = S.check_one_vs_rest(S.real_device, iris['X_train'], iris['y_train'], iris['X_test'], ProjectedGradient,
                              max_iter=1000, yt=iris['y_test'], inside=against_goldens)
    assert ovr.test_score_ >= 0.97


def test_one_vs_rest_frank_wolfe_and_adagrad_recipes(golden):
    """ml/tests/test_svc.py:113 (FrankWolfe) and :134-140 (AdaGrad, learning_rate=1) under one-vs-rest"""
    from optiml_b200.opti.constrained import FrankWolfe
    from optiml_b200.opti.unconstrained.stochastic import AdaGrad
    iris, fw = golden('iris_ovr'), golden('frank_wolfe')
    ovr = S.check_one_vs_rest(S.real_device, iris['X_train'], iris['y_train'], iris['X_test'], FrankWolfe, max_iter=1000,
                              yt=iris['y_test'])
    assert ovr.test_score_ >= 0.97
    for c, e in enumerate(ovr.estimators_):
        assert np.abs(e.alphas_ - fw[f'iris_c{c}_alphas']).max() <= 1e-8
        assert np.array_equal(e.support_, fw[f'iris_c{c}_support'])
    ovr = S.check_one_vs_rest(S.real_device, iris['X_train'], iris['y_train'], iris['X_test'], AdaGrad, max_iter=1000,
                              yt=iris['y_test'])   # seeded start point (the goldens use one seed per class)
    assert ovr.test_score_ >= 0.97


def test_meta_estimators_against_the_reference_wrapped_in_sklearn(golden):
    """tests/golden/shared_gram.npz: OneVsRestClassifier / MultiOutputRegressor over the REAL reference's SVC / SVR
    (FrankWolfe, 400 iterations): alphas to 1e-8, support sets, intercepts, predictions, decision values"""
    S.check_against_reference_meta_estimators(S.real_device, golden('shared_gram'), max_iter=400)


def test_one_vs_rest_c1_sized_four_classes():
    """C1-sized inputs (n = 2000, d = 20) with four classes: four binary problems, one pass over M per iteration"""
    from sklearn.datasets import make_classification
    from optiml_b200.opti.constrained import ProjectedGradient
    X, y = make_classification(n_samples=2000, n_features=20, n_informative=6, n_classes=4, random_state=0)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ovr = S.check_one_vs_rest(S.real_device, X, y, X[:200], ProjectedGradient, max_iter=300)
    passes = [e.optimizer.q_passes for e in ovr.estimators_]
    assert max(passes) <= 302   # one multi-vector pass per iteration for the four problems together


def test_multi_output_regressor_shares_the_gram_matrix():
    S.check_multi_output(S.real_device, n=1200, max_iter=200)


def test_sharded_one_vs_rest_is_bitwise_equal_to_single_gpu():
    """row shards + lockstep batch + fused peer exchange for the whole batch (tests/multigpu_check.py, opt-in section)"""
    import os
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs >= 2 GPUs')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={2 if n < 4 else 4}',
           '--master-addr', '127.0.0.1', '--master-port', '29519', os.path.join(root, 'tests', 'multigpu_check.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, SVMB200_CHECK_SHARED_GRAM='1'))
    assert out.returncode == 0 and 'MULTIGPU_CHECK PASS' in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count('shared-Gram one-vs-rest') == 3
