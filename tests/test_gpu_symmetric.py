"""The opt-in symmetric pass (K2s, optiml_b200/csrc/k2_symv.cuh; runtime.use_symmetric_pass / SVMB200_SYMMETRIC=1) on the
B200: the product Q d of projected_gradient.py:113 from the upper triangle of the matrix alone.

The checks of tests/symv_checks.py (run on the host emulation in the CPU suite with small tile shapes) with the shipped
tile shape, and the full-size stable configurations against the REAL reference's golden runs: north_star's bar -- alpha
within 1e-8, identical support set, same predictions -- must hold in this mode too ("next" 9).
"""
import numpy as np
import pytest

import symv_checks as SY
from shared_gram_checks import real_device
from optiml_b200.configs import make_config
from test_gpu_estimators import api, check_stable_fit

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('n', [130, 1000, 2049, 4500, 9001])
def test_symmetric_pass_product(n):
    w0 = SY.check_symmetric_product(real_device, n)
    w1 = SY.check_symmetric_product(real_device, n)
    assert np.array_equal(w0, w1)   # reproducible run to run


@pytest.mark.parametrize('n', [300, 5000])
def test_symmetric_pass_never_reads_below_the_diagonal_blocks(n):
    from optiml_b200 import _native as N
    SY.check_lower_triangle_is_never_read(real_device, n, 128)   # the shipped band height (SVMB200_SYMV_TR x _NRB)


def test_symmetric_pass_solvers_follow_the_default_pass():
    SY.check_symmetric_solves(real_device, n=2600, max_iter=60)
    SY.check_symmetric_solves(real_device, n=333, max_iter=40)


def test_symmetric_pass_estimators():
    SY.check_symmetric_fit(real_device, n=2500)


def test_c1_full_parity_symmetric(golden):
    """config C1 against the real reference's run.  n = 2000 runs on the persistent shared-memory loop by default (which
    does not stream the matrix at all); the profile flag keeps it on the two-kernel loop so that K2s is what is tested"""
    A = api()
    g = golden('c1_svc_gaussian')
    spec, X, y = make_config('C1')
    with SY.symmetric_pass():
        m = A['SVC'](loss=A['hinge'], kernel=A['GaussianKernel'](), C=1, dual=True, reg_intercept=True,
                     optimizer=A['ProjectedGradient'])
        m.profile_matvec = True
        m.fit(X, y)
    assert m.optimizer.symmetric_pass
    check_stable_fit(m, g)
    assert len(m.support_) == 1015
    assert np.abs(m.decision_function(X[:256]) - g['decision']).max() <= 1e-9
    assert np.array_equal(m.predict(X[:256]), g['predict'])


def test_c4_headline_full_parity_symmetric(golden):
    """BASELINE headline config (n = 50 000, 20 GB Hessian, 10 GB streamed per iteration in this mode) against the
    unmodified reference's 1000 iterations: alpha within 1e-8, identical support set, intercept, loss history"""
    A = api()
    g = golden('c4_full_svc_gaussian')
    spec, X, y = make_config('C4')
    with SY.symmetric_pass():
        m = A['DualSVC'](kernel=A['GaussianKernel'](), C=1).fit(X, y)
    assert m.optimizer.symmetric_pass
    check_stable_fit(m, g)
    assert len(m.support_) == 49020
    fh = np.array(m.train_loss_history)
    assert np.all(np.diff(fh) <= 1e-9 * np.abs(fh[:-1]))
    g_fresh = m.obj.jacobian(m.alphas_)  # the FULL pass over the resident Q: an independent check of the gradient
    assert np.abs(g_fresh - m.optimizer.g_x).max() <= 1e-9 * np.abs(g_fresh).max()
    m.obj.release()


def test_c3_linear_full_size_symmetric(golden):
    """config C3 (linear kernel, n = 20 000): symmetric pass against the default pass on the same resident Gram matrix"""
    A = api()
    spec, X, y = make_config('C3')
    fits = {}
    for sym in (False, True):
        with SY.symmetric_pass(sym):
            m = A['DualSVC'](kernel=A['LinearKernel'](), C=1, max_iter=300).fit(X, y)
            assert m.optimizer.symmetric_pass is sym
            fits[sym] = (m.alphas_.copy(), m.support_.copy(), m.intercept_)
            m.obj.release()
    assert np.abs(fits[True][0] - fits[False][0]).max() <= 1e-8
    assert np.array_equal(fits[True][1], fits[False][1])
    assert abs(fits[True][2] - fits[False][2]) <= 1e-8
