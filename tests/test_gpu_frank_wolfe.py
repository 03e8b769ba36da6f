"""Widening step 1: FrankWolfe on the same boundary, parity vs the reference's
frank_wolfe.py through golden vectors (tests/golden/make_golden_fw.py).  Frank-Wolfe trajectories are stable
(every step is a convex combination with a box vertex), so the 1e-8 bar holds on every case, including the
ones that are chaotic under ProjectedGradient."""
import numpy as np
import pytest

from oracle import svm_oracle as O
from optiml_b200.configs import make_config

pytestmark = pytest.mark.gpu


def fw_solve(Q, q, ub, **kw):
    from optiml_b200.opti import Quadratic
    from optiml_b200.opti.constrained import FrankWolfe
    return FrankWolfe(quad=Quadratic(Q, q), ub=ub, **kw).minimize()


@pytest.mark.parametrize('key,p,t', [('p2', 'p2', 0.), ('p5', 'p5', 0.), ('p64', 'p64', 0.), ('p200', 'p200', 0.),
                                     ('p64_t05', 'p64', 0.5)])
def test_bcqp_golden(golden, key, p, t):
    g, fw = golden('bcqp'), golden('frank_wolfe')
    lb = g[p + '_lb'] if p + '_lb' in g else None
    opt = fw_solve(g[p + '_Q'], g[p + '_q'], g[p + '_ub'], lb=lb, t=t)
    assert opt.iter == int(fw[key + '_iter']) and opt.status == str(fw[key + '_status'])
    assert np.abs(opt.x - fw[key + '_x']).max() <= 1e-8
    assert np.abs(opt.g_x - fw[key + '_g']).max() <= 1e-8 * max(1., np.abs(fw[key + '_g']).max())
    hist = opt.f_hist if hasattr(opt, 'f_hist') else np.array(opt.f_x_history)
    assert np.abs(hist - fw[key + '_f_hist']).max() <= 1e-9 * max(1., np.abs(fw[key + '_f_hist']).max())
    assert np.all(opt.x >= opt.lb - 1e-9) and np.all(opt.x <= opt.ub + 1e-9)


def test_lower_bound_recipe_like_reference(golden):
    """opti/constrained/tests/test_lower_bound.py:9-25 for FrankWolfe: feasible and within 1e-2 of the optimum"""
    from optiml_b200.opti import Quadratic
    from optiml_b200.opti.constrained import FrankWolfe
    g = golden('bcqp')
    quad = Quadratic(g['p5_Q'], g['p5_q'])
    bcqp = FrankWolfe(quad=quad, ub=g['p5_ub'], lb=g['p5_lb'])
    x = bcqp.minimize().x
    assert np.all(x >= g['p5_lb'] - 1e-6) and np.all(x <= g['p5_ub'] + 1e-6)
    f_opt = quad.function(bcqp.x_star())
    assert (quad.function(x) - f_opt) / max(abs(f_opt), 1) <= 1e-2


@pytest.mark.parametrize('c', [0, 1, 2])
def test_iris_ovr(golden, c):
    """ml/tests/test_svc.py:113 recipe"""
    from optiml_b200.ml.svm import SVC
    from optiml_b200.ml.svm.kernels import gaussian
    from optiml_b200.ml.svm.losses import hinge
    from optiml_b200.opti.constrained import FrankWolfe
    iris, fw = golden('iris_ovr'), golden('frank_wolfe')
    m = SVC(loss=hinge, kernel=gaussian, reg_intercept=True, dual=True, optimizer=FrankWolfe).fit(
        iris['X_train'], (iris['y_train'] == c).astype(int))
    p = f'iris_c{c}_'
    assert m.optimizer.iter == int(fw[p + 'iter']) and m.optimizer.status == str(fw[p + 'status'])
    assert np.abs(m.alphas_ - fw[p + 'alphas']).max() <= 1e-8
    assert np.array_equal(m.support_, fw[p + 'support'])
    assert abs(m.intercept_ - float(fw[p + 'intercept'])) <= 1e-8
    assert np.array_equal(m.predict(iris['X_test']), fw[p + 'predict'])
    assert np.abs(np.array(m.train_loss_history) - fw[p + 'f_hist']).max() <= 1e-9


def test_c1_and_reduced_c2(golden):
    from optiml_b200.ml.svm import SVC, SVR
    from optiml_b200.ml.svm.kernels import GaussianKernel, PolyKernel
    from optiml_b200.ml.svm.losses import hinge, epsilon_insensitive
    from optiml_b200.opti.constrained import FrankWolfe
    fw = golden('frank_wolfe')
    spec, X, y = make_config('C1')
    m = SVC(loss=hinge, kernel=GaussianKernel(), C=1, reg_intercept=True, dual=True, optimizer=FrankWolfe).fit(X, y)
    assert m.optimizer.iter == 1000 and m.optimizer.status == 'stopped'
    assert np.abs(m.alphas_ - fw['c1_alphas']).max() <= 1e-8 and np.array_equal(m.support_, fw['c1_support'])
    assert abs(m.intercept_ - float(fw['c1_intercept'])) <= 1e-8
    assert np.abs(np.array(m.train_loss_history) - fw['c1_f_hist']).max() <= 1e-9 * np.abs(fw['c1_f_hist']).max()
    spec, X, y = make_config('C2', n=600)
    m = SVR(loss=epsilon_insensitive, epsilon=0.1, kernel=PolyKernel(degree=3), C=1, reg_intercept=True, dual=True,
            optimizer=FrankWolfe).fit(X, y)
    assert np.abs(m.alphas_ - fw['c2small_alphas']).max() <= 1e-8 and np.array_equal(m.support_, fw['c2small_support'])
    assert abs(m.intercept_ - float(fw['c2small_intercept'])) <= 1e-8


def test_protocol(golden, capsys):
    from optiml_b200.opti import Quadratic
    from optiml_b200.opti.constrained import FrankWolfe
    g = golden('bcqp')
    with pytest.raises(ValueError):
        FrankWolfe(quad=Quadratic(g['p5_Q'], g['p5_q']), ub=g['p5_ub'], t=1.)  # frank_wolfe.py:84-85
    seen = []
    opt = fw_solve(g['p5_Q'], g['p5_q'], g['p5_ub'], lb=g['p5_lb'], max_iter=12, callback=lambda o: seen.append(o.iter))
    assert seen == list(range(13)) and opt.status == 'stopped'
    want = O.frank_wolfe(g['p5_Q'], g['p5_q'], g['p5_ub'], lb=g['p5_lb'], max_iter=12)
    assert np.abs(opt.x - want.x).max() <= 1e-12
    fw_solve(g['p2_Q'], g['p2_q'], g['p2_ub'], verbose=True)
    out = capsys.readouterr().out
    assert out.startswith('iter\t cost\t\t lb\t\t gap') and out.endswith('\n\n') and '\n   0\t' in out
