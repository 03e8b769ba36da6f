"""Golden vectors for the augmented-Lagrangian dual path, produced by the REAL reference:
SVC/SVR(dual=True, optimizer=<StochasticOptimizer>, reg_intercept in {True, False}) -- the recipe of the reference's
own tests ml/tests/test_svc.py:134-147 and test_svr.py:150-163 (AdaGrad, learning_rate=1.) plus the other six update
rules and the momentum variants:   python tests/golden/make_golden_al.py"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle.ref_shim import load_reference  # noqa: E402
from al_cases import CASES, SVC_RUNS, TOL_RUNS, SVR_RUNS, svc_key, svr_key  # noqa: E402

ref = load_reference()
from optiml.opti.unconstrained.stochastic import (AdaGrad, StochasticGradientDescent, RMSProp, AdaDelta, Adam,  # noqa: E402
                                                  AMSGrad, AdaMax)
from sklearn.datasets import load_diabetes  # noqa: E402
from sklearn.preprocessing import StandardScaler  # noqa: E402

warnings.simplefilter('ignore')
OUT = os.path.dirname(os.path.abspath(__file__))
iris = dict(np.load(os.path.join(OUT, 'iris_ovr.npz')))
out = {}

REF_CLASS = dict(adagrad=AdaGrad, sgd=StochasticGradientDescent, rmsprop=RMSProp, adadelta=AdaDelta, adam=Adam,
                 amsgrad=AMSGrad, adamax=AdaMax)

def record(key, m, X_test):
    o = m.optimizer
    out.update({f'{key}_alphas': m.alphas_, f'{key}_dual_x': o.f.dual_x, f'{key}_iter': o.iter, f'{key}_status': o.status,
                f'{key}_f_x': o.f_x, f'{key}_g_x': o.g_x, f'{key}_pf_hist': np.array(m.train_loss_history),
                f'{key}_support': m.support_, f'{key}_intercept': m.intercept_,
                f'{key}_decision': m.decision_function(X_test)})
    print(key, o.iter, o.status, o.f_x, len(m.support_), m.intercept_)


# --- SVC on the iris one-vs-rest split of the reference's test (class 1 vs rest for every rule, all classes for AdaGrad)
for name, ri, c in SVC_RUNS:
    rule, lr, kw, iters = CASES[name]
    m = ref.SVC(loss=ref.hinge, kernel=ref.gaussian, reg_intercept=ri, dual=True, optimizer=REF_CLASS[rule], learning_rate=lr,
                max_iter=iters, random_state=c + 1, **kw).fit(iris['X_train'], (iris['y_train'] == c).astype(int))
    record(svc_key(name, ri, c), m, iris['X_test'])

# --- an optimality exit: a loose tolerance ends the run through check_lagrangian_dual_optimality (opti/_base.py:143-147)
for name, tol in TOL_RUNS:
    rule, lr, kw, _ = CASES[name]
    m = ref.SVC(loss=ref.hinge, kernel=ref.gaussian, reg_intercept=False, dual=True, optimizer=REF_CLASS[rule], learning_rate=lr,
                tol=tol, max_iter=1000, random_state=7, **kw).fit(iris['X_train'], (iris['y_train'] == 2).astype(int))
    record(f'svc_{name}_tol_c2', m, iris['X_test'])

# --- SVR: first 150 rows of diabetes, standardised (the reference's Boston test cannot run offline)
Xd, yd = load_diabetes(return_X_y=True)
Xd = StandardScaler().fit_transform(Xd)[:150]
yd = ((yd - yd.mean()) / yd.std())[:150]
out.update(svr_X=Xd, svr_y=yd, svr_X_test=Xd[:40] + 0.05)
for kname, name, ri in SVR_RUNS:
    rule, lr, kw, iters = CASES[name]
    m = ref.SVR(loss=ref.epsilon_insensitive, epsilon=0.1, kernel=getattr(ref, kname), reg_intercept=ri, dual=True,
                optimizer=REF_CLASS[rule], learning_rate=lr, max_iter=min(iters, 400), random_state=3, **kw).fit(Xd, yd)
    record(svr_key(kname, name, ri), m, out['svr_X_test'])

# --- the reference's own unit test of AugmentedLagrangianQuadratic (opti/constrained/tests/test_lagrangian_quadratic.py:18-22):
#     the 2-variable generator problem with the equality row A = [2, 7], b = 0 and 0 <= x <= ub; the optimum is x = 0
from optiml.opti.constrained import AugmentedLagrangianQuadratic  # noqa: E402
bc = dict(np.load(os.path.join(OUT, 'bcqp.npz')))
for seed in (0, 1, 2):
    ld = AugmentedLagrangianQuadratic(primal=ref.Quadratic(bc['p2_Q'], bc['p2_q']), A=[2, 7], b=np.zeros(1),
                                      lb=np.zeros(2), ub=bc['p2_ub'], rho=1)
    o = AdaGrad(ld, step_size=1, epochs=15000, random_state=seed).minimize()
    out.update({f'alq2d_s{seed}_x': o.x, f'alq2d_s{seed}_iter': o.iter, f'alq2d_s{seed}_status': o.status,
                f'alq2d_s{seed}_dual_x': ld.dual_x, f'alq2d_s{seed}_f_x': o.f_x, f'alq2d_s{seed}_g_x': o.g_x,
                f'alq2d_s{seed}_pf_hist': np.array(o.f_x_history)})
    print('alq2d', seed, o.x, o.iter, o.status)

np.savez_compressed(os.path.join(OUT, 'al_stochastic.npz'), **out)
print('wrote', len(out), 'arrays')
