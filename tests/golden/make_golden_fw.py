"""Golden vectors for the Frank-Wolfe widening, produced by the REAL reference
(optiml/opti/constrained/frank_wolfe.py; reference tests test_frank_wolfe.py, test_lower_bound.py:9-25,
ml/tests/test_svc.py:113, test_svr.py:129):   python tests/golden/make_golden_fw.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_shim import load_reference  # noqa: E402
from optiml_b200.configs import make_config  # noqa: E402

ref = load_reference()
OUT = os.path.dirname(os.path.abspath(__file__))
bc = dict(np.load(os.path.join(OUT, 'bcqp.npz')))
out = {}


def run_fw(Q, q, ub, lb=None, t=0., max_iter=1000):
    hist = []
    opt = ref.FrankWolfe(quad=ref.Quadratic(Q, q), ub=ub, lb=lb, t=t, max_iter=max_iter,
                         callback=lambda o: hist.append(o.f_x)).minimize()
    return dict(x=opt.x, g=opt.g_x, iter=opt.iter, status=opt.status, f_hist=np.array(hist))


for p, t in (('p2', 0.), ('p5', 0.), ('p64', 0.), ('p200', 0.), ('p64', 0.5)):
    lb = bc[p + '_lb'] if p + '_lb' in bc else None
    r = run_fw(bc[p + '_Q'], bc[p + '_q'], bc[p + '_ub'], lb=lb, t=t)
    key = p + ('_t05' if t else '')
    out.update({f'{key}_{k}': v for k, v in r.items()})
    print(key, r['iter'], r['status'], r['f_hist'][-1])

iris = dict(np.load(os.path.join(OUT, 'iris_ovr.npz')))
for c in range(3):
    m = ref.SVC(loss=ref.hinge, kernel=ref.GaussianKernel(), reg_intercept=True, dual=True,
                optimizer=ref.FrankWolfe).fit(iris['X_train'], (iris['y_train'] == c).astype(int))
    out.update({f'iris_c{c}_alphas': m.alphas_, f'iris_c{c}_iter': m.optimizer.iter, f'iris_c{c}_status': m.optimizer.status,
                f'iris_c{c}_f_hist': np.array(m.train_loss_history), f'iris_c{c}_support': m.support_,
                f'iris_c{c}_intercept': m.intercept_, f'iris_c{c}_predict': m.predict(iris['X_test'])})
    print('iris', c, m.optimizer.iter, m.optimizer.status)

spec, X, y = make_config('C1')
m = ref.SVC(loss=ref.hinge, kernel=ref.GaussianKernel(), C=1, reg_intercept=True, dual=True,
            optimizer=ref.FrankWolfe).fit(X, y)
out.update(c1_alphas=m.alphas_, c1_iter=m.optimizer.iter, c1_status=m.optimizer.status, c1_support=m.support_,
           c1_f_hist=np.array(m.train_loss_history), c1_intercept=m.intercept_)
print('C1', m.optimizer.iter, m.optimizer.status, m.optimizer.f_x, len(m.support_))
spec, X, y = make_config('C2', n=600)
m = ref.SVR(loss=ref.epsilon_insensitive, epsilon=0.1, kernel=ref.PolyKernel(degree=3), C=1, reg_intercept=True, dual=True,
            optimizer=ref.FrankWolfe).fit(X, y)
out.update(c2small_alphas=m.alphas_, c2small_iter=m.optimizer.iter, c2small_status=m.optimizer.status,
           c2small_support=m.support_, c2small_f_hist=np.array(m.train_loss_history), c2small_intercept=m.intercept_)
print('C2 small', m.optimizer.iter, m.optimizer.status, m.optimizer.f_x)
np.savez_compressed(os.path.join(OUT, 'frank_wolfe.npz'), **out)
