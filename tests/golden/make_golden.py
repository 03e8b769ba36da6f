"""Generate the golden vectors under tests/golden/ by running the REAL reference
(/root/reference, dmeoli/optiml 1.8) in this container.  Run from the repo root:

    python tests/golden/make_golden.py

The .npz files are committed; the GPU box (where /root/reference does not exist) only reads them.
Environment used: numpy 2.3.5, scipy 1.18.1, scikit-learn 1.9.0, OpenBLAS (8 threads).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_shim import load_reference  # noqa: E402
from optiml_b200.configs import make_config  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
ref = load_reference()


def save(name, **arrays):
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **arrays)
    print('wrote', name, {k: np.shape(v) for k, v in arrays.items()})


def run_pg(Q, q, ub, lb=None, max_iter=1000):
    hist = {'f': [], 'ng': []}

    def cb(opt):
        hist['f'].append(opt.f_x)

    opt = ref.ProjectedGradient(quad=ref.Quadratic(Q, q), ub=ub, lb=lb, max_iter=max_iter, callback=cb).minimize()
    return dict(x=opt.x, f_x=opt.f_x, g_x=opt.g_x, iter=opt.iter, status=opt.status, f_hist=np.array(hist['f']))


# ---------------------------------------------------------------- kernels (kernels.py:49-51, 91-95, 125-129)
rng = np.random.default_rng(42)
Xk = rng.standard_normal((50, 7)) * 1.7 + 0.3
Yk = rng.standard_normal((30, 7)) * 0.9 - 0.2
kern = dict(X=Xk, Y=Yk)
cases = {
    'linear': ref.LinearKernel(),
    'poly_d3_scale': ref.PolyKernel(),
    'poly_d2_auto_c1': ref.PolyKernel(degree=2, gamma='auto', coef0=1.),
    'poly_d4_g05_c05': ref.PolyKernel(degree=4, gamma=0.5, coef0=0.5),
    'gauss_scale': ref.GaussianKernel(),
    'gauss_auto': ref.GaussianKernel(gamma='auto'),
    'gauss_g03': ref.GaussianKernel(gamma=0.3),
}
for name, k in cases.items():
    kern[name + '_XX'] = k(Xk)
    kern[name + '_XY'] = k(Xk, Yk)
save('kernels', **kern)

# ---------------------------------------------------------------- BCQP known-answer problems
# opti/constrained/tests/test_projected_gradient.py:9-12 and test_lower_bound.py:9-25
bc = {}
Q, q, ub = ref.generate_box_constrained_quadratic(ndim=2)
r = run_pg(Q, q, ub)
bc.update({'p2_Q': Q, 'p2_q': q, 'p2_ub': ub, 'p2_x': r['x'], 'p2_iter': r['iter'], 'p2_status': r['status'],
           'p2_f_hist': r['f_hist'], 'p2_g': r['g_x']})
Q, q, ub = ref.generate_box_constrained_quadratic(ndim=5, seed=7)
r = run_pg(Q, q, ub, lb=ub / 4)
bc.update({'p5_Q': Q, 'p5_q': q, 'p5_ub': ub, 'p5_lb': ub / 4, 'p5_x': r['x'], 'p5_iter': r['iter'],
           'p5_status': r['status'], 'p5_f_hist': r['f_hist'], 'p5_g': r['g_x']})
for nd, seed in ((64, 1), (200, 3)):
    Q, q, ub = ref.generate_box_constrained_quadratic(ndim=nd, seed=seed)
    r = run_pg(Q, q, ub)
    p = f'p{nd}_'
    bc.update({p + 'Q': Q, p + 'q': q, p + 'ub': ub, p + 'x': r['x'], p + 'iter': r['iter'],
               p + 'status': r['status'], p + 'f_hist': r['f_hist'], p + 'g': r['g_x']})
save('bcqp', **bc)


# ---------------------------------------------------------------- estimator fits
def fit_svc(X, y, kernel, Xt, **kw):
    m = ref.SVC(loss=ref.hinge, kernel=kernel, reg_intercept=True, dual=True,
                optimizer=ref.ProjectedGradient, **kw).fit(X, y)
    return dict(alphas=m.alphas_, support=m.support_, dual_coef=m.dual_coef_, intercept=m.intercept_,
                f_hist=np.array(m.train_loss_history), iter=m.optimizer.iter, status=m.optimizer.status,
                f_x=m.optimizer.f_x, g_x=m.optimizer.g_x, decision=m.decision_function(Xt), predict=m.predict(Xt),
                **({'coef': m.coef_} if isinstance(kernel, ref.LinearKernel) else {}))


def fit_svr(X, y, kernel, Xt, **kw):
    m = ref.SVR(loss=ref.epsilon_insensitive, kernel=kernel, reg_intercept=True, dual=True,
                optimizer=ref.ProjectedGradient, **kw).fit(X, y)
    return dict(alphas=m.alphas_, support=m.support_, dual_coef=m.dual_coef_, intercept=m.intercept_,
                f_hist=np.array(m.train_loss_history), iter=m.optimizer.iter, status=m.optimizer.status,
                f_x=m.optimizer.f_x, g_x=m.optimizer.g_x, decision=m.decision_function(Xt), predict=m.predict(Xt),
                **({'coef': m.coef_} if isinstance(kernel, ref.LinearKernel) else {}))


def pack(prefix, d):
    return {prefix + k: v for k, v in d.items()}


# iris one-vs-rest, ml/tests/test_svc.py:96-103
from sklearn.datasets import load_iris, load_diabetes  # noqa: E402
from sklearn.model_selection import train_test_split  # noqa: E402
from sklearn.preprocessing import MinMaxScaler, StandardScaler  # noqa: E402

X, y = load_iris(return_X_y=True)
Xs = MinMaxScaler().fit_transform(X)
Xtr, Xte, ytr, yte = train_test_split(Xs, y, train_size=0.75, random_state=123456)
iris = dict(X_train=Xtr, X_test=Xte, y_train=ytr, y_test=yte)
for c in range(3):
    iris.update(pack(f'c{c}_', fit_svc(Xtr, (ytr == c).astype(int), ref.GaussianKernel(), Xte)))
save('iris_ovr', **iris)

# diabetes SVR (stand-in for the Boston test ml/tests/test_svr.py:112-119 that needs a download)
X, y = load_diabetes(return_X_y=True)
Xs = StandardScaler().fit_transform(X)
y = (y - y.mean()) / y.std()
Xtr, Xte, ytr, yte = train_test_split(Xs, y, train_size=0.75, random_state=123456)
dia = dict(X_train=Xtr, X_test=Xte, y_train=ytr, y_test=yte)
for name, k in (('linear', ref.LinearKernel()), ('poly', ref.PolyKernel(degree=3)), ('gauss', ref.GaussianKernel())):
    dia.update(pack(name + '_', fit_svr(Xtr, ytr, k, Xte, epsilon=0.1, C=1)))
save('diabetes_svr', **dia)

# C1 at full size (the reference's own CPU-runnable case): inputs are regenerated from the seed
spec, X, y = make_config('C1')
r = fit_svc(X, y, ref.GaussianKernel(), X[:256], C=1)
K = ref.GaussianKernel()(X)
r['K_sub'] = K[::97, ::89].copy()
r['K_row0'] = K[0].copy()
r['gamma'] = 1. / (X.shape[1] * X.var())
r['X_checksum'] = np.array([X.sum(), (X * X).sum()])
save('c1_svc_gaussian', **r)

# reduced-size C2 / C3 recipes (SVR poly, SVC linear)
spec, X, y = make_config('C2', n=600)
save('c2small_svr_poly', **fit_svr(X, y, ref.PolyKernel(degree=3), X[:128], epsilon=0.1, C=1))
spec, X, y = make_config('C3', n=500)
save('c3small_svc_linear', **fit_svc(X, y, ref.LinearKernel(), X[:128], C=1))
# reduced-size C4 recipe with fewer iterations and a different C
spec, X, y = make_config('C4', n=1200)
save('c4small_svc_gaussian', **fit_svc(X, y, ref.GaussianKernel(), X[:128], C=2.5, max_iter=300))
