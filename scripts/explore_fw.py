import os, sys
sys.path.insert(0, '.')
import numpy as np
from optiml_b200.opti import Quadratic
from optiml_b200.opti.constrained import FrankWolfe
from optiml_b200.ml.svm import SVC, SVR
from optiml_b200.ml.svm.kernels import GaussianKernel, PolyKernel
from optiml_b200.ml.svm.losses import hinge, epsilon_insensitive
from optiml_b200.configs import make_config
from oracle import svm_oracle as O
def G(n):
    z = np.load(f'tests/golden/{n}.npz'); return {k: z[k] for k in z.files}
bc, fw, iris = G('bcqp'), G('frank_wolfe'), G('iris_ovr')
for key, p, t in (('p2', 'p2', 0.), ('p5', 'p5', 0.), ('p64', 'p64', 0.), ('p200', 'p200', 0.), ('p64_t05', 'p64', 0.5)):
    lb = bc.get(p + '_lb')
    opt = FrankWolfe(quad=Quadratic(bc[p + '_Q'], bc[p + '_q']), ub=bc[p + '_ub'], lb=lb, t=t).minimize()
    fh = getattr(opt, 'f_hist', np.array(getattr(opt, 'f_x_history', [])))
    gh = fw[key + '_f_hist']; k = min(len(fh), len(gh))
    dev = np.abs(fh[:k] - gh[:k]) / np.maximum(1, np.abs(gh[:k]))
    print(key, 'iter', opt.iter, int(fw[key + '_iter']), opt.status, fw[key + '_status'], 'max|dx|', np.abs(opt.x - fw[key + '_x']).max(),
          'first f dev>1e-9', int(np.argmax(dev > 1e-9)) if (dev > 1e-9).any() else -1, 'final df', fh[-1] - gh[-1])
    # iteration map from the start: 5 iterations
    want = O.frank_wolfe(bc[p + '_Q'], bc[p + '_q'], bc[p + '_ub'], lb=lb, t=t, max_iter=5)
    got = FrankWolfe(quad=Quadratic(bc[p + '_Q'], bc[p + '_q']), ub=bc[p + '_ub'], lb=lb, t=t, max_iter=5).minimize()
    print('    5-iter map: iter', got.iter, want.iter, got.status, want.status, 'max|dx|', np.abs(got.x - want.x).max())
for c in range(3):
    m = SVC(loss=hinge, kernel=GaussianKernel(), reg_intercept=True, dual=True, optimizer=FrankWolfe).fit(iris['X_train'], (iris['y_train'] == c).astype(int))
    fh = np.array(m.train_loss_history); gh = fw[f'iris_c{c}_f_hist']
    dev = np.abs(fh - gh) / np.maximum(1, np.abs(gh))
    print('iris', c, m.optimizer.iter, m.optimizer.status, 'max|da|', np.abs(m.alphas_ - fw[f'iris_c{c}_alphas']).max(), 'same sv', np.array_equal(m.support_, fw[f'iris_c{c}_support']),
          'first dev', int(np.argmax(dev > 1e-9)) if (dev > 1e-9).any() else -1, 'pred eq', np.array_equal(m.predict(iris['X_test']), fw[f'iris_c{c}_predict']), 'b', m.intercept_, float(fw[f'iris_c{c}_intercept']))
spec, X, y = make_config('C1')
m = SVC(loss=hinge, kernel=GaussianKernel(), C=1, reg_intercept=True, dual=True, optimizer=FrankWolfe).fit(X, y)
fh = np.array(m.train_loss_history); gh = fw['c1_f_hist']; dev = np.abs(fh - gh) / np.maximum(1, np.abs(gh))
print('C1 FW', m.optimizer.iter, m.optimizer.status, 'max|da|', np.abs(m.alphas_ - fw['c1_alphas']).max(), 'same sv', np.array_equal(m.support_, fw['c1_support']), 'first dev', int(np.argmax(dev > 1e-9)) if (dev > 1e-9).any() else -1, 'final f', fh[-1], gh[-1], 'device ms', m.optimizer.device_ms)
spec, X, y = make_config('C2', n=600)
m = SVR(loss=epsilon_insensitive, epsilon=0.1, kernel=PolyKernel(degree=3), C=1, reg_intercept=True, dual=True, optimizer=FrankWolfe).fit(X, y)
fh = np.array(m.train_loss_history); gh = fw['c2small_f_hist']; dev = np.abs(fh - gh) / np.maximum(1, np.abs(gh))
print('C2small FW', m.optimizer.iter, m.optimizer.status, 'max|da|', np.abs(m.alphas_ - fw['c2small_alphas']).max(), 'first dev', int(np.argmax(dev > 1e-9)) if (dev > 1e-9).any() else -1, 'final f', fh[-1], gh[-1])
spec, X, y = make_config('C4')
m = SVC(loss=hinge, kernel=GaussianKernel(), C=1, reg_intercept=True, dual=True, optimizer=FrankWolfe).fit(X, y)
print('C4 FW full', m.optimizer.iter, m.optimizer.status, 'f', m.optimizer.f_x, 'pg ms', m.optimizer.device_ms, 'it/s', m.optimizer.iter / m.optimizer.device_ms * 1e3)
