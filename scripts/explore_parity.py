"""Exploratory: deviations of the CUDA path from every golden case (prints a table)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from optiml_b200.ml.svm import SVC, SVR, DualSVC, DualSVR
from optiml_b200.ml.svm.kernels import GaussianKernel, PolyKernel, LinearKernel
from optiml_b200.ml.svm.losses import hinge, epsilon_insensitive
from optiml_b200.opti import Quadratic
from optiml_b200.opti.constrained import ProjectedGradient
from optiml_b200.configs import make_config

def G(name):
    z = np.load(os.path.join(ROOT, 'tests', 'golden', name + '.npz')); return {k: z[k] for k in z.files}

def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))

g = G('kernels')
cases = {'linear': LinearKernel(), 'poly_d3_scale': PolyKernel(), 'poly_d2_auto_c1': PolyKernel(degree=2, gamma='auto', coef0=1.),
         'poly_d4_g05_c05': PolyKernel(degree=4, gamma=0.5, coef0=0.5), 'gauss_scale': GaussianKernel(),
         'gauss_auto': GaussianKernel(gamma='auto'), 'gauss_g03': GaussianKernel(gamma=0.3)}
for name, k in cases.items():
    a, b = k(g['X']), k(g['X'], g['Y'])
    ea = np.abs(a - g[name + '_XX']) / np.maximum(np.abs(g[name + '_XX']), 1e-300)
    print(f'kernel {name:18s} XX maxrel(global)={rel(a, g[name+"_XX"]):.2e} elemwise max={ea.max():.2e}  XY={rel(b, g[name+"_XY"]):.2e}')

g = G('bcqp')
for p in ['p2', 'p5', 'p64', 'p200']:
    lb = g.get(p + '_lb')
    opt = ProjectedGradient(quad=Quadratic(g[p + '_Q'], g[p + '_q']), ub=g[p + '_ub'], lb=lb).minimize()
    fh = getattr(opt, 'f_hist', None)
    k = min(len(fh), len(g[p+'_f_hist'])) if fh is not None else 0
    print(f'bcqp {p}: iter {opt.iter} vs {int(g[p+"_iter"])} status {opt.status} vs {g[p+"_status"]} '
          f'max|dx|={np.abs(opt.x - g[p+"_x"]).max():.2e} f_hist prefix maxdiff={np.abs(fh[:k]-g[p+"_f_hist"][:k]).max() if k else -1:.2e}')

def cmp_fit(tag, m, g, prefix=''):
    fh = np.array(m.train_loss_history); gh = g[prefix + 'f_hist']; k = min(len(fh), len(gh))
    d = np.abs(fh[:k] - gh[:k]) / np.maximum(1, np.abs(gh[:k]))
    first_bad = int(np.argmax(d > 1e-9)) if (d > 1e-9).any() else -1
    print(f'{tag}: iter {m.optimizer.iter} vs {int(g[prefix+"iter"])} {m.optimizer.status} vs {g[prefix+"status"]} '
          f'max|dalpha|={np.abs(m.alphas_ - g[prefix+"alphas"]).max():.2e} nSV {len(m.support_)} vs {len(g[prefix+"support"])} '
          f'same_sv={np.array_equal(m.support_, g[prefix+"support"])} b {m.intercept_:.12f} vs {float(g[prefix+"intercept"]):.12f} '
          f'f_hist first dev>1e-9 at {first_bad} (len {k}) final df={fh[-1]-gh[-1]:.2e}')

g = G('iris_ovr')
for c in range(3):
    m = SVC(loss=hinge, kernel=GaussianKernel(), reg_intercept=True, dual=True, optimizer=ProjectedGradient).fit(g['X_train'], (g['y_train'] == c).astype(int))
    cmp_fit(f'iris c{c}', m, g, f'c{c}_')
    dec = m.decision_function(g['X_test'])
    print('    decision maxdiff', np.abs(dec - g[f'c{c}_decision']).max(), 'pred equal', np.array_equal(m.predict(g['X_test']), g[f'c{c}_predict']))
g = G('diabetes_svr')
for name, k in (('linear', LinearKernel()), ('poly', PolyKernel(degree=3)), ('gauss', GaussianKernel())):
    m = SVR(loss=epsilon_insensitive, kernel=k, reg_intercept=True, dual=True, optimizer=ProjectedGradient, epsilon=0.1, C=1).fit(g['X_train'], g['y_train'])
    cmp_fit(f'diabetes {name}', m, g, name + '_')
    print('    decision maxdiff', np.abs(m.decision_function(g['X_test']) - g[name + '_decision']).max())
for cfg, nn, gold, mk in (('C1', None, 'c1_svc_gaussian', lambda: DualSVC(kernel=GaussianKernel(), C=1)),
                          ('C2', 600, 'c2small_svr_poly', lambda: DualSVR(kernel=PolyKernel(degree=3), epsilon=0.1, C=1)),
                          ('C3', 500, 'c3small_svc_linear', lambda: DualSVC(kernel=LinearKernel(), C=1)),
                          ('C4', 1200, 'c4small_svc_gaussian', lambda: DualSVC(kernel=GaussianKernel(), C=2.5, max_iter=300))):
    spec, X, y = make_config(cfg, n=nn)
    g = G(gold)
    t = time.time(); m = mk().fit(X, y); dt = time.time() - t
    cmp_fit(f'{cfg} n={len(y)} ({dt:.2f}s)', m, g)
    print('    decision maxdiff', np.abs(m.decision_function(X[:len(g["decision"])]) - g['decision']).max(),
          'device_ms', m.optimizer.device_ms, 'passes', m.optimizer.q_passes)
if os.path.exists(os.path.join(ROOT, 'tests', 'golden', 'c4_full_svc_gaussian.npz')) and '--full' in sys.argv:
    g = G('c4_full_svc_gaussian')
    spec, X, y = make_config('C4')
    for rep in range(2):
        t = time.time(); m = DualSVC(kernel=GaussianKernel(), C=1); m.fit(X, y); dt = time.time() - t
        print(f'C4 full fit wall {dt:.2f}s pg device_ms {m.optimizer.device_ms:.1f} passes {m.optimizer.q_passes} '
              f'=> {8*50000**2*m.optimizer.q_passes/m.optimizer.device_ms/1e6:.0f} GB/s')
    cmp_fit('C4 full', m, g)
