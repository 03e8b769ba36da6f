"""Golden vector for the HEADLINE config C4 (DualSVC Gaussian n=50 000 d=128, C=1, 1000 PG iterations),
produced by the REAL reference's solver code on this container's host cores (~20 min, ~25 GB RAM):

    python tests/golden/make_golden_c4_full.py [n]

* Q = yy^T o (K+1) is assembled in row blocks with the oracle's restatement of
  kernels.py:125-129 / sklearn euclidean_distances (the reference's own ``kernel(X)`` needs three
  n x n temporaries = 60 GB at this size); the diagonal is forced to exp(0) = 1 as the reference does.
* The solver is the reference's ``ProjectedGradient.minimize`` (projected_gradient.py:76-143) on the
  reference's ``Quadratic`` (opti/_base.py:228-300).  The ``Quadratic`` instance is created without its
  constructor's ``np.array(Q)`` copy (20 GB) -- attributes Q, q, ndim are set directly -- everything
  the loop executes is the reference's code.
* SV selection / intercept follow ml/svm/_base.py:867-880 with K rows recomputed blockwise.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_shim import load_reference  # noqa: E402
from oracle import svm_oracle as O  # noqa: E402
from optiml_b200.configs import make_config  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
ref = load_reference()
spec, X, y = make_config('C4', n=n)
classes, ys = O.binarize_labels(y)
gamma = 1. / (X.shape[1] * X.var())
XX = np.einsum('ij,ij->i', X, X)
t0 = time.time()
Q = np.empty((n, n))
B = 2000


def k_rows(r0, r1):
    D = -2 * (X[r0:r1] @ X.T)
    D += XX[r0:r1, None]
    D += XX[None, :]
    np.maximum(D, 0, out=D)
    D[np.arange(r1 - r0), np.arange(r0, r1)] = 0
    return np.exp(-gamma * D)


for r0 in range(0, n, B):
    r1 = min(n, r0 + B)
    Kb = k_rows(r0, r1)
    yy = np.outer(ys[r0:r1], ys)
    Qb = Kb * yy
    Qb += yy
    Q[r0:r1] = Qb
print('Q built', time.time() - t0, flush=True)

quad = object.__new__(ref.Quadratic)
quad.Q, quad.q, quad.ndim = Q, -np.ones(n), n
f_hist = []
t0 = time.time()


def cb(opt):
    f_hist.append(opt.f_x)
    if opt.iter % 50 == 0:
        print(opt.iter, opt.f_x, time.time() - t0, flush=True)


opt = ref.ProjectedGradient(quad=quad, ub=np.ones(n) * 1., max_iter=1000, callback=cb).minimize()
pg_s = time.time() - t0
alphas = opt.x
sv = alphas > 1e-6
support = np.arange(n)[sv]
sv_y, a = ys[sv], alphas[sv]
dual_coef = a * sv_y
b = 0.
for r0 in range(0, n, B):
    r1 = min(n, r0 + B)
    rows = support[(support >= r0) & (support < r1)]
    if len(rows) == 0:
        continue
    Kb = k_rows(r0, r1)
    for r in rows:
        b += ys[r]
        b -= np.sum(dual_coef * Kb[r - r0, sv])
b /= len(a)
name = 'c4_full_svc_gaussian' if n == 50000 else f'c4_n{n}_svc_gaussian'
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), name + '.npz'),
                    alphas=alphas, support=support.astype(np.int64), intercept=b, f_hist=np.array(f_hist),
                    iter=opt.iter, status=opt.status, f_x=opt.f_x, gamma=gamma, pg_seconds=pg_s,
                    X_checksum=np.array([X.sum(), (X * X).sum()]), cores=os.cpu_count())
print('done', opt.iter, opt.status, opt.f_x, len(support), b, 'pg_s', pg_s)
