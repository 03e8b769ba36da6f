"""States of the REAL reference's trajectory on C2 at full size (DualSVR poly, n = 10 000 -> 20 000 variables), for the
iteration-map parity test of the regime where the loss histories part (tests/test_gpu_estimators.py), plus the
reference algorithm's own sensitivity on this problem (how far a +-1 ulp change of the Gram matrix moves alpha).

    python tests/golden/make_golden_c2_states.py        (~4 min and ~10 GB on 8 host cores)

Stored: x at iterations K and K + STEPS of the reference run (callback of ProjectedGradient, projected_gradient.py:95-98),
f at those points; `env_*`: the oracle's single-pass block form (bit-identical arithmetic to nothing in particular -- it is
only compared with itself) on M = K + 1 and on M perturbed by +-1 ulp: max |delta alpha| after 1000 iterations and the
relative distance of the two loss histories at every iteration.
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

KS = (0, 90, 100, 250, 500, 750, 995)
STEPS = 5

ref = load_reference()
spec, X, y = make_config('C2')
n = len(y)
want = sorted(set(KS) | {k + STEPS for k in KS})
snaps, fs = {}, {}


def grab(opt):
    if opt.iter in want:
        snaps[opt.iter] = opt.x.copy()
        fs[opt.iter] = float(opt.f_x)


t0 = time.time()
m = ref.SVR(loss=ref.epsilon_insensitive, epsilon=0.1, kernel=ref.PolyKernel(degree=3), C=1, reg_intercept=True,
            dual=True, optimizer=ref.ProjectedGradient)
# SVR.fit passes callback=self._store_train_info (ml/svm/_base.py:1180-1185); wrap it so the history is kept as well
orig = m._store_train_info


def both(opt):
    orig(opt)
    grab(opt)


m._store_train_info = both
m.fit(X, y)
fit_s = time.time() - t0
g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'c2_full_svr_poly.npz'))
assert np.array_equal(m.alphas_, g['alphas']), 'the run does not reproduce c2_full_svr_poly.npz'
del m

# the reference algorithm's own sensitivity at this size (oracle, single-pass block form on M = K + 1)
K = O.poly_kernel(X, degree=3)
M = K + 1.0
del K
q = np.hstack((-y, y)) + 0.1
ub = np.ones(2 * n)
base = O.projected_gradient(O.SVRBlockOperator(M), q, ub, passes=1)
rng = np.random.default_rng(0)
E = rng.integers(-1, 2, size=M.shape)
E = np.triu(E) + np.triu(E, 1).T
pert = O.projected_gradient(O.SVRBlockOperator(M * (1 + E * 2.2e-16)), q, ub, passes=1)
env_dalpha = float(np.abs(pert.x - base.x).max())
env_f_rel = np.abs(pert.f_hist - base.f_hist) / np.abs(base.f_hist)

out = {f'x_{k}': snaps[k] for k in want}
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'c2_full_states.npz'),
                    ks=np.array(KS), steps=STEPS, f_at=np.array([fs[k] for k in want]), k_at=np.array(want),
                    env_dalpha=env_dalpha, env_f_rel=env_f_rel, env_f_base=base.f_hist, env_f_pert=pert.f_hist,
                    env_clipped=np.array([base.n_clipped, pert.n_clipped]), fit_seconds=fit_s, **out)
print('done', fit_s, 'envelope max|dalpha|', env_dalpha, 'f', base.f_x, pert.f_x, 'clipped steps', base.n_clipped,
      pert.n_clipped)
