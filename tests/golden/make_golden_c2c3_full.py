"""Full-size golden vectors for C2 (DualSVR poly n=10 000 d=32) and C3 (DualSVC linear n=20 000 d=784)
from the REAL reference's SVR.fit / SVC.fit (~3 min and ~10 GB each on 8 host cores):

    python tests/golden/make_golden_c2c3_full.py C2|C3
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_shim import load_reference  # noqa: E402
from optiml_b200.configs import make_config  # noqa: E402

ref = load_reference()
which = sys.argv[1]
spec, X, y = make_config(which)
t0 = time.time()
if which == 'C2':
    m = ref.SVR(loss=ref.epsilon_insensitive, epsilon=0.1, kernel=ref.PolyKernel(degree=3), C=1, reg_intercept=True,
                dual=True, optimizer=ref.ProjectedGradient).fit(X, y)
    name = 'c2_full_svr_poly'
else:
    m = ref.SVC(loss=ref.hinge, kernel=ref.LinearKernel(), C=1, reg_intercept=True, dual=True,
                optimizer=ref.ProjectedGradient).fit(X, y)
    name = 'c3_full_svc_linear'
fit_s = time.time() - t0
extra = {'coef': m.coef_} if which == 'C3' else {}
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), name + '.npz'),
                    alphas=m.alphas_, support=m.support_.astype(np.int64), intercept=m.intercept_,
                    f_hist=np.array(m.train_loss_history), iter=m.optimizer.iter, status=m.optimizer.status,
                    f_x=m.optimizer.f_x, decision=m.decision_function(X[:256]), fit_seconds=fit_s,
                    X_checksum=np.array([X.sum(), (X * X).sum()]), cores=os.cpu_count(), **extra)
print('done', which, m.optimizer.iter, m.optimizer.status, m.optimizer.f_x, len(m.support_), m.intercept_, fit_s)
