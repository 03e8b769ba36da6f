"""Golden Gram matrices of the reference's LaplacianKernel / SigmoidKernel (kernels.py:132-201) on the inputs of
kernels.npz, plus an iris SVC fit with each (widening 8f-2):   python tests/golden/make_golden_kernels_extra.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_shim import load_reference  # noqa: E402

ref = load_reference()
OUT = os.path.dirname(os.path.abspath(__file__))
k = dict(np.load(os.path.join(OUT, 'kernels.npz')))
X, Y = k['X'], k['Y']
out = {}
for name, kern in (('lap_scale', ref.LaplacianKernel()), ('lap_g03', ref.LaplacianKernel(gamma=0.3)),
                   ('sig_scale', ref.SigmoidKernel()), ('sig_g01_c05', ref.SigmoidKernel(gamma=0.01, coef0=0.5))):
    out[name + '_XX'] = kern(X)
    out[name + '_XY'] = kern(X, Y)
iris = dict(np.load(os.path.join(OUT, 'iris_ovr.npz')))
yb = (iris['y_train'] == 0).astype(int)
for name, kern in (('lap', ref.LaplacianKernel()), ('sig', ref.SigmoidKernel(gamma=0.05, coef0=0.))):
    m = ref.SVC(loss=ref.hinge, kernel=kern, reg_intercept=True, dual=True, optimizer=ref.FrankWolfe, max_iter=300).fit(
        iris['X_train'], yb)
    out.update({f'iris_{name}_alphas': m.alphas_, f'iris_{name}_support': m.support_, f'iris_{name}_intercept': m.intercept_,
                f'iris_{name}_decision': m.decision_function(iris['X_test']), f'iris_{name}_f_hist': np.array(m.train_loss_history)})
    print(name, m.optimizer.iter, m.optimizer.status, len(m.support_))
np.savez_compressed(os.path.join(OUT, 'kernels_extra.npz'), **out)
