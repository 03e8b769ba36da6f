"""Where does the C2 (DualSVR, PolyKernel(3), n = 10 000) trajectory leave the reference's?  

Three full-length solves of the same problem, compared pairwise:
  A  device loop on the DEVICE Gram matrix      (the product path: K1 + K2/K3)
  B  device loop on the HOST   Gram matrix      (NumPy `(gamma X X' + c0) ** 3 + 1`, uploaded: same bits as the oracle's)
  C  oracle (NumPy, single-pass block form) on the HOST Gram matrix
B vs C differ only in the summation order of the matvec / reductions; A vs B only in the ulps of the Gram matrix
(`pow` and the FP64 contraction order).  For each pair: max |delta alpha| at the end and the first iteration at which the
loss histories differ by more than 1e-9 relative.  One JSON line; CPU part ~30 s.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import svm_oracle as O  # noqa: E402  (checker only)
from optiml_b200.configs import make_config  # noqa: E402
from optiml_b200.ml.svm import DualSVR  # noqa: E402
from optiml_b200.ml.svm.kernels import PolyKernel  # noqa: E402
from optiml_b200.opti import Quadratic  # noqa: E402
from optiml_b200.opti.constrained import ProjectedGradient  # noqa: E402
from optiml_b200.runtime import DeviceHessian, default_context  # noqa: E402


def onset(fa, fb, tol=1e-9):
    k = min(len(fa), len(fb))
    rel = np.abs(fa[:k] - fb[:k]) / np.abs(fb[:k])
    bad = np.nonzero(rel > tol)[0]
    return int(bad[0]) if len(bad) else None


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else None
    spec, X, y = make_config('C2', n=n)
    n = len(y)
    ctx = default_context()
    # A: product path
    mA = DualSVR(kernel=PolyKernel(degree=3), epsilon=0.1, C=1).fit(X, y)
    fA, xA = np.array(mA.train_loss_history), mA.alphas_.copy()
    MA = mA.obj.device_hessian().shard_to_host()
    mA.obj.release()
    # host Gram matrix, as the reference builds it (kernels.py:91-95) + the bias term (ml/svm/_base.py:1178)
    M = O.poly_kernel(X, degree=3) + 1.0
    gram_rel = float((np.abs(MA - M) / np.abs(M)).max())
    gram_differs = int((MA != M).sum())
    q, ub = np.hstack((-y, y)) + 0.1, np.ones(2 * n)
    # B: device loop on the host matrix
    H = DeviceHessian(ctx, n, 'svr')
    block = np.zeros((H.nrows, H.ld))
    block[:, :n] = M
    ctx.h2d(H.matrix.dptr, block)
    del block
    sB = ProjectedGradient(quad=Quadratic(H, q), ub=ub, max_iter=1000).minimize()
    fB, xB = np.array(sB.f_hist), sB.x.copy()
    H.release()
    # C: oracle on the host matrix
    rC = O.projected_gradient(O.SVRBlockOperator(M), q, ub, passes=1)
    fC, xC = rC.f_hist, rC.x
    out = {'n': n, 'gram_max_rel_diff_device_vs_host': gram_rel, 'gram_entries_that_differ': gram_differs,
           'A_vs_B_same_loop_different_gram_ulps': {'max_abs_dalpha': float(np.abs(xA - xB).max()), 'onset_iter': onset(fA, fB),
                                                    'f_end': [float(fA[-1]), float(fB[-1])]},
           'B_vs_C_same_gram_bits_different_summation': {'max_abs_dalpha': float(np.abs(xB - xC).max()),
                                                         'onset_iter': onset(fB, fC), 'f_end': [float(fB[-1]), float(fC[-1])]},
           'A_vs_C': {'max_abs_dalpha': float(np.abs(xA - xC).max()), 'onset_iter': onset(fA, fC)}}
    gpath = os.path.join(ROOT, 'tests', 'golden', 'c2_full_svr_poly.npz')
    if n == 10000 and os.path.exists(gpath):
        g = np.load(gpath)
        for name, x, f in (('A', xA, fA), ('B', xB, fB), ('C', xC, fC)):
            out[name + '_vs_reference_golden'] = {'max_abs_dalpha': float(np.abs(x - g['alphas']).max()),
                                                  'onset_iter': onset(f, g['f_hist']), 'f_end': [float(f[-1]), float(g['f_hist'][-1])]}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
