"""Device timing of the augmented-Lagrangian dual path (widening 8f-3) at the headline size:
SVC(hinge, Gaussian, dual=True, optimizer=AdaGrad) on C4 (n = 50 000, d = 128), both formulations.
    python scripts/bench_al.py [--n N] [--iters K]"""
import argparse
import json
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=None)
    ap.add_argument('--iters', type=int, default=300)
    args = ap.parse_args()
    from sklearn.exceptions import ConvergenceWarning
    from optiml_b200.configs import make_config
    from optiml_b200.ml.svm import SVC
    from optiml_b200.ml.svm.kernels import GaussianKernel
    from optiml_b200.ml.svm.losses import hinge
    from optiml_b200.opti.unconstrained.stochastic import AdaGrad, Adam
    warnings.simplefilter('ignore', ConvergenceWarning)
    spec, X, y = make_config('C4', n=args.n)
    n = len(y)
    for name, opt, kw in (('adagrad', AdaGrad, dict(learning_rate=1.)),
                          ('adam_nesterov', Adam, dict(learning_rate=0.001, momentum_type='nesterov', momentum=0.5))):
        for ri in (True, False):
            for rep in range(2):  # first fit warms the buffer pools
                m = SVC(loss=hinge, kernel=GaussianKernel(), C=1, reg_intercept=ri, dual=True, optimizer=opt,
                        max_iter=args.iters, random_state=0, **kw)
                m.profile_matvec = True
                t0 = time.perf_counter()
                m.fit(X, y)
                fit_s = time.perf_counter() - t0
                o = m.optimizer
                m.obj.release()
            passes = max(o.profile_samples, 1)
            print(json.dumps(dict(case=f'C4 n={n} SVC {name} reg_intercept={ri}', iters=o.iter + 1, status=o.status,
                                  fit_s=round(fit_s, 4), loop_ms=round(o.device_ms, 2),
                                  its_per_s=round((o.iter + 1) / (o.device_ms / 1e3), 1),
                                  matvec_us=round(1e3 * o.matvec_ms / passes, 2), vector_us=round(1e3 * o.vector_ms / passes, 2),
                                  hbm_gbps=round(8.0 * n * n * o.q_passes / (o.device_ms / 1e3) / 1e9, 1),
                                  primal_cost=float(m.train_loss_history[-1]), n_sv=int(len(m.support_)))), flush=True)


if __name__ == '__main__':
    main()
