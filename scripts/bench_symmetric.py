"""A/B of the opt-in symmetric pass (K2s) against the default full pass (K2) on a BASELINE config, through the public API:
the same fits, device-resident X, each against the real reference's golden run (C4 / C1).  One JSON line per mode.

    python scripts/bench_symmetric.py [--config C4] [--steps 3] [--warmup 1]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', default='C4')
    ap.add_argument('--n', type=int, default=None)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=1)
    ap.add_argument('--max-iter', type=int, default=1000)
    args = ap.parse_args()
    from optiml_b200 import runtime
    from optiml_b200.configs import make_config
    from optiml_b200.ml.svm import DualSVC
    from optiml_b200.ml.svm.kernels import GaussianKernel, LinearKernel
    import bench
    ctx = runtime.default_context()
    spec, X, y = make_config(args.config, n=args.n)
    n, d = X.shape
    kernel = LinearKernel() if spec['kernel'] == 'linear' else GaussianKernel()
    dX = ctx.upload_matrix(X)
    ref = None
    for sym in (False, True):
        runtime.use_symmetric_pass(sym)
        fits = []
        for i in range(args.warmup + args.steps):
            ctx.sync()
            t0 = time.perf_counter()
            m = DualSVC(kernel=kernel, C=1, max_iter=args.max_iter)
            m.profile_matvec = True
            m.fit(X, y, X_device=dX)
            ctx.sync()
            dt = time.perf_counter() - t0
            if i >= args.warmup:
                fits.append((dt, m.optimizer.iter, m.optimizer.device_ms, m.optimizer.matvec_ms, m.optimizer.vector_ms,
                             m.optimizer.profile_samples, m.optimizer.q_passes))
            m.obj.release()
        assert m.optimizer.symmetric_pass is sym
        tot_s = sum(f[0] for f in fits)
        iters = sum(f[1] for f in fits)
        samples = max(1, sum(f[5] for f in fits))
        mv_ms = sum(f[3] for f in fits) / samples
        line = {'mode': 'symmetric pass (K2s: upper triangle)' if sym else 'full pass (K2)', 'config': args.config, 'n': n, 'd': d,
                'fits': len(fits), 'fit_s': tot_s / len(fits), 'pg_its_per_s_whole_fit': iters / tot_s,
                'pg_its_per_s_loop': iters / (sum(f[2] for f in fits) / 1e3),
                'product_ms_per_iteration': mv_ms, 'vector_phase_us': 1e3 * sum(f[4] for f in fits) / samples,
                'full_matrix_bytes_per_product_ms_GBps': 8.0 * n * n / mv_ms / 1e6,
                'status': m.optimizer.status, 'f_x': m.optimizer.f_x, 'n_sv': int(len(m.support_))}
        ns = argparse.Namespace(config=args.config)
        line['parity_vs_reference_golden'] = bench.parity_block(ns, m, n) if args.max_iter == 1000 else None
        if ref is None:
            ref = (m.alphas_.copy(), m.support_.copy(), m.intercept_)
        else:
            line['vs_full_pass'] = {'max_abs_dalpha': float(np.abs(m.alphas_ - ref[0]).max()),
                                    'same_support': bool(np.array_equal(m.support_, ref[1])),
                                    'intercept_abs': float(abs(m.intercept_ - ref[2]))}
        print(json.dumps(line), flush=True)
    runtime.use_symmetric_pass(False)


if __name__ == '__main__':
    main()
