"""K6 at scale: decision_function of a fitted C4 model on its own training set (nSV ~ 49 000 support vectors x m = 50 000
test points x d = 128): K1 blocks of K(X_chunk, SV) (<= 2 GiB each) contracted with dual_coef_ by K2.  One JSON line:
seconds per call, Gram TFLOP/s (2 m nSV d), streamed GB/s (8 m nSV written + read back), with the support vectors
resident in HBM (what fit leaves behind) and re-uploaded from the host copy (the round-1 path)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from optiml_b200.configs import make_config  # noqa: E402
from optiml_b200.ml.svm import DualSVC  # noqa: E402
from optiml_b200.ml.svm.kernels import GaussianKernel  # noqa: E402
from optiml_b200.runtime import default_context  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else None
    spec, X, y = make_config('C4', n=n)
    m = DualSVC(kernel=GaussianKernel(), C=1, max_iter=100).fit(X, y)
    m.obj.release()
    default_context().trim()
    nsv, mtest, d = len(m.support_), X.shape[0], X.shape[1]

    def timed(reps=3):
        m.decision_function(X[:256])
        ts = []
        for _ in range(reps):
            t = time.perf_counter()
            out = m.decision_function(X)
            ts.append(time.perf_counter() - t)
        return min(ts), out

    t_dev, dec_dev = timed()
    m.support_vectors_ = m.support_vectors_.copy()   # drops the device copy: host-SV path
    t_host, dec_host = timed()
    acc = float(np.mean(m.predict(X) == y))
    for name, t in (('device_sv', t_dev), ('host_sv', t_host)):
        print(json.dumps({'case': f'decision_function C4 train set: m={mtest} nSV={nsv} d={d}', 'support_vectors': name,
                          'seconds': round(t, 4), 'gram_tflops': round(2.0 * mtest * nsv * d / t / 1e12, 2),
                          'block_gbps_write_plus_read': round(2 * 8.0 * mtest * nsv / t / 1e9, 1),
                          'bitwise_equal_paths': bool(np.array_equal(dec_dev, dec_host)), 'train_accuracy': acc}), flush=True)


if __name__ == '__main__':
    main()
