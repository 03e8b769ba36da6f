"""Runs tests/multigpu_check.py under torchrun when the box has >= 2 GPUs (skipped otherwise)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_fit_is_bitwise_equal_to_single_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs >= 2 GPUs (run tests/multigpu_check.py under torchrun on a multi-GPU box)')
    n = 2 if n < 4 else (4 if n < 8 else 8)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={n}', '--master-addr',
           '127.0.0.1', '--master-port', '29517', os.path.join(ROOT, 'tests', 'multigpu_check.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and 'MULTIGPU_CHECK PASS' in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def test_single_process_device_group_is_bitwise_equal_to_single_gpu():
    """One Python process, all visible GPUs (runtime.use_devices: no torchrun, no NCCL): alpha, intercept, loss history
    and decision values of a sharded fit equal the one-GPU fit bit for bit; sklearn's GridSearchCV runs on it unchanged."""
    import numpy as np
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip('needs >= 2 GPUs')
    from sklearn.model_selection import GridSearchCV
    from optiml_b200 import runtime
    from optiml_b200.configs import make_config
    from optiml_b200.ml.svm import DualSVC, DualSVR
    from optiml_b200.ml.svm.kernels import GaussianKernel, PolyKernel
    cases = [('C4', 8192 + 37, lambda: DualSVC(kernel=GaussianKernel(), C=1, max_iter=200)),
             ('C2', 6000, lambda: DualSVR(kernel=PolyKernel(degree=3), epsilon=0.1, C=1, max_iter=120))]

    def fit_all():
        out = []
        for cfg, n, mk in cases:
            spec, X, y = make_config(cfg, n=n)
            m = mk().fit(X, y)
            out.append((type(m.obj.device_hessian()).__name__, m.alphas_.copy(), m.intercept_,
                        np.array(m.train_loss_history), m.decision_function(X[:64])))
            m.obj.release()
        return out

    solo = fit_all()
    runtime.use_devices(list(range(min(ngpu, 8))))
    try:
        grouped = fit_all()
        spec, X, y = make_config('C4', n=6000)
        gs = GridSearchCV(DualSVC(kernel=GaussianKernel(), max_iter=60), {'C': [0.5, 2.0]}, cv=2).fit(X, y)
        assert type(gs.best_estimator_.obj.device_hessian()).__name__ == 'GroupHessian'
    finally:
        runtime.use_devices(None)
    for s, g in zip(solo, grouped):
        assert s[0] == 'DeviceHessian' and g[0] == 'GroupHessian'
        assert np.array_equal(s[1], g[1]) and s[2] == g[2] and np.array_equal(s[3], g[3]) and np.array_equal(s[4], g[4])
