"""Wall-clock time of every native call of one DualSVC fit (C4 by default), in call order: where `fit - PG loop` goes."""
import sys, time, json
sys.path.insert(0, '.')
import numpy as np
from optiml_b200 import _native as N
from optiml_b200.configs import make_config
from optiml_b200.ml.svm import DualSVC
from optiml_b200.ml.svm.kernels import GaussianKernel
from optiml_b200.runtime import default_context

cfg = sys.argv[1] if len(sys.argv) > 1 else 'C4'
spec, X, y = make_config(cfg)
ctx = default_context()
dX = ctx.upload_matrix(X)
for _ in range(2):
    m = DualSVC(kernel=GaussianKernel(), C=1).fit(X, y, X_device=dX); m.obj.release()
log = []
orig = N.call
def timed(name, *a):
    t = time.perf_counter(); r = orig(name, *a); log.append((name, (time.perf_counter() - t) * 1e3)); return r
N.call = timed
import optiml_b200.runtime as R, optiml_b200.ml.svm._base as B, optiml_b200.ml.svm.kernels as K, optiml_b200.opti.constrained._device_loop as D
for mod in (R, B, K, D):
    mod.N.call = timed
t0 = time.perf_counter()
m = DualSVC(kernel=GaussianKernel(), C=1).fit(X, y, X_device=dX)
total = (time.perf_counter() - t0) * 1e3
native = sum(ms for _, ms in log)
print(json.dumps({'config': cfg, 'fit_ms': round(total, 2), 'native_ms': round(native, 2), 'python_ms': round(total - native, 2),
                  'pg_device_ms': round(m.optimizer.device_ms, 2)}))
for name, ms in log:
    print(f'{ms:9.3f} ms  {name}')
