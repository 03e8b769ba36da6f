import cProfile, pstats, sys, time, io
sys.path.insert(0, '.')
import numpy as np
from optiml_b200.configs import make_config
from optiml_b200.ml.svm import DualSVC
from optiml_b200.ml.svm.kernels import GaussianKernel
spec, X, y = make_config('C4')
for i in range(2):
    m = DualSVC(kernel=GaussianKernel(), C=1).fit(X, y); m.obj.release()
for prof in (False, True):
    m = DualSVC(kernel=GaussianKernel(), C=1)
    m.profile_matvec = prof
    pr = cProfile.Profile()
    t = time.perf_counter(); pr.enable(); m.fit(X, y); pr.disable(); dt = time.perf_counter() - t
    print('profile_matvec', prof, 'fit wall', dt, m.fit_times_, 'pg device ms', m.optimizer.device_ms)
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(28); print(s.getvalue()[:6000])
    m.obj.release()
