"""cProfile of one C4 fit on the GPU: Python-side time by function (native calls excluded by name)."""
import cProfile, io, pstats, sys
sys.path.insert(0, '.')
from optiml_b200.configs import make_config
from optiml_b200.ml.svm import DualSVC
from optiml_b200.ml.svm.kernels import GaussianKernel
from optiml_b200.runtime import default_context
spec, X, y = make_config('C4')
ctx = default_context()
dX = ctx.upload_matrix(X)
for _ in range(3):
    m = DualSVC(kernel=GaussianKernel(), C=1).fit(X, y, X_device=dX); m.obj.release()
pr = cProfile.Profile(); pr.enable()
m = DualSVC(kernel=GaussianKernel(), C=1).fit(X, y, X_device=dX)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(30); print(s.getvalue()[:6000])
