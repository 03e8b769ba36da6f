"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys
sys.path.insert(0, '.')
import numpy as np
from optiml_b200.configs import make_config
from optiml_b200.ml.svm import DualSVC, DualSVR
from optiml_b200.ml.svm.kernels import GaussianKernel, PolyKernel, LinearKernel
from optiml_b200.opti import Quadratic
from optiml_b200.opti.constrained import ProjectedGradient

spec, X, y = make_config('C1', n=333)
m = DualSVC(kernel=GaussianKernel(), C=1, max_iter=25).fit(X, y)
print('svc', m.optimizer.iter, m.optimizer.status, len(m.support_), m.decision_function(X[:7]))
spec, X, y = make_config('C2', n=201)
m = DualSVR(kernel=PolyKernel(degree=3), epsilon=0.1, max_iter=25).fit(X, y)
print('svr', m.optimizer.iter, m.optimizer.status, len(m.support_), m.predict(X[:5]))
rng = np.random.default_rng(0)
A = rng.standard_normal((37, 9)); B = rng.standard_normal((5, 9))
print('kern', LinearKernel()(A, B).shape, GaussianKernel()(A).shape)
G = rng.standard_normal((70, 67)); Q = G.T @ G
opt = ProjectedGradient(quad=Quadratic(Q, rng.standard_normal(67)), ub=np.ones(67), callback=lambda o: None, max_iter=10).minimize()
print('pg', opt.iter, opt.status)
