"""Box-constrained QP solver base (reference: optiml/opti/constrained/_base.py:10-85)."""
import numpy as np

from .. import Optimizer, Quadratic
from .._base import OptimizationFunction


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


class AugmentedLagrangianQuadratic(Quadratic):
    r"""Augmented-Lagrangian relaxation of  min x'Qx/2 + q'x : A x = b, lb <= x <= ub
    (reference: optiml/opti/constrained/_base.py:224-410):

        L(x; mu, lambda) = x'Qx/2 + q'x + mu (A x - b) + lambda' (G x - h) + rho/2 (|A x - b|^2 + |max(G x - h, 0)|^2)

    with G' = [-I  I], h = [-lb  ub].  The reference stores AG = [A; G] as a dense (2n+1) x n matrix and pays a dense
    n x n product per gradient; here the two identity blocks are implicit, ``Q`` stays in HBM (``primal`` holds the
    device handle) and one streaming pass gives both x'Qx and Qx.  The multipliers ``dual_x = [mu, lambda_lb,
    lambda_ub]`` are updated by the optimiser, once per iteration (optiml/opti/_base.py:129-149).

    Scope: at most ONE equality row (the SVM duals have exactly one, ml/svm/_base.py:644-649, 1194-1199), both bounds
    given, no general ``G x <= h`` block -- anything else raises ``NotImplementedError``.
    """

    def __init__(self, primal, A=None, b=None, G=None, h=None, lb=None, ub=None, rho=1):
        if not isinstance(primal, Quadratic):
            raise TypeError(f'{primal} is not an allowed quadratic function')
        if G is None and h is not None:
            raise ValueError('incomplete inequality constraint (missing G)')
        if G is not None and h is None:
            raise ValueError('incomplete inequality constraint (missing h)')
        if A is None and b is not None:
            raise ValueError('incomplete equality constraint (missing A)')
        if A is not None and b is None:
            raise ValueError('incomplete equality constraint (missing b)')
        if not rho > 0:
            raise ValueError('rho must be must > 0')
        if G is not None:
            raise NotImplementedError('general inequality constraints G x <= h are outside the SVM dual path')
        if lb is None or ub is None:
            raise NotImplementedError('both lb and ub are required on the SVM dual path')
        # share the primal's storage: no second copy of Q (the reference copies it, constrained/_base.py:242)
        OptimizationFunction.__init__(self, primal.ndim)
        self._device, self._host_Q = primal._device, primal._host_Q
        self.q = primal.q
        self.primal = primal
        self.A = np.atleast_2d(A).astype(float) if A is not None else None
        if self.A is not None and self.A.shape != (1, self.ndim):
            raise NotImplementedError('exactly one equality row of length ndim is supported')
        self.b = np.atleast_1d(np.asarray(b, dtype=float)) if b is not None else None
        self.lb = np.asarray(lb, dtype=float)
        self.ub = np.asarray(ub, dtype=float)
        if self.lb.size != self.ndim or self.ub.size != self.ndim:
            raise ValueError('bounds size does not match with Q')
        self.rho = rho
        self.n_eq = 1 if self.A is not None else 0
        self.dual_x = np.zeros(self.n_eq + 2 * self.ndim)  # mu_lmbda, constrained/_base.py:307
        self.past_dual_x = self.dual_x.copy()

    # Q lives with the primal (it may be uploaded lazily there)
    def device_hessian(self, ctx=None):
        return self.primal.device_hessian(ctx)

    @property
    def Q(self):
        return self.primal.Q

    def release(self):
        self.primal.release()

    # dense forms of the reference's attributes, materialised on request only
    @property
    def G(self):
        return np.concatenate((-np.eye(self.ndim), np.eye(self.ndim)), axis=0)

    @property
    def h(self):
        return np.concatenate((-self.lb, self.ub))

    @property
    def AG(self):
        return np.concatenate((self.A, self.G)) if self.n_eq else self.G

    @property
    def bh(self):
        return np.concatenate((self.b, self.h)) if self.n_eq else self.h

    def f_star(self):
        return np.inf  # the reference asks cvxopt through qpsolvers (constrained/_base.py:316-325); not a dependency here

    def constraints(self, x):
        """AG @ x - bh with the identity blocks applied implicitly (constrained/_base.py:327-335)."""
        x = np.asarray(x, dtype=float)
        parts = [self.A @ x - self.b] if self.n_eq else []
        return np.concatenate(parts + [-x - (-self.lb), x - self.ub])

    def _clipped(self, c):
        cc = c.copy()
        cc[self.n_eq:] = np.clip(c[self.n_eq:], a_min=0, a_max=None)
        return cc

    def function(self, x):
        """constrained/_base.py:337-352"""
        return self.function_jacobian(x)[0]

    def jacobian(self, x):
        """constrained/_base.py:373-393"""
        return self.function_jacobian(x)[1]

    def function_jacobian(self, x):
        """constrained/_base.py:395-407; one pass over Q on the device, the rest is O(n) on the host."""
        x = np.asarray(x, dtype=float)
        n, n_eq, rho = self.ndim, self.n_eq, self.rho
        Qx = self.primal._Qx(x)
        c = self.constraints(x)
        cc = self._clipped(c)
        fun = 0.5 * (x @ Qx) + self.q @ x + self.dual_x @ c + 0.5 * rho * np.linalg.norm(cc) ** 2
        act = cc != 0
        act_lb, act_ub = act[n_eq:n_eq + n], act[n_eq + n:]
        jac = Qx + self.q
        t2 = -self.dual_x[n_eq:n_eq + n] + self.dual_x[n_eq + n:]
        t3 = rho * (act_lb * x + act_ub * x)
        t4 = rho * (act_lb * self.lb + act_ub * self.ub)
        if n_eq:
            a = self.A[0]
            t2 = self.dual_x[0] * a + t2
            if act[0]:
                t3 = rho * a * (a @ x) + t3
                t4 = rho * self.b[0] * a + t4
        return fun, jac + t2 + t3 - t4

    def hessian(self, x):
        raise NotImplementedError('the Hessian of the augmented Lagrangian is not used on the SVM dual path')
