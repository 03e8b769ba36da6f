"""CPU oracle for the augmented-Lagrangian dual path  --  TEST INFRASTRUCTURE ONLY.

NumPy FP64 restatement of what the reference does for
``SVC/SVR(dual=True, optimizer=<StochasticOptimizer>, reg_intercept in {True, False})``:

  * ``AugmentedLagrangianQuadratic`` (optiml/opti/constrained/_base.py:224-410): the box (and, with
    ``reg_intercept=False``, the equality ``A x = b``) are relaxed with multipliers and a quadratic penalty,
  * the multiplier update / optimality test of ``Optimizer.check_lagrangian_dual_optimality``
    (optiml/opti/_base.py:129-149) and the Lagrangian branch of ``Optimizer.callback`` (:96-117),
  * the full-batch loop of the stochastic optimisers (optiml/opti/unconstrained/stochastic/adagrad.py:84-125 and
    its siblings).

The reference materialises ``AG = [A; -I; I]`` ((2n+1) x n, dense) and evaluates the penalty gradient with a dense
n x n product ``(rho AG[idx]' AG[idx]) x`` (O(n^3) per iteration).  The restatement uses the structure of AG
(two identity blocks and one dense row), which is the same arithmetic up to the summation order of BLAS:
it is pinned to the real reference at 1e-9 on the golden vectors of ``tests/golden/make_golden_al.py`` (small n,
where the reference can run), not bit-exactly.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module.
"""
import numpy as np

from . import svm_oracle as O


class ALResult:
    __slots__ = ('x', 'f_x', 'g_x', 'iter', 'status', 'f_hist', 'pf_hist', 'dual_x', 'primal_f_x', 'dgap')


def _constraints(x, A, b, lb, ub):
    """AG @ x - bh for AG = [A; -I; I], bh = [b; -lb; ub] (constrained/_base.py:248-263, 327-335)."""
    parts = []
    if A is not None:
        parts.append(np.atleast_1d(A @ x - b))
    parts.append(-x - (-lb))
    parts.append(x - ub)
    return np.concatenate(parts)


def al_function_jacobian(Qx, x, q, dual_x, c, A, b, lb, ub, rho):
    """constrained/_base.py:395-407 with the structure of AG; ``Qx`` = Q @ x, ``c`` = constraints(x)."""
    n = len(x)
    n_eq = 0 if A is None else 1
    cc = c.copy()
    cc[n_eq:] = np.clip(c[n_eq:], a_min=0, a_max=None)
    pf = 0.5 * (x @ Qx) + q @ x                                   # opti/_base.py:282
    f = pf + dual_x @ c + 0.5 * rho * np.linalg.norm(cc) ** 2
    act = cc != 0
    act_lb, act_ub = act[n_eq:n_eq + n], act[n_eq + n:]
    lam_lb, lam_ub = dual_x[n_eq:n_eq + n], dual_x[n_eq + n:]
    t2 = -lam_lb + lam_ub                                         # dual_x @ AG
    t3 = rho * (act_lb * x + act_ub * x)                          # rho AG[idx]' AG[idx] x
    t4 = rho * (act_lb * lb + act_ub * ub)                        # rho bh[idx] @ AG[idx]
    if n_eq:
        t2 = dual_x[0] * A + t2
        if act[0]:
            t3 = rho * A * (A @ x) + t3
            t4 = rho * b * A + t4
    g = (Qx + q) + t2 + t3 - t4
    return f, g, pf


RULES = ('adagrad', 'sgd', 'rmsprop', 'adadelta', 'adam', 'amsgrad', 'adamax')


def al_stochastic(Qmul, q, lb, ub, x0, A=None, b=0., rho=1., rule='adagrad', step_size=1., offset=None,
                  momentum_type='none', momentum=0.9, decay=0.9, beta1=0.9, beta2=0.999, tol=1e-8, epochs=1000):
    """Full-batch stochastic optimiser on the augmented Lagrangian (one multiplier update per iteration).

    ``Qmul(x)`` returns Q @ x.  ``step_size`` is a scalar or a sequence with one value per iteration.
    Loop shape: stochastic/adagrad.py:84-125 (and gradient_descent.py, rmsprop.py, adadelta.py, adam.py,
    amsgrad.py, adamax.py for the other rules); multiplier update and optimality test: opti/_base.py:129-149.
    """
    if rule not in RULES:
        raise ValueError(rule)
    if offset is None:
        offset = 1e-6 if rule == 'adadelta' else 1e-8
    q, lb, ub = (np.asarray(v, dtype=np.float64) for v in (q, lb, ub))
    A = None if A is None else np.asarray(A, dtype=np.float64)
    x = np.array(x0, dtype=np.float64)
    n = len(x)
    n_eq = 0 if A is None else 1
    dual_x = np.zeros(n_eq + 2 * n)
    s1, s2 = np.zeros(n), np.zeros(n)          # optimiser state (gms / moments / sms)
    if rule == 'rmsprop':
        s1 = np.ones(n)                        # rmsprop.py:93 starts the moving mean of g^2 at one
    amax = 0.                                  # amsgrad running maximum
    step = np.zeros(n) if momentum_type != 'none' or rule == 'adadelta' else 0
    lrs = np.broadcast_to(np.asarray(step_size, dtype=np.float64), (epochs,)) if np.ndim(step_size) == 0 \
        else np.asarray(step_size, dtype=np.float64)
    res = ALResult()
    res.status, res.iter = 'unknown', 0
    f_hist, pf_hist = [], []
    epoch = 0
    c = None
    while True:
        jump = 0
        if momentum_type == 'nesterov' and rule != 'adagrad' and rule != 'adadelta':
            jump = momentum * step
            x += jump
            c = None                            # constraints cache is keyed on x (constrained/_base.py:327-335)
        if c is None:
            c = _constraints(x, A, b, lb, ub)
        f_x, g, pf = al_function_jacobian(Qmul(x), x, q, dual_x, c, A, b, lb, ub, rho)
        f_hist.append(f_x)
        pf_hist.append(pf)
        past_x = x.copy()                       # opti/_base.py:117
        epoch += 1
        if epoch >= epochs:
            res.status = 'stopped'
            break
        lr = lrs[res.iter]
        d = -g
        if rule == 'adagrad':
            s1 += g ** 2
            step = lr * d / np.sqrt(s1 + offset)
            x += step
        elif rule == 'adadelta':
            s1 = decay * s1 + (1. - decay) * g ** 2
            step = lr * d * (np.sqrt(s2 + offset) / np.sqrt(s1 + offset))
            x += step
        else:
            if rule == 'sgd':
                step2 = lr * d
            elif rule == 'rmsprop':
                s1 = decay * s1 + (1. - decay) * g ** 2
                step2 = lr * d / np.sqrt(s1 + offset)
            else:
                t = res.iter + 1
                s1 = beta1 * s1 + (1. - beta1) * d
                if rule == 'adamax':
                    s2 = np.maximum(beta2 * s2, np.abs(g))
                    step2 = lr * (s1 / (1. - beta1 ** t)) / (s2 + offset)
                else:
                    s2 = beta2 * s2 + (1. - beta2) * g ** 2
                    if rule == 'adam':
                        step2 = lr * (s1 / (1. - beta1 ** t)) / (np.sqrt(s2 / (1. - beta2 ** t)) + offset)
                    else:
                        amax = np.maximum(s2, amax)
                        step2 = lr * s1 / (np.sqrt(amax) + offset)
            if momentum_type == 'polyak':
                # gradient_descent.py adds `lr d + m step`, the adaptive rules `m step + step2`
                step = step2 + momentum * step if rule == 'sgd' else momentum * step + step2
                x += step
            elif momentum_type == 'nesterov':
                x += step2
                step = jump + step2
            else:
                step = step2
                x += step
        # opti/_base.py:129-149
        c = _constraints(x, A, b, lb, ub)
        past_dual = dual_x.copy()
        dual_x += rho * c
        dual_x[n_eq:] = np.clip(dual_x[n_eq:], a_min=0, a_max=None)
        if (np.linalg.norm(dual_x - past_dual) + np.linalg.norm(x - past_x) <= tol) or np.linalg.norm(c) <= tol:
            res.status = 'optimal'
            break
        if rule == 'adadelta':
            s2 = decay * s2 + (1. - decay) * step ** 2
        res.iter += 1
    res.x, res.f_x, res.g_x, res.dual_x = x, f_x, g, dual_x
    res.f_hist, res.pf_hist = np.array(f_hist), np.array(pf_hist)
    res.primal_f_x = pf
    res.dgap = abs((pf - f_x) / max(abs(pf), 1))
    return res


def start_point(n, random_state):
    """opti/_base.py:36-42, 58-60: the primal variable starts at a uniform random point."""
    return (np.random.RandomState(random_state).uniform if random_state is not None else np.random.uniform)(size=n)


def svc_dual_al_fit(X, y, kind='gaussian', C=1., rho=1., reg_intercept=True, degree=3, gamma='scale', coef0=0.,
                    max_iter=1000, tol=1e-4, learning_rate=1., random_state=None, x0=None, **opt):
    """optiml/ml/svm/_base.py:547-556, 638-725, 867-880 (loss=hinge, dual=True, optimizer=<StochasticOptimizer>)."""
    X = np.asarray(X, dtype=np.float64)
    classes, ys = O.binarize_labels(y)
    n = len(ys)
    K = O.kernel_matrix(kind, X, None, degree=degree, gamma=gamma, coef0=coef0)
    yy = np.outer(ys, ys)
    Q = K * yy
    if reg_intercept:
        Q += yy
    q = -np.ones(n)
    lb, ub = np.zeros(n), np.ones(n) * C
    A = None if reg_intercept else ys.astype(float)
    x0 = start_point(n, random_state) if x0 is None else x0
    al = al_stochastic(lambda v: Q @ v, q, lb, ub, x0, A=A, b=0., rho=rho, step_size=learning_rate, tol=tol,
                       epochs=max_iter, **opt)
    out = O.FitResult()
    out.classes_, out.al, out.alphas_ = classes, al, al.x
    sv = al.x > 1e-6
    out.support_ = np.arange(n)[sv]
    out.support_vectors_ = X[sv]
    sv_y, a = ys[sv], al.x[sv]
    out.dual_coef_ = a * sv_y
    out.coef_ = out.dual_coef_ @ out.support_vectors_ if kind == 'linear' else None
    bsum = 0.
    for i in range(len(a)):
        bsum += sv_y[i]
        bsum -= np.sum(out.dual_coef_ * K[out.support_[i], sv])
    out.intercept_ = bsum / len(a)
    out.kernel = dict(kind=kind, degree=degree, gamma=gamma, coef0=coef0)
    out.K, out.Q = K, Q
    return out


def svr_dual_al_fit(X, y, kind='linear', C=1., epsilon=0.1, rho=1., reg_intercept=True, degree=3, gamma='scale',
                    coef0=0., max_iter=1000, tol=1e-4, learning_rate=1., random_state=None, x0=None, **opt):
    """optiml/ml/svm/_base.py:1091-1104, 1188-1270, 1423-1437 (loss=epsilon_insensitive, dual=True,
    optimizer=<StochasticOptimizer>)."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n = len(y)
    K = O.kernel_matrix(kind, X, None, degree=degree, gamma=gamma, coef0=coef0)
    Q = np.vstack((np.hstack((K, -K)), np.hstack((-K, K))))
    q = np.hstack((-y, y)) + epsilon
    lb, ub = np.zeros(2 * n), np.ones(2 * n) * C
    e = np.hstack((np.ones(n), -np.ones(n)))
    if reg_intercept:
        Q += np.outer(e, e)
    x0 = start_point(2 * n, random_state) if x0 is None else x0
    al = al_stochastic(lambda v: Q @ v, q, lb, ub, x0, A=None if reg_intercept else e, b=0., rho=rho,
                       step_size=learning_rate, tol=tol, epochs=max_iter, **opt)
    out = O.FitResult()
    out.al, out.alphas_ = al, al.x
    ap, an = np.split(al.x, 2)
    sv = np.logical_or(ap > 1e-6, an > 1e-6)
    out.support_ = np.arange(n)[sv]
    out.support_vectors_ = X[sv]
    sv_y, ap, an = y[sv], ap[sv], an[sv]
    out.dual_coef_ = ap - an
    out.coef_ = out.dual_coef_ @ out.support_vectors_ if kind == 'linear' else None
    bsum = 0.
    for i in range(len(ap)):
        bsum += sv_y[i]
        bsum -= np.sum(out.dual_coef_ * K[out.support_[i], sv])
    bsum -= epsilon
    out.intercept_ = bsum / len(ap)
    out.kernel = dict(kind=kind, degree=degree, gamma=gamma, coef0=coef0)
    out.K, out.Q = K, Q
    return out
