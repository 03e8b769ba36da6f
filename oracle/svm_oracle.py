"""CPU oracle for the kernel-SVM dual training path  --  TEST INFRASTRUCTURE ONLY.

This module is a NumPy FP64 restatement of the reference algorithm (dmeoli/optiml 1.8,
pure Python/NumPy).  It exists so that the CUDA path can be checked on machines where
``/root/reference`` is absent (the GPU box).  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; nothing under
``optiml_b200/`` does, and the product path fails loudly when the CUDA library is missing.

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the *real* reference (imported
from ``/root/reference`` with four stub modules for absent third-party packages that the path
never calls) and stores its outputs under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
checks every function below against those vectors (bit-exact where the arithmetic is the same
sequence of NumPy calls, <=1e-12 otherwise).

Each function cites the reference lines it follows (paths relative to ``/root/reference``).
Arithmetic done by un-vendored third-party code is restated from its published algorithm:
  * scikit-learn 1.9.0 ``euclidean_distances`` (``sklearn/metrics/pairwise.py`` ``_euclidean_distances``):
    ``D = -2 X Y^T; D += |x_i|^2; D += |y_j|^2; D = max(D, 0); diag(D) = 0 when Y is X``.
  * scikit-learn ``LabelBinarizer(neg_label=-1)``: sorted classes, classes_[1] -> +1, classes_[0] -> -1.
  * NumPy 2.3.5 ``@``/``dot``/``exp``/``power``/``var``/``linalg.norm``.
"""
import numpy as np

# ----------------------------------------------------------------------------- kernels


def resolve_gamma(gamma, X):
    """optiml/ml/svm/kernels.py:93-94, 127-128: 'scale' -> 1/(d*X.var()), 'auto' -> 1/d,
    computed from the FIRST argument of every call."""
    if isinstance(gamma, str):
        if gamma == 'scale':
            return 1. / (X.shape[1] * X.var())
        if gamma == 'auto':
            return 1. / X.shape[1]
        raise ValueError(f'unknown gamma type {gamma}')
    return gamma


def linear_kernel(X, Y=None):
    """optiml/ml/svm/kernels.py:49-51  K = X Y^T."""
    X = np.asarray(X, dtype=np.float64)
    Y = X if Y is None else np.asarray(Y, dtype=np.float64)
    return X @ Y.T


def poly_kernel(X, Y=None, degree=3, gamma='scale', coef0=0.):
    """optiml/ml/svm/kernels.py:91-95  K = (gamma X Y^T + coef0) ** degree."""
    X = np.asarray(X, dtype=np.float64)
    Y = X if Y is None else np.asarray(Y, dtype=np.float64)
    g = resolve_gamma(gamma, X)
    return (g * (X @ Y.T) + coef0) ** degree


def squared_distances(X, Y=None):
    """sklearn 1.9.0 pairwise._euclidean_distances (float64 branch), squared=True."""
    X = np.asarray(X, dtype=np.float64)
    same = Y is None or Y is X
    Y = X if same else np.asarray(Y, dtype=np.float64)
    XX = np.einsum('ij,ij->i', X, X)[:, None]
    YY = XX.T if same else np.einsum('ij,ij->i', Y, Y)[None, :]
    D = -2 * (X @ Y.T)
    D += XX
    D += YY
    np.maximum(D, 0, out=D)
    if same:
        np.fill_diagonal(D, 0)
    return D


def gaussian_kernel(X, Y=None, gamma='scale'):
    """optiml/ml/svm/kernels.py:125-129  K = exp(-gamma * ||x - y||^2)."""
    X = np.asarray(X, dtype=np.float64)
    g = resolve_gamma(gamma, X)
    return np.exp(-g * squared_distances(X, Y))


def manhattan_distances(X, Y=None):
    """sklearn manhattan_distances (dense) = scipy cdist 'cityblock': s += |u_k - v_k|, k ascending."""
    X = np.asarray(X, dtype=np.float64)
    Y = X if Y is None else np.asarray(Y, dtype=np.float64)
    D = np.zeros((X.shape[0], Y.shape[0]))
    for k in range(X.shape[1]):
        D += np.abs(X[:, k, None] - Y[None, :, k])
    return D


def laplacian_kernel(X, Y=None, gamma='scale'):
    """optiml/ml/svm/kernels.py:159-163  K = exp(-gamma * ||x - y||_1)  (widening 8f-2)."""
    X = np.asarray(X, dtype=np.float64)
    g = resolve_gamma(gamma, X)
    return np.exp(-g * manhattan_distances(X, Y))


def sigmoid_kernel(X, Y=None, gamma='scale', coef0=0.):
    """optiml/ml/svm/kernels.py:197-201  K = tanh(gamma X Y^T + coef0)  (widening 8f-2)."""
    X = np.asarray(X, dtype=np.float64)
    Y = X if Y is None else np.asarray(Y, dtype=np.float64)
    g = resolve_gamma(gamma, X)
    return np.tanh(g * (X @ Y.T) + coef0)


def kernel_matrix(kind, X, Y=None, degree=3, gamma='scale', coef0=0.):
    if kind == 'laplacian':
        return laplacian_kernel(X, Y, gamma=gamma)
    if kind == 'sigmoid':
        return sigmoid_kernel(X, Y, gamma=gamma, coef0=coef0)
    if kind == 'linear':
        return linear_kernel(X, Y)
    if kind == 'poly':
        return poly_kernel(X, Y, degree=degree, gamma=gamma, coef0=coef0)
    if kind == 'gaussian':
        return gaussian_kernel(X, Y, gamma=gamma)
    raise ValueError(kind)


# ----------------------------------------------------------------------------- projected gradient


class PGResult:
    __slots__ = ('x', 'f_x', 'g_x', 'iter', 'status', 'f_hist', 'ng_hist', 'n_clipped')


class SVRBlockOperator:
    """[[M, -M], [-M, M]] @ v without the 2n x 2n matrix (optiml/ml/svm/_base.py:1098-1099 builds it densely): for
    sizes where three passes over the materialised block matrix are impractical.  ``passes=1`` only."""

    def __init__(self, M):
        self.M = np.asarray(M, dtype=np.float64)
        self.n = self.M.shape[0]

    def __matmul__(self, v):
        w = self.M @ (v[:self.n] - v[self.n:])
        return np.hstack((w, -w))


def projected_gradient(Q, q, ub, lb=None, x0=None, eps=1e-6, max_iter=1000, passes=3, callback=None):
    """optiml/opti/constrained/projected_gradient.py:76-143 with the start point of
    optiml/opti/constrained/_base.py:59-65 (lb = 0, x0 = (lb+ub)/2).

    ``passes=3`` evaluates f, the gradient and d'Qd exactly as the reference does
    (``opti/_base.py:282, 291`` and ``projected_gradient.py:121``): three products with Q per
    iteration.  ``passes=1`` is the algebraically equivalent single-product form used for sizes
    where three passes over a host-resident Q are impractical: ``g <- g + t*(Q d)``,
    ``f = x.(g+q)/2``; it is validated against ``passes=3`` in the tests (|dx| ~ 1e-14).

    ``callback(k, x, f, ng)`` mirrors the per-iteration callback point
    (``projected_gradient.py:95-98``: after the norm, before the stopping tests).
    """
    if not isinstance(Q, SVRBlockOperator):
        Q = np.asarray(Q, dtype=np.float64)
    elif passes != 1:
        raise ValueError('the block operator offers the single-pass form only')
    q = np.asarray(q, dtype=np.float64)
    ub = np.asarray(ub, dtype=np.float64)
    lb = np.zeros_like(ub) if lb is None else np.asarray(lb, dtype=np.float64)
    x = ((lb + ub) / 2) if x0 is None else np.array(x0, dtype=np.float64)
    res = PGResult()
    res.status, res.iter, res.n_clipped = 'unknown', 0, 0
    f_hist, ng_hist = [], []
    g = None
    if passes == 1:
        g = Q @ x + q
    while True:
        if passes == 3:
            f_x = 0.5 * x @ Q @ x + q @ x
            g = Q @ x + q
        else:
            f_x = 0.5 * (x @ (g + q))
        d = -g
        d[np.logical_and(ub - x <= 1e-12, d > 0)] = 0
        d[np.logical_and(x - lb <= 1e-12, d < 0)] = 0
        ng = np.sqrt(d.dot(d))
        f_hist.append(f_x)
        ng_hist.append(ng)
        if callback is not None:
            callback(res.iter, x, f_x, ng)
        if ng <= eps:
            res.status = 'optimal'
            break
        if res.iter >= max_iter:
            res.status = 'stopped'
            break
        pos = d > 0
        max_t = np.min((ub[pos] - x[pos]) / d[pos]) if pos.any() else np.inf
        neg = d < 0
        if neg.any():
            max_t = min(max_t, np.min((lb[neg] - x[neg]) / d[neg]))
        w = Q @ d if passes == 1 else d.dot(Q)
        den = w.dot(d)
        if den <= 1e-16:
            t = max_t
        else:
            t = min(-g.dot(d) / den, max_t)
            res.n_clipped += int(t == max_t)
        x += t * d
        if passes == 1:
            g = g + t * w
        res.iter += 1
    res.x, res.f_x, res.g_x = x, f_x, g
    res.f_hist, res.ng_hist = np.array(f_hist), np.array(ng_hist)
    return res


def frank_wolfe(Q, q, ub, lb=None, x0=None, t=0., eps=1e-6, max_iter=1000, passes=3, callback=None):
    """optiml/opti/constrained/frank_wolfe.py:88-165 (start point opti/constrained/_base.py:59-65).
    ``passes`` as in :func:`projected_gradient`.  ``ng_hist`` holds the relative gap."""
    Q = np.asarray(Q, dtype=np.float64)
    q = np.asarray(q, dtype=np.float64)
    ub = np.asarray(ub, dtype=np.float64)
    lb = np.zeros_like(ub) if lb is None else np.asarray(lb, dtype=np.float64)
    x = ((lb + ub) / 2) if x0 is None else np.array(x0, dtype=np.float64)
    res = PGResult()
    res.status, res.iter, res.n_clipped = 'unknown', 0, 0
    f_hist, gap_hist = [], []
    best_lb = -np.inf
    g = Q @ x + q if passes == 1 else None
    while True:
        if passes == 3:
            f_x = 0.5 * x @ Q @ x + q @ x
            g = Q @ x + q
        else:
            f_x = 0.5 * (x @ (g + q))
        y = np.where(g < 0, ub, lb)
        lbv = f_x + g.dot(y - x)
        if lbv > best_lb:
            best_lb = lbv
        gap = (f_x - best_lb) / max(abs(f_x), 1)
        f_hist.append(f_x)
        gap_hist.append(gap)
        if callback is not None:
            callback(res.iter, x, f_x, gap)
        if gap <= eps:
            res.status = 'optimal'
            break
        if res.iter >= max_iter:
            res.status = 'stopped'
            break
        if t > 0:
            radius = t * (ub - lb)
            y = np.clip(y, x - radius, x + radius)
        d = y - x
        w = Q @ d if passes == 1 else d.dot(Q)
        den = w.dot(d)
        if den <= 1e-16:
            a = 1
        else:
            a = min(-g.dot(d) / den, 1)
            res.n_clipped += int(a == 1)
        x += a * d
        if passes == 1:
            g = g + a * w
        res.iter += 1
    res.x, res.f_x, res.g_x = x, f_x, g
    res.f_hist, res.ng_hist = np.array(f_hist), np.array(gap_hist)
    return res


# ----------------------------------------------------------------------------- estimators


def binarize_labels(y):
    """sklearn LabelBinarizer(neg_label=-1) as used at optiml/ml/svm/_base.py:419, 436-440."""
    classes = np.unique(y)
    if len(classes) > 2:
        raise ValueError('more than two labels')
    if len(classes) == 1:
        # LabelBinarizer maps a single class to neg_label
        return classes, -np.ones(len(y), dtype=np.int64)
    return classes, np.where(np.asarray(y) == classes[1], 1, -1).astype(np.int64)


class FitResult:
    pass


def svc_dual_fit(X, y, kind='gaussian', C=1., degree=3, gamma='scale', coef0=0., max_iter=1000,
                 eps=1e-6, passes=3):
    """optiml/ml/svm/_base.py:435-440, 547-559, 619-636, 725, 867-880
    (loss=hinge, dual=True, reg_intercept=True, optimizer=ProjectedGradient)."""
    X = np.asarray(X, dtype=np.float64)
    classes, ys = binarize_labels(y)
    n = len(ys)
    K = kernel_matrix(kind, X, None, degree=degree, gamma=gamma, coef0=coef0)
    yy = np.outer(ys, ys)
    Q = K * yy
    Q += yy
    q = -np.ones(n)
    ub = np.ones(n) * C
    pg = projected_gradient(Q, q, ub, eps=eps, max_iter=max_iter, passes=passes)
    out = FitResult()
    out.classes_, out.pg, out.alphas_ = classes, pg, pg.x
    sv = pg.x > 1e-6
    out.support_ = np.arange(n)[sv]
    out.support_vectors_ = X[sv]
    sv_y, a = ys[sv], pg.x[sv]
    out.dual_coef_ = a * sv_y
    out.coef_ = out.dual_coef_ @ out.support_vectors_ if kind == 'linear' else None
    b = 0.
    for i in range(len(a)):
        b += sv_y[i]
        b -= np.sum(out.dual_coef_ * K[out.support_[i], sv])
    out.intercept_ = b / len(a)
    out.kernel = dict(kind=kind, degree=degree, gamma=gamma, coef0=coef0)
    out.K, out.Q = K, Q
    return out


def svr_dual_fit(X, y, kind='poly', C=1., epsilon=0.1, degree=3, gamma='scale', coef0=0., max_iter=1000,
                 eps=1e-6, passes=3, materialize=True):
    """optiml/ml/svm/_base.py:979-983, 1091-1104, 1126, 1169-1186, 1275-1277, 1423-1437
    (loss=epsilon_insensitive, dual=True, reg_intercept=True, optimizer=ProjectedGradient)."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n = len(y)
    K = kernel_matrix(kind, X, None, degree=degree, gamma=gamma, coef0=coef0)
    Q = np.vstack((np.hstack((K, -K)), np.hstack((-K, K))))
    q = np.hstack((-y, y)) + epsilon
    ub = np.ones(2 * n) * C
    e = np.hstack((np.ones(n), -np.ones(n)))
    Q += np.outer(e, e)
    pg = projected_gradient(Q, q, ub, eps=eps, max_iter=max_iter, passes=passes)
    out = FitResult()
    out.pg, out.alphas_ = pg, pg.x
    ap, an = np.split(pg.x, 2)
    sv = np.logical_or(ap > 1e-6, an > 1e-6)
    out.support_ = np.arange(n)[sv]
    out.support_vectors_ = X[sv]
    sv_y, ap, an = y[sv], ap[sv], an[sv]
    out.dual_coef_ = ap - an
    out.coef_ = out.dual_coef_ @ out.support_vectors_ if kind == 'linear' else None
    b = 0.
    for i in range(len(ap)):
        b += sv_y[i]
        b -= np.sum(out.dual_coef_ * K[out.support_[i], sv])
    b -= epsilon
    out.intercept_ = b / len(ap)
    out.kernel = dict(kind=kind, degree=degree, gamma=gamma, coef0=coef0)
    out.K, out.Q = K, Q
    return out


def decision_function(fit, X):
    """optiml/ml/svm/_base.py:284-287.  gamma='scale' is resolved from support_vectors_ (first
    argument of the kernel call), not from the training matrix."""
    X = np.asarray(X, dtype=np.float64)
    if fit.kernel['kind'] == 'linear':
        return X @ fit.coef_ + fit.intercept_
    k = fit.kernel
    Ksx = kernel_matrix(k['kind'], fit.support_vectors_, X, degree=k['degree'], gamma=k['gamma'], coef0=k['coef0'])
    return fit.dual_coef_ @ Ksx + fit.intercept_


def svc_predict(fit, X):
    """optiml/ml/svm/_base.py:884-885: LabelBinarizer.inverse_transform thresholds at > 0."""
    dec = decision_function(fit, X)
    return fit.classes_[(dec > 0).astype(int)] if len(fit.classes_) == 2 else np.repeat(fit.classes_[0], len(dec))
