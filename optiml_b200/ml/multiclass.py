"""One-vs-rest / multi-target meta-estimators whose binary problems share one Gram matrix in HBM.

The reference's multi-class recipe is ``sklearn.multiclass.OneVsRestClassifier(SVC(...))`` (ml/tests/test_svc.py:5,
101-147; ml/svm/_base.py:437-439 tells the user to do so): sklearn clones the estimator once per class and fits the
clones one after the other, so the same n x n Gram matrix is built ``n_classes`` times and every solver streams a
private copy ``Q_c = (y_c y_c') o (K + 1)``.  The classes differ in the label signs only.  The drop-ins below keep
sklearn's contract -- ``estimators_`` holds fitted clones, ``predict`` / ``decision_function`` / ``score`` are
sklearn's own -- but build ``M = K + bias`` ONCE, leave it unsigned, and advance all binary solvers in lockstep with
one streaming pass over M per iteration for up to four problems (``svmb200_pg_run_batch``).  Each clone ends
bit-identical to what ``clone(estimator).fit(X, y_c)`` produces on the same GPU.

``MultiOutputRegressor`` is the same idea for ``SVR``: the targets only change the linear term q.

Anything the lockstep driver cannot take (a generic callback, ``verbose``, an estimator of another type, fit
parameters, a degenerate class column) goes through sklearn's own ``fit`` -- still on the GPU, one clone after
the other.
"""
import time

import numpy as np
from sklearn import multiclass as _skmc
from sklearn import multioutput as _skmo
from sklearn.base import clone
from sklearn.preprocessing import LabelBinarizer

from .svm._base import SVC, SVR
from .svm.kernels import _dense_f64
from ..opti.batch import minimize_batch

__all__ = ['OneVsRestClassifier', 'MultiOutputRegressor', 'fit_shared_gram']


def fit_shared_gram(estimator, X, targets):
    """Fit one clone of ``estimator`` (an ``SVC`` or ``SVR`` of this package on the dual path) per entry of ``targets``
    (label / target vectors over the same rows of X) on ONE resident Gram matrix.  Returns the fitted clones."""
    if not isinstance(estimator, (SVC, SVR)):
        raise TypeError(f'{estimator} is not an SVC / SVR of optiml_b200')
    targets = [np.asarray(t) for t in targets]
    if not targets:
        return []
    clones = [clone(estimator) for _ in targets]
    X, _ = _dense_f64(X, None, finite=False)  # the all-finite test runs on the device copy (_build_hessian)
    head = clones[0]
    head._bcqp_solver_class()  # raises for the branches outside the dual path before any device work
    # K1 once: M = K + bias without label signs (SVC: the solvers apply them; SVR: the block signs are implicit)
    shared = head._build_hessian(X, None, 'svr' if isinstance(estimator, SVR) else 'plain', None, head._bias())
    gram_s = head.fit_times_['gram_s']
    dX, owned = head._train_X_device   # every clone gathers its support vectors from this one copy of X in HBM
    try:
        plans = [est._plan_fit(X, t, shared=shared) for est, t in zip(clones, targets)]
        t0 = time.perf_counter()
        minimize_batch([p['solver'] for p in plans])
        solve_s = time.perf_counter() - t0
        for est, plan in zip(clones, plans):
            # wall-clock of the shared steps: every clone reports the whole batch
            est.fit_times_ = {'gram_s': gram_s, 'solve_s': solve_s, 'batch': len(clones)}
            est._train_X_device = (dX, False)
            est._finish_fit(plan)
    finally:
        for est in clones:
            est._train_X_device = None
        if owned:
            dX.release()
    return clones


def _lockstep_capable(estimator, n_samples=None):
    """The estimator runs a device-resident solver that nothing on the host has to watch (and the problem is not one a
    single-process device group shards over its GPUs: there the clones are fitted one after the other, each on all GPUs)."""
    if not isinstance(estimator, (SVC, SVR)) or estimator.verbose:
        return False
    if n_samples is not None:
        from ..runtime import default_context
        group = default_context().group
        if group is not None and group.wants(int(n_samples)):
            return False
    try:
        estimator._bcqp_solver_class()
    except (NotImplementedError, TypeError):
        return False
    return True


class OneVsRestClassifier(_skmc.OneVsRestClassifier):
    """``sklearn.multiclass.OneVsRestClassifier`` whose binary ``SVC`` problems share one Gram matrix."""

    def fit(self, X, y, **fit_params):
        if fit_params or not isinstance(self.estimator, SVC) or not _lockstep_capable(self.estimator, np.shape(X)[0]):
            return super().fit(X, y, **fit_params)
        self._validate_params()
        # label handling of sklearn/multiclass.py OneVsRestClassifier.fit
        label_binarizer = LabelBinarizer(sparse_output=True)
        Y = label_binarizer.fit_transform(y).tocsc()
        columns = [col.toarray().ravel() for col in Y.T]
        if len(columns) < 2 or any(len(np.unique(c)) == 1 for c in columns):
            return super().fit(X, y)  # a single problem, or a constant column (sklearn's _ConstantPredictor)
        self.label_binarizer_ = label_binarizer
        self.classes_ = label_binarizer.classes_
        self.estimators_ = fit_shared_gram(self.estimator, X, columns)
        if hasattr(self.estimators_[0], 'n_features_in_'):
            self.n_features_in_ = self.estimators_[0].n_features_in_
        if hasattr(self.estimators_[0], 'feature_names_in_'):
            self.feature_names_in_ = self.estimators_[0].feature_names_in_
        return self


class MultiOutputRegressor(_skmo.MultiOutputRegressor):
    """``sklearn.multioutput.MultiOutputRegressor`` whose per-target ``SVR`` problems share one Gram matrix."""

    def fit(self, X, y, sample_weight=None, **fit_params):
        y_arr = np.asarray(y)
        if sample_weight is not None or fit_params or not isinstance(self.estimator, SVR) or \
                not _lockstep_capable(self.estimator, np.shape(X)[0]) or y_arr.ndim != 2 or y_arr.shape[1] < 2:
            return super().fit(X, y, sample_weight=sample_weight, **fit_params)
        self._validate_params()
        self.estimators_ = fit_shared_gram(self.estimator, X, [y_arr[:, i] for i in range(y_arr.shape[1])])
        if hasattr(self.estimators_[0], 'n_features_in_'):
            self.n_features_in_ = self.estimators_[0].n_features_in_
        if hasattr(self.estimators_[0], 'feature_names_in_'):
            self.feature_names_in_ = self.estimators_[0].feature_names_in_
        return self
