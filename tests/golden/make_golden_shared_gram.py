"""Golden vectors for the shared-Gram widening, produced by the REAL reference wrapped in sklearn's
own meta-estimators, exactly as the reference's tests do (ml/tests/test_svc.py:5, 101-147):
    OneVsRestClassifier(optiml SVC(dual=True, reg_intercept=True, optimizer=FrankWolfe))   4 classes, n = 240
    MultiOutputRegressor(optiml SVR(dual=True, reg_intercept=True, optimizer=FrankWolfe))  3 targets, n = 200
Frank-Wolfe trajectories are stable, so alpha can be compared to 1e-8.
    python tests/golden/make_golden_shared_gram.py"""
import os
import sys

import numpy as np
from sklearn.datasets import make_classification
from sklearn.multiclass import OneVsRestClassifier
from sklearn.multioutput import MultiOutputRegressor

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_shim import load_reference  # noqa: E402

ref = load_reference()
OUT = os.path.dirname(os.path.abspath(__file__))
out = {}

X, y = make_classification(n_samples=300, n_features=8, n_informative=5, n_redundant=1, n_classes=4, n_clusters_per_class=1,
                           class_sep=1.5, random_state=3)
Xtr, ytr, Xte, yte = X[:240], y[:240], X[240:], y[240:]
ovr = OneVsRestClassifier(ref.SVC(loss=ref.hinge, kernel=ref.GaussianKernel(), C=1, reg_intercept=True, dual=True,
                                  optimizer=ref.FrankWolfe, max_iter=400)).fit(Xtr, ytr)
out.update(ovr_X_train=Xtr, ovr_y_train=ytr, ovr_X_test=Xte, ovr_y_test=yte, ovr_predict=ovr.predict(Xte),
           ovr_decision=ovr.decision_function(Xte), ovr_score=ovr.score(Xte, yte))
for c, e in enumerate(ovr.estimators_):
    out.update({f'ovr_c{c}_alphas': e.alphas_, f'ovr_c{c}_support': e.support_, f'ovr_c{c}_intercept': e.intercept_,
                f'ovr_c{c}_iter': e.optimizer.iter, f'ovr_c{c}_status': e.optimizer.status,
                f'ovr_c{c}_f_hist': np.array(e.train_loss_history)})
    print('ovr class', c, e.optimizer.iter, e.optimizer.status, len(e.support_), e.intercept_)
print('ovr score', out['ovr_score'])

rng = np.random.default_rng(12)
Xr = rng.standard_normal((230, 5))
Yr = np.stack((np.sin(Xr[:, 0]) + 0.3 * Xr[:, 1], Xr[:, 2] * Xr[:, 3], np.tanh(Xr[:, 4]) - 0.5 * Xr[:, 0]), axis=1)
Yr = (Yr - Yr.mean(axis=0)) / Yr.std(axis=0)
mor = MultiOutputRegressor(ref.SVR(loss=ref.epsilon_insensitive, epsilon=0.1, kernel=ref.GaussianKernel(), C=1,
                                   reg_intercept=True, dual=True, optimizer=ref.FrankWolfe, max_iter=400)).fit(Xr[:200], Yr[:200])
out.update(mor_X_train=Xr[:200], mor_Y_train=Yr[:200], mor_X_test=Xr[200:], mor_Y_test=Yr[200:],
           mor_predict=mor.predict(Xr[200:]), mor_score=mor.score(Xr[200:], Yr[200:]))
for t, e in enumerate(mor.estimators_):
    out.update({f'mor_t{t}_alphas': e.alphas_, f'mor_t{t}_support': e.support_, f'mor_t{t}_intercept': e.intercept_,
                f'mor_t{t}_iter': e.optimizer.iter, f'mor_t{t}_status': e.optimizer.status,
                f'mor_t{t}_f_hist': np.array(e.train_loss_history)})
    print('mor target', t, e.optimizer.iter, e.optimizer.status, len(e.support_), e.intercept_)
print('mor score', out['mor_score'])
np.savez_compressed(os.path.join(OUT, 'shared_gram.npz'), **out)
