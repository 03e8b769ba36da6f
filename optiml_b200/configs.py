"""Deterministic synthetic inputs for the five BASELINE.json configurations.

All arrays are float64, C-contiguous.  ``scale`` shrinks n (never d) so that the same recipe can
be used for parity cases that the CPU oracle finishes in seconds.
"""
import numpy as np

CONFIGS = {
    'C1': dict(task='svc', kernel='gaussian', n=2000, d=20),
    'C2': dict(task='svr', kernel='poly', n=10000, d=32, degree=3, epsilon=0.1),
    'C3': dict(task='svc', kernel='linear', n=20000, d=784),
    'C4': dict(task='svc', kernel='gaussian', n=50000, d=128),
    'C5': dict(task='svc', kernel='gaussian', n=120000, d=64),
}


def make_config(name, n=None):
    """Return ``(spec, X, y)`` for config ``name``; ``n`` overrides the sample count."""
    from sklearn.datasets import make_classification, make_regression
    spec = dict(CONFIGS[name])
    if n is not None:
        spec['n'] = int(n)
    n, d = spec['n'], spec['d']
    if name in ('C1', 'C4', 'C5'):
        X, y = make_classification(n_samples=n, n_features=d, random_state=0)
    elif name == 'C2':
        X, y = make_regression(n_samples=n, n_features=d, noise=0.1, random_state=0)
        y = (y - y.mean()) / y.std()
    elif name == 'C3':
        rng = np.random.default_rng(0)
        X = rng.random((n, d)) * (rng.random((n, d)) < 0.19)
        w = rng.standard_normal(d)
        s = X @ w
        y = (s > np.median(s)).astype(int)
    else:
        raise KeyError(name)
    return spec, np.ascontiguousarray(X, dtype=np.float64), np.ascontiguousarray(y)
