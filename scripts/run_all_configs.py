"""Fit every BASELINE config that fits one GPU (C1-C4), compare with the reference's full-size golden vectors,
print one JSON line per config (timings on this GPU next to the reference's timings from the golden files)."""
import json, os, sys, time
sys.path.insert(0, '.')
import numpy as np
from optiml_b200.configs import make_config
from optiml_b200.ml.svm import DualSVC, DualSVR
from optiml_b200.ml.svm.kernels import GaussianKernel, PolyKernel, LinearKernel

def G(name):
    z = np.load(os.path.join('tests', 'golden', name + '.npz')); return {k: z[k] for k in z.files}

CASES = [('C1', 'c1_svc_gaussian', lambda: DualSVC(kernel=GaussianKernel(), C=1)),
         ('C2', 'c2_full_svr_poly', lambda: DualSVR(kernel=PolyKernel(degree=3), epsilon=0.1, C=1)),
         ('C3', 'c3_full_svc_linear', lambda: DualSVC(kernel=LinearKernel(), C=1)),
         ('C4', 'c4_full_svc_gaussian', lambda: DualSVC(kernel=GaussianKernel(), C=1))]
for cfg, gold, mk in CASES:
    spec, X, y = make_config(cfg)
    g = G(gold)
    mk().fit(X, y).obj.release()  # warm-up
    t = time.perf_counter(); m = mk().fit(X, y); fit_s = time.perf_counter() - t
    fh = np.array(m.train_loss_history); gh = g['f_hist']
    dev = np.abs(fh - gh) / np.maximum(1., np.abs(gh))
    first_dev = int(np.argmax(dev > 1e-9)) if (dev > 1e-9).any() else -1
    n = len(y)
    ref_s = float(g['fit_seconds']) if 'fit_seconds' in g else (float(g['pg_seconds']) if 'pg_seconds' in g else None)
    out = dict(config=cfg, n=n, d=X.shape[1], task=spec['task'], kernel=spec['kernel'], fit_s=round(fit_s, 4),
               gram_s=round(m.fit_times_['gram_s'], 4), pg_ms=round(m.optimizer.device_ms, 2), iters=m.optimizer.iter,
               status=m.optimizer.status, pg_its_per_s=round(m.optimizer.iter / (m.optimizer.device_ms / 1e3), 1),
               q_gbps=round(8.0 * n * n * m.optimizer.q_passes / (m.optimizer.device_ms / 1e3) / 1e9, 1),
               max_abs_dalpha=float(np.abs(m.alphas_ - g['alphas']).max()),
               same_support=bool(np.array_equal(m.support_, g['support'])), n_sv=len(m.support_),
               intercept=m.intercept_, ref_intercept=float(g['intercept']),
               f_hist_first_dev_gt_1e9=first_dev, final_f=float(fh[-1]), ref_final_f=float(gh[-1]),
               reference_seconds_8_host_cores=ref_s,
               reference_seconds_kind=('whole fit' if 'fit_seconds' in g else 'PG loop only' if ref_s else None))
    print('CONFIG_RESULT ' + json.dumps(out), flush=True)
    m.obj.release()
