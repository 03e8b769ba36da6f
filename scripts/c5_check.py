"""C5 (DualSVC Gaussian n=120 000 d=64: a 115 GB Hessian that only fits row-sharded over 8 B200s): one fit,
size-independent properties, timing.  torchrun --nproc-per-node 8 scripts/c5_check.py"""
import hashlib, json, os, sys, time
sys.path.insert(0, '.')
import numpy as np, torch, torch.distributed as dist
lr = int(os.environ.get('LOCAL_RANK', '0')); torch.cuda.set_device(lr)
dist.init_process_group(backend='nccl', device_id=torch.device('cuda', lr))
from optiml_b200 import runtime
from optiml_b200.configs import make_config
from optiml_b200.ml.svm import DualSVC
from optiml_b200.ml.svm.kernels import GaussianKernel
ctx = runtime.default_context()
spec, X, y = make_config('C5')
n = len(y)
res = {}
for rep in range(2):
    dist.barrier()
    t = time.perf_counter()
    m = DualSVC(kernel=GaussianKernel(), C=1)
    m.profile_matvec = True
    m.fit(X, y)
    dt = time.perf_counter() - t
    res = dict(n=n, d=X.shape[1], n_gpus=ctx.nranks, exchange=ctx.exchange, fit_s=dt, pg_ms=m.optimizer.device_ms,
               iters=m.optimizer.iter, status=m.optimizer.status, f_x=m.optimizer.f_x, n_sv=len(m.support_),
               intercept=m.intercept_, its_per_s=m.optimizer.iter / (m.optimizer.device_ms / 1e3),
               matvec_us=1e3 * m.optimizer.matvec_ms / max(m.optimizer.profile_samples, 1),
               hbm_gbps_per_gpu=8.0 * n * n / ctx.nranks / (m.optimizer.matvec_ms / max(m.optimizer.profile_samples, 1) / 1e3) / 1e9,
               hessian_gb_per_gpu=8.0 * n * n / ctx.nranks / 1e9,
               symmetric_pass=bool(m.optimizer.symmetric_pass))   # SVMB200_SYMMETRIC=1: hbm_gbps_per_gpu is then the FULL-matrix equivalent
    if rep == 0:
        m.obj.release()
fh = np.array(m.train_loss_history)
g_fresh = m.obj.jacobian(m.alphas_)
# against the CPU oracle (the reference cannot run this size: 346 GB of host matrices): 128 rows spread over all shards of
# the Hessian, rebuilt with the oracle's kernel -- checks the Gram shards, the streaming product and the gradient the
# solver carried through 1000 iterations at full C5 size
from oracle import svm_oracle as O
rows = np.unique(np.linspace(0, n - 1, 128).astype(np.int64))
ys = np.where(y == np.unique(y)[1], 1.0, -1.0)
Krows = O.gaussian_kernel(X[rows], X, gamma=1.0 / (X.shape[1] * X.var()))
Krows[np.arange(len(rows)), rows] = 1.0                       # the self-Gram's exact diagonal (sklearn semantics)
g_oracle = (ys[rows, None] * ys[None, :] * (Krows + 1.0)) @ m.alphas_ - 1.0
sv_mask = m.alphas_ > 1e-6
# decision_function resolves gamma='scale' from the SUPPORT VECTORS (first argument of the kernel call, ml/svm/_base.py:286)
Ksv = O.gaussian_kernel(X[sv_mask], X[rows], gamma=1.0 / (X.shape[1] * X[sv_mask].var()))
dec_oracle = (m.alphas_[sv_mask] * ys[sv_mask]) @ Ksv + m.intercept_
dec_dev = m.decision_function(X[rows])
digest = hashlib.sha256(m.alphas_.tobytes()).hexdigest()
digs = [None] * ctx.nranks
dist.all_gather_object(digs, digest)
res.update(monotone_descent=bool(np.all(np.diff(fh) <= 1e-9 * np.abs(fh[:-1]))),
           feasible=bool(m.alphas_.min() >= -1e-12 and m.alphas_.max() <= 1 + 1e-12),
           grad_drift=float(np.abs(g_fresh - m.optimizer.g_x).max() / np.abs(g_fresh).max()),
           f_consistency=float(abs(0.5 * m.alphas_ @ (g_fresh - 1.) - m.optimizer.f_x) / abs(m.optimizer.f_x)),
           oracle_rows=int(len(rows)),
           grad_vs_oracle_rows=float(np.abs(m.optimizer.g_x[rows] - g_oracle).max() / np.abs(g_oracle).max()),
           fresh_grad_vs_oracle_rows=float(np.abs(g_fresh[rows] - g_oracle).max() / np.abs(g_oracle).max()),
           decision_vs_oracle_rows=float(np.abs(dec_dev - dec_oracle).max() / max(1.0, np.abs(dec_oracle).max())),
           ranks_identical=len(set(digs)) == 1, train_acc_first_4096=float(m.score(X[:4096], y[:4096])))
if dist.get_rank() == 0:
    print('C5_RESULT ' + json.dumps(res), flush=True)
dist.destroy_process_group()
