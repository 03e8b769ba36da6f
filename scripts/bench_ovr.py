"""Device timing of the shared-Gram one-vs-rest fit (widening 8f-4) at the headline size: C4's inputs
(n = 50 000, d = 128) relabelled into `--classes` classes, OneVsRestClassifier(SVC(hinge, Gaussian, dual=True,
reg_intercept=True, optimizer=ProjectedGradient)).

Arms:  `shared`  optiml_b200.ml.multiclass.OneVsRestClassifier -- one M = K + 1 in HBM, lockstep solvers, one pass
                 over M per iteration for up to four classes;
       `cloned`  sklearn.multiclass.OneVsRestClassifier over the same estimator -- one Gram build and one private Q per
                 class, one pass per class and iteration (what the reference's recipe does, on the GPU).
Prints one JSON line per arm (fit seconds, problem-iterations/s, passes over the matrix, bytes streamed per
problem-iteration) and whether the two arms agree bitwise.

    python scripts/bench_ovr.py [--n N] [--classes K] [--iters I] [--solver pg|fw]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/bench_ovr.py   (row-sharded M)
"""
import argparse
import json
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=None)
    ap.add_argument('--classes', type=int, default=4)
    ap.add_argument('--iters', type=int, default=300)
    ap.add_argument('--solver', default='pg', choices=['pg', 'fw'])
    ap.add_argument('--skip-cloned', action='store_true')
    args = ap.parse_args()
    from sklearn.multiclass import OneVsRestClassifier as SklearnOVR
    from optiml_b200.configs import make_config
    from optiml_b200.ml.multiclass import OneVsRestClassifier
    from optiml_b200.ml.svm import SVC
    from optiml_b200.ml.svm.kernels import GaussianKernel
    from optiml_b200.ml.svm.losses import hinge
    from optiml_b200.opti.constrained import FrankWolfe, ProjectedGradient
    from optiml_b200.runtime import default_context
    warnings.simplefilter('ignore')
    rank, world = 0, int(os.environ.get('WORLD_SIZE', '1'))
    if world > 1:   # under torchrun: one rank per GPU, M row-sharded, fused peer exchange for the whole batch
        import torch
        import torch.distributed as dist
        local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend='nccl', device_id=torch.device('cuda', local_rank))
        rank = dist.get_rank()
    spec, X, y2 = make_config('C4', n=args.n)
    n = len(y2)
    # synthetic multi-class labels on C4's inputs: quantiles of a fixed random projection
    rng = np.random.default_rng(0)
    score = X @ rng.standard_normal(X.shape[1])
    y = np.searchsorted(np.quantile(score, np.linspace(0, 1, args.classes + 1)[1:-1]), score)
    est = SVC(loss=hinge, kernel=GaussianKernel(), C=1, reg_intercept=True, dual=True,
              optimizer=ProjectedGradient if args.solver == 'pg' else FrankWolfe, max_iter=args.iters)
    ctx = default_context()
    fitted = {}
    arms = [('shared', OneVsRestClassifier)] + ([] if args.skip_cloned else [('cloned', SklearnOVR)])
    for arm, cls in arms:
        for rep in range(2):  # first fit warms the buffer pools
            ctx.trim()
            launches = ctx.launch_count()
            t0 = time.perf_counter()
            model = cls(est).fit(X, y)
            fit_s = time.perf_counter() - t0
            launches = ctx.launch_count() - launches
            if rep == 0:
                for e in model.estimators_:
                    e.obj.release()
                del model
        ests = model.estimators_
        iters = sum(e.optimizer.iter for e in ests)
        if arm == 'shared':
            loop_ms, passes = ests[0].optimizer.device_ms, ests[0].optimizer.q_passes   # whole batch
        else:
            loop_ms, passes = sum(e.optimizer.device_ms for e in ests), sum(e.optimizer.q_passes for e in ests)
        fitted[arm] = [e.alphas_.copy() for e in ests]
        if rank == 0:
            print(json.dumps(dict(arm=arm, n_gpus=world, exchange=ctx.exchange,
                              case=f'C4 inputs n={n}, {args.classes} classes, {args.solver}, max_iter={args.iters}',
                              fit_s=round(fit_s, 4), loop_ms=round(loop_ms, 2), problem_iterations=int(iters),
                              problem_its_per_s=round(iters / (loop_ms / 1e3), 1), passes_over_matrix=int(passes),
                              us_per_pass=round(1e3 * loop_ms / max(passes, 1), 1),
                              hbm_gbps_all_gpus=round(8.0 * n * n * passes / (loop_ms / 1e3) / 1e9, 1),
                              gb_per_problem_iteration=round(8.0 * n * n * passes / max(iters, 1) / 1e9, 2),
                              kernel_launches=int(launches), n_sv=[int(len(e.support_)) for e in ests])), flush=True)
        for e in ests:
            e.obj.release()
        del model, ests
    if len(fitted) == 2 and rank == 0:
        same = all(np.array_equal(a, b) for a, b in zip(fitted['shared'], fitted['cloned']))
        print(json.dumps(dict(bitwise_equal_shared_vs_cloned=bool(same))), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
