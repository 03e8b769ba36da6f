#!/usr/bin/env python
"""Benchmark of the kernel-SVM dual training path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config C4] [--n N_SAMPLES]

A *step* is one complete ``DualSVC(GaussianKernel(), C=1).fit`` of config C4 (n = 50 000, d = 128,
1000 projected-gradient iterations): Gram/Hessian build into HBM + the PG loop + support set and
intercept.  With N > 1 (launched by torchrun, one rank per GPU) Q is row-block sharded and every
iteration ends with one NCCL all-gather, total work fixed (strong scaling).

``value``   PG iterations per second over the whole step with X already resident in HBM (device
            timed, CUDA events on the library's stream, max over ranks).
``e2e``     the same through the public API ``DualSVC.fit(X, y)`` with pinned HOST arrays: the
            host->device copies of X/y/q/bounds and the device->host reads of alpha, gradient, loss
            history and the intercept product are inside the timed region.
``roofline`` the streaming matvec (one launch per iteration): 8 n^2 / N bytes per launch / mean
            launch duration from CUDA events around every launch, vs the measured HBM copy peak.
``parity``  the last timed fit against the golden vectors of the REAL reference's run of the same config
            (tests/golden/c4_full_svc_gaussian.npz): max |delta alpha|, support set, objective, intercept.
``cpu_baseline`` / ``--impl reference``: the UNMODIFIED reference (pip-installed into the git-ignored
            ``baseline/_ref``) on the host cores, on the FULL problem (same n, d, kernel, C) for a
            bounded number of iterations: one ``SVC.fit(max_iter=m)`` with a time-stamping callback gives the
            set-up time (Gram, Q, first evaluation), the time per PG iteration (iteration-independent: three
            passes over Q) and the post-processing time; the reported it/s is max_iter / (set-up +
            max_iter * per-iteration + post), the whole-fit figure the GPU arm reports.  Falls back to the
            NumPy oracle port on a row sample when the reference copy or the host memory (~90 GB) is missing.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'DualSVC Gaussian n=50k d=128 C=1: projected-gradient iterations/s over the whole fit'
UNIT = 'PG it/s'


def metric_name(config, n, d):
    if config == 'C4' and n == 50000:
        return METRIC
    return f'DualSVC Gaussian n={n} d={d} C=1 ({config}): projected-gradient iterations/s over the whole fit'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='C4')
    ap.add_argument('--n', type=int, default=None, help='override the sample count (debug)')
    ap.add_argument('--max-iter', type=int, default=1000)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--symmetric', action='store_true',
                    help='run the whole bench with the opt-in symmetric pass (K2s: products from the upper triangle of Q '
                         'alone); without it the default full pass is benched and, on one GPU, a short symmetric-pass leg '
                         'is added to the line as "symmetric_pass"')
    ap.add_argument('--no-symmetric-leg', action='store_true')
    ap.add_argument('--devices', default=None,
                    help="ONE process driving several GPUs (comma-separated indices or 'all'): the single-process device "
                         "group instead of torchrun ranks; not a driver mode, used to compare the two launch models")
    return ap.parse_args()


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs, copy)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.device), f'--query-gpu={self.FIELDS}',
                                          '--format=csv,noheader,nounits', '-lms', '200'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'power_w_max': max(power) if power else None, 'samples': len(sm), 'reasons': sorted(reasons)}


# ----------------------------------------------------------------------------- CPU arms
def workload_name(config, n, d, max_iter):
    return f'{config} DualSVC GaussianKernel C=1 n={n} d={d} max_iter={max_iter} (make_classification random_state=0)'


def bench_config(config, n, d, max_iter, world):
    """The `config` object of the JSON line: identical for the GPU arm and the reference arm."""
    return {'workload': workload_name(config, n, d, max_iter), 'parallelism': f'row-block x{world}',
            'l2': f'inputs larger than L2: Q shard = {8.0 * n * n / world / 1e9:.2f} GB per GPU, streamed once per iteration'}


def oracle_sample(config, n_full, n_sample, iters, max_iter):
    """Fallback: time the oracle (reference algorithm, NumPy FP64, 3 products with Q per iteration) on the first
    n_sample rows of the workload for `iters` PG iterations; scale it/s to n_full by (n_sample/n_full)^2
    (every iteration is 3 passes over the n x n matrix: cost is proportional to n^2)."""
    from oracle import svm_oracle as O
    from optiml_b200.configs import make_config
    spec, X, y = make_config(config, n=n_full)
    X, y = X[:n_sample], y[:n_sample]
    t0 = time.perf_counter()
    classes, ys = O.binarize_labels(y)
    K = O.gaussian_kernel(X)
    yy = np.outer(ys, ys)
    Q = K * yy
    Q += yy
    gram_s = time.perf_counter() - t0
    t1 = time.perf_counter()
    res = O.projected_gradient(Q, -np.ones(n_sample), np.ones(n_sample), max_iter=iters, passes=3)
    pg_s = time.perf_counter() - t1
    its_sample = res.iter / pg_s
    scale = (n_sample / n_full) ** 2
    # whole-fit equivalent at n_full with max_iter iterations: Gram scales with n^2 as well
    fit_full_s = gram_s / scale + max_iter / (its_sample * scale)
    return dict(its_sample=its_sample, gram_s=gram_s, pg_s=pg_s, iters=res.iter, scale=scale,
                its_full_pg_only=its_sample * scale, its_full_whole_fit=max_iter / fit_full_s, fit_full_s=fit_full_s)


def host_threads():
    try:
        from threadpoolctl import threadpool_info
        nt = [p.get('num_threads', 1) for p in threadpool_info() if p.get('user_api') == 'blas']
        return max(nt) if nt else (os.cpu_count() or 1)
    except Exception:
        return os.cpu_count() or 1


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arms are meant to use every host core the BLAS can."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        pass


def available_host_bytes():
    try:
        import psutil
        return int(psutil.virtual_memory().available)
    except Exception:
        return 0


class ReferenceRun:
    """The unmodified reference (baseline/_ref) on the full problem.  `fit_sample(m)` runs the reference's own
    ``SVC(loss=hinge, kernel=GaussianKernel(), C=1, dual=True, reg_intercept=True, optimizer=ProjectedGradient,
    max_iter=m).fit(X, y)`` with a time-stamping wrapper around its per-iteration callback; `loop_sample(m)` re-runs the
    reference's ``ProjectedGradient(quad, ub, max_iter=m).minimize()`` on the Hessian that fit left in ``model.obj``."""

    # K, the int64 outer(y, y), Q = K * yy and the copy Quadratic takes are alive together (ml/svm/_base.py:552-554, 628;
    # opti/_base.py:243): ~4 n^2 doubles at the peak, plus NumPy temporaries
    @staticmethod
    def bytes_needed(n):
        return int(4.6 * 8 * n * n)

    def __init__(self, config, n):
        from oracle.ref_shim import load_reference, REFERENCE_ROOT
        from optiml_b200.configs import make_config
        self.ref = load_reference()
        self.root = REFERENCE_ROOT
        self.spec, self.X, self.y = make_config(config, n=n)
        self.n = len(self.y)
        self.model = None

    def fit_sample(self, m):
        ref = self.ref
        model = ref.SVC(loss=ref.hinge, kernel=ref.GaussianKernel(), C=1, reg_intercept=True, dual=True,
                        optimizer=ref.ProjectedGradient, max_iter=m)
        stamps = []
        inner = model._store_train_info  # the callback SVC.fit hands to the solver (ml/svm/_base.py:631-636)

        def stamped(opt):
            inner(opt)
            stamps.append(time.perf_counter())

        model._store_train_info = stamped
        t0 = time.perf_counter()
        model.fit(self.X, self.y)
        t1 = time.perf_counter()
        self.model = model
        iters = len(stamps) - 1
        return dict(setup_s=stamps[0] - t0, iter_s=(stamps[-1] - stamps[0]) / max(iters, 1), post_s=t1 - stamps[-1],
                    iters=iters, fit_s=t1 - t0, n_sv=len(model.support_), f_x=float(model.optimizer.f_x))

    def loop_sample(self, m):
        ref, quad = self.ref, self.model.obj
        stamps = []
        opt = ref.ProjectedGradient(quad=quad, ub=np.ones(self.n), max_iter=m, callback=lambda o: stamps.append(time.perf_counter()))
        opt.minimize()
        iters = len(stamps) - 1
        return (stamps[-1] - stamps[0]) / max(iters, 1), iters


def reference_measure(config, n_full, max_iter, steps=0, warmup=0, fit_iters=8):
    """One reference fit on the full problem (+ `warmup + steps` solver-only samples).  Returns the dict both
    `cpu_baseline` and the `--impl reference` line are built from; falls back to the oracle port on a row sample."""
    from optiml_b200.configs import CONFIGS
    use_all_host_threads()
    d = CONFIGS[config]['d']
    why = None
    try:
        from oracle.ref_shim import reference_available
        if not reference_available():
            why = 'no reference copy under baseline/_ref'
    except Exception as e:  # pragma: no cover
        why = f'reference not importable: {e}'
    need, have = ReferenceRun.bytes_needed(n_full), available_host_bytes()
    if why is None and have < need:
        why = f'host memory: {have / 1e9:.0f} GB available, the reference needs ~{need / 1e9:.0f} GB at n={n_full}'
    if why is None:
        try:
            run = ReferenceRun(config, n_full)
            first = run.fit_sample(fit_iters)
            iter_s = [first['iter_s']]
            # per-step samples: the reference's own solver on the Hessian its fit built, sized for ~3 s each
            m = int(min(max(3, round(3.0 / first['iter_s'])), 50))
            extra = []
            for i in range(warmup + steps):
                t, _ = run.loop_sample(m)
                if i >= warmup:
                    extra.append(t)
            if extra:
                iter_s = extra
            it = float(np.mean(iter_s))
            fit_full_s = first['setup_s'] + max_iter * it + first['post_s']
            sample = (f'UNMODIFIED reference (optiml 1.8 from {os.path.relpath(run.root, ROOT)}) on the FULL workload n={n_full} d={d}: '
                      f'SVC(hinge, GaussianKernel, C=1, dual, reg_intercept, ProjectedGradient).fit with max_iter={first["iters"]}: '
                      f'set-up (Gram, Q, first evaluation) {first["setup_s"]:.1f} s, {first["iter_s"] * 1e3:.0f} ms per PG iteration '
                      f'(3 passes over Q), support set + intercept {first["post_s"]:.1f} s'
                      + (f'; then {len(extra)} timed samples of ProjectedGradient.minimize(max_iter={m}) on the same Hessian: '
                         f'{it * 1e3:.0f} ms per iteration' if extra else '')
                      + f'; value = {max_iter} / (set-up + {max_iter} x per-iteration + post) = whole-fit it/s')
            return dict(kind='reference', value=max_iter / fit_full_s, pg_only_its=1.0 / it, sample=sample, fit_full_s=fit_full_s,
                        step_ms=(m * it * 1e3) if extra else first['fit_s'] * 1e3, parts=first, iter_ms=[round(t * 1e3, 2) for t in iter_s])
        except MemoryError:
            why = 'MemoryError inside the reference fit'
    n_sample = min(n_full, 16000)
    vals = [oracle_sample(config, n_full, n_sample, 20, max_iter) for _ in range(max(1, min(steps, 3)))]
    sample = (f'FALLBACK ({why}): first {n_sample} rows of {config} (n={n_full}); NumPy oracle = reference algorithm '
              f'(Gram + 20 PG iterations, 3 passes over Q each); it/s scaled by ({n_sample}/{n_full})^2 '
              f'to the full problem, whole-fit equivalent incl. Gram')
    return dict(kind='port', value=float(np.mean([v['its_full_whole_fit'] for v in vals])),
                pg_only_its=float(np.mean([v['its_full_pg_only'] for v in vals])), sample=sample,
                fit_full_s=float(np.mean([v['fit_full_s'] for v in vals])),
                step_ms=float(np.mean([v['gram_s'] + v['pg_s'] for v in vals]) * 1e3), parts=None, iter_ms=None)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from optiml_b200.configs import CONFIGS
    n_full = args.n or CONFIGS[args.config]['n']
    d = CONFIGS[args.config]['d']
    world = int(os.environ.get('WORLD_SIZE', '1'))
    r = reference_measure(args.config, n_full, args.max_iter, steps=args.steps, warmup=args.warmup)
    line = {'impl': 'reference', 'metric': metric_name(args.config, n_full, d), 'value': r['value'], 'unit': UNIT,
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': r['step_ms'],
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': bench_config(args.config, n_full, d, args.max_iter, max(world, args.gpus)),
            'cpu_baseline': {'value': r['value'], 'unit': UNIT, 'cores': host_threads(), 'kind': r['kind'], 'sample': r['sample'],
                             'pg_only_its': r['pg_only_its'], 'whole_fit_s': r['fit_full_s'], 'parts': r['parts'],
                             'iter_ms_samples': r['iter_ms']},
            'e2e': {'value': r['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def pinned_array(shape, dtype=np.float64):
    import ctypes as C
    from optiml_b200 import _native as N
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    N.call('svmb200_host_alloc_pinned', nbytes, C.byref(p))
    buf = (C.c_char * nbytes).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def parity_block(args, m, n):
    """The last fit against the REAL reference's own run of this config (golden vectors made by
    tests/golden/make_golden_c4_full.py from the imported reference): north_star's bar is max |delta alpha| <= 1e-8 and an
    identical support set."""
    gpath = os.path.join(ROOT, 'tests', 'golden', 'c4_full_svc_gaussian.npz')
    if not (args.config == 'C4' and n == 50000 and os.path.exists(gpath)):
        return None
    g = np.load(gpath)
    k = min(len(m.train_loss_history), len(g['f_hist'])) - 1
    if k < 1:
        return None
    out = {'against': 'tests/golden/c4_full_svc_gaussian.npz (unmodified reference, 1000 iterations, 954.6 s on 8 host cores)',
           'iterations_compared': int(k),
           'f_hist_max_rel': float(np.max(np.abs(np.array(m.train_loss_history)[:k + 1] - g['f_hist'][:k + 1]) /
                                          np.maximum(1.0, np.abs(g['f_hist'][:k + 1]))))}
    if m.optimizer.iter == int(g['iter']):
        out.update({'max_abs_dalpha': float(np.abs(m.alphas_ - g['alphas']).max()),
                    'same_support': bool(np.array_equal(m.support_, g['support'])),
                    'f_x_rel': float(abs(m.optimizer.f_x - float(g['f_hist'][-1])) / abs(float(g['f_hist'][-1]))),
                    'intercept_abs': float(abs(m.intercept_ - float(g['intercept']))),
                    'meets_north_star': bool(np.abs(m.alphas_ - g['alphas']).max() <= 1e-8 and
                                             np.array_equal(m.support_, g['support']))})
    return out


def symv_streamed_bytes(n):
    import ctypes as C
    from optiml_b200 import _native as N
    b = C.c_int64(0)
    N.call('svmb200_symv_geometry', int(n), int(N.padded_ld(n)), C.byref(b), None, None, None)
    return int(b.value)


def symmetric_leg(args, ctx, make_model, X, y, dX, n, peak, barrier, max_over_ranks, gpus):
    """The same fits with the opt-in symmetric pass (runtime.use_symmetric_pass): device-resident leg + end-to-end leg, two
    fits each after one warm-up, its own parity block against the reference's golden run and its own roofline (bytes =
    what K2s streams per GPU: half the matrix + the lower halves of the diagonal blocks).  Every rank takes part."""
    from optiml_b200.runtime import use_symmetric_pass
    use_symmetric_pass(True)
    try:
        m = make_model().fit(X, y, X_device=dX)
        m.obj.release()
        fits = max(1, min(args.steps, 2))
        iters, mv_ms, samples, pg_ms, vec_ms = 0, 0.0, 0, 0.0, 0.0
        barrier()
        launches0 = ctx.launch_count()
        ctx.timer_start()
        for _ in range(fits):
            m = make_model().fit(X, y, X_device=dX)
            iters += m.optimizer.iter
            mv_ms += m.optimizer.matvec_ms
            vec_ms += m.optimizer.vector_ms
            samples += m.optimizer.profile_samples
            pg_ms += m.optimizer.device_ms
            m.obj.release()
        barrier()
        dev_ms = max_over_ranks(ctx.timer_stop_ms())
        launches = ctx.launch_count() - launches0
        t0 = time.perf_counter()
        e2e_iters = 0
        for _ in range(fits):
            m = make_model(False).fit(X, y)
            e2e_iters += m.optimizer.iter
            m.obj.release()
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        used = bool(m.optimizer.symmetric_pass)
        streamed = float(symv_streamed_bytes(n)) / gpus   # the ranks stream equal shares (within a band)
        avg_ms = mv_ms / max(samples, 1)
        achieved = streamed / (avg_ms / 1e3) / 1e9
        return {'what': 'the same workload with runtime.use_symmetric_pass(True) / SVMB200_SYMMETRIC=1: every product Q d read '
                        'from the upper triangle of Q alone (K2s; on row blocks every pair of off-diagonal blocks is read by '
                        'one of its two owners); reproducible, not bit-identical to the default pass',
                'used': used, 'value': iters / (dev_ms / 1e3), 'unit': UNIT, 'fits': fits, 'fit_s': dev_ms / fits / 1e3,
                'e2e': {'value': e2e_iters / e2e_s, 'unit': UNIT, 'fit_s': e2e_s / fits},
                'pg_its_per_s': iters / (pg_ms / 1e3),
                'per_iteration_us': {'product (tile pass + sends + combine)': 1e3 * avg_ms,
                                     'vector_phase': 1e3 * vec_ms / max(samples, 1)},
                'roofline': {'bound': 'hbm', 'kernel': 'symv_tile_kernel (+ symv_send_kernel, symv_combine_kernel) (K2s)',
                             'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                             'traffic': symv_traffic(n, gpus),
                             'bytes_per_launch': streamed, 'full_matrix_equivalent_gbs': 8.0 * n * n / gpus / (avg_ms / 1e3) / 1e9,
                             'frac_of_dram_theoretical': achieved / 8184.0},
                'gpu_launches': int(launches), 'parity': parity_block(args, m, n)}
    finally:
        use_symmetric_pass(False)


def symv_traffic(n, gpus):
    """DRAM bytes per product of the symmetric pass from the committed ncu capture (profiles/symv_traffic.json), or None"""
    tpath = os.path.join(ROOT, 'profiles', 'symv_traffic.json')
    if not os.path.exists(tpath):
        return None
    with open(tpath) as fh:
        tj = json.load(fh)
    if tj.get('n') == n and str(gpus) in tj.get('per_gpus', {}):
        return tj['per_gpus'][str(gpus)]['dram_bytes_per_launch']
    return None


def run_b200(args):
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend='nccl', device_id=torch.device('cuda', local_rank))
    from optiml_b200.configs import make_config
    from optiml_b200.ml.svm import DualSVC
    from optiml_b200.ml.svm.kernels import GaussianKernel
    from optiml_b200.runtime import default_context, use_devices

    group_size = 1
    if args.devices and world == 1:
        if args.devices.strip().lower() == 'all':
            import ctypes as C
            from optiml_b200 import _native as N
            cnt = C.c_int(0)
            N.call('svmb200_device_count', C.byref(cnt))
            devs = list(range(cnt.value))
        else:
            devs = [int(t) for t in args.devices.split(',')]
        use_devices(devs)
        group_size = len(devs)
    ctx = default_context()
    if args.symmetric:
        from optiml_b200.runtime import use_symmetric_pass
        use_symmetric_pass(True)
    spec, X0, y0 = make_config(args.config, n=args.n)
    n, d = X0.shape
    X = pinned_array(X0.shape)
    X[:] = X0
    y = y0
    dX = ctx.upload_matrix(X)  # resident copy for the device-timed leg

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()
        ctx.sync()

    def make_model(profile=True):
        m = DualSVC(kernel=GaussianKernel(), C=1, max_iter=args.max_iter)
        m.profile_matvec = profile  # CUDA events around one K2 / K3 launch in 16 (value leg: roofline, per_iteration_us)
        return m

    def max_over_ranks(v):
        if dist is None:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident leg: `value`
    for _ in range(args.warmup):
        m = make_model().fit(X, y, X_device=dX)
        m.obj.release()
    # a full Python GC pass costs ~0.2 s with torch + scikit-learn imported and would land in a random timed
    # step: collect now and move the survivors to the permanent generation (host-side hygiene only)
    import gc
    gc.collect()
    gc.freeze()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count()
    iters_total, mv_ms, mv_launches, pg_ms, comm_ms, vec_ms, step_s, step_parts, mv_samples = 0, 0.0, 0, 0.0, 0.0, 0.0, [], [], 0
    ctx.timer_start()
    t_wall = time.perf_counter()
    for _ in range(args.steps):
        ts = time.perf_counter()
        m = make_model().fit(X, y, X_device=dX)
        iters_total += m.optimizer.iter
        mv_ms += m.optimizer.matvec_ms
        comm_ms += m.optimizer.comm_ms
        vec_ms += m.optimizer.vector_ms
        mv_launches += m.optimizer.q_passes
        mv_samples += m.optimizer.profile_samples
        pg_ms += m.optimizer.device_ms
        m.obj.release()
        step_s.append(round(time.perf_counter() - ts, 4))
        step_parts.append({'gram_s': round(m.fit_times_['gram_s'], 4), 'solve_s': round(m.fit_times_['solve_s'], 4),
                           'pg_device_s': round(m.optimizer.device_ms / 1e3, 4)})
    barrier()
    dev_ms = ctx.timer_stop_ms()
    wall_s = time.perf_counter() - t_wall
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = max_over_ranks(dev_ms)
    value = iters_total / (dev_ms / 1e3)
    status, fx, nsv = m.optimizer.status, m.optimizer.f_x, len(m.support_)

    # ---- end-to-end leg through the public API with host buffers
    for _ in range(min(args.warmup, 1)):
        m = make_model(False).fit(X, y)
        m.obj.release()
    barrier()
    t0 = time.perf_counter()
    e2e_iters, e2e_steps = 0, []
    for _ in range(args.steps):
        ts = time.perf_counter()
        m = make_model(False).fit(X, y)
        e2e_iters += m.optimizer.iter
        m.obj.release()
        e2e_steps.append(round(time.perf_counter() - ts, 4))
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    nvars = n
    h2d = X.nbytes + 8 * n + 4 * 8 * nvars + 8 * n  # X, label signs, q/lb/ub/x0, intercept mask vector
    d2h = 2 * 8 * nvars + 2 * 8 * (args.max_iter + 1) + 8 * n  # alpha, gradient, f/|d| history, masked product

    sym_used = bool(getattr(m.optimizer, 'symmetric_pass', False))
    sym_leg = None
    if not args.symmetric and not args.no_symmetric_leg:
        try:
            sym_leg = symmetric_leg(args, ctx, make_model, X, y, dX, n, measured_peak()[0], barrier, max_over_ranks,
                                    world * group_size)
        except Exception as exc:   # the extra leg must never cost the line of the default pass
            sym_leg = {'error': repr(exc)}
    if rank != 0:
        return
    parity = parity_block(args, m, n)
    peak, peak_src = measured_peak()
    gpus = world * group_size
    bytes_per_launch = 8.0 * n * n / gpus
    kernel_name = 'matvec_seg_kernel (K2)'
    if sym_used:
        # K2s streams the upper triangle in band geometry (diagonal blocks in full): these are its algorithmic bytes
        bytes_per_launch = float(symv_streamed_bytes(n)) / gpus
        kernel_name = 'symv_tile_kernel + symv_combine_kernel (K2s, upper triangle; the timed pair)'
    mv_avg_ms = mv_ms / max(mv_samples, 1)  # CUDA events bracket one K2 launch in 16 (they serialise programmatic launches)
    achieved = bytes_per_launch / (mv_avg_ms / 1e3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'matvec_traffic.json')
    if os.path.exists(tpath):
        with open(tpath) as fh:
            tj = json.load(fh)
        if tj.get('n') == n and str(gpus) in tj.get('per_gpus', {}):
            traffic = tj['per_gpus'][str(gpus)]['dram_bytes_per_launch']   # ncu on this GPU count's shard shape
        elif tj.get('n') == n and tj.get('n_gpus', 1) == gpus:
            traffic = tj.get('dram_bytes_per_launch')
    line = {
        'metric': metric_name(args.config, n, d), 'value': value, 'unit': UNIT, 'n_gpus': gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': dev_ms / args.steps, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': bench_config(args.config, n, d, args.max_iter, gpus),
        'exchange': ctx.exchange if group_size == 1 else 'p2p',
        'launch_model': 'one process, one host thread, %d GPUs (device group)' % group_size if group_size > 1 else
                        ('torchrun, one process per GPU' if world > 1 else 'one process, one GPU'),
        'host_note': 'gc.collect() + gc.freeze() after warm-up (a full Python GC pass is ~0.2 s with torch/sklearn loaded); '
                     'in the value leg one iteration in 16 is bracketed by CUDA events (roofline / per_iteration_us); the '
                     'e2e leg runs without them',
        'parity': parity,
        'fit_s': dev_ms / args.steps / 1e3, 'pg_its_per_s': iters_total / (pg_ms / 1e3),
        'hbm_gbps_pg_loop': 8.0 * n * n * mv_launches / (pg_ms / 1e3) / 1e9,
        'frac_of_8TBps_nominal': 8.0 * n * n * mv_launches / (pg_ms / 1e3) / 1e9 / gpus / 8000.0,
        'iters_per_step': iters_total / args.steps, 'status': status, 'f_x': fx, 'n_sv': nsv,
        'wall_s_value_leg': wall_s, 'step_wall_s': step_s, 'step_parts': step_parts,
        'per_iteration_us': {'matvec': 1e3 * mv_ms / max(mv_samples, 1), 'allgather': 1e3 * comm_ms / max(mv_samples, 1),
                             'vector_phase': 1e3 * vec_ms / max(mv_samples, 1),
                             'pg_loop_total': 1e3 * pg_ms / max(mv_launches, 1)},
        'roofline': {'bound': 'hbm', 'kernel': kernel_name, 'achieved': achieved, 'peak': peak,
                     'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                     'bytes_per_launch': bytes_per_launch, 'avg_launch_ms': mv_avg_ms, 'launches_timed': mv_samples, 'launches_total': mv_launches,
                     'dram_theoretical_gbs': 8184.0, 'frac_of_dram_theoretical': achieved / 8184.0,
                     'note': 'peak = measured copy bandwidth (read+write stream); a read-only stream can exceed it; '
                             'dram_theoretical = 2048 B/clk x 3.996 GHz as reported by ncu'},
        'e2e': {'value': e2e_iters / e2e_s, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                'fit_s': e2e_s / args.steps, 'step_wall_s': e2e_steps, 'api': 'optiml_b200.ml.svm.DualSVC.fit(X_host_pinned, y_host)'},
        'gpu_launches': int(launches), 'clocks': clocks,
    }
    line['product_pass'] = 'symmetric (K2s: upper triangle of Q only, opt-in)' if sym_used else 'full (K2: every row of Q, the default)'
    if sym_used:
        line['hbm_gbps_pg_loop'] = bytes_per_launch * mv_launches / (pg_ms / 1e3) / 1e9
        line['frac_of_8TBps_nominal'] = line['hbm_gbps_pg_loop'] / gpus / 8000.0
        line['roofline']['traffic'] = symv_traffic(n, gpus)
        line['roofline']['full_matrix_equivalent_gbs'] = 8.0 * n * n / (mv_avg_ms / 1e3) / 1e9
    if sym_leg is not None:
        line['symmetric_pass'] = sym_leg
    if world == 1 and not args.no_cpu_baseline:
        # release the GPU arm's host copies first: the reference needs ~4.6 x 8 n^2 bytes of host memory
        r = reference_measure(args.config, n, args.max_iter)
        line['cpu_baseline'] = {
            'value': r['value'], 'unit': UNIT, 'cores': host_threads(), 'kind': r['kind'], 'sample': r['sample'],
            'pg_only_its': r['pg_only_its'], 'whole_fit_s': r['fit_full_s'], 'parts': r['parts'],
            'reference_full_run_note': 'the unmodified reference solver needed 954.6 s for the 1000 PG iterations of '
                                       'this config on the 8 host cores of the build container '
                                       '(tests/golden/c4_full_svc_gaussian.npz)'}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)
        world = int(os.environ.get('WORLD_SIZE', '1'))
        if world > 1 and int(os.environ.get('RANK', '0')) != 0:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()


if __name__ == '__main__':
    main()
