"""Shape sweep of the multi-vector streaming pass K2 x NB (widening 8f-4) at the headline size.

For every (R, U, MINB[, H]) variant: rebuild the library with -DSVMB200_MULTI_R/U/MINB/H into its own directory (nvcc is on
the GPU box), then, in a subprocess that loads that variant, time svmb200_matvec_multi with NB = 1..4 vectors against a
resident n x n matrix with CUDA events (inputs larger than L2: 20 GB at n = 50 000) and check every result bitwise
against svmb200_matvec.  One JSON line per (variant, NB): ms per pass, HBM GB/s of the pass (8 n^2 bytes), and the
per-problem rate NB x that.

    python scripts/sweep_multi.py [--n 50000] [--reps 20] [--variants R:U:MINB[:H],...]
    python scripts/sweep_multi.py --child ...   (internal)
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(n, reps, tag):
    from optiml_b200 import _native as N
    from optiml_b200.runtime import default_context
    ctx = default_context()
    ld = N.padded_ld(n)
    rng = np.random.default_rng(0)
    dQ = ctx.malloc(8 * n * ld)
    block = 4096
    for r0 in range(0, n, block):   # fill the matrix in row blocks (host RAM stays small)
        rows = min(block, n - r0)
        chunk = np.zeros((rows, ld))
        chunk[:, :n] = rng.standard_normal((rows, n))
        N.call('svmb200_h2d', ctx.handle, C.c_void_p(dQ + 8 * r0 * ld), chunk.ctypes.data_as(C.c_void_p), chunk.nbytes)
    dus, dws, single = [], [], []
    for b in range(4):
        u = np.zeros(ld)
        u[:n] = rng.standard_normal(n)
        dus.append(ctx.malloc(u.nbytes))
        ctx.h2d(dus[-1], u)
        dws.append(ctx.malloc(8 * n))
        N.call('svmb200_matvec', ctx.handle, C.c_void_p(dQ), n, ld, C.c_void_p(dus[b]), C.c_void_p(dws[b]))
        w = np.empty(n)
        ctx.d2h(w, dws[b])
        single.append(w)
    for nb in (1, 2, 3, 4):
        du, dw = (C.c_void_p * nb)(*dus[:nb]), (C.c_void_p * nb)(*dws[:nb])
        for _ in range(3):
            N.call('svmb200_matvec_multi', ctx.handle, C.c_void_p(dQ), n, ld, du, dw, nb)
        ctx.timer_start()
        for _ in range(reps):
            N.call('svmb200_matvec_multi', ctx.handle, C.c_void_p(dQ), n, ld, du, dw, nb)
        ms = ctx.timer_stop_ms() / reps
        same = True
        for b in range(nb):
            w = np.empty(n)
            ctx.d2h(w, dws[b])
            same = same and bool(np.array_equal(w, single[b]))
        gbs = 8.0 * n * n / (ms / 1e3) / 1e9
        print(json.dumps(dict(variant=tag, n=n, nb=nb, ms_per_pass=round(ms, 4), hbm_gbps=round(gbs, 1),
                              per_problem_gbps=round(nb * gbs, 1), bitwise_equal_to_single=same)), flush=True)
    # the single-vector kernel for reference
    ctx.timer_start()
    for _ in range(reps):
        N.call('svmb200_matvec', ctx.handle, C.c_void_p(dQ), n, ld, C.c_void_p(dus[0]), C.c_void_p(dws[0]))
    ms = ctx.timer_stop_ms() / reps
    print(json.dumps(dict(variant='matvec_seg_kernel', n=n, nb=1, ms_per_pass=round(ms, 4),
                          hbm_gbps=round(8.0 * n * n / (ms / 1e3) / 1e9, 1))), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=50000)
    ap.add_argument('--reps', type=int, default=20)
    ap.add_argument('--variants', default='4:2:0,8:1:2,8:1:1,2:4:3,16:1:1,4:2:1:2,2:4:1:2')
    ap.add_argument('--child', default=None)
    args = ap.parse_args()
    if args.child is not None:
        return child(args.n, args.reps, args.child)
    from optiml_b200.csrc import build
    for v in args.variants.split(','):
        fields = [int(t) for t in v.split(':')]
        r, u, minb = fields[:3]
        h = fields[3] if len(fields) > 3 else 1   # thread groups per CTA (L1 reuse of the vector operands)
        lib_dir = os.path.join(ROOT, 'optiml_b200', '_lib_sweep', f'R{r}_U{u}_B{minb}_H{h}')
        try:
            lib = build.build(defines=(f'SVMB200_MULTI_R={r}', f'SVMB200_MULTI_U={u}', f'SVMB200_MULTI_MINB={minb}',
                                       f'SVMB200_MULTI_H={h}'), lib_dir=lib_dir)
        except subprocess.CalledProcessError as exc:
            print(json.dumps(dict(variant=v, error=f'build failed: {exc}')), flush=True)
            continue
        env = dict(os.environ, SVMB200_LIB=lib)
        subprocess.run([sys.executable, os.path.abspath(__file__), '--child', f'R{r}_U{u}_B{minb}_H{h}', '--n', str(args.n),
                        '--reps', str(args.reps)], env=env, check=False)


if __name__ == '__main__':
    main()
