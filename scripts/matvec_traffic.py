"""DRAM traffic of the streaming matvec (K2) on the shard shapes of C4 at 1, 2, 4 and 8 GPUs: run under
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:matvec_seg_kernel --csv
(one GPU: the kernel a rank of a P-GPU job launches is `svmb200_matvec` on its svmb200_shard_rows block; the peer stores of the
fused exchange add 16 bytes per row and peer, < 0.1 % of the shard).  `python scripts/matvec_traffic.py parse <csv>` turns the
ncu log into profiles/matvec_traffic.json, which bench.py reports as `roofline.traffic` for the matching GPU count."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
N, REPS = 50000, 3


def run():
    import ctypes as C
    import numpy as np
    from optiml_b200 import _native as N_
    from optiml_b200.runtime import default_context, shard_rows
    ctx = default_context()
    ld = N_.padded_ld(N)
    du, dw = ctx.malloc(ld * 8), ctx.malloc(N * 8)
    ctx.h2d(du, np.random.default_rng(0).standard_normal(ld))
    for P in (1, 2, 4, 8):
        rows = shard_rows(N, 0, P)[1]
        dQ = ctx.malloc(rows * ld * 8)
        ctx.memset(dQ, 0, rows * ld * 8)
        for _ in range(REPS):
            N_.call('svmb200_matvec', ctx.handle, C.c_void_p(dQ), rows, ld, C.c_void_p(du), C.c_void_p(dw))
        ctx.sync()
        if len(sys.argv) > 1 and sys.argv[1] == 'time':   # back-to-back launches, CUDA events (not under ncu)
            reps = 100 * P
            ctx.timer_start()
            for _ in range(reps):
                N_.call('svmb200_matvec', ctx.handle, C.c_void_p(dQ), rows, ld, C.c_void_p(du), C.c_void_p(dw))
            ms = ctx.timer_stop_ms() / reps
            print(json.dumps({'gpus': P, 'shard_rows': rows, 'us_per_pass': round(ms * 1e3, 2),
                              'gbps': round(8.0 * rows * N / ms / 1e6, 1), 'ideal_us_at_P1_rate': None}), flush=True)
        ctx.free(dQ, rows * ld * 8)
        ctx.trim()
        print(f'P={P} rows={rows}', flush=True)


def parse(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    H = rows[h]
    name, val, kern, idc = H.index('Metric Name'), H.index('Metric Value'), H.index('Kernel Name'), H.index('ID')
    launches = {}
    for r in rows[h + 1:]:
        if len(r) > val and 'matvec_seg_kernel' in r[kern]:
            launches.setdefault(int(r[idc]), {})[r[name]] = float(r[val].replace(',', ''))
    ids = sorted(launches)
    unit = {r[name]: r[H.index('Metric Unit')] for r in rows[h + 1:] if len(r) > val}
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    out = {'n': N, 'kernel': 'matvec_seg_kernel', 'per_gpus': {},
           'source': f'{os.path.basename(path)}: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum, {REPS} launches per shard '
                     f'shape (scripts/matvec_traffic.py), mean per launch'}
    for k, P in enumerate((1, 2, 4, 8)):
        grp = [launches[i] for i in ids[k * REPS:(k + 1) * REPS]]
        rd = sum(g['dram__bytes_read.sum'] * scale[unit['dram__bytes_read.sum']] for g in grp) / len(grp)
        wr = sum(g['dram__bytes_write.sum'] * scale[unit['dram__bytes_write.sum']] for g in grp) / len(grp)
        out['per_gpus'][str(P)] = {'dram_bytes_per_launch': rd + wr, 'dram_bytes_read': rd, 'dram_bytes_write': wr,
                                   'algorithmic_bytes': 8.0 * N * N / P}
    out['n_gpus'] = 1
    out['dram_bytes_per_launch'] = out['per_gpus']['1']['dram_bytes_per_launch']
    with open(os.path.join(ROOT, 'profiles', 'matvec_traffic.json'), 'w') as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out['per_gpus']))


if __name__ == '__main__':
    if len(sys.argv) > 2 and sys.argv[1] == 'parse':
        parse(sys.argv[2])
    else:
        run()
