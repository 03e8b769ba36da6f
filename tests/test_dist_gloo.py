"""World-size-2 checks of the multi-rank HOST logic on CPU (gloo): the row partition from the C library,
the unique-id broadcast plumbing, and the sharded algorithm itself (row-block products exchanged by an
all-gather, vector phase replicated on every rank) emulated with NumPy -- it must reproduce the
single-process oracle and leave every rank with the identical iterate."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from optiml_b200 import runtime
        from oracle import svm_oracle as O
        # 1. unique-id broadcast plumbing
        uid = runtime.broadcast_unique_id(dist, lambda: bytes(range(128)))
        assert uid == bytes(range(128))
        # 2. row partition from the C library is a partition, identical on every rank
        n = 333
        shards = [runtime.shard_rows(n, r, world) for r in range(world)]
        assert sum(s[1] for s in shards) == n and shards[0][0] == 0 and all(s[0] % 64 == 0 for s in shards)
        # 3. sharded PG: w_shard = Q[rows] u, all-gather, replicated vector phase
        rng = np.random.default_rng(0)
        G = rng.standard_normal((n + 5, n))
        Q = G.T @ G / n
        q = rng.standard_normal(n)
        ub = rng.uniform(0.5, 2., n)
        lb = np.zeros(n)
        r0, nr = shards[rank]
        Qs = Q[r0:r0 + nr]

        def product(u):
            mine = torch.from_numpy(np.ascontiguousarray(Qs @ u))
            rpr = shards[0][1]
            buf = torch.zeros(rpr, dtype=torch.float64)
            buf[:nr] = mine
            parts = [torch.zeros(rpr, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(parts, buf)
            return torch.cat(parts).numpy()[:n]

        x = (lb + ub) / 2
        g = product(x) + q
        iters = 0
        for k in range(60):
            d = -g
            d[np.logical_and(ub - x <= 1e-12, d > 0)] = 0
            d[np.logical_and(x - lb <= 1e-12, d < 0)] = 0
            s = d.dot(d)
            if np.sqrt(s) <= 1e-6 or k >= 50:
                break
            pos, neg = d > 0, d < 0
            mt = np.min((ub[pos] - x[pos]) / d[pos]) if pos.any() else np.inf
            if neg.any():
                mt = min(mt, np.min((lb[neg] - x[neg]) / d[neg]))
            w = product(d)
            den = w.dot(d)
            t = mt if den <= 1e-16 else min(s / den, mt)
            x = x + t * d
            g = g + t * w
            iters += 1
        ref = O.projected_gradient(Q, q, ub, max_iter=50, passes=1)
        assert iters == ref.iter
        assert np.abs(x - ref.x).max() <= 1e-12
        others = [None] * world
        dist.all_gather_object(others, x.tobytes())
        assert len(set(others)) == 1  # identical iterate on every rank
        # 4. augmented-Lagrangian dual on the sharded matrix: an unseeded start point is drawn per rank and must be
        #    replaced by rank 0's (runtime.Context.broadcast_array); the replicated loop then agrees everywhere
        from oracle import al_oracle as AL
        ctx = runtime.Context.__new__(runtime.Context)  # no device needed for the host plumbing
        ctx.rank, ctx.nranks = rank, world
        x0 = ctx.broadcast_array(np.random.default_rng(100 + rank).uniform(size=n))
        assert np.array_equal(x0, np.random.default_rng(100).uniform(size=n))
        A = np.where(np.arange(n) % 3 == 0, 1., -1.)
        al = AL.al_stochastic(product, q, lb, ub, x0, A=A, rho=1., rule='adagrad', step_size=1., tol=1e-9, epochs=40)
        solo = AL.al_stochastic(lambda v: Q @ v, q, lb, ub, x0, A=A, rho=1., rule='adagrad', step_size=1., tol=1e-9, epochs=40)
        assert al.iter == solo.iter and np.abs(al.x - solo.x).max() <= 1e-12 and np.abs(al.dual_x - solo.dual_x).max() <= 1e-11
        dist.all_gather_object(others, al.x.tobytes() + al.dual_x.tobytes())
        assert len(set(others)) == 1
        with open(os.path.join(out_dir, f'ok{rank}'), 'w') as fh:
            fh.write('ok')
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ['ok0', 'ok1']
