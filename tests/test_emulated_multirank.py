"""SURVEY.md 8(e) on the host emulation: row-sharded solves with ranks as THREADS of the test process.

Each thread owns a context (= one emulated GPU), joins the communicator through the product's own comm.cu (NCCL and
CUDA IPC replaced by in-process stand-ins, tests/cuda_emu/emu_runtime.cpp) and runs the real solver kernels on its
row shard.  Both exchange paths are driven: the all-gather, and the fused peer-memory exchange (K2 stores tagged
entries straight into every rank's arena, K3 waits on them).  The property asserted is the one the GPU runs showed at
2 and 8 GPUs: the iterate is BIT-IDENTICAL to the single-rank solve for every rank count, including ragged and empty
shards.  What the emulation cannot show: memory ordering over NVLink, timing.
"""
import ctypes as C
import gc
import threading
import time
import warnings

import numpy as np
import pytest

import shared_gram_checks as S
from emu import emulated_device
from optiml_b200 import _native as N


def run_ranks(nranks, exchange, body):
    """body(ctx) on `nranks` threads, each with its own context attached to one communicator; returns the results in
    rank order.  exchange: 'nccl' (all-gather) or 'p2p' (fused peer-memory exchange; one rank has nothing to exchange)."""
    from optiml_b200.runtime import Context
    uid = Context.new_unique_id()
    gate = threading.Barrier(nranks)
    handles, results, errors = [None] * nranks, [None] * nranks, []

    def main(rank):
        try:
            ctx = Context(device=0)
            ctx.attach_communicator(rank, nranks, uid)
            if exchange == 'p2p' and nranks > 1:
                buf = C.create_string_buffer(64)
                N.call('svmb200_comm_p2p_export', ctx.handle, 1 << 20, C.cast(buf, C.c_void_p))
                handles[rank] = buf.raw
                gate.wait(timeout=60)
                blob = C.create_string_buffer(b''.join(handles), 64 * nranks)
                N.call('svmb200_comm_p2p_attach', ctx.handle, C.cast(blob, C.c_void_p), nranks)
                gate.wait(timeout=60)
                assert ctx.exchange == 'p2p'
            else:
                assert ctx.exchange == ('nccl' if nranks > 1 else 'none')
            results[rank] = body(ctx)
            gc.collect()   # cyclic garbage of the body holds device buffers of this context: finalise it before the context goes
            assert N.load_library().emu_sticky_error() == 0
            gate.wait(timeout=120)   # nobody tears its arena down while a peer may still store into it
            ctx._finalizer()
        except BaseException as exc:  # noqa: surfaced in the main thread below
            errors.append((rank, exc))
            gate.abort()

    # daemon threads: a rank that is stuck (a protocol bug would show up like that) fails the test, it cannot hang the run
    threads = [threading.Thread(target=main, args=(r,), daemon=True) for r in range(nranks)]
    for t in threads:
        t.start()
    deadline = time.monotonic() + 240
    for t in threads:
        t.join(timeout=max(0.1, deadline - time.monotonic()))
    if errors:
        raise errors[0][1]
    assert not any(t.is_alive() for t in threads), 'a rank is stuck'
    return results


def shard_hessian(ctx, M, layout='plain'):
    from optiml_b200.runtime import DeviceHessian
    n = M.shape[0]
    H = DeviceHessian(ctx, n, layout)
    block = np.zeros((max(H.nrows, 1), H.ld))
    block[:H.nrows, :n] = M[H.row0:H.row0 + H.nrows]
    ctx.h2d(H.matrix.dptr, block)
    return H


def solve(kind, H, q, ub, max_iter):
    from optiml_b200.opti import Quadratic
    from optiml_b200.opti.constrained import AugmentedLagrangianQuadratic, FrankWolfe, ProjectedGradient
    from optiml_b200.opti.unconstrained.stochastic import Adam
    quad = Quadratic(H, q)
    if kind == 'pg':
        s = ProjectedGradient(quad=quad, ub=ub, max_iter=max_iter)
    elif kind == 'fw':
        s = FrankWolfe(quad=quad, ub=ub, max_iter=max_iter, t=0.1)
    else:
        f = AugmentedLagrangianQuadratic(primal=quad, lb=np.zeros_like(ub), ub=ub, rho=1.)
        s = Adam(f=f, step_size=0.05, epochs=max_iter, random_state=3, tol=1e-6, momentum_type='polyak', momentum=0.3)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        s.minimize()
    return S.solver_state(s)


@pytest.mark.parametrize('kind,layout,n,nranks,exchange', [
    ('pg', 'plain', 200, 2, 'nccl'), ('pg', 'plain', 200, 2, 'p2p'), ('pg', 'plain', 200, 3, 'p2p'),
    ('pg', 'plain', 70, 4, 'p2p'),       # 64-row shards: ranks 2 and 3 own nothing
    ('pg', 'svr', 150, 2, 'p2p'), ('fw', 'plain', 130, 3, 'nccl'), ('fw', 'plain', 130, 2, 'p2p'),
    ('adam', 'plain', 130, 2, 'p2p'), ('adam', 'svr', 100, 2, 'nccl'),
])
def test_sharded_solve_is_bit_identical_to_one_rank(kind, layout, n, nranks, exchange):
    rng = np.random.default_rng(n + nranks)
    M = S.psd(rng, n)
    nv = 2 * n if layout == 'svr' else n
    q, ub = rng.standard_normal(nv), np.full(nv, 1.5)
    max_iter = 10
    with emulated_device() as lib:
        one = run_ranks(1, 'nccl', lambda ctx: solve(kind, shard_hessian(ctx, M, layout), q, ub, max_iter))[0]
        gathers = lib.emu_allgather_calls()
        many = run_ranks(nranks, exchange, lambda ctx: solve(kind, shard_hessian(ctx, M, layout), q, ub, max_iter))
        gathers = lib.emu_allgather_calls() - gathers
        # the fused exchange needs no collective per ITERATION -- only the one-double barrier at solver creation; a
        # problem that leaves a rank without rows falls back to the all-gather (an empty rank publishes nothing, nobody
        # would wait for it, and it could be lapped: csrc/pg.cu, bcqp_create)
        rows_per_rank = -(-(-(-n // nranks)) // 64) * 64     # ceil(n / P) rounded up to the 64-row group
        fused = exchange == 'p2p' and (nranks - 1) * rows_per_rank < n
        assert gathers == nranks if fused else gathers >= nranks * max_iter
    for rank_state in many:
        for a, b in zip(one, rank_state):
            assert np.array_equal(a, b)


@pytest.mark.parametrize('kind,count,nranks', [('pg', 3, 2), ('fw', 5, 2), ('adagrad', 2, 3)])
def test_sharded_lockstep_batch_is_bit_identical_to_one_rank(kind, count, nranks):
    """the batched (one-vs-rest) driver on row shards, with the all-gather per problem and with the fused peer exchange
    for the whole batch (own arena region, one tag per iteration)"""
    from optiml_b200.opti import Quadratic
    from optiml_b200.opti.batch import minimize_batch
    rng = np.random.default_rng(count + nranks)
    n = 150
    M = S.psd(rng, n, shift=0.0) + 1.0
    signs = [np.where(rng.random(n) < 0.4, 1.0, -1.0) for _ in range(count)]
    q, ub = -np.ones(n), np.ones(n)

    def body(ctx):
        shared = shard_hessian(ctx, M)
        solvers = S.make_solvers(kind, lambda c: Quadratic(shared.with_signs(signs[c]), q), signs, ub, 8, [1e-6])
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            minimize_batch(solvers)
        assert all(s.batch_size_ == count for s in solvers)
        return [S.solver_state(s) for s in solvers]

    with emulated_device() as lib:
        one = run_ranks(1, 'nccl', body)[0]
        for exchange in ('nccl', 'p2p'):
            gathers = lib.emu_allgather_calls()
            states = run_ranks(nranks, exchange, body)
            gathers = lib.emu_allgather_calls() - gathers
            # the fused exchange needs no collective per pass (one creation barrier per member); the fallback gathers
            # every problem's shard every pass
            assert gathers == nranks * count if exchange == 'p2p' else gathers >= nranks * count * 8
            for rank_states in states:
                for sa, sb in zip(one, rank_states):
                    for a, b in zip(sa, sb):
                        assert np.array_equal(a, b)


def _fuzz_case(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.choice([2, 3, 5, 17, 63, 64, 65, 100, 129, 200]))
    case = dict(n=n, kind=str(rng.choice(['pg', 'fw'])), layout=str(rng.choice(['plain', 'svr'])),
                nranks=int(rng.choice([1, 2, 3, 4])), exchange=str(rng.choice(['nccl', 'p2p'])),
                matrix=str(rng.choice(['low rank', 'full rank', 'zero'])))
    if case['matrix'] == 'zero':
        M = np.zeros((n, n))
    else:
        G = rng.standard_normal((n, max(1, n // 3) if case['matrix'] == 'low rank' else n + 2))
        M = G @ G.T / G.shape[1]
    nv = 2 * n if case['layout'] == 'svr' else n
    lb = np.where(rng.random(nv) < 0.5, 0.0, -rng.random(nv))
    ub = lb + rng.uniform(0.1, 2.0, nv)
    case.update(M=M, q=rng.standard_normal(nv), lb=lb, ub=ub,
                x0=lb + rng.random(nv) * (ub - lb) if rng.random() < 0.5 else None, iters=int(rng.choice([1, 2, 6, 15])))
    return case


@pytest.mark.parametrize('block', range(4))
def test_random_problems_shards_exchanges_and_schedules(block):
    """Seeded fuzz: sizes from 2 to 200 (below, at and above the 64-row group), 1-4 ranks with both exchanges, plain
    and SVR layouts, zero / low-rank / full-rank matrices, lower bounds != 0, given or default start points, 1-15
    iterations, shuffled thread and block schedules.  Every case: sharded == one rank bitwise, and == the oracle."""
    from optiml_b200.opti import Quadratic
    from optiml_b200.opti.constrained import FrankWolfe, ProjectedGradient
    from oracle import svm_oracle as O

    def run(ctx, c):
        H = shard_hessian(ctx, c['M'], c['layout'])
        cls, kw = (ProjectedGradient, {}) if c['kind'] == 'pg' else (FrankWolfe, dict(t=0.25))
        s = cls(quad=Quadratic(H, c['q']), ub=c['ub'], lb=c['lb'], x=None if c['x0'] is None else c['x0'].copy(),
                max_iter=c['iters'], **kw).minimize()
        hist = np.asarray(getattr(s, 'f_hist', getattr(s, 'f_x_history', [])))   # ndim <= 3 takes the step-wise path
        return [np.asarray(s.x), np.asarray(s.g_x), np.array([s.iter, s.f_x]), np.array([s.status == 'optimal']), hist]

    for seed in range(15 * block, 15 * block + 15):
        c = _fuzz_case(seed)
        label = {k: v for k, v in c.items() if not isinstance(v, np.ndarray)}
        with emulated_device(order=2, seed=seed):
            one = run_ranks(1, 'nccl', lambda ctx: run(ctx, c))[0]
            many = run_ranks(c['nranks'], c['exchange'], lambda ctx: run(ctx, c)) if c['nranks'] > 1 else []
        for state in many:
            assert all(np.array_equal(a, b) for a, b in zip(one, state)), label
        M = c['M']
        Q = M if c['layout'] == 'plain' else np.vstack((np.hstack((M, -M)), np.hstack((-M, M))))
        solver, kw = (O.projected_gradient, {}) if c['kind'] == 'pg' else (O.frank_wolfe, dict(t=0.25))
        want = solver(Q, c['q'], c['ub'], lb=c['lb'], x0=c['x0'], max_iter=c['iters'], **kw)
        assert int(one[2][0]) == want.iter, label
        assert np.abs(one[0] - want.x).max() <= 1e-9 * max(1., np.abs(want.x).max()), label


def _batch_case(seed):
    rng = np.random.default_rng(5000 + seed)
    n = int(rng.choice([4, 17, 64, 65, 100, 130]))
    c = dict(n=n, kind=str(rng.choice(['pg', 'fw', 'adagrad', 'adam'])), layout=str(rng.choice(['plain', 'plain', 'svr'])),
             count=int(rng.choice([2, 3, 4, 5, 7, 9])), nranks=int(rng.choice([1, 2, 3])),
             exchange=str(rng.choice(['nccl', 'p2p'])), iters=int(rng.choice([1, 2, 7, 12])))
    G = rng.standard_normal((n, n // 2 + 2))
    nv = 2 * n if c['layout'] == 'svr' else n
    signed = c['layout'] == 'plain' and rng.random() < 0.8
    c.update(M=G @ G.T / G.shape[1] + 1.0, signed=signed,
             signs=[np.where(rng.random(n) < 0.5, 1.0, -1.0) for _ in range(c['count'])] if signed else None,
             qs=[rng.standard_normal(nv) for _ in range(c['count'])], ubs=[rng.uniform(0.3, 2.0, nv) for _ in range(c['count'])],
             eps=[float(rng.choice([1e-6, 0.3, 5.0])) for _ in range(c['count'])])
    return c


@pytest.mark.parametrize('block', range(3))
def test_random_lockstep_batches(block):
    """Seeded fuzz of the batched driver: 2-9 problems (one to three multi-vector launches per iteration), with and
    without label signs, SVR blocks, per-problem right-hand sides, bounds and stopping thresholds (some problems stop
    at once, some never), PG / FW / AdaGrad / Adam-Nesterov with and without an equality row, 1-3 ranks with both
    exchanges.  Every case: lockstep batch (sharded) == the solvers one after the other on one rank, bitwise."""
    from optiml_b200.opti import Quadratic
    from optiml_b200.opti.batch import batchable, minimize_batch
    from optiml_b200.opti.constrained import AugmentedLagrangianQuadratic, FrankWolfe, ProjectedGradient
    from optiml_b200.opti.unconstrained.stochastic import AdaGrad, Adam

    def run(ctx, c, batch):
        H = shard_hessian(ctx, c['M'], c['layout'])
        solvers = []
        for i in range(c['count']):
            quad, ub = Quadratic(H.with_signs(c['signs'][i]) if c['signed'] else H, c['qs'][i]), c['ubs'][i]
            if c['kind'] == 'pg':
                solvers.append(ProjectedGradient(quad=quad, ub=ub, max_iter=c['iters'], eps=c['eps'][i]))
            elif c['kind'] == 'fw':
                solvers.append(FrankWolfe(quad=quad, ub=ub, max_iter=c['iters'], eps=c['eps'][i], t=0.1))
            else:
                A = np.where(np.arange(len(ub)) % 2 == 0, 1.0, -1.0) if i % 2 == 0 else None
                f = AugmentedLagrangianQuadratic(primal=quad, A=A, b=None if A is None else np.zeros(1), lb=np.zeros(len(ub)),
                                                 ub=ub, rho=1.5)
                solvers.append(AdaGrad(f=f, step_size=0.5, epochs=c['iters'], tol=1e-3 if i == 1 else 1e-12, random_state=i)
                               if c['kind'] == 'adagrad' else
                               Adam(f=f, step_size=0.02, epochs=c['iters'], tol=1e-12, random_state=i, momentum_type='nesterov',
                                    momentum=0.4))
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            if batch:
                assert batchable(solvers)
                minimize_batch(solvers)
                assert all(s.batch_size_ == c['count'] for s in solvers)
            else:
                for s in solvers:
                    s.minimize()
        return [S.solver_state(s) for s in solvers]

    for seed in range(8 * block, 8 * block + 8):
        c = _batch_case(seed)
        label = {k: v for k, v in c.items() if not isinstance(v, (np.ndarray, list))}
        with emulated_device(order=2, seed=seed):
            sequential = run_ranks(1, 'nccl', lambda ctx: run(ctx, c, False))[0]
            batched = run_ranks(c['nranks'], c['exchange'], lambda ctx: run(ctx, c, True))
        for rank_states in batched:
            for sa, sb in zip(sequential, rank_states):
                assert all(np.array_equal(a, b) for a, b in zip(sa, sb)), label


def test_rank_without_rows_cannot_be_lapped():
    """Regression for a hazard the fuzz found: with n <= 64 (P - 1) some ranks own no rows; under the fused exchange
    nobody waited for them, so a rank with rows could run two products ahead and overwrite tagged entries an empty
    rank had not read yet (intermittent: the empty rank then spun until its 20 s time-out).  Such problems now take
    the all-gather; many short solves with a deliberately slow empty rank must all agree with one rank."""
    from optiml_b200.opti import Quadratic
    from optiml_b200.opti.constrained import ProjectedGradient
    rng = np.random.default_rng(428)
    n = 5
    G = rng.standard_normal((n, n + 2))
    M, q, ub = G @ G.T / (n + 2), rng.standard_normal(n), np.full(n, 2.0)

    def body(ctx):
        out = []
        for rep in range(12):
            if ctx.rank == ctx.nranks - 1:
                time.sleep(0.002 * (rep % 3))       # the empty rank dawdles
            s = ProjectedGradient(quad=Quadratic(shard_hessian(ctx, M), q), ub=ub, max_iter=40).minimize()
            out.append(np.concatenate((s.x, [s.iter])))
        return np.array(out)

    with emulated_device() as lib:
        one = run_ranks(1, 'nccl', body)[0]
        gathers = lib.emu_allgather_calls()
        many = run_ranks(3, 'p2p', body)
        assert lib.emu_allgather_calls() - gathers > 0          # fell back to the all-gather
    for state in many:
        assert np.array_equal(state, one)


def test_barrier_before_the_first_fused_product(monkeypatch):
    """One tiny all-gather per solver creation (the guarantee that every rank's previous solve has left the arena), then
    the fused exchange -- same bits, exactly `ranks x solves` collectives; SVMB200_P2P_CREATE_BARRIER=0 removes it (A/B
    timing only)"""
    from optiml_b200.opti import Quadratic
    from optiml_b200.opti.constrained import ProjectedGradient
    rng = np.random.default_rng(6)
    sizes = (200, 150, 260, 130)     # solves of different size back to back: the layouts of the arena change
    problems = [(S.psd(rng, n), rng.standard_normal(n), np.full(n, 1.5)) for n in sizes]

    def body(ctx):
        return [ProjectedGradient(quad=Quadratic(shard_hessian(ctx, M), q), ub=ub, max_iter=6).minimize().x for M, q, ub in problems]

    with emulated_device() as lib:
        one = run_ranks(1, 'nccl', body)[0]
        monkeypatch.delenv('SVMB200_P2P_CREATE_BARRIER', raising=False)
        gathers = lib.emu_allgather_calls()
        many = run_ranks(2, 'p2p', body)
        assert lib.emu_allgather_calls() - gathers == 2 * len(sizes)
        monkeypatch.setenv('SVMB200_P2P_CREATE_BARRIER', '0')
        gathers = lib.emu_allgather_calls()
        unguarded = run_ranks(2, 'p2p', body)
        assert lib.emu_allgather_calls() - gathers == 0
    for state in many + unguarded:
        assert all(np.array_equal(a, b) for a, b in zip(one, state))


def test_many_back_to_back_solves_of_alternating_sizes():
    """The hardware stress of tests/multigpu_check.py in small: 48 solves back to back on three ranks, sizes alternating so
    that consecutive arena layouts overlap across parities and one size leaves the last rank without rows (that one
    takes the all-gather) -- every repeat bit-identical to one rank, nobody stalls."""
    from optiml_b200.opti import Quadratic
    from optiml_b200.opti.constrained import ProjectedGradient
    rng = np.random.default_rng(12)
    sizes = (200, 128, 330, 90)      # 128 on 3 ranks: 64-row shards, rank 2 owns nothing
    problems = [(S.psd(rng, n), rng.standard_normal(n), np.full(n, 1.5)) for n in sizes]
    iters = (3, 1, 5)

    def body(ctx):
        quads = [Quadratic(shard_hessian(ctx, M), q) for M, q, _ in problems]
        return [ProjectedGradient(quad=quads[i % 4], ub=problems[i % 4][2], max_iter=iters[i % 3]).minimize().x for i in range(48)]

    with emulated_device():
        one = run_ranks(1, 'nccl', body)[0]
        many = run_ranks(3, 'p2p', body)
    for state in many:
        assert all(np.array_equal(a, b) for a, b in zip(one, state))


# ---------------------------------------------------------------------------------------------- one process, N GPUs
@pytest.fixture
def device_group():
    """a single-process device group of three emulated GPUs that shards problems from 64 rows per GPU on"""
    from optiml_b200 import runtime
    with emulated_device() as lib:
        saved = runtime.DeviceGroup.MIN_ROWS_PER_GPU
        runtime.DeviceGroup.MIN_ROWS_PER_GPU = 64
        runtime.use_devices([0, 1, 2])
        try:
            yield runtime
        finally:
            runtime.use_devices(None)
            runtime.DeviceGroup.MIN_ROWS_PER_GPU = saved


def test_single_process_group_is_bit_identical_to_one_gpu(device_group):
    """``SVC.fit`` in ONE Python process on a group of GPUs (runtime.use_devices / SVMB200_DEVICES: no torchrun, no NCCL):
    same partition and kernels as the torchrun ranks, one host thread enqueuing iteration-major -- alpha, intercept,
    histories and decision values equal the one-GPU fit bit for bit, for every solver family"""
    runtime = device_group
    from optiml_b200.ml.svm import SVC, DualSVC, DualSVR
    from optiml_b200.ml.svm.kernels import GaussianKernel, PolyKernel
    from optiml_b200.ml.svm.losses import hinge
    from optiml_b200.opti import Quadratic
    from optiml_b200.opti.constrained import FrankWolfe, ProjectedGradient
    from optiml_b200.opti.unconstrained.stochastic import AdaGrad
    rng = np.random.default_rng(0)
    X = rng.standard_normal((300, 5))
    y = (X[:, 0] + 0.2 * rng.standard_normal(300) > 0).astype(int)
    t = X @ rng.standard_normal(5)
    makers = [lambda: DualSVC(kernel=GaussianKernel(), C=1, max_iter=30),
              lambda: DualSVR(kernel=PolyKernel(degree=2), C=1, max_iter=25),
              lambda: DualSVC(kernel=GaussianKernel(), C=1, max_iter=25, optimizer=FrankWolfe),
              lambda: SVC(loss=hinge, kernel=GaussianKernel(), C=1, reg_intercept=False, dual=True, optimizer=AdaGrad,
                          learning_rate=1., max_iter=25, random_state=5)]

    def fit_all():
        out = []
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            for mk in makers:
                m = mk()
                m.fit(X, t if isinstance(m, DualSVR) else y)
                out.append((type(m.obj.device_hessian()).__name__, m.alphas_.copy(), m.intercept_,
                            np.array(m.train_loss_history), m.decision_function(X[:40]), len(m.support_)))
                m.obj.release()
        return out

    grouped = fit_all()
    assert all(g[0] == 'GroupHessian' for g in grouped)
    # a generic callback drives the group step by step; a host-resident Q is sharded on upload
    G = rng.standard_normal((205, 200))
    Q, q, ub = G.T @ G / 200, rng.standard_normal(200), np.full(200, 1.5)
    seen = []
    stepwise = ProjectedGradient(quad=Quadratic(Q, q), ub=ub, max_iter=7, callback=lambda o: seen.append(o.f_x)).minimize()
    assert len(seen) == stepwise.iter + 1 == 8
    # small problems stay on the solo context
    small = DualSVC(kernel=GaussianKernel(), C=1, max_iter=5).fit(X[:100], y[:100])
    assert type(small.obj.device_hessian()).__name__ == 'DeviceHessian'
    runtime.use_devices(None)
    solo = fit_all()
    assert all(s[0] == 'DeviceHessian' for s in solo)
    for g, s in zip(grouped, solo):
        assert np.array_equal(g[1], s[1]) and g[2] == s[2] and np.array_equal(g[3], s[3]) and np.array_equal(g[4], s[4])
        assert g[5] == s[5]
    one = ProjectedGradient(quad=Quadratic(Q, q), ub=ub, max_iter=7).minimize()
    assert np.array_equal(one.x, stepwise.x) and np.allclose(seen, one.f_hist, rtol=0, atol=0)


def test_single_process_group_under_sklearn_meta_estimators(device_group):
    """GridSearchCV / OneVsRestClassifier over the drop-in work unchanged on a device group (clones are fitted one after
    the other, each on all GPUs; the lockstep batch is a one-context feature)"""
    from sklearn.model_selection import GridSearchCV
    from optiml_b200.ml.multiclass import OneVsRestClassifier
    from optiml_b200.ml.svm import DualSVC
    from optiml_b200.ml.svm.kernels import GaussianKernel
    rng = np.random.default_rng(1)
    X = rng.standard_normal((420, 4))
    y = (X[:, 0] > 0).astype(int)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        gs = GridSearchCV(DualSVC(kernel=GaussianKernel(), max_iter=15), {'C': [0.5, 2.0]}, cv=2).fit(X, y)
        assert gs.best_estimator_.score(X, y) > 0.8
        assert type(gs.best_estimator_.obj.device_hessian()).__name__ == 'GroupHessian'
        y3 = np.digitize(X[:, 1], [-0.5, 0.5])
        ovr = OneVsRestClassifier(DualSVC(kernel=GaussianKernel(), C=1, max_iter=15)).fit(X, y3)
        assert len(ovr.estimators_) == 3 and ovr.score(X, y3) > 0.6
        assert all(type(e.obj.device_hessian()).__name__ == 'GroupHessian' for e in ovr.estimators_)


# ---------------------------------------------------------------------------------------------- symmetric pass, sharded
# (K2s on row blocks: every pair of off-diagonal blocks is read once, by one of its two owners; the column sums travel
# to the other owner as tagged entries -- optiml_b200/csrc/k2_symv.cuh)
import symv_checks as SY   # noqa: E402

_SYMV_SMALL = ('SVMB200_SYMV_TR=4', 'SVMB200_SYMV_NRB=2', 'SVMB200_SYMV_NCH=1', 'SVMB200_SYMV_LB=2', 'SVMB200_SYMV_STAGES=2')
_SYMV_B32 = ('SVMB200_SYMV_TR=16', 'SVMB200_SYMV_NRB=2', 'SVMB200_SYMV_NCH=1', 'SVMB200_SYMV_LB=8', 'SVMB200_SYMV_STAGES=2')


@pytest.mark.parametrize('kind,layout,n,nranks,defines', [
    ('pg', 'plain', 200, 2, _SYMV_SMALL),     # the halved pair only
    ('pg', 'plain', 380, 3, _SYMV_SMALL),     # odd rank count: whole blocks only
    ('pg', 'plain', 500, 4, _SYMV_SMALL),     # whole blocks and halved pairs, ragged last block
    ('pg', 'svr', 460, 4, _SYMV_B32),         # 128-row blocks of 32-row bands, SVR block Hessian
    ('fw', 'plain', 330, 3, _SYMV_B32), ('adam', 'plain', 200, 2, _SYMV_SMALL),
    ('pg', 'plain', 600, 5, _SYMV_B32),
])
def test_sharded_symmetric_pass_follows_the_one_rank_solve(kind, layout, n, nranks, defines):
    """every rank ends with the SAME bits (the finished product is published once, by its owner), and the iterate stays
    within rounding of the default full pass on one rank"""
    rng = np.random.default_rng(n + nranks)
    M = S.psd(rng, n)
    nv = 2 * n if layout == 'svr' else n
    q, ub = rng.standard_normal(nv), np.full(nv, 1.5)
    max_iter = 10
    flags = []

    def body(ctx):
        H = shard_hessian(ctx, M, layout)
        state = solve(kind, H, q, ub, max_iter)
        return state

    with emulated_device(defines=defines, order=2, seed=n) as lib:
        one = run_ranks(1, 'nccl', body)[0]
        with SY.symmetric_pass():
            gathers = lib.emu_allgather_calls()
            many = run_ranks(nranks, 'p2p', body)
            assert lib.emu_allgather_calls() - gathers == nranks   # the creation barrier only: no collective per iteration
            again = run_ranks(nranks, 'p2p', body)
            sym_one = run_ranks(1, 'nccl', body)[0]
    for rank_state in many[1:]:
        for a, b in zip(many[0], rank_state):
            assert np.array_equal(a, b)
    for a, b in zip(many[0], again[0]):
        assert np.array_equal(a, b)          # reproducible run to run
    for a, b, c in zip(one, many[0], sym_one):
        scale = max(1.0, np.abs(a).max())
        assert np.abs(np.asarray(a, dtype=float) - np.asarray(b, dtype=float)).max() <= 1e-10 * scale
        assert np.abs(np.asarray(a, dtype=float) - np.asarray(c, dtype=float)).max() <= 1e-10 * scale


def test_sharded_symmetric_pass_really_runs_and_reads_half(monkeypatch):
    """the solver reports the mode, and every rank's plan covers each unordered pair of rows exactly once (checked on the
    host plan through the product itself: a matrix with NaN in every block a rank must not read)"""
    from optiml_b200.opti import Quadratic
    from optiml_b200.opti.constrained import ProjectedGradient
    rng = np.random.default_rng(5)
    n, nranks = 500, 4
    M = S.psd(rng, n)
    q, ub = rng.standard_normal(n), np.full(n, 1.5)
    rpr = 128
    seen = []

    def body(ctx):
        H = shard_hessian(ctx, M)
        # poison what this rank must not touch: the lower triangle of its diagonal block below the 8-row bands, and every
        # block at cyclic distance > P/2 behind ... simply: blocks (p, q) with (q - p) mod P in {3} are read by q, not p
        block = np.zeros((max(H.nrows, 1), H.ld))
        block[:H.nrows, :n] = M[H.row0:H.row0 + H.nrows]
        p = H.row0 // rpr
        for qq in range(nranks):
            if (qq - p) % nranks == 3:
                block[:H.nrows, qq * rpr:min(n, (qq + 1) * rpr)] = np.nan
        rows, cols = np.indices((H.nrows, n))
        own = (cols >= H.row0) & (cols < H.row0 + H.nrows) & (cols - H.row0 < (rows // 8) * 8)
        block[:H.nrows, :n][own] = np.nan
        ctx.h2d(H.matrix.dptr, block)
        s = ProjectedGradient(quad=Quadratic(H, q), ub=ub, max_iter=6)
        s.minimize()
        seen.append(s.symmetric_pass)
        return [np.asarray(s.x), np.asarray(s.g_x)]

    with emulated_device(defines=_SYMV_SMALL):
        with SY.symmetric_pass():
            many = run_ranks(nranks, 'p2p', body)
    assert seen == [True] * nranks
    from oracle import svm_oracle as O
    want = O.projected_gradient(M, q, ub, max_iter=6)
    for st in many:
        assert np.all(np.isfinite(st[0])) and np.abs(st[0] - want.x).max() <= 1e-10


def test_single_process_group_with_the_symmetric_pass(device_group):
    """one host thread drives all ranks: the tile passes and sends of every rank are issued before any rank's combine"""
    runtime = device_group
    from optiml_b200.ml.svm import DualSVC, DualSVR
    from optiml_b200.ml.svm.kernels import GaussianKernel, PolyKernel
    from optiml_b200.opti.constrained import FrankWolfe
    rng = np.random.default_rng(1)
    X = rng.standard_normal((420, 5))
    y = (X[:, 0] + 0.2 * rng.standard_normal(420) > 0).astype(int)
    t = X @ rng.standard_normal(5)
    makers = [lambda: DualSVC(kernel=GaussianKernel(), C=1, max_iter=30),
              lambda: DualSVR(kernel=PolyKernel(degree=2), C=1, max_iter=20),
              lambda: DualSVC(kernel=GaussianKernel(), C=1, max_iter=20, optimizer=FrankWolfe)]

    def fit_all(expect_sym):
        out = []
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            for mk in makers:
                m = mk()
                m.fit(X, t if isinstance(m, DualSVR) else y)
                assert m.optimizer.symmetric_pass is expect_sym
                out.append((type(m.obj.device_hessian()).__name__, m.alphas_.copy(), m.intercept_, m.support_.copy()))
                m.obj.release()
        return out

    with SY.symmetric_pass():
        grouped = fit_all(True)
    assert all(g[0] == 'GroupHessian' for g in grouped)
    runtime.use_devices(None)
    solo = fit_all(False)
    for g, s in zip(grouped, solo):
        assert np.abs(g[1] - s[1]).max() <= 1e-10 and abs(g[2] - s[2]) <= 1e-9 and np.array_equal(g[3], s[3])


def test_symmetric_and_default_solves_share_the_arena_back_to_back():
    """eight ranks, problems of alternating sizes and passes back to back: a size that leaves the last rank without rows
    takes the all-gather (and the full pass) even when the symmetric pass is requested; everything else alternates
    between the two fused exchanges, whose regions live in the two halves of the arena.  No stall, every repeat
    bit-identical on every rank."""
    from optiml_b200 import runtime
    from optiml_b200.opti import Quadratic
    from optiml_b200.opti.constrained import ProjectedGradient
    rng = np.random.default_rng(8)
    nranks = 8
    sizes = (450, 667, 1000, 1400)    # 667: 128-row blocks, ranks 6 and 7 own nothing
    problems = [(S.psd(rng, n), rng.standard_normal(n), np.full(n, 1.5)) for n in sizes]
    fused = [(nranks - 1) * (-(-(-(-n // nranks)) // 64) * 64) < n for n in sizes]
    assert fused == [True, False, True, True]

    def body(ctx):
        quads = [Quadratic(shard_hessian(ctx, M), q) for M, q, _ in problems]
        out, flags = [], []
        for i in range(40):
            j, sym, it = i % len(problems), (i // 2) % 2 == 0, (2, 5, 1, 3, 4)[i % 5]
            N.call('svmb200_ctx_set_symmetric', ctx.handle, int(sym))
            s = ProjectedGradient(quad=quads[j], ub=problems[j][2], max_iter=it).minimize()
            flags.append(s.symmetric_pass == (sym and fused[j]))
            out.append(((j, sym, it), s.x.copy()))
        assert all(flags)
        return out

    with emulated_device(defines=_SYMV_SMALL):
        many = run_ranks(nranks, 'p2p', body)
    first = {}
    for key, x in many[0]:
        assert np.array_equal(first.setdefault(key, x), x)
    for state in many[1:]:
        for (ka, xa), (kb, xb) in zip(many[0], state):
            assert ka == kb and np.array_equal(xa, xb)


_SYMV_B32_FEW_SLOTS = _SYMV_B32 + ('SVMB200_SYMV_PLAN_SLOTS=4', 'SVMB200_SYMV_PLAN_OVERHEAD=64.0')


@pytest.mark.parametrize('n,nranks', [(900, 4), (700, 3)])
def test_sharded_symmetric_pass_with_short_bands(n, nranks):
    """the planner cuts the end of every shard into short bands (few resident CTAs assumed): the halved block pair, the sends
    and the owners' combines must agree on the band tables"""
    import ctypes as C
    rng = np.random.default_rng(n)
    M = S.psd(rng, n)
    q, ub = rng.standard_normal(n), np.full(n, 1.5)

    def body(ctx):
        return solve('pg', shard_hessian(ctx, M), q, ub, 8)

    with emulated_device(defines=_SYMV_B32_FEW_SLOTS):
        short = C.c_int64()
        N.call('svmb200_symv_plan_info', n, N.padded_ld(n), 1, nranks, 148, None, C.byref(short), None, None)
        assert short.value >= 3
        one = run_ranks(1, 'nccl', body)[0]
        with SY.symmetric_pass():
            many = run_ranks(nranks, 'p2p', body)
    for state in many[1:]:
        for a, b in zip(many[0], state):
            assert np.array_equal(a, b)
    for a, b in zip(one, many[0]):
        assert np.abs(np.asarray(a, dtype=float) - np.asarray(b, dtype=float)).max() <= 1e-10 * max(1.0, np.abs(a).max())


@pytest.mark.parametrize('n,nranks,defines', [(1100, 1, _SYMV_B32_FEW_SLOTS), (900, 4, _SYMV_B32_FEW_SLOTS), (700, 3, _SYMV_B32_FEW_SLOTS),
                                              (1000, 2, _SYMV_B32_FEW_SLOTS), (1400, 8, _SYMV_SMALL), (1000, 8, _SYMV_B32_FEW_SLOTS)])
def test_graded_plans_cover_every_pair_of_rows_exactly_once(n, nranks, defines):
    """the coverage property of tests/test_host_logic.py on plans that DO end on short bands and cut panels (few resident CTAs
    assumed, small tiles)"""
    import ctypes as C
    from test_host_logic import _symv_plan_cover
    with emulated_device(defines=defines):
        if 'SVMB200_SYMV_PLAN_SLOTS=4' in defines:
            shorts = []
            for r in range(nranks):
                short = C.c_int64()
                N.call('svmb200_symv_plan_info', n, N.padded_ld(n), r, nranks, 148, None, C.byref(short), None, None)
                shorts.append(short.value)
            assert nranks == 1 or max(shorts) >= 1, shorts   # (a grid of many waves is not graded: nothing to gain)
        count, _ = _symv_plan_cover(N.call, n, nranks)
    assert count.min() == 1 and count.max() == 1
