"""K2/K3 parity: the device projected-gradient loop vs the reference's solver (golden vectors) and vs
the CPU oracle.  Exact line search on the box is a chaotic map on problems where free (unclipped)
steps dominate: a 1-ulp perturbation of Q moves the REFERENCE's own result by up to 1e-2 after a few
hundred iterations (tests/test_oracle_sensitivity.py).  Parity is therefore asserted (a) to 1e-8 on
trajectories that are stable, (b) on the iteration map itself -- a few iterations from reference
states -- to 1e-11 everywhere, and (c) against the oracle's own 1-ulp sensitivity envelope otherwise."""
import ctypes as C

import numpy as np
import pytest

from oracle import svm_oracle as O

pytestmark = pytest.mark.gpu


def solve(Q, q, ub, lb=None, x=None, **kw):
    from optiml_b200.opti import Quadratic
    from optiml_b200.opti.constrained import ProjectedGradient
    return ProjectedGradient(quad=Quadratic(Q, q), ub=ub, lb=lb, x=x, **kw).minimize()


def envelope(Q, q, ub, lb=None, max_iter=1000, seed=0):
    """max |x - x'| between the oracle on Q and on Q perturbed by +-1 ulp (symmetric)."""
    rng = np.random.default_rng(seed)
    E = rng.integers(-1, 2, size=Q.shape)
    E = np.triu(E) + np.triu(E, 1).T
    r0 = O.projected_gradient(Q, q, ub, lb=lb, max_iter=max_iter)
    r1 = O.projected_gradient(Q * (1 + E * 2.2e-16), q, ub, lb=lb, max_iter=max_iter)
    return r0, float(np.abs(r0.x - r1.x).max())


@pytest.mark.parametrize('p', ['p2', 'p5', 'p64'])
def test_bcqp_golden_stable(golden, p):
    """reference tests opti/constrained/tests/test_projected_gradient.py:9-12, test_lower_bound.py:9-25"""
    g = golden('bcqp')
    lb = g[p + '_lb'] if p + '_lb' in g else None
    opt = solve(g[p + '_Q'], g[p + '_q'], g[p + '_ub'], lb=lb)
    assert opt.iter == int(g[p + '_iter']) and opt.status == str(g[p + '_status'])
    assert np.abs(opt.x - g[p + '_x']).max() <= 1e-8
    assert np.all(opt.x >= opt.lb - 1e-6) and np.all(opt.x <= opt.ub + 1e-6)
    f_ref = g[p + '_f_hist'][-1]
    assert abs(opt.f_x - f_ref) <= 1e-9 * max(1., abs(f_ref))
    # same acceptance test as the reference: allclose to the box-constrained optimum
    assert np.allclose(opt.x, opt.x_star(), atol=1e-5)


def test_bcqp_p200_within_sensitivity_envelope(golden):
    g = golden('bcqp')
    Q, q, ub = g['p200_Q'], g['p200_q'], g['p200_ub']
    opt = solve(Q, q, ub)
    r0, env = envelope(Q, q, ub)
    assert opt.status == 'optimal' and abs(opt.iter - int(g['p200_iter'])) <= 5
    assert np.abs(opt.x - g['p200_x']).max() <= max(1e-8, 20 * env)
    assert abs(opt.f_x - float(g['p200_f_hist'][-1])) <= 1e-9 * abs(float(g['p200_f_hist'][-1]))


@pytest.mark.parametrize('p,steps', [('p5', 3), ('p64', 4), ('p200', 4)])
def test_iteration_map_from_reference_states(golden, p, steps):
    """`steps` iterations started from points of the reference trajectory agree with the oracle to 1e-11:
    chaos cannot amplify in a handful of iterations, so this pins every branch of the iteration map."""
    g = golden('bcqp')
    Q, q, ub = g[p + '_Q'], g[p + '_q'], g[p + '_ub']
    lb = g[p + '_lb'] if p + '_lb' in g else None
    for k0 in (0, 7, 40):
        start = O.projected_gradient(Q, q, ub, lb=lb, max_iter=max(k0, 1)).x if k0 else None
        want = O.projected_gradient(Q, q, ub, lb=lb, x0=start, max_iter=steps)
        got = solve(Q, q, ub, lb=lb, x=None if start is None else start.copy(), max_iter=steps)
        assert got.iter == want.iter and got.status == want.status
        assert np.abs(got.x - want.x).max() <= 1e-11 * max(1., np.abs(want.x).max())
        assert np.abs(got.g_x - want.g_x).max() <= 1e-10 * max(1., np.abs(want.g_x).max())
        assert abs(got.f_x - want.f_x) <= 1e-11 * max(1., abs(want.f_x))


def test_random_psd_problems_vs_oracle():
    rng = np.random.default_rng(5)
    for n in (2, 3, 17, 130, 1000):
        G = rng.standard_normal((n + 3, n))
        Q = G.T @ G / n
        q = rng.standard_normal(n)
        ub = rng.uniform(0.5, 2., n)
        lb = -rng.uniform(0., 1., n)
        want = O.projected_gradient(Q, q, ub, lb=lb, max_iter=25)
        got = solve(Q, q, ub, lb=lb, max_iter=25)
        assert got.iter == want.iter and got.status == want.status
        assert np.abs(got.x - want.x).max() <= 1e-10
        k = min(len(want.f_hist), 26)
        hist = got.f_hist if hasattr(got, 'f_hist') else np.array(got.f_x_history)
        assert np.allclose(hist[:k], want.f_hist[:k], rtol=1e-10, atol=1e-12)


def test_callback_protocol_and_histories(golden, capsys):
    g = golden('bcqp')
    Q, q, ub = g['p5_Q'], g['p5_q'], g['p5_ub']
    seen = []

    def cb(opt, tag):
        seen.append((opt.iter, opt.f_x, opt.x.copy(), tag))

    opt = solve(Q, q, ub, lb=g['p5_lb'], callback=cb, callback_args=('t',))
    assert [s[0] for s in seen] == list(range(opt.iter + 1))  # iterations + 1 callback points
    assert np.allclose([s[1] for s in seen], g['p5_f_hist'], rtol=1e-12, atol=1e-12)
    assert all(s[3] == 't' for s in seen)

    # StopIteration from the callback stops the loop with status 'unknown' (projected_gradient.py:95-98)
    def stopper(opt):
        if opt.iter == 3:
            raise StopIteration

    opt = solve(Q, q, ub, lb=g['p5_lb'], callback=stopper)
    assert opt.iter == 3 and opt.status == 'unknown'

    # ndim <= 3 histories (opti/_base.py:78-82, 121-124) and verbose output format
    opt = solve(g['p2_Q'], g['p2_q'], g['p2_ub'], verbose=True)
    out = capsys.readouterr().out
    assert out.startswith('iter\t cost\t\t gnorm') and out.endswith('\n\n')
    assert len(opt.f_x_history) == opt.iter + 1 == 3
    assert np.allclose(opt.f_x_history, g['p2_f_hist'], atol=1e-12)
    assert '\n   0\t' in out


def test_solver_argument_errors():
    from optiml_b200.opti import Quadratic
    from optiml_b200.opti.constrained import ProjectedGradient
    with pytest.raises(ValueError):
        Quadratic([[1.]], [1.])  # n <= 1 (opti/_base.py:249)
    with pytest.raises(ValueError):
        Quadratic(np.eye(3), np.ones(2))
    with pytest.raises(TypeError):
        ProjectedGradient(quad=np.eye(2), ub=np.ones(2))
    with pytest.raises(ValueError):
        ProjectedGradient(quad=Quadratic(np.eye(2), np.ones(2)), ub=np.ones(2), max_iter=0)


def test_quadratic_values_on_device():
    from optiml_b200.opti import Quadratic
    rng = np.random.default_rng(1)
    A = rng.standard_normal((37, 37))
    Q = A @ A.T
    q = rng.standard_normal(37)
    x = rng.standard_normal(37)
    quad = Quadratic(Q, q)
    assert np.allclose(quad.function(x), 0.5 * x @ Q @ x + q @ x, rtol=1e-13)
    assert np.allclose(quad.jacobian(x), Q @ x + q, rtol=1e-12, atol=1e-12)
    assert np.array_equal(quad.hessian(x), Q)


def test_matvec_matches_numpy_and_is_row_order_independent():
    """K2 alone: row results do not depend on how many rows a launch covers (the property that makes
    alpha bit-identical for any number of GPUs)."""
    from optiml_b200 import _native as N
    from optiml_b200.runtime import default_context
    ctx = default_context()
    rng = np.random.default_rng(3)
    n = 1234
    ld = N.padded_ld(n)
    Q = np.zeros((n, ld))
    Q[:, :n] = rng.standard_normal((n, n))
    u = np.zeros(ld)
    u[:n] = rng.standard_normal(n)
    dQ, du, dw = ctx.malloc(Q.nbytes), ctx.malloc(u.nbytes), ctx.malloc(8 * n)
    ctx.h2d(dQ, Q)
    ctx.h2d(du, u)
    full = np.empty(n)
    N.call('svmb200_matvec', ctx.handle, C.c_void_p(dQ), n, ld, C.c_void_p(du), C.c_void_p(dw))
    ctx.d2h(full, dw)
    assert np.allclose(full, Q @ u, rtol=1e-12, atol=1e-12)
    parts = np.empty(n)
    for r0, r1 in ((0, 309), (309, 618), (618, 1234)):
        N.call('svmb200_matvec', ctx.handle, C.c_void_p(dQ + r0 * ld * 8), r1 - r0, ld, C.c_void_p(du),
               C.c_void_p(dw + r0 * 8))
    ctx.d2h(parts, dw)
    assert np.array_equal(parts, full)
    for p in (dQ, du, dw):
        ctx.free(p)


def test_persistent_loop_is_bit_identical_to_the_two_kernel_loop(monkeypatch):
    """csrc/k_persistent.cuh on hardware: problems whose matrix fits the shared memory of the 148 SMs (n <= 2016: 14 rows of
    2016 doubles per CTA; plain layout) run as ONE cooperative launch; SVMB200_PERSISTENT_GRID=0 forces the K2 + K3 launch pairs -- same bits, on a run
    that stops at the iteration limit, one that reaches optimality and the ragged sizes around the limits"""
    from optiml_b200.runtime import default_context
    ctx = default_context()
    rng = np.random.default_rng(8)
    for n, iters in ((2000, 300), (2016, 40), (1531, 60), (257, 400), (9, 50)):
        G = rng.standard_normal((n + 5, n))
        Q, q, ub = G.T @ G / n, rng.standard_normal(n), rng.uniform(0.5, 2., n)
        runs = []
        for grid in ('0', None):
            if grid is None:
                monkeypatch.delenv('SVMB200_PERSISTENT_GRID', raising=False)
            else:
                monkeypatch.setenv('SVMB200_PERSISTENT_GRID', grid)
            before = ctx.launch_count()
            o = solve(Q, q, ub, max_iter=iters)
            runs.append((o.x, o.g_x, o.f_hist, o.ng_hist, o.iter, o.status, ctx.launch_count() - before))
        two_kernel, persistent = runs
        assert persistent[6] <= 5 < two_kernel[6]
        assert persistent[4:6] == two_kernel[4:6]
        for a, b in zip(two_kernel[:4], persistent[:4]):
            assert np.array_equal(a, b)
    # just beyond the scope (n = 2032: 14 rows no longer fit 221 KB): the two-kernel loop, silently
    n = 2032
    G = rng.standard_normal((n + 5, n))
    before = ctx.launch_count()
    solve(G.T @ G / n, rng.standard_normal(n), np.ones(n), max_iter=20)
    assert ctx.launch_count() - before > 20
