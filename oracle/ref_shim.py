"""Import the REAL reference (dmeoli/optiml)  --  TEST / BENCHMARK INFRASTRUCTURE ONLY.

Where it is looked for, in order: ``$OPTIML_REFERENCE_ROOT``; ``baseline/_ref`` (the unmodified reference installed by
``pip install --no-deps --target baseline/_ref``, git-ignored, travels to the GPU box: what ``bench.py --impl reference``
and its ``cpu_baseline`` leg time); ``/root/reference`` (the build container only).

The reference's dual-BCQP path is pure NumPy, but ``import optiml.ml.svm`` also imports four
third-party packages that are absent from this image and that the path never calls
(``autograd``, ``qpsolvers``, ``wurlitzer``, ``cvxpy``).  ``load_reference()`` registers inert
stand-ins for those names and returns the reference's modules.  It is used by
``tests/golden/make_golden.py`` (to generate the golden vectors) and by the optional live
cross-checks in ``tests/`` that are skipped when ``/root/reference`` does not exist (GPU box).
Nothing here ships in the product package.
"""
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root():
    for cand in (os.environ.get('OPTIML_REFERENCE_ROOT'), os.path.join(_REPO, 'baseline', '_ref'), '/root/reference'):
        if cand and os.path.isdir(os.path.join(cand, 'optiml')):
            return cand
    return os.environ.get('OPTIML_REFERENCE_ROOT', '/root/reference')


REFERENCE_ROOT = _find_root()


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'optiml'))


def _install_stubs():
    import numpy
    if 'autograd' not in sys.modules:
        ag = types.ModuleType('autograd')
        ag.jacobian = ag.hessian = lambda f: (lambda *a, **k: None)
        sys.modules['autograd'] = ag
        sys.modules['autograd.numpy'] = numpy
    if 'qpsolvers' not in sys.modules:
        qp = types.ModuleType('qpsolvers')

        def solve_qp(*a, **k):
            raise RuntimeError('qpsolvers is not installed (stub)')

        qp.solve_qp = solve_qp
        sys.modules['qpsolvers'] = qp
    if 'wurlitzer' not in sys.modules:
        w = types.ModuleType('wurlitzer')
        w.pipes = w.STDOUT = None
        sys.modules['wurlitzer'] = w
    if 'cvxpy' not in sys.modules:
        cv = types.ModuleType('cvxpy')
        for name in ('Variable', 'Problem', 'Minimize', 'sum_squares'):
            setattr(cv, name, None)
        sys.modules['cvxpy'] = cv


def load_reference():
    """Return a namespace with the reference classes of the hot path."""
    if not reference_available():
        raise RuntimeError(f'reference not found under {REFERENCE_ROOT}')
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from optiml.ml.svm import SVC, SVR
    from optiml.ml.svm.kernels import (GaussianKernel, PolyKernel, LinearKernel, LaplacianKernel, SigmoidKernel,
                                       gaussian, poly, linear)
    from optiml.ml.svm.losses import hinge, epsilon_insensitive
    from optiml.opti import Quadratic
    from optiml.opti.constrained import ProjectedGradient, FrankWolfe
    ns = types.SimpleNamespace(SVC=SVC, SVR=SVR, GaussianKernel=GaussianKernel, PolyKernel=PolyKernel,
                               LinearKernel=LinearKernel, LaplacianKernel=LaplacianKernel,
                               SigmoidKernel=SigmoidKernel, gaussian=gaussian, poly=poly, linear=linear,
                               hinge=hinge, epsilon_insensitive=epsilon_insensitive, Quadratic=Quadratic,
                               ProjectedGradient=ProjectedGradient, FrankWolfe=FrankWolfe)
    ns.generate_box_constrained_quadratic = _load_bcqp_generator()
    return ns


def _load_bcqp_generator():
    """optiml/opti/utils.py imports matplotlib and casadi at module level; execute only the
    generator function's source (pure NumPy, lines 54-161) in a fresh namespace."""
    import ast
    import numpy
    path = os.path.join(REFERENCE_ROOT, 'optiml', 'opti', 'utils.py')
    with open(path) as fh:
        tree = ast.parse(fh.read())
    fn = [node for node in tree.body if isinstance(node, ast.FunctionDef)
          and node.name == 'generate_box_constrained_quadratic']
    mod = ast.Module(body=fn, type_ignores=[])
    glb = {'np': numpy}
    exec(compile(mod, path, 'exec'), glb)
    return glb['generate_box_constrained_quadratic']
