"""Projected gradient with exact line search on the box, run on the GPU(s).

Host mirror of optiml/opti/constrained/projected_gradient.py:27-143: same constructor, same
``minimize() -> self`` contract, same ``x f_x g_x iter status`` attributes, same callback point
(once per iteration, before the stopping tests) and the same ``verbose`` output.  The loop itself is
``svmb200_pg_*`` (csrc/pg.cu): Q stays in HBM, one streaming pass per iteration.
"""
from . import BoxConstrainedQuadraticOptimizer
from ._device_loop import DeviceLoopMixin


class ProjectedGradient(DeviceLoopMixin, BoxConstrainedQuadraticOptimizer):
    _create_symbol = 'svmb200_pg_create'
    _verbose_header = 'iter\t cost\t\t gnorm'

    def __init__(self, quad, ub, lb=None, x=None, eps=1e-6, tol=1e-8, max_iter=1000, callback=None,
                 callback_args=(), verbose=False):
        super(ProjectedGradient, self).__init__(quad=quad, ub=ub, lb=lb, x=x, eps=eps, tol=tol, max_iter=max_iter,
                                                callback=callback, callback_args=callback_args, verbose=verbose)

    def _print_iteration(self, scalars):
        # projected_gradient.py:92-93
        print('\n{:4d}\t{: 1.4e}\t{: 1.4e}'.format(self.iter, self.f_x, scalars[1]), end='')
