"""(Stabilised) Frank-Wolfe with exact line search on the box, run on the GPU(s).

Host mirror of optiml/opti/constrained/frank_wolfe.py:30-165 -- the first widening step after the projected
gradient: identical solver protocol, and on the device the identical streaming pass over Q
per iteration; only the O(n) vector phase differs (``fw_vector_kernel`` in csrc/k3_vector.cuh)."""
from . import BoxConstrainedQuadraticOptimizer
from ._device_loop import DeviceLoopMixin


class FrankWolfe(DeviceLoopMixin, BoxConstrainedQuadraticOptimizer):
    _create_symbol = 'svmb200_fw_create'
    _verbose_header = 'iter\t cost\t\t lb\t\t gap'

    def __init__(self, quad, ub, lb=None, x=None, t=0., eps=1e-6, tol=1e-8, max_iter=1000, callback=None,
                 callback_args=(), verbose=False):
        super(FrankWolfe, self).__init__(quad=quad, ub=ub, lb=lb, x=x, eps=eps, tol=tol, max_iter=max_iter,
                                         callback=callback, callback_args=callback_args, verbose=verbose)
        if not 0 <= t < 1:
            raise ValueError('t has to lie in [0, 1)')
        self.t = t

    def _extra_create_args(self):
        return (float(self.t),)

    def _print_iteration(self, scalars):
        # frank_wolfe.py:111-112: iteration, cost, best lower bound, relative gap
        print('\n{:4d}\t{: 1.4e}\t{: 1.4e}\t{: 1.4e}'.format(self.iter, self.f_x, scalars[2], scalars[1]), end='')
