"""The seven update rules of optiml/opti/unconstrained/stochastic/*.py: constructors and validation only -- the
arithmetic of every rule is in csrc/al_math.cuh (al_step)."""
import warnings

import numpy as np

from ._base import StochasticOptimizer, StochasticMomentumOptimizer


def _common(kw):
    return {k: kw[k] for k in ('f', 'x', 'step_size', 'batch_size', 'eps', 'tol', 'epochs', 'callback',
                               'callback_args', 'shuffle', 'random_state', 'verbose')}


class StochasticGradientDescent(StochasticMomentumOptimizer):
    """gradient_descent.py:4-133"""
    _rule = 'sgd'

    def __init__(self, f, x=None, batch_size=None, eps=1e-6, tol=1e-8, epochs=1000, step_size=0.01,
                 momentum_type='none', momentum=0.9, callback=None, callback_args=(), shuffle=True,
                 random_state=None, verbose=False):
        super(StochasticGradientDescent, self).__init__(momentum_type=momentum_type, momentum=momentum,
                                                        **_common(locals()))


class AdaGrad(StochasticOptimizer):
    """adagrad.py:6-125"""
    _rule = 'adagrad'

    def __init__(self, f, x=None, batch_size=None, eps=1e-6, tol=1e-8, epochs=1000, step_size=1., offset=1e-8,
                 callback=None, callback_args=(), shuffle=True, random_state=None, verbose=False):
        super(AdaGrad, self).__init__(**_common(locals()))
        if not offset > 0:
            raise ValueError('offset must be > 0')
        self.offset = offset


class AdaDelta(StochasticOptimizer):
    """adadelta.py"""
    _rule = 'adadelta'

    def __init__(self, f, x=None, batch_size=None, eps=1e-6, tol=1e-8, epochs=1000, step_size=1., decay=0.9,
                 offset=1e-6, callback=None, callback_args=(), shuffle=True, random_state=None, verbose=False):
        super(AdaDelta, self).__init__(**_common(locals()))
        if not 0 <= decay < 1:
            raise ValueError('decay has to lie in [0, 1)')
        self.decay = decay
        if not offset > 0:
            raise ValueError('offset must be > 0')
        self.offset = offset


class RMSProp(StochasticMomentumOptimizer):
    """rmsprop.py"""
    _rule = 'rmsprop'

    def __init__(self, f, x=None, step_size=0.001, momentum_type='none', momentum=0.9, batch_size=None, eps=1e-6,
                 tol=1e-8, epochs=1000, decay=0.9, offset=1e-8, callback=None, callback_args=(), shuffle=True,
                 random_state=None, verbose=False):
        super(RMSProp, self).__init__(momentum_type=momentum_type, momentum=momentum, **_common(locals()))
        if not 0 <= decay < 1:
            raise ValueError('decay has to lie in [0, 1)')
        self.decay = decay
        if not offset > 0:
            raise ValueError('offset must be > 0')
        self.offset = offset


class _MomentEstimator(StochasticMomentumOptimizer):
    """shared constructor of adam.py, amsgrad.py, adamax.py"""
    _default_step = 0.001

    def __init__(self, f, x=None, batch_size=None, eps=1e-6, tol=1e-8, epochs=1000, step_size=None,
                 momentum_type='none', momentum=0.9, beta1=0.9, beta2=0.999, offset=1e-8, callback=None,
                 callback_args=(), shuffle=True, random_state=None, verbose=False):
        if step_size is None:
            step_size = self._default_step
        super(_MomentEstimator, self).__init__(momentum_type=momentum_type, momentum=momentum, **_common(locals()))
        if not 0 <= beta1 < 1:
            raise ValueError('beta1 has to lie in [0, 1)')
        self.beta1 = beta1
        if not 0 <= beta2 < 1:
            raise ValueError('beta2 has to lie in [0, 1)')
        self.beta2 = beta2
        if not self.beta1 < np.sqrt(self.beta2):
            warnings.warn('constraint from convergence analysis for adam not satisfied')
        if not offset > 0:
            raise ValueError('offset must be > 0')
        self.offset = offset


class Adam(_MomentEstimator):
    _rule = 'adam'


class AMSGrad(_MomentEstimator):
    _rule = 'amsgrad'


class AdaMax(_MomentEstimator):
    _rule = 'adamax'
    _default_step = 0.002
