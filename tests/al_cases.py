"""Case table shared by tests/golden/make_golden_al.py (reference side), the CPU tests (oracle, host emulation) and the
GPU tests of the augmented-Lagrangian dual path.  name -> (rule, learning rate, estimator keywords, max_iter)."""

CASES = {
    'adagrad': ('adagrad', 1., {}, 1000),
    'sgd': ('sgd', 0.001, {}, 300),
    'sgd_polyak': ('sgd', 0.001, dict(momentum_type='polyak', momentum=0.5), 300),
    'sgd_nesterov': ('sgd', 0.001, dict(momentum_type='nesterov', momentum=0.5), 300),
    'rmsprop': ('rmsprop', 0.01, {}, 300),
    'rmsprop_nesterov': ('rmsprop', 0.001, dict(momentum_type='nesterov', momentum=0.5), 300),
    'adadelta': ('adadelta', 1., {}, 300),
    'adam': ('adam', 0.01, {}, 300),
    'adam_nesterov': ('adam', 0.001, dict(momentum_type='nesterov', momentum=0.5), 300),
    'amsgrad_polyak': ('amsgrad', 0.001, dict(momentum_type='polyak', momentum=0.5), 300),
    'adamax': ('adamax', 0.01, {}, 300),
}
OPTIMIZER_CLASS = {'adagrad': 'AdaGrad', 'sgd': 'StochasticGradientDescent', 'rmsprop': 'RMSProp', 'adadelta': 'AdaDelta',
                   'adam': 'Adam', 'amsgrad': 'AMSGrad', 'adamax': 'AdaMax'}

# SVC on the iris one-vs-rest split: (case, reg_intercept, class) ; AdaGrad (the reference's own test) on every class
SVC_RUNS = [(name, ri, c) for name in CASES for ri in (True, False) for c in ((0, 1, 2) if name == 'adagrad' else (1,))]
# loose tolerances that end the run through the optimality test of the multiplier update
TOL_RUNS = [('adagrad', 0.1), ('sgd_nesterov', 0.1), ('rmsprop_nesterov', 0.2)]
SVR_RUNS = [(k, name, ri) for k in ('linear', 'gaussian') for name in ('adagrad', 'adam_nesterov', 'adadelta')
            for ri in (True, False)]


def svc_key(name, ri, c):
    return f'svc_{name}_ri{int(ri)}_c{c}'


def svr_key(kernel, name, ri):
    return f'svr_{kernel}_{name}_ri{int(ri)}'
