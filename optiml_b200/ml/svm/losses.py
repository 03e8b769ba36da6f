"""Loss tags of the dual path.

In the reference (optiml/ml/svm/losses.py) these classes are the primal objectives; the dual
branches of ``SVC.fit`` / ``SVR.fit`` only use them as *tags* (``self.loss == Hinge``,
ml/svm/_base.py:557, 1102) and for the ``_loss_type`` check (:417, :959).  The primal formulations
are out of scope here, so the classes carry the tag and nothing else.
"""


class SVMLoss:
    _loss_type = None

    def __init__(self, *args, **kwargs):
        raise NotImplementedError('optiml_b200 implements the dual (kernel) formulation only; '
                                  'primal SVM losses are tags here')


class Hinge(SVMLoss):
    _loss_type = 'classifier'


class SquaredHinge(SVMLoss):
    _loss_type = 'classifier'


class EpsilonInsensitive(SVMLoss):
    _loss_type = 'regressor'


class SquaredEpsilonInsensitive(SVMLoss):
    _loss_type = 'regressor'


hinge = Hinge
squared_hinge = SquaredHinge
epsilon_insensitive = EpsilonInsensitive
squared_epsilon_insensitive = SquaredEpsilonInsensitive
