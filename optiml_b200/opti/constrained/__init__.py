__all__ = ['BoxConstrainedQuadraticOptimizer', 'AugmentedLagrangianQuadratic', 'ProjectedGradient', 'FrankWolfe']

from ._base import BoxConstrainedQuadraticOptimizer, AugmentedLagrangianQuadratic
from .projected_gradient import ProjectedGradient
from .frank_wolfe import FrankWolfe
