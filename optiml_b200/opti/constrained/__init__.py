__all__ = ['BoxConstrainedQuadraticOptimizer', 'ProjectedGradient', 'FrankWolfe']

from ._base import BoxConstrainedQuadraticOptimizer
from .projected_gradient import ProjectedGradient
from .frank_wolfe import FrankWolfe
