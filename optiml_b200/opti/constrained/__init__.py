__all__ = ['BoxConstrainedQuadraticOptimizer', 'ProjectedGradient']

from ._base import BoxConstrainedQuadraticOptimizer
from .projected_gradient import ProjectedGradient
