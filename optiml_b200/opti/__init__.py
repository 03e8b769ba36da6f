__all__ = ['Optimizer', 'OptimizationFunction', 'Quadratic']

from ._base import Optimizer, OptimizationFunction, Quadratic
