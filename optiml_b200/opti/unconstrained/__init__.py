"""Unconstrained optimisers of the reference that the SVM dual path can drive (optiml/opti/unconstrained):
only the full-batch stochastic family on the augmented-Lagrangian dual is implemented."""
