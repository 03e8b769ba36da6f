__all__ = ['SVM', 'SVC', 'SVR', 'DualSVC', 'DualSVR']

from ._base import SVM, SVC, SVR, DualSVC, DualSVR
