"""B200-native drop-in for deepdow's allocation hot path.

Same class names, constructor and ``forward`` signatures as ``deepdow.layers`` for
``NumericalMarkowitz``, ``SparsemaxAllocator``, ``SoftmaxAllocator`` and ``CovarianceMatrix``
(reference deepdow/layers/allocate.py:198-273, 414-595 and deepdow/layers/misc.py:33-201), backed by
hand-written sm_100a CUDA kernels behind the C ABI of ``include/ddw.h``.
"""
from .layers import (CovarianceMatrix, NumericalMarkowitz, NumericalRiskBudgeting, Resample, SoftmaxAllocator, SolverError,
                     SparsemaxAllocator)

__all__ = ["CovarianceMatrix", "NumericalMarkowitz", "NumericalRiskBudgeting", "Resample", "SoftmaxAllocator", "SparsemaxAllocator",
           "SolverError"]
__version__ = "0.1.0"
