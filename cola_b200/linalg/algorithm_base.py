"""Algorithm objects (cola/linalg/algorithm_base.py:11-34)."""
from types import SimpleNamespace

from ..ops import LinearOperator


class Algorithm:
    pass


class Auto(SimpleNamespace, Algorithm):
    pass


class IterativeOperatorWInfo(LinearOperator):
    """Lazy A^{-1}: `_matmat(X)` runs the solver and stores `info` (algorithm_base.py:16-29)."""
    def __init__(self, A, alg):
        super().__init__(A.dtype, A.shape)
        self.A = A
        self.alg = alg
        self.info = {}
        self.device = A.device

    def _matmat(self, X):
        Y, self.info = self.alg(self.A, X)
        return Y

    def __str__(self):
        return f"{self.alg}({str(self.A)})"
