"""Algorithm objects of the dispatch surface (reference: cola/linalg/algorithm_base.py:11-34).

Three names are part of the interface the Krylov path is called through:
  Algorithm                marker base of CG / GMRES / Lanczos / Arnoldi / Hutch / PowerIteration / ...
  Auto(**options)          "pick for me"; its options are forwarded to the algorithm the dispatch rules choose
                           (`CG(**alg.__dict__)`, inv.py:72-92), so it is nothing but an attribute bag
  IterativeOperatorWInfo   the lazy inverse returned by `inv(A, CG())` / `inv(A, GMRES())`: applying it runs the
                           solver on the given right-hand sides and keeps the solver's `info` dict
"""
from ..ops import LinearOperator


class Algorithm:
    """Marker base class; concrete algorithms are dataclasses whose fields are the solver options."""


class Auto(Algorithm):
    def __init__(self, **options):
        vars(self).update(options)

    def __repr__(self):
        return "Auto(" + ", ".join(f"{k}={v!r}" for k, v in vars(self).items()) + ")"

    def __eq__(self, other):
        return isinstance(other, Auto) and vars(self) == vars(other)


class IterativeOperatorWInfo(LinearOperator):
    def __init__(self, A, alg):
        LinearOperator.__init__(self, dtype=A.dtype, shape=A.shape)
        self.A, self.alg = A, alg
        self.info = {}              # filled by the most recent application
        self.device = A.device

    def _matmat(self, X):
        solution, info = self.alg(self.A, X)
        self.info = info
        return solution

    def __str__(self):
        return "%s(%s)" % (self.alg, self.A)
