"""Power iteration on the device (cola/linalg/eig/power_iteration.py:10-81).

`PowerIteration(tol, max_iter, pbar, key)(A) -> (v, eigmax, info)`.  One step is the operator matmat with the Rayleigh
quotient <v, A v> fused into it, one column-dot sweep for ||A v||^2 and one scaling sweep; the quotient is read back
once per step for the reference's stopping rule `|eig_prev - eig| / eig > tol` (power_iteration.py:66-72).
"""
import time
from dataclasses import dataclass
from typing import Any, Optional

import numpy as np
import torch

from .. import backend as be
from .. import rng
from ..ops import LinearOperator
from .algorithm_base import Algorithm

PRNGKey = Any


@dataclass
class PowerIteration(Algorithm):
    """cola/linalg/eig/power_iteration.py:10-32"""
    tol: float = 1e-06
    max_iter: int = 100
    pbar: bool = False
    key: Optional[PRNGKey] = None

    def __call__(self, A: LinearOperator):
        return power_iteration(A, tol=self.tol, max_iter=self.max_iter, pbar=self.pbar, key=self.key)


def power_iteration(A: LinearOperator, tol=1e-6, max_iter=1000, pbar=False, key=None, momentum=None):
    """cola/linalg/eig/power_iteration.py:35-81 -> (v (n,), eigmax (0-d), info)."""
    if momentum is not None:
        raise NotImplementedError("power iteration with momentum is outside the Krylov hot path")
    key = rng.PRNGKey(42) if key is None else key
    dt = A.dtype
    v = rng.randn(A.shape[-1], dtype=dt, device=A.device, key=key)
    be.require_cuda(v, "the operator")
    A.plan()                                               # validate / compile once; the loop uses matmat_into
    v = v.reshape(-1, 1).contiguous()                      # the start vector is NOT normalised (:56)
    p = torch.empty_like(v)
    acc = torch.zeros((2, 1), dtype=torch.float64, device=v.device)   # <v, A v>, ||A v||^2
    cast = np.float32 if dt == torch.float32 else np.float64
    eig, eigprev = cast(10.), cast(1.)                     # :76-77
    samples, evals, i = [], 0, 0
    t0 = time.time()
    while True:
        err = abs(eigprev - eig) / eig                     # :66-68
        samples.append(float(err))
        evals += 1
        if not (i < max_iter and err > tol):               # :70-72
            break
        acc.zero_()
        A.matmat_into(v, p, dots=acc[0])                   # p = A v ; eig = v . p   (:60-61)
        be.col_dots(p, p, acc[1])
        be.col_scale(p, v, acc[1], take_sqrt=True, mode=2)  # v = p / ||p||          (:65)
        eigprev, eig = eig, cast(acc[0, 0].item())         # host poll for the stop rule
        i += 1
    elapsed = time.time() - t0
    samples.append(samples[-1])
    info = {"iterations": evals, "errors": np.array(samples[2:]), "iteration_time": elapsed / evals}
    return v[:, 0], torch.tensor(eig, dtype=dt, device=v.device), info
