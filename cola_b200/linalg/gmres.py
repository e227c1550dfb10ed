"""GMRES on the device Arnoldi (cola/linalg/inverse/gmres.py:10-124).

`GMRES(tol, max_iters, pbar, x0, P)(A, b) -> (soln, info)`, `gmres(A, rhs, x0, max_iters, tol, P, ...)`.
The Krylov work is the MGS Arnoldi of arnoldi.py (`mgs_link` kernels); what is added here is
  * the residual `b - A x0` (one matmat + one axpby sweep; skipped for the default zero guess),
  * the reference's small solve, restated verbatim: it drops the last row of the (m+1, m) Hessenberg and solves the
    normal equations of the square part, `(H^H H + D) y = H^H[:, 0] * beta` with D = identity on all-zero rows
    (gmres.py:110-118) -- a batched m x m dense solve, done with torch.linalg on the device (library code for a
    tiny dense problem, as the Lanczos path does for `eigh(T)`),
  * `x0 + Q y`: the tall-skinny combination is the reorth "update" kernel with sign +1.
As in the reference, P is accepted and not used (gmres.py:92-124 never touches it); the Householder / triangular
variants are outside the hot path.
"""
from dataclasses import dataclass
from typing import Any

import torch

from .. import backend as be
from ..ops import LinearOperator
from .algorithm_base import Algorithm
from .arnoldi import arnoldi_fact


@dataclass
class GMRES(Algorithm):
    """cola/linalg/inverse/gmres.py:10-37"""
    tol: float = 1e-6
    max_iters: int = 1000
    pbar: bool = False
    x0: Any = None
    P: Any = None

    def __call__(self, A, b):
        return gmres(A, b, **self.__dict__)


def gmres(A: LinearOperator, rhs, x0=None, max_iters=100, tol=1e-7, P=None, use_householder=False,
          use_triangular=False, pbar=False):
    """cola/linalg/inverse/gmres.py:41-71"""
    if use_householder or use_triangular:
        raise NotImplementedError("the Householder / triangular-QR GMRES variants are outside the Krylov hot path")
    is_vector = len(rhs.shape) == 1
    if is_vector:
        rhs = rhs[..., None]
        x0 = x0[..., None] if x0 is not None else None
    soln, infodict = gmres_fwd(A, rhs, x0, max_iters, tol, P, pbar)
    if is_vector:
        soln = soln[:, 0]
    return soln, infodict


def gmres_fwd(A, rhs, x0, max_iters, tol, P=None, pbar=False):
    """cola/linalg/inverse/gmres.py:92-124.  rhs (n, b) -> (soln (n, b), info)."""
    be.require_cuda(rhs, "right-hand sides")
    A.plan()                                               # validate / compile once
    dt = A.dtype
    rhs = rhs.to(dt).contiguous()
    n, b = rhs.shape
    if x0 is None:
        res = rhs                                                  # zero guess: r0 = b exactly
    else:
        x0 = x0.to(dt).contiguous()
        res = torch.empty_like(rhs)
        A.matmat_into(x0, res)
        be.axpby(rhs, res, 1.0, -1.0)                              # res = rhs - A x0
    m = int(max_iters)
    Q, H, idx, info = arnoldi_fact(A, res, m, tol, pbar)           # Q (m+1, n, b), H (b, m+1, m)
    H = H[:, :-1, :]                                               # (b, m, m): the reference drops the last row
    nrm = torch.zeros(b, dtype=torch.float64, device=rhs.device)
    be.col_dots(res, res, nrm)
    beta = torch.sqrt(nrm).to(dt)
    HT = torch.conj(H.permute(0, 2, 1))
    largest = torch.max(torch.abs(H), -1)[0]
    overall = torch.max(largest.reshape(largest.shape[0], -1), -1)[0]
    thresh = 10 * tol * overall[:, None]
    padding = torch.where(largest < thresh, torch.ones_like(largest), torch.zeros_like(largest))
    y = torch.linalg.solve(HT @ H + torch.diag_embed(padding), HT[..., 0, None]).squeeze(-1) * beta[:, None]
    y = torch.where(largest < thresh, torch.zeros_like(y), y)      # (b, m)
    # soln = x0 + sum_j y[:, j] Q[j]
    soln = torch.zeros_like(rhs) if x0 is None else x0.clone()
    if m > 0:
        C = y.T.to(torch.float64).contiguous()                     # (m, b)
        be.reorth_update(Q, 0, m, soln, C, sign=1.0)
    return soln, info
