"""f(A) V for operators that are not self-adjoint, on the device Arnoldi (cola/linalg/unary/unary.py:63-91).

`ArnoldiUnary(A, f, **arnoldi_kwargs)._matmat(V)` runs one batched Arnoldi factorisation with the columns of V
as start vectors (the MGS chain of cola_b200/linalg/arnoldi.py), takes the small (m x m) eigen-decomposition
H = P diag(lam) P^-1 per column (library call on the device, as in the reference), and forms
        f(A) v  ~=  ||v|| * Q_m P f(lam) P^-1 e_1.
The only n-sized work after the factorisation, Q_m @ coef, is the reorthogonalisation "update" kernel with
sign +1 -- once for the real and once for the imaginary part of the coefficients, because the Krylov basis is
real while eig(H) is complex.  The result is complex, like the reference's.
"""
from typing import Callable

import torch

from .. import backend as be
from ..ops import LinearOperator
from .arnoldi import arnoldi_fact
from .stochastic import friendly_chunks, probe_chunk

_COMPLEX = {torch.float32: torch.complex64, torch.float64: torch.complex128}


class ArnoldiUnary(LinearOperator):
    """cola/linalg/unary/unary.py:63-91"""
    def __init__(self, A: LinearOperator, f: Callable, **kwargs):
        super().__init__(A.dtype, A.shape)
        self.A, self.f, self.kwargs = A, f, kwargs
        self.info = {}
        self.device = A.device

    def _matmat(self, V):
        if "start_vector" in self.kwargs.keys():
            self.kwargs.pop("start_vector")
        kw = dict(self.kwargs)
        kw.pop("key", None)
        if kw.pop("use_householder", False):
            raise NotImplementedError("Householder Arnoldi is outside the Krylov hot path")
        m = int(kw.pop("max_iters", 100))
        tol = kw.pop("tol", 1e-7)
        V = V.to(self.dtype).contiguous()
        n, k = V.shape
        out = torch.empty((n, k), dtype=_COMPLEX[self.dtype], device=V.device)
        cb = probe_chunk(n, m, self.dtype, V.device, kw.pop("probe_chunk", None))
        for c0, c1 in friendly_chunks(k, cb, self.dtype):
            blk = V[:, c0:c1].contiguous()
            b = blk.shape[1]
            Q, H, _, info = arnoldi_fact(self.A, blk, max_iters=m, tol=tol)     # Q (m+1, n, b), H (b, m+1, m)
            self.info.update(info)
            # (b, m, m): tiny, library call -- on the host: torch's CUDA eig is MAGMA's hybrid geev (it round-trips through
            # the host anyway and costs 30-40 ms for 16 matrices of 30 x 30; LAPACK on the copies: ~2 ms)
            eigvals, P = torch.linalg.eig(H[:, :-1].cpu())
            eigvals, P = eigvals.to(V.device), P.to(V.device)
            nrm = torch.zeros(b, dtype=torch.float64, device=V.device)
            be.col_dots(blk, blk, nrm)
            norms = torch.sqrt(nrm).to(self.dtype)
            e0 = torch.zeros((b, m, 1), dtype=P.dtype, device=V.device)
            e0[:, 0] = 1.0
            coef = torch.linalg.solve(P, e0).squeeze(-1) * norms[:, None]      # P^-1 e_1 ||v||
            thresh = 10 * torch.finfo(self.dtype).eps * torch.max(torch.abs(eigvals), dim=1, keepdim=True)[0]
            f_eig = torch.where(torch.abs(eigvals) > thresh, self.f(eigvals), torch.zeros_like(eigvals))
            coef = (P @ (f_eig * coef)[..., None])[..., 0]        # (b, m) complex weights of q_0..q_{m-1}
            parts = []
            for c in (coef.real, coef.imag):
                C = torch.zeros((m + 1, b), dtype=torch.float64, device=V.device)
                C[:m] = c.T.to(torch.float64)
                w = torch.zeros_like(blk)
                be.reorth_update(Q, 0, m, w, C, sign=1.0)
                parts.append(w)
            out[:, c0:c1] = torch.complex(parts[0], parts[1])
            del Q
        return out
