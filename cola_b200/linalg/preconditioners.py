"""Nystrom preconditioner for CG (cola/linalg/preconditioning/preconditioners.py:97-170).

`NystromPrecond(A, rank, mu, eps, adjust_mu, key)` is `P = U diag(s) U^T + I` with (Lambda, U) the rank-r Nystrom
approximation of A and `s = (min Lambda + mu') / (Lambda + mu') - 1`.  Construction follows `get_nys_approx`
(:145-157) step by step; its small dense factorizations (thin QR of the n x r sketch, r x r Cholesky, triangular
solve, thin SVD) are library calls on the device, as in the reference, and `A @ Omega` is this package's matmat.
Applying P is two tall-skinny products and lives in the operator plan as `Product(Dense(U), Dense(s * U^T)) + I`,
so the preconditioned CG loop (linalg/cg.py) treats it like any other operator, with <r, P r> fused into the apply.
"""
import torch

from .. import rng
from ..ops import Dense, I_like, LinearOperator, Product, Sum


def get_nys_approx(A, Omega, eps):
    """preconditioners.py:145-157"""
    Omega, _ = torch.linalg.qr(Omega, mode="reduced")
    Omega = Omega.contiguous()
    Y = A @ Omega
    nu = eps * torch.linalg.norm(Y)
    Y = Y + nu * Omega
    C = torch.linalg.cholesky(Omega.T @ Y)
    aux = torch.linalg.solve_triangular(C, Y.T, upper=False)
    B = aux.T
    U, Sigma, _ = torch.linalg.svd(B, full_matrices=False)
    Lambda = torch.clip(Sigma**2.0 - nu, min=0.0)
    return Lambda, U


class NystromPrecond(LinearOperator):
    """preconditioners.py:97-130"""
    def __init__(self, A, rank, mu=1e-7, eps=1e-8, adjust_mu=True, key=None):
        super().__init__(dtype=A.dtype, shape=A.shape)
        self.device = A.device
        key = rng.PRNGKey(42) if key is None else key
        Omega = rng.randn(A.shape[0], rank, dtype=A.dtype, device=A.device, key=key)
        self._create_approx(A=A, Omega=Omega, mu=mu, eps=eps, adjust_mu=adjust_mu)

    def _create_approx(self, A, Omega, mu, eps, adjust_mu):
        self.Lambda, self.U = get_nys_approx(A=A, Omega=Omega, eps=eps)
        self.adjusted_mu = amu = mu * torch.max(self.Lambda) if adjust_mu else mu
        self.subspace_num = torch.min(self.Lambda) + amu
        self.subspace_denom = self.Lambda + amu
        self.subspace_scaling = (self.subspace_num / self.subspace_denom - 1)[:, None]
        self.preconditioned_eigmax = torch.min(self.Lambda) + amu
        self.preconditioned_eigmin = amu
        U = self.U.contiguous()
        SUt = (self.subspace_scaling * U.T).contiguous()            # (r, n)
        self._op = Sum(Product(Dense(U), Dense(SUt)), I_like(self))

    def plan(self):
        return self._op.plan()
