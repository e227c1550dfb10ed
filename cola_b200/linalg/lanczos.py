"""Batched Lanczos with double classical Gram-Schmidt (CGS2) full reorthogonalisation on the device.

Interface of the reference (cola/linalg/decompositions/lanczos.py): `lanczos(A, start_vector, max_iters, tol,
pbar, key) -> (Q, T, info)`, `lanczos_eigs(...) -> (eigvals, V, info)`; `lanczos_fact` keeps its role as the
inner loop.  Differences underneath (DESIGN.md "Lanczos"):
  * the Krylov basis is stored (m+2, n, b) -- vector j is an (n, b) block, the matmat operand layout -- instead
    of (b, n, m+2) with the Krylov index fastest (lanczos.py:280); the API hands back a strided *view* with
    the reference's logical shape (b, n, iters), no copy;
  * per step: matmat with <w, v_i> fused; one three-term sweep; CGS2 as two `V^T w` / `w -= V c` kernel pairs
    restricted to the filled vectors (the reference sweeps all m+2 columns, filled or not, through four
    (b, n, m+2) temporaries), with ||w||^2 fused into the last update.
"""
import time

import numpy as np
import torch

from .. import backend as be
from .. import rng
from ..ops import Dense, LinearOperator, SelfAdjoint, Tridiagonal, Unitary, lazify


class LanczosState:
    """Device-resident result of lanczos_fact."""
    def __init__(self, V, alpha_acc, sub_sq, i, info):
        self.V, self.alpha_acc, self.sub_sq, self.i, self.info = V, alpha_acc, sub_sq, i, info

    @property
    def iters(self):
        return self.i - 1


def lanczos_fact(A: LinearOperator, rhs, max_iters=100, tol=1e-7, pbar=False):
    """rhs (n, b) on the device.  Mirrors lanczos.py:235-284 (init_lanczos + lanczos_fact)."""
    be.require_cuda(rhs, "start vectors")
    A.plan()                                               # validate / compile once; the loop uses matmat_into
    dt = A.dtype
    rhs = rhs.to(dt).contiguous()
    n, b = rhs.shape
    m = int(max_iters)
    dev = rhs.device
    V = torch.empty((m + 2, n, b), dtype=dt, device=dev)
    V[0].zero_()
    alpha_acc = torch.zeros((m, b), dtype=torch.float64, device=dev)      # diag   (<w, v_i>)
    sub_sq = torch.zeros((m + 1, b), dtype=torch.float64, device=dev)     # subdiag^2 (||w||^2)
    CC = torch.zeros((2, m + 2, b), dtype=torch.float64, device=dev)    # coefficients of the two Gram-Schmidt passes
    C, C2 = CC[0], CC[1]
    nrm = torch.zeros(b, dtype=torch.float64, device=dev)
    # init_lanczos: V[1] = rhs / ||rhs||   (lanczos.py:281-283)
    be.col_dots(rhs, rhs, nrm)
    be.col_scale(rhs, V[1], nrm, take_sqrt=True, mode=2)

    samples = []
    t0 = time.time()
    i = 1
    evals = 0
    sub_host = np.zeros((m + 1, b))
    while True:
        # error + cond_fun (lanczos.py:256-268)
        base = np.maximum(sub_host[1], 1e-30)
        samples.append(float(np.max(sub_host[i - 1] / base) + (1.0 if i <= 1 else 0.0)))
        evals += 1
        with np.errstate(invalid="ignore"):
            large = (_cast(sub_host[i - 1], dt) > _cast(tol * _cast(sub_host[1], dt), dt)) | (i <= 1)
        if not ((i <= m) and bool(np.any(large))):
            break
        # body (lanczos.py:238-254)
        vi = V[i]
        nrm.zero_()
        if i == 1:
            be.col_dots(vi, vi, nrm)                       # the reference renormalises V[1] as well
            be.col_scale(vi, vi, nrm, take_sqrt=True, mode=2)
        else:
            be.col_scale(vi, vi, sub_sq[i - 1], take_sqrt=True, mode=2)   # ||V[i]|| is the stored subdiag
        w = V[i + 1]
        A.matmat_into(vi, w, dots=alpha_acc[i - 1])        # w = A v_i ; diag[i-1] = <w, v_i>
        be.lanczos_three_term(w, vi, V[i - 1] if i > 1 else None, alpha_acc[i - 1], sub_sq[i - 1] if i > 1 else None)
        # do_double_gram (lanczos.py:287-296): dots, update, dots, update -- the middle two share one sweep
        CC.zero_()                                         # one fill for both passes' accumulators
        be.reorth_dots(V, 1, i + 1, w, C)
        if not be.reorth_update_dots(V, 1, i + 1, w, C, C2, sign=-1.0):
            be.reorth_update(V, 1, i + 1, w, C, sign=-1.0)
            be.reorth_dots(V, 1, i + 1, w, C2)
        be.reorth_update(V, 1, i + 1, w, C2, sign=-1.0, wnorm2=sub_sq[i])
        sub_host[i] = np.sqrt(be.read_small(sub_sq[i]).numpy())     # poll: the stop rule needs subdiag[i]
        if i == 1:
            pass
        i += 1
    elapsed = time.time() - t0
    samples.append(samples[-1])
    info = {"iterations": evals, "errors": np.array(samples[2:]), "iteration_time": elapsed / evals}
    return LanczosState(V, alpha_acc, sub_sq, i, info)


def _cast(x, dt):
    return np.asarray(x, dtype=np.float32 if dt == torch.float32 else np.float64)


def lanczos(A: LinearOperator, start_vector=None, max_iters=100, tol=1e-7, pbar=False, key=None):
    """cola/linalg/decompositions/lanczos.py:185-232.  Returns (Q, T, info):
       1-D start vector: Q = Unitary(Dense (n, iters)), T = Tridiagonal(alpha, beta, alpha);
       (n, b) start block: Q = Unitary(Dense) whose `.A` is the (b, n, iters) view, T = Tridiagonal with
       batched leaves alpha (b, iters-1, 1), beta (b, iters, 1) (what vmap(Tridiagonal) builds)."""
    max_iters = min(max_iters, A.shape[0])
    if start_vector is None:
        key = rng.PRNGKey(42) if key is None else key
        start_vector = rng.randn(A.shape[0], dtype=A.dtype, device=A.device, key=key)
    rhs = start_vector[:, None] if len(start_vector.shape) == 1 else start_vector
    st = lanczos_fact(A, rhs, max_iters, tol, pbar)
    iters = st.iters
    dt = A.dtype
    alpha = torch.sqrt(st.sub_sq[1:iters]).to(dt).T.contiguous()          # (b, iters-1) off-diagonal
    beta = st.alpha_acc[:iters].to(dt).T.contiguous()                      # (b, iters)   diagonal
    Qv = st.V[1:iters + 1].permute(2, 1, 0)                                # (b, n, iters) view, no copy
    if len(start_vector.shape) == 1:
        T = Tridiagonal(alpha[0], beta[0], alpha[0])
        Q = Unitary(Dense(Qv[0]))
        return Q, T, st.info
    T = BatchedTridiagonal(alpha, beta)
    Q = Unitary(BatchedDense(Qv))
    return Q, T, st.info


class BatchedDense(LinearOperator):
    """What `vmap(Dense)(Q)` is in the reference: a Dense whose leaf `.A` carries a leading batch dim."""
    def __init__(self, A):
        self.A = A
        super().__init__(dtype=A.dtype, shape=tuple(A.shape[-2:]))

    def to_dense(self):
        return self.A


class BatchedTridiagonal(LinearOperator):
    """What `vmap(Tridiagonal)(alpha, beta, alpha)` is in the reference: leaves (b, m-1, 1), (b, m, 1)."""
    def __init__(self, alpha, beta):
        self.alpha, self.beta, self.gamma = alpha[..., None], beta[..., None], alpha[..., None]
        super().__init__(dtype=beta.dtype, shape=(beta.shape[-1], beta.shape[-1]))

    def to_dense(self):
        T = torch.diag_embed(self.beta[..., 0])
        if self.alpha.shape[-2] > 0:
            T = T + torch.diag_embed(self.alpha[..., 0], offset=-1) + torch.diag_embed(self.gamma[..., 0], offset=1)
        return T


def lanczos_eigs(A: LinearOperator, start_vector=None, max_iters=100, tol=1e-7, pbar=False, key=None):
    """cola/linalg/decompositions/lanczos.py:34-61: Ritz pairs, ascending; eigenvectors stay lazy (Q @ S)."""
    Q, T, info = lanczos(A=A, start_vector=start_vector, max_iters=max_iters, tol=tol, pbar=pbar, key=key)
    eigvals, eigvectors = torch.linalg.eigh(T.to_dense())      # (iters x iters): not a hot spot
    idx = torch.argsort(eigvals, dim=-1)
    V = Q @ lazify(eigvectors[:, idx])
    eigvals = eigvals[..., idx]
    return eigvals, V, info
