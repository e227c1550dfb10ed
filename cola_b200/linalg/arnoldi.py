"""Batched Arnoldi with modified Gram-Schmidt on the device (cola/linalg/decompositions/arnoldi.py:166-205,
289-335).  `arnoldi(A, start_vector, max_iters, tol, use_householder, pbar, key) -> (Q, H, info)`.

The reference's Python `for_loop` over basis vectors (arnoldi.py:304-311: one dot, one store, one axpy per j, each
a separate eager op with its own temporaries) becomes ONE cooperative launch per step (`mgs_chain`: pass j subtracts
h_{j-1} q_{j-1} from w and accumulates h_j = <q_j, w> in the same sweep, grid-wide sync between passes, w resident in
L2, each q_j read from DRAM once), in exact MGS order (the factorisation matches the reference to rounding, not just
to CGS2 accuracy).  Blocks the chain kernel declines run the same passes as separate `mgs_link` launches.
The basis is stored (m+1, n, b), the matmat operand layout.
"""
import time

import numpy as np
import torch

from .. import backend as be
from .. import rng
from ..ops import Dense, LinearOperator, Stiefel, lazify
from .lanczos import BatchedDense

USE_CHAIN = True     # False: one mgs_link launch per link (A/B checks)


def arnoldi_fact(A: LinearOperator, rhs, max_iters, tol, pbar=False):
    """rhs (n, b) on the device -> (Q (m+1, n, b), H (b, m+1, m) in A.dtype, idx, info)."""
    be.require_cuda(rhs, "start vectors")
    A.plan()                                               # validate / compile once; the loop uses matmat_into
    dt = A.dtype
    rhs = rhs.to(dt).contiguous()
    n, b = rhs.shape
    m = int(max_iters)
    dev = rhs.device
    m_eff = min(m, A.shape[0])
    Q = torch.zeros((m + 1, n, b), dtype=dt, device=dev)
    H64 = torch.zeros((m, m + 1, b), dtype=torch.float64, device=dev)   # [idx][j][c]  (column idx of H)
    nrm_sq = torch.zeros((m + 1, b), dtype=torch.float64, device=dev)
    be.col_dots(rhs, rhs, nrm_sq[0])
    be.col_scale(rhs, Q[0], nrm_sq[0], take_sqrt=True, mode=2)         # init_arnoldi (arnoldi.py:327-335)
    norm_host = np.sqrt(be.read_small(nrm_sq[0]).numpy())
    h10 = None
    samples, evals, idx = [], 0, 0
    t0 = time.time()
    while True:
        samples.append(float(norm_host[0]))
        evals += 1
        with np.errstate(invalid="ignore"):
            ref = np.zeros_like(norm_host) if h10 is None else h10
            large = (_cast(norm_host, dt) > _cast(tol * _cast(ref, dt), dt)) | (idx <= 0)
        if not ((idx < m_eff) and bool(np.any(large))):
            break
        w = Q[idx + 1]
        A.matmat_into(Q[idx], w)
        h = H64[idx]
        # MGS chain (arnoldi.py:304-311): one cooperative launch per step; link by link where the library declines
        if not (USE_CHAIN and be.mgs_chain(w, Q, idx + 1, h, wnorm2=nrm_sq[idx + 1])):
            be.mgs_link(w, None, None, Q[0], h[0])
            for j in range(1, idx + 1):
                be.mgs_link(w, Q[j - 1], h[j - 1], Q[j], h[j])
            be.mgs_link(w, Q[idx], h[idx], None, None, wnorm2=nrm_sq[idx + 1])
        be.col_scale(w, w, nrm_sq[idx + 1], take_sqrt=True, mode=3, a=tol / 2.)   # w /= clip(norm, tol/2)
        norm_host = np.sqrt(be.read_small(nrm_sq[idx + 1]).numpy())                         # poll for the stop rule
        if idx == 0:
            h10 = norm_host.copy()                                                # H[:, 1, 0]
        idx += 1
    elapsed = time.time() - t0
    samples.append(samples[-1])
    info = {"iterations": evals, "errors": np.array(samples[2:]), "iteration_time": elapsed / evals}
    # assemble H (b, m+1, m): column j holds h_0..h_j and the norm at row j+1
    H = torch.zeros((b, m + 1, m), dtype=dt, device=dev)
    if idx > 0:
        Hc = H64[:idx].clone()                                  # (idx, m+1, b)
        rows = torch.arange(idx, device=dev)
        Hc[rows, rows + 1] = torch.sqrt(nrm_sq[1:idx + 1])
        H[:, :, :idx] = Hc.permute(2, 1, 0).to(dt)
    return Q, H, idx, info


def _cast(x, dt):
    return np.asarray(x, dtype=np.float32 if dt == torch.float32 else np.float64)


def arnoldi(A: LinearOperator, start_vector=None, max_iters=100, tol=1e-7, use_householder=False, pbar=False,
            key=None):
    """cola/linalg/decompositions/arnoldi.py:166-205."""
    if use_householder:
        raise NotImplementedError("Householder Arnoldi is outside the Krylov hot path (disabled in the reference's "
                                  "own tests, tests/algorithms/test_arnoldi.py:187)")
    if start_vector is None:
        key = rng.PRNGKey(42) if key is None else key
        start_vector = rng.randn(A.shape[-1], dtype=A.dtype, device=A.device, key=key)
    rhs = start_vector[:, None] if len(start_vector.shape) == 1 else start_vector
    Q, H, _, info = arnoldi_fact(A, rhs, max_iters=max_iters, tol=tol, pbar=pbar)
    Qv = Q.permute(2, 1, 0)                                     # (b, n, m+1) view
    if len(start_vector.shape) == 1:
        return Stiefel(Dense(Qv[0])), Dense(H[0]), info
    return Stiefel(BatchedDense(Qv)), BatchedDense(H), info


def arnoldi_eigs(A: LinearOperator, start_vector=None, max_iters=100, tol=1e-7, use_householder=False, pbar=False,
                 key=None):
    """cola/linalg/decompositions/arnoldi.py:35-62."""
    Q, H, info = arnoldi(A=A, start_vector=start_vector, max_iters=max_iters, tol=tol,
                         use_householder=use_householder, pbar=pbar, key=key)
    Qd, Hd = Q.to_dense()[:, :-1], H.to_dense()[:-1]
    eigvals, vs = torch.linalg.eig(Hd.cpu())                    # (m x m) library call, on the host (CUDA eig is a hybrid that goes there anyway)
    eigvals, vs = eigvals.to(Hd.device), vs.to(Hd.device)
    # complex Ritz vectors: the product Q @ vs is a small-k dense contraction, done on real/imag parts
    Qc = Qd.contiguous()
    re = Dense(Qc) @ vs.real.contiguous()
    im = Dense(Qc) @ vs.imag.contiguous()
    eigvectors = lazify(torch.complex(re, im))
    return eigvals, eigvectors, info
