"""Stochastic trace / diagonal / log-determinant estimators on top of the device Lanczos:
  * LanczosUnary            cola/linalg/unary/unary.py:37-60      f(A) V ~= Q P f(Lambda) P^H e_1 ||v||
  * hutchinson_diag_estimate cola/linalg/trace/diagonal_estimation.py:158-210
  * stochastic_lanczos_quad / slq_fwd  cola/linalg/tbd/slq.py:37-75
Probe blocks are processed in chunks that fit the Krylov basis in HBM (the reference allocates
(probes, n, m+2) at once, lanczos.py:280: 438 GB for BASELINE config 4); per-probe work is independent, so
chunking does not change any per-probe value.  Probes/RHS can be sharded over ranks (cola_b200.sharding).
"""
from dataclasses import dataclass
from typing import Any, Callable, Optional

import numpy as np
import torch

from .. import backend as be
from .. import rng
from ..ops import I_like, LinearOperator, SelfAdjoint
from .algorithm_base import Algorithm
from .lanczos import lanczos_fact


def _tridiag_dense(st):
    """(b, iters, iters) dense T from a LanczosState."""
    iters = st.iters
    dt = st.V.dtype
    beta = st.alpha_acc[:iters].to(dt).T                       # diagonal (b, iters)
    T = torch.diag_embed(beta)
    if iters > 1:
        alpha = torch.sqrt(st.sub_sq[1:iters]).to(dt).T        # off-diagonal (b, iters-1)
        T = T + torch.diag_embed(alpha, offset=1) + torch.diag_embed(alpha, offset=-1)
    return T


def friendly_chunks(k, cb, dtype, start=0):
    """Column ranges covering probes [start, start + k) in blocks no wider than cb whose widths are powers of two
    (>= one 16-byte vector): 100 probes -> 64 + 32 + 4.  The reorthogonalisation kernels take their 16-byte vector
    path for those widths only; a 100- or 52-wide block falls to scalar loads and ran the API form of cfg4 at
    2.7 s per 100 probes instead of 1.6 s."""
    vec = 16 // (torch.finfo(dtype).bits // 8)
    out, c0, end = [], start, start + k
    while c0 < end:
        left = end - c0
        w = 1 << (min(left, cb).bit_length() - 1)        # largest power of two <= min(left, cb)
        if w < vec:
            w = left                                       # tail narrower than one vector: one scalar block
        out.append((c0, c0 + w))
        c0 += w
    return out


def probe_chunk(n, m, dtype, device, requested=None):
    """Largest probe-block width whose (m+2, n, b) basis fits comfortably in free HBM (<= 256, multiple of 32
    when possible so the reorth kernels take their 16-byte vector path)."""
    if requested is not None:
        return int(requested)
    free, _ = torch.cuda.mem_get_info(device)
    per_probe = (m + 6) * n * torch.finfo(dtype).bits // 8
    b = int(0.6 * free // max(per_probe, 1))
    b = max(1, min(b, 128))
    if b >= 32:
        b -= b % 32
    return b


def lanczos_chunks_lockstep(A, blocks, max_iters, tol, pbar, process, group=None):
    """lanczos_fact over column chunks with the iteration count of the reference's ONE batched factorisation.
    Its stop rule is an any() over all columns (lanczos.py:256-268), so the batch runs as long as its slowest
    column needs; a chunk on its own may stop earlier and would hand back a smaller T.  First pass: every chunk with
    the rule on.  Chunks that stopped before the longest one are redone with the rule off (tol = -1) and exactly
    that many steps.  `blocks` are callables returning an (n, b) block, `process(i, state)` the per-chunk result.
    With max_iters binding (the usual SLQ / f(A)v setting) there is no second pass.  With `group` (probe columns
    sharded over ranks) the longest count is agreed on with one MAX all-reduce."""
    results, iters = [], []
    for i, get in enumerate(blocks):
        st = lanczos_fact(A, get(), max_iters=max_iters, tol=tol, pbar=pbar)
        iters.append(st.iters)
        results.append(process(i, st))
        del st
    longest = max(iters) if iters else 0
    if group is not None:
        import torch.distributed as dist
        agreed = torch.tensor([longest], dtype=torch.int64, device=A.device)
        dist.all_reduce(agreed, op=dist.ReduceOp.MAX, group=group)
        longest = int(agreed[0])
    for i, get in enumerate(blocks):
        if iters[i] < longest:
            st = lanczos_fact(A, get(), max_iters=longest, tol=-1.0, pbar=pbar)
            results[i] = process(i, st)
            del st
    return results


class LanczosUnary(LinearOperator):
    """cola/linalg/unary/unary.py:37-60"""
    def __init__(self, A: LinearOperator, f: Callable, **kwargs):
        super().__init__(A.dtype, A.shape, annotations={SelfAdjoint})
        self.A, self.f, self.kwargs = A, f, kwargs
        self.info = {}
        self.device = A.device

    def _matmat(self, V):
        if "start_vector" in self.kwargs.keys():
            self.kwargs.pop("start_vector")
        kw = dict(self.kwargs)
        kw.pop("key", None)
        max_iters = min(kw.pop("max_iters", 100), self.A.shape[0])
        V = V.to(self.dtype).contiguous()
        n, k = V.shape
        out = torch.empty_like(V)
        cb = probe_chunk(n, max_iters, self.dtype, V.device, kw.pop("probe_chunk", None))
        chunks = friendly_chunks(k, cb, self.dtype)
        tol, pbar = kw.pop("tol", 1e-7), kw.pop("pbar", False)
        assert not kw, f"unexpected Lanczos arguments {sorted(kw)}"

        def process(i, st):
            c0, c1 = chunks[i]
            blk = V[:, c0:c1].contiguous()
            self.info.update(st.info)
            T = _tridiag_dense(st)
            eigvals, P = torch.linalg.eigh(T)                  # (b, iters, iters): tiny, library call
            nrm = torch.zeros(blk.shape[1], dtype=torch.float64, device=V.device)
            be.col_dots(blk, blk, nrm)
            norms = torch.sqrt(nrm).to(self.dtype)
            thresh = 10 * torch.finfo(self.dtype).eps * torch.max(torch.abs(eigvals), dim=1, keepdim=True)[0]
            f_eig = torch.where(torch.abs(eigvals) > thresh, self.f(eigvals), torch.zeros_like(eigvals))
            coef = P[:, 0, :] * norms[:, None]                 # conj(P)[:,0,:] * ||v||
            coef = (P @ (f_eig * coef)[..., None])[..., 0]     # (b, iters): weights of the Krylov vectors
            # out = sum_j coef[:, j] * V[j+1]  -> the reorth "update" kernel with sign +1
            C = torch.zeros((st.iters + 1, blk.shape[1]), dtype=torch.float64, device=V.device)
            C[1:] = coef.T.to(torch.float64)
            w = torch.zeros_like(blk)
            be.reorth_update(st.V, 1, st.iters + 1, w, C, sign=1.0)
            out[:, c0:c1] = w

        lanczos_chunks_lockstep(self.A, [lambda c=c: V[:, c[0]:c[1]].contiguous() for c in chunks], max_iters, tol, pbar,
                                process)
        return out


PRNGKey = Any


@dataclass
class Hutch(Algorithm):
    """cola/linalg/trace/diagonal_estimation.py:34-58"""
    tol: float = 3e-2
    max_iters: int = 10_000
    bs: int = 100
    rand: str = 'normal'
    pbar: bool = False
    key: Optional[PRNGKey] = None

    def __call__(self, A, k):
        return hutchinson_diag_estimate(A, k, **self.__dict__)[0]


def hutchinson_diag_estimate(A: LinearOperator, k=0, bs=100, tol=3e-2, max_iters=10000, pbar=False, rand='normal',
                             key=None, group=None):
    """cola/linalg/trace/diagonal_estimation.py:158-210 (k-th diagonal, numpy offset convention).  With `group`
    (a torch.distributed process group) each rank handles a contiguous slice of every 100-probe block and the
    running sums are all-reduced once per block (the stopping rule needs the global statistics)."""
    import time
    bs = min(100, A.shape[0])
    assert tol > 1e-3, "tolerance chosen too high for stochastic diagonal estimation"
    assert rand in ['normal', 'rademacher'], "rand must be 'normal' or 'rademacher'"
    key = rng.PRNGKey(42) if key is None else key
    n = A.shape[0]
    dev = A.device
    rank, world = 0, 1
    if group is not None:
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    sums = torch.zeros((2, n - abs(k)), dtype=A.dtype, device=dev)   # diag_sum, diag_sumsq

    def err(i):
        mean = sums[0] / (i * bs)
        stderr = torch.sqrt((sums[1] / (i * bs) - mean**2) / (i * bs))
        return float(torch.mean(stderr / torch.maximum(torch.abs(mean), .1 * torch.ones_like(mean))))

    samples, evals, i = [], 0, 0
    t0 = time.time()
    while True:
        e = err(i) if i > 0 else float("nan")
        samples.append(e)
        evals += 1
        if not ((i == 0) or ((i < max_iters) and (e > tol))):
            break
        key = rng.next_key(key)
        z = rng.randn(n, bs, dtype=A.dtype, key=key, device=dev)
        if rand == 'rademacher':
            z = torch.sign(z)
        lo, hi = (rank * bs) // world, ((rank + 1) * bs) // world
        zl = z[:, lo:hi].contiguous()
        Az = A @ zl
        # sum_probes est and est^2 with est = (A z)[r] * z[r + k] (roll by -k, :190-194; k < 0: (A z)[j + |k|] * z[j]): one
        # pass of the row-dots kernel over the two blocks, no (n, bs) temporaries
        part = sums if group is None else torch.zeros_like(sums)
        if hi > lo:
            be.row_dots(Az, max(-k, 0), zl, max(k, 0), n - abs(k), 1.0, part[0], out_sq=part[1], accumulate=group is None)
        if group is not None:
            dist.all_reduce(part, group=group)
            sums += part
        i += 1
    samples.append(samples[-1])
    info = {"iterations": evals, "errors": np.array(samples[2:]), "iteration_time": (time.time() - t0) / evals}
    return sums[0] / (i * bs), info


USE_TRIDIAG_QL = True   # False: dense batched torch.linalg.eigh on the (b, m, m) tridiagonals, as the reference does


def _tridiag_eig_first_row(st):
    """(eigvals (b, iters), tau (b, iters)) of the Lanczos tridiagonals: eigenvalues and first eigenvector
    components, all SLQ needs (slq.py:44-46), from the O(m^2) QL kernel instead of a dense O(m^3) eigh.  The
    entries of T are first rounded to the operator dtype, which is what the reference's T holds."""
    iters = st.iters
    dt = st.V.dtype
    d = st.alpha_acc[:iters].to(dt).to(torch.float64).contiguous()                # (iters, b)
    e = torch.zeros_like(d)
    if iters > 1:
        e[:iters - 1] = torch.sqrt(st.sub_sq[1:iters]).to(dt).to(torch.float64)
    z = torch.empty_like(d)
    status = torch.zeros(d.shape[1], dtype=torch.int32, device=d.device)
    be.tridiag_eig_first_row(d, e, z, status)
    if bool(status.any()):
        raise RuntimeError("tridiagonal QL iteration did not converge")
    return d.T.to(dt), z.T.to(dt)


def slq_per_probe(A, fun, Z, max_iters, tol, pbar=False):
    """n * sum_j tau_j^2 f(lambda_j) for every probe column of Z (n, b)  (slq.py:42-51)."""
    return _quadrature(A, fun, lanczos_fact(A, Z, max_iters, tol, pbar))


def _quadrature(A, fun, st):
    eps = torch.finfo(A.dtype).eps
    if USE_TRIDIAG_QL:
        eigvals, tau = _tridiag_eig_first_row(st)
    else:
        eigvals, Q = torch.linalg.eigh(_tridiag_dense(st))
        tau = Q[..., 0, :]
    const = 10 * eps * torch.max(eigvals, dim=1, keepdim=True)[0]
    fn_vals = torch.where(torch.abs(eigvals) > const, fun(eigvals), torch.zeros_like(eigvals))
    return A.shape[-2] * torch.sum(tau**2 * fn_vals, dim=-1)


def slq_fwd(A, fun, num_samples, max_iters, tol, pbar, key, probe_chunk_size=None, group=None, probes=None):
    """cola/linalg/tbd/slq.py:34-52: the estimate, with slq_bwd (slq.py:10-31) as its backward when a parameter of A
    requires grad (cola_b200/autograd.py)."""
    from .. import autograd as ag
    kw = dict(probe_chunk_size=probe_chunk_size, group=group, probes=probes)
    if ag.needs_grad(A):
        return ag.slq_with_grad(A, lambda: _slq_fwd(A, fun, num_samples, max_iters, tol, pbar, key, **kw),
                                dict(num_samples=num_samples, key=key, probe_chunk_size=probe_chunk_size))
    return _slq_fwd(A, fun, num_samples, max_iters, tol, pbar, key, **kw)


def _slq_fwd(A, fun, num_samples, max_iters, tol, pbar, key, probe_chunk_size=None, group=None, probes=None):
    """cola/linalg/tbd/slq.py:37-52.  The full (n, num_samples) probe block is drawn exactly as the reference
    draws it (one randn call), then processed in column chunks; with `group`, rank r takes the contiguous
    column range [r*P/G, (r+1)*P/G) and a single all-reduce combines (sum, count)."""
    n = A.shape[1]
    max_iters = min(max_iters, A.shape[0])
    rank, world = 0, 1
    if group is not None:
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = (rank * num_samples) // world, ((rank + 1) * num_samples) // world
    if probes is None:
        probes = DeferredProbes(n, num_samples, A.dtype, A.device, key)
    cb = probe_chunk(n, max_iters, A.dtype, A.device, probe_chunk_size)
    total = torch.zeros(2, dtype=torch.float64, device=A.device)
    chunks = friendly_chunks(hi - lo, cb, A.dtype, start=lo)
    ests = lanczos_chunks_lockstep(A, [lambda c=c: probes.columns(c[0], c[1]) for c in chunks], max_iters, tol, pbar,
                                   lambda i, st: _quadrature(A, fun, st), group=group)
    for est in ests:
        total[0] += est.to(torch.float64).sum()
        total[1] += est.numel()
    if group is not None:
        dist.all_reduce(total, group=group)          # the one collective of the sharded estimate
    return (total[0] / total[1]).to(A.dtype)


class DeferredProbes:
    """The reference's `randn(n, num_samples, key)` block, materialised lazily by column range.
    With rng.PROBE_DEVICE == 'cpu' the whole block is drawn on the host generator (parity with the CPU
    reference) and only the requested columns are moved to the GPU; otherwise it is drawn on the device."""
    def __init__(self, n, num, dtype, device, key):
        self.n, self.num, self.dtype, self.device, self.key = n, num, dtype, device, key
        self._full = None

    def columns(self, c0, c1):
        if self._full is None:
            draw_on = rng.PROBE_DEVICE if rng.PROBE_DEVICE is not None else self.device
            self._full = rng.randn(self.n, self.num, dtype=self.dtype, key=self.key, device=draw_on)
        return self._full[:, c0:c1].to(self.device).contiguous()


def stochastic_lanczos_quad(A: LinearOperator, fun: Callable, max_iters: int = 100, tol: float = 1e-5, vtol=0.1,
                            pbar: bool = False, key=None, **kw):
    """cola/linalg/tbd/slq.py:55-75"""
    num_samples = max(int(1 / vtol**2), 1)
    return slq_fwd(A, fun, num_samples=num_samples, max_iters=max_iters, tol=tol, pbar=pbar, key=key, **kw)
