"""The cola.linalg dispatch surface for the Krylov path: solve / inv / eig / logdet / slogdet / trace / diag with
the reference's signatures (cola/linalg/inverse/inv.py, eig/eigs.py, logdet/logdet.py, trace/diag_trace.py,
unary/unary.py:249-262).  The reference resolves these with plum multiple dispatch; here the same rules are
plain isinstance chains (plum is not a dependency of this package).  Only rules that lead to the Krylov loops
or to closed forms of the hot-path operators are restated; dense small-matrix algorithms (Cholesky/LU/Eigh) are
torch library calls exactly as in the reference and are not part of the accelerated path."""
from dataclasses import dataclass
from typing import Any, Optional

import numpy as np
import torch

from .. import rng
from ..ops import (PSD, BlockDiag, Dense, Diagonal, I_like, Identity, Kronecker, KronSum, LinearOperator, Product,
                   ScalarMul, SelfAdjoint, Sum, Transpose, Triangular, Unitary, lazify)
from .algorithm_base import Algorithm, Auto, IterativeOperatorWInfo
from .arnoldi import arnoldi, arnoldi_eigs
from .cg import CG
from .gmres import GMRES
from .power_iteration import PowerIteration
from .lanczos import lanczos, lanczos_eigs
from .stochastic import Hutch, LanczosUnary, hutchinson_diag_estimate
from .unary import ArnoldiUnary


@dataclass
class Lanczos(Algorithm):
    """cola/linalg/decompositions/decompositions.py:116-144"""
    start_vector: Any = None
    max_iters: int = 1_000
    tol: float = 1e-6
    pbar: bool = False
    key: Optional[Any] = None

    def __call__(self, A: LinearOperator):
        return lanczos(A, **self.__dict__)


@dataclass
class Arnoldi(Algorithm):
    """cola/linalg/decompositions/decompositions.py:60-88"""
    start_vector: Any = None
    max_iters: int = 1_000
    tol: float = 1e-6
    pbar: bool = False
    key: Optional[Any] = None

    def __call__(self, A: LinearOperator):
        return arnoldi(A, **self.__dict__)


@dataclass
class Cholesky(Algorithm):
    pass


@dataclass
class Eigh(Algorithm):
    pass


@dataclass
class LU(Algorithm):
    pass


@dataclass
class Exact(Algorithm):
    """cola/linalg/trace/diagonal_estimation.py:13-31"""
    bs: int = 100
    pbar: bool = False

    def __call__(self, A, k):
        return exact_diag(A, k, self.bs)


def exact_diag(A, k, bs, group=None):
    """diagonal_estimation.py:84-128: blocks of 100 identity columns through the fused matmat.  Column c of A holds
    entry (c - k, c) of the k-th diagonal, so each block is read off directly instead of being masked with a shifted
    identity block and reduced over columns (the reference's form; it also fails on a ragged last block for k != 0,
    this one does not).  With `group` (a torch.distributed process group) the column blocks are dealt round-robin
    to the ranks and the disjoint pieces are combined by one all-reduce (independent units, operator
    replicated)."""
    n = A.shape[0]
    bs = min(100, n)
    out = torch.zeros(n - abs(k), dtype=A.dtype, device=A.device)
    rank, world = 0, 1
    if group is not None:
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    for b, i in enumerate(range(0, n, bs)):
        if b % world != rank:
            continue
        w = min(bs, n - i)
        chunk = torch.zeros((n, w), dtype=A.dtype, device=A.device)
        chunk[i:i + w] = torch.eye(w, dtype=A.dtype, device=A.device)
        AE = A @ chunk
        cols = torch.arange(i, i + w, device=A.device)
        rows = cols - k
        ok = (rows >= 0) & (rows < n)
        out[(rows if k >= 0 else cols)[ok]] = AE[rows[ok], (cols - i)[ok]]
    if world > 1:
        dist.all_reduce(out, group=group)              # every entry was written by exactly one rank
    return out


# ---------------------------------------------------------------------------------------------------- inverse
def solve(A, b, alg=Auto()):
    """cola/linalg/inverse/inv.py:23-39"""
    return inv(A, alg) @ b


class TriangularInv(LinearOperator):
    """cola/linalg/inverse/inv.py:154-165: a triangular solve per application (library trsm on the device, as the
    reference's xnp.solvetri); inside an operator tree it is an opaque core of the plan."""
    def __init__(self, A: Triangular):
        super().__init__(A.dtype, A.shape)
        self.A = A.to_dense()
        self.lower = A.lower
        self.device = A.device

    def _matmat(self, X):
        return torch.linalg.solve_triangular(self.A, X, upper=not self.lower)

    def _rmatmat(self, X):
        return torch.linalg.solve_triangular(self.A.T, X.T, upper=self.lower).T


class _DenseInverse(LinearOperator):
    """inv(A, Cholesky) for small PSD operators: torch.linalg.cholesky on the dense matrix (library, as in the
    reference: decompositions.py:147-175 + TriangularInv)."""
    def __init__(self, A):
        super().__init__(A.dtype, A.shape)
        self.L = torch.linalg.cholesky(A.to_dense())
        self.device = A.device

    def _matmat(self, X):
        return torch.cholesky_solve(X, self.L)


class _DenseLUInverse(LinearOperator):
    """inv(A, LU) for small general operators: torch.linalg.lu_factor / lu_solve on the dense matrix (library, as in
    the reference: decompositions.py:178-186 + TriangularInv / Permutation products, inv.py:102-105)."""
    def __init__(self, A):
        super().__init__(A.dtype, A.shape)
        self.LU, self.piv = torch.linalg.lu_factor(A.to_dense())
        self.device = A.device

    def _matmat(self, X):
        return torch.linalg.lu_solve(self.LU, self.piv, X)

    def _rmatmat(self, X):
        return torch.linalg.lu_solve(self.LU, self.piv, X.T.contiguous(), adjoint=True).T


def inv(A: LinearOperator, alg: Algorithm = Auto()):
    """cola/linalg/inverse/inv.py:42-151"""
    # structure rules first (inv.py:108-151)
    if A.isa(Unitary):
        return Unitary(A.H)
    if isinstance(A, Identity):
        return A
    if isinstance(A, ScalarMul):
        return ScalarMul(1 / A.c, shape=A.shape, dtype=A.dtype, device=A.c.device)
    if type(A) is Product and all(M.shape[-2] == M.shape[-1] for M in A.Ms):
        return Product(*reversed([inv(M, alg) for M in A.Ms]))
    if isinstance(A, BlockDiag):
        return BlockDiag(*[inv(M, alg) for M in A.Ms], multiplicities=A.multiplicities)
    if isinstance(A, Kronecker):
        return Kronecker(*[inv(M, alg) for M in A.Ms])
    if isinstance(A, Diagonal):
        return Diagonal(1. / A.diag)
    if isinstance(A, Triangular):   # inv.py:149-151
        return TriangularInv(A)
    # base cases
    if isinstance(alg, Auto):   # inv.py:72-92
        small = bool(np.prod(A.shape) <= 1e6)
        if A.isa(PSD):
            alg = Cholesky() if small else CG(**alg.__dict__)
        elif small:
            alg = LU()
        else:
            alg = GMRES(**alg.__dict__)
    if isinstance(alg, CG):     # inv.py:66-69
        assert A.isa(PSD), "CG only valid for PSD matrices, wrap in cola.PSD if desired"
        return IterativeOperatorWInfo(A, alg)
    if isinstance(alg, GMRES):  # inv.py:60-62
        return IterativeOperatorWInfo(A, alg)
    if isinstance(alg, Cholesky):
        assert A.isa(PSD), "Cholesky only valid for PSD matrices, wrap in cola.PSD if desired"
        return _DenseInverse(A)
    if isinstance(alg, LU):         # inv.py:102-105
        return _DenseLUInverse(A)
    raise NotImplementedError(f"inv with {type(alg).__name__} is outside the Krylov hot path")


# ---------------------------------------------------------------------------------------------------- pseudo-inverse
@dataclass
class LSTSQ(Algorithm):
    """cola/linalg/inverse/pinv.py:14-19"""
    def __call__(self, A: LinearOperator):
        return LSTSQSolve(A)


class LSTSQSolve(LinearOperator):
    """cola/linalg/inverse/pinv.py:22-28: dense least squares (library), for small operators."""
    def __init__(self, A: LinearOperator):
        super().__init__(A.dtype, (A.shape[-1], A.shape[-2]))
        self.A = A.to_dense()
        self.device = A.device

    def _matmat(self, X):
        return torch.linalg.lstsq(self.A, X).solution


def pinv(A: LinearOperator, alg: Algorithm = Auto()):
    """cola/linalg/inverse/pinv.py:31-96.  The CG rule is the hot-path one: CG on the normal equations A^H A (a
    Product chain of the plan: transposed core, core), plus eps * I, applied to A^H b."""
    if isinstance(A, Identity):
        return A
    if isinstance(A, ScalarMul):
        return ScalarMul(1 / A.c, shape=A.shape, dtype=A.dtype, device=A.c.device)
    if isinstance(A, Diagonal):
        return Diagonal(1. / A.diag)
    if isinstance(alg, Auto):          # pinv.py:50-62
        alg = LSTSQ() if bool(np.prod(A.shape) <= 1e6) else CG(**alg.__dict__)
    if isinstance(alg, CG):            # pinv.py:65-71
        M = A.H @ A
        cons = (1e-6 if A.dtype == torch.float32 else 1e-15) * max(A.shape)
        Op = IterativeOperatorWInfo(M, alg)
        return PSD(Op + cons * I_like(M)) @ A.H
    if isinstance(alg, LSTSQ):
        return LSTSQSolve(A)
    raise NotImplementedError(f"pinv with {type(alg).__name__} is outside the Krylov hot path")


# ---------------------------------------------------------------------------------------------------- eig
def get_slice(num, which):
    """cola/linalg/decompositions/decompositions.py:214-224"""
    if num == -1:
        raise ValueError(f"Number of eigenvalues {num} must be explicitly specified")
    if which == "SM":
        return slice(0, num, None)
    if which == "LM":
        return slice(-1 if num is None else -num, None, None)
    raise NotImplementedError(f"which={which} is not implemented")


def eigmax(A: LinearOperator, alg: Algorithm = Auto()):
    """cola/linalg/eig/eigs.py:44-57"""
    es, _ = eig(A, k=1, which="LM", alg=alg)
    return es[0]


def eigmin(A: LinearOperator, alg: Algorithm = Auto()):
    """cola/linalg/eig/eigs.py:60-73"""
    es, _ = eig(A, k=1, which="SM", alg=alg)
    return es[0]


def eig(A: LinearOperator, k: int, which: str = "LM", alg: Algorithm = Auto()):
    """cola/linalg/eig/eigs.py:19-182 (Lanczos, Arnoldi and Auto/Eigh rules)."""
    eig_slice = get_slice(k, which)
    # closed forms (eigs.py:143-149, 175-182).  With an explicit Lanczos / Arnoldi / ... the reference's rules for
    # (LinearOperator, that algorithm) and (Identity | Diagonal, Algorithm) are ambiguous; the algorithm is honoured.
    if isinstance(A, Identity) and isinstance(alg, Auto):
        vals = torch.ones(A.shape[0], dtype=A.dtype, device=A.device)
        vecs = torch.eye(A.shape[0], dtype=A.dtype, device=A.device)
        return vals[eig_slice], Unitary(lazify(vecs[:, eig_slice]))
    if isinstance(A, Diagonal) and isinstance(alg, Auto):
        order = torch.argsort(A.diag)
        vecs = torch.eye(A.shape[0], dtype=A.dtype, device=A.device)[:, order]
        return A.diag[order][eig_slice], Unitary(lazify(vecs[:, eig_slice]))
    if isinstance(alg, Auto):   # eigs.py:76-96
        small = bool(np.prod(A.shape) <= 1e6)
        if k == 1 and which == "LM":
            alg = PowerIteration(**alg.__dict__)
        elif A.isa(SelfAdjoint):
            alg = Eigh() if small else Lanczos(**alg.__dict__)
        else:
            alg = Arnoldi(**alg.__dict__)
    if isinstance(alg, PowerIteration):   # eigs.py:136-140
        assert k == 1 and which == 'LM', "PowerIteration only valid for k=1 and which='LM'"
        v, emax, _ = alg(A)
        return emax[None], v[:, None]
    if isinstance(alg, Lanczos):   # eigs.py:106-111
        assert A.isa(SelfAdjoint)
        eig_vals, eig_vecs, _ = lanczos_eigs(A, **alg.__dict__)
        return eig_vals[eig_slice], eig_vecs[:, eig_slice]
    if isinstance(alg, Arnoldi):   # eigs.py:99-103
        eig_vals, eig_vecs, _ = arnoldi_eigs(A, **alg.__dict__)
        return eig_vals[eig_slice], eig_vecs[:, eig_slice]
    if isinstance(alg, Eigh):
        vals, vecs = torch.linalg.eigh(A.to_dense())
        return vals[eig_slice], Unitary(lazify(vecs[:, eig_slice]))
    raise NotImplementedError(f"eig with {type(alg).__name__} is outside the Krylov hot path")


# ---------------------------------------------------------------------------------------------------- trace / diag
def diag(A: LinearOperator, k: int = 0, alg: Algorithm = Auto()):
    """cola/linalg/trace/diag_trace.py:22-120: the structure rules come first (a Sum is estimated term by term, so
    Dense / Diagonal / Identity / Kronecker-of-those terms contribute their exact diagonals whatever `alg` is)."""
    if isinstance(A, Dense):                         # :58-61
        return torch.diagonal(A.A, offset=k)
    if isinstance(A, (Identity, Diagonal)):          # :64-78
        if k == 0:
            return A.diag if isinstance(A, Diagonal) else torch.ones(A.shape[0], dtype=A.dtype, device=A.device)
        return torch.zeros(A.shape[0] - abs(k), dtype=A.dtype, device=A.device)
    if type(A) is Sum:                               # :81-84
        return sum(diag(M, k, alg) for M in A.Ms)
    if isinstance(A, BlockDiag):                     # :87-91
        assert k == 0, "Havent filled this case yet, need to pad with 0s"
        return torch.cat([d for M, m in zip(A.Ms, A.multiplicities) for d in [diag(M, k, alg)] * m])
    if isinstance(A, ScalarMul):                     # :94-96
        return A.c * diag(I_like(A), k, alg)
    if isinstance(A, (Kronecker, KronSum)):          # :103-119: outer product / outer sum of the factors' diagonals
        assert k == 0, "Need to verify correctness of rule for off diagonal case"
        ds = [diag(M, k, alg) for M in A.Ms]
        out = None
        for i, d in enumerate(ds):
            d = d.reshape([-1 if j == i else 1 for j in range(len(ds))])
            out = d if out is None else (out * d if isinstance(A, Kronecker) else out + d)
        return out.reshape(-1)
    if isinstance(alg, Auto):   # diag_trace.py:43-50
        tol = alg.__dict__.get("tol", 1e-6)
        use_exact = bool(tol < 1 / np.sqrt(10 * np.prod(A.shape)))
        alg = Exact(**{kk: v for kk, v in alg.__dict__.items() if kk in ("bs", "pbar")}) if use_exact else \
            Hutch(**alg.__dict__)
    return alg(A, k)


def trace(A: LinearOperator, alg: Algorithm = Auto()):
    """cola/linalg/trace/diag_trace.py:122-147"""
    assert A.shape[0] == A.shape[1], "Can't trace non square matrix"
    if isinstance(A, Kronecker):
        out = 1.0
        for M in A.Ms:
            out = out * trace(M, alg)
        return out
    return diag(A, 0, alg).sum()


# ---------------------------------------------------------------------------------------------------- f(A)
def apply_unary(f, A: LinearOperator, alg: Algorithm = Auto()):
    """cola/linalg/unary/unary.py:94-212: structure rules first, then Auto / Lanczos / Arnoldi / Eigh."""
    if isinstance(A, Diagonal):                      # unary.py:181-183
        return Diagonal(f(A.diag))
    if isinstance(A, BlockDiag):                     # :186-189
        return BlockDiag(*[apply_unary(f, a, alg) for a in A.Ms], multiplicities=A.multiplicities)
    if isinstance(A, Identity):                      # :192-195
        return f(torch.ones((), dtype=A.dtype, device=A.device)) * A
    if isinstance(A, ScalarMul):                     # :198-200
        return f(A.c) * I_like(A)
    if isinstance(A, Transpose):                     # :203-205
        return Transpose(apply_unary(f, A.A, alg))
    if isinstance(alg, Auto):                        # :113-131
        psd, small = A.isa(PSD), bool(np.prod(A.shape) <= 1e6)
        if psd:
            alg = Eigh() if small else Lanczos(**alg.__dict__)
        elif small:
            raise NotImplementedError("small non-PSD operators route to a dense complex eig in the reference "
                                      "(unary.py:171-178), which is outside the Krylov hot path")
        else:
            alg = Arnoldi(**alg.__dict__)
    if isinstance(alg, Lanczos):                     # :134-137
        assert A.isa(SelfAdjoint), "Lanczos only valid for SelfAdjoint, wrap in cola.SelfAdjoint if desired"
        return LanczosUnary(A, f, **alg.__dict__)
    if isinstance(alg, Arnoldi):                     # :140-142
        return ArnoldiUnary(A, f, **alg.__dict__)
    if isinstance(alg, Eigh):                        # :160-168: dense eigh (library), lazy V f(D) V^T
        assert A.isa(SelfAdjoint), "Eigh only valid for SelfAdjoint, wrap in cola.SelfAdjoint if desired"
        eigs, V = torch.linalg.eigh(A.to_dense())
        V = lazify(V)
        return V @ Diagonal(f(eigs)) @ V.H
    raise NotImplementedError(f"apply_unary with {type(alg).__name__} is outside the Krylov hot path")


def exp(A: LinearOperator, alg: Algorithm = None):
    """cola/linalg/unary/unary.py:229-246.  The KronSum rule is declared on (KronSum, Algorithm) without a default,
    so it only applies when an algorithm is passed; `exp(A)` is the generic rule with Auto."""
    if alg is None:
        return apply_unary(torch.exp, A, Auto())
    if isinstance(A, KronSum):                       # exp(A (+) B) = exp(A) (x) exp(B)
        return Kronecker(*[exp(a, alg) for a in A.Ms])
    return apply_unary(torch.exp, A, alg)


def log(A: LinearOperator, alg: Algorithm = Auto()):
    """cola/linalg/unary/unary.py:249-262"""
    return apply_unary(torch.log, A, alg)


def pow(A: LinearOperator, alpha, alg: Algorithm = None):
    """cola/linalg/unary/unary.py:265-305: integer powers are products / inverses, the rest is f(A) = A^alpha.  The
    Kronecker rule (:303-305) is declared without a default algorithm: it applies to the three-argument call only
    (which is what sqrt / isqrt make)."""
    if alg is None:
        alg = Auto()
    elif isinstance(A, Kronecker):
        return Kronecker(*[pow(a, alpha, alg) for a in A.Ms])
    k = int(np.round(alpha))
    if np.isclose(alpha, k):
        if k == 0:
            return I_like(A)
        if 0 < k < 10:
            out = A
            for _ in range(k - 1):
                out = out @ A
            return out
        if k == -1:
            if isinstance(alg, Lanczos):
                new_alg = CG(**{kk: v for kk, v in alg.__dict__.items() if kk in ("tol", "max_iters", "pbar")})
            elif isinstance(alg, Arnoldi):
                new_alg = GMRES(**{kk: v for kk, v in alg.__dict__.items() if kk in ("tol", "max_iters", "pbar")})
            elif isinstance(alg, Eigh):
                new_alg = Cholesky()
            else:
                new_alg = alg
            return inv(A, new_alg)
    return apply_unary(lambda x: x**alpha, A, alg)


def sqrt(A: LinearOperator, alg: Algorithm = Auto()):
    """cola/linalg/unary/unary.py:308-320"""
    return pow(A, 0.5, alg)


def isqrt(A: LinearOperator, alg: Algorithm = Auto()):
    """cola/linalg/unary/unary.py:323-335"""
    return pow(A, -0.5, alg)


def slogdet(A: LinearOperator, log_alg: Algorithm = Auto(), trace_alg: Algorithm = Auto()):
    """cola/linalg/logdet/logdet.py:52-184: structure rules, then trace(log(A, Lanczos), trace_alg)."""
    one = torch.ones((), dtype=A.dtype, device=A.device)
    if isinstance(A, Identity):
        return one, torch.zeros((), dtype=A.dtype, device=A.device)
    if isinstance(A, Diagonal):
        return torch.prod(torch.sign(A.diag)), torch.sum(torch.log(torch.abs(A.diag)))
    if type(A) is Product and all(M.shape[-2] == M.shape[-1] for M in A.Ms):   # logdet.py:121-125
        signs, logdets = zip(*[slogdet(M, log_alg, trace_alg) for M in A.Ms])
        sg = one
        for sgn in signs:
            sg = sg * sgn
        return sg, sum(logdets)
    if isinstance(A, ScalarMul):   # logdet.py:134-139, restated as it is: log|c|, not n log|c|
        return A.c / torch.abs(A.c), torch.log(torch.abs(A.c))
    if isinstance(A, Triangular):  # logdet.py:170-176
        d = torch.diagonal(A.A)
        return torch.prod(d / torch.abs(d)), torch.sum(torch.log(torch.abs(d)))
    if isinstance(A, Kronecker):   # logdet.py:147-160
        n = A.shape[0]
        signs, logdets = zip(*[slogdet(M, log_alg, trace_alg) for M in A.Ms])
        ld = sum(l * (n // M.shape[-1]) for l, M in zip(logdets, A.Ms))
        sg = one
        for s, M in zip(signs, A.Ms):
            sg = sg * s**(n // M.shape[-1])
        return sg, ld
    if isinstance(A, BlockDiag):   # logdet.py:163-170
        signs, logdets = zip(*[slogdet(M, log_alg, trace_alg) for M in A.Ms])
        ld = sum(l * c for l, c in zip(logdets, A.multiplicities))
        sg = one
        for s, c in zip(signs, A.multiplicities):
            sg = sg * s**c
        return sg, ld
    if isinstance(log_alg, Auto):   # logdet.py:82-94
        small = bool(np.prod(A.shape) <= 1e6)
        if small:
            s, l = torch.linalg.slogdet(A.to_dense())
            return s, l
        log_alg = Lanczos(**log_alg.__dict__) if A.isa(PSD) else Arnoldi(**log_alg.__dict__)
    if isinstance(log_alg, (Lanczos, Arnoldi)):   # logdet.py:111-117
        tr = trace(log(A, log_alg), trace_alg)
        mag = torch.abs(tr)                        # the reference returns (phase, |tr log A|), restated as is
        return tr / mag, mag
    raise NotImplementedError(f"slogdet with {type(log_alg).__name__} is outside the Krylov hot path")


def logdet(A: LinearOperator, log_alg: Algorithm = Auto(), trace_alg: Algorithm = Auto()):
    """cola/linalg/logdet/logdet.py:30-49"""
    _, ld = slogdet(A, log_alg, trace_alg)
    return ld
