"""Structured linear operators with the reference's interface (cola/ops/operator_base.py, cola/ops/operators.py)
whose `_matmat` runs on the hand-written sm_100a kernels.

Same class names, constructor arguments, shapes and error behaviour as the reference for the operators on the
Krylov hot path: Dense, Sparse, Kronecker, BlockDiag, Diagonal, ScalarMul, Identity, Sum, Product (+ Tridiagonal
as the container Lanczos returns).  What differs is underneath: instead of one eager torch op (and one
temporary) per node of the operator tree, `compile_plan` flattens the tree into
        A X = sum_t  scale_t * K_t X   +  (shift + diag) o X
and executes it as one fused kernel per core operator K_t (shift / Diagonal / pAp-dots ride in its epilogue).

CUDA float32/float64 only; CPU tensors raise (there is no fallback path).
"""
import copy
import os
import ctypes
from functools import reduce
from numbers import Number

import numpy as np
import torch

from . import backend as be


# ----------------------------------------------------------------------------------------------------
# annotations (cola/annotations.py:26-74)
# ----------------------------------------------------------------------------------------------------
class _WrapMeta(type):
    def __str__(cls):
        return cls.__name__

    __repr__ = __str__

    def __call__(cls, obj):
        new = copy.copy(obj)
        new.annotations = set(obj.annotations) | {cls}
        new._plan = None
        return new


class Annotation(metaclass=_WrapMeta):
    pass


class SelfAdjoint(Annotation):
    pass


Hermitian = SelfAdjoint


class PSD(SelfAdjoint):
    pass


class Stiefel(Annotation):
    pass


class Unitary(Stiefel):
    pass


def _intersect(ops):
    return reduce(lambda x, y: x & y, (set(op.annotations) for op in ops))


# ----------------------------------------------------------------------------------------------------
# base class (cola/ops/operator_base.py:16-219)
# ----------------------------------------------------------------------------------------------------
def _find_device(obj):
    if torch.is_tensor(obj):
        return obj.device
    if isinstance(obj, LinearOperator):
        return obj.device
    if isinstance(obj, (tuple, list)):
        for o in obj:
            d = _find_device(o)
            if d is not None:
                return d
    if isinstance(obj, dict):
        for o in obj.values():
            d = _find_device(o)
            if d is not None:
                return d
    return None


class LinearOperator:
    """Linear operator base class: `A @ X` -> `_matmat(X)` with X (d, k) (operator_base.py:97-106)."""
    __array_ufunc__ = None

    def __new__(cls, *args, **kwargs):
        obj = super().__new__(cls)
        obj.device = _find_device([args, kwargs])
        obj._plan = None
        return obj

    def __init__(self, dtype, shape, matmat=None, annotations=()):
        self.dtype = dtype
        self.shape = tuple(shape)
        if matmat is not None:
            self._matmat = matmat
        self.annotations = self._infer_annotations() | set(annotations)
        self.device = self.device or torch.device("cpu")

    def _infer_annotations(self):  # cola/annotations.py:80-193 (rules for the hot-path operators)
        return set()

    def isa(self, annotation):
        return any(issubclass(a, annotation) for a in self.annotations)

    # ---- application ---------------------------------------------------------------------------------
    def plan(self):
        """The compiled plan, recompiled when a leaf tensor was replaced or written in place since the last
        compile (an optimizer step on a parameter): plans hold derived copies (scaled diagonals, folded scalars,
        densified factors, the CSR form of a Tridiagonal) that would otherwise go stale.  The Krylov loops validate
        once at entry and then use `matmat_into`, which does not re-check."""
        token = _version_token(self)
        if torch.is_grad_enabled() and _token_requires_grad(token):
            raise RuntimeError("a parameter of this operator requires grad, but cola_b200's native API is not "
                               "differentiable: detach it / use torch.no_grad(), or keep the reference's operators and "
                               "call cola_b200.install() (the reference's backward rules then run with the solves on "
                               "the kernels)")
        if self._plan is None or self.__dict__.get("_plan_token") != token:
            self._plan = compile_plan(self)
            self._plan_token = token
            self.__dict__.pop("_cg_workspace", None)   # a captured CG batch points at the old plan's buffers
        return self._plan

    def _matmat(self, X):
        X = _as_operand(self, X)
        Y = torch.empty((self.shape[0], X.shape[1]), dtype=X.dtype, device=X.device)
        self.plan().apply(X, Y)
        return Y

    def _rmatmat(self, X):
        if self.isa(SelfAdjoint):
            return self._matmat(X.T.contiguous()).T
        return (self.T._matmat(X.T.contiguous())).T

    def matmat_into(self, X, Y, dots=None, dots_row=None, gate=None):
        """Fused form used by the Krylov loops: Y = A X and, if given, dots[row] += colsum(X * Y) in the same
        kernel; `gate`/`dots_row` are device int32 scalars (see include/cola_b200.h)."""
        (self._plan or self.plan()).apply(X, Y, dots=dots, dots_row=dots_row, gate=gate)

    def __matmul__(self, X):
        assert X.shape[0] == self.shape[-1], f"dimension mismatch {self.shape} vs {X.shape}"
        if isinstance(X, LinearOperator):
            return dot(self, X)
        if len(X.shape) == 1:
            return self._matmat(X.reshape(-1, 1)).reshape(-1)
        if len(X.shape) >= 2:
            return self._matmat(X)
        raise NotImplementedError

    def __rmatmul__(self, X):
        assert X.shape[-1] == self.shape[-2], f"dimension mismatch {self.shape} vs {X.shape}"
        if isinstance(X, LinearOperator):
            return dot(X, self)
        if len(X.shape) == 1:
            return self._rmatmat(X.reshape(1, -1)).reshape(-1)
        return self._rmatmat(X)

    def to_dense(self):
        eye = torch.eye(self.shape[-1], dtype=self.dtype, device=self.device)
        return self @ eye

    @property
    def T(self):
        return transpose(self)

    @property
    def H(self):
        return transpose(self)  # real dtypes only on this path

    def to(self, device, dtype=None):
        new = copy.copy(self)
        new._plan = None
        for key, val in vars(self).items():
            if torch.is_tensor(val):
                setattr(new, key, val.to(device=device, dtype=dtype if val.is_floating_point() else None))
            elif isinstance(val, LinearOperator):
                setattr(new, key, val.to(device, dtype))
            elif isinstance(val, tuple) and val and all(isinstance(v, LinearOperator) for v in val):
                setattr(new, key, tuple(v.to(device, dtype) for v in val))
        new.device = torch.device(device)
        return new

    # ---- algebra (cola/fns.py:63-137) -----------------------------------------------------------------
    def __add__(self, other):
        if isinstance(other, Number) and other == 0:
            return self
        return add(self, other)

    __radd__ = __add__

    def __mul__(self, c):
        return mul(self, c)

    __rmul__ = __mul__

    def __neg__(self):
        return -1 * self

    def __sub__(self, x):
        return self.__add__(-x)

    def __truediv__(self, x):
        return self.__mul__(1 / x)

    def __str__(self):
        return self.__class__.__name__

    def __repr__(self):
        return "<%dx%d %s with dtype=%s>" % (self.shape[0], self.shape[1], self.__class__.__name__, self.dtype)


def _version_token(op):
    """(id, in-place version counter, requires_grad) of every tensor in the operator tree."""
    tok = []
    for val in vars(op).values():
        if torch.is_tensor(val):
            tok.append((id(val), val._version, bool(val.requires_grad)))
        elif isinstance(val, LinearOperator):
            tok.append(_version_token(val))
        elif isinstance(val, (tuple, list)) and val and all(isinstance(v, LinearOperator) for v in val):
            tok.append(tuple(_version_token(v) for v in val))
    return tuple(tok)


def _token_requires_grad(tok):
    for item in tok:
        if isinstance(item, tuple):
            if len(item) == 3 and isinstance(item[2], bool) and not isinstance(item[0], tuple):
                if item[2]:
                    return True
            elif _token_requires_grad(item):
                return True
    return False


def _as_operand(A, X):
    if not torch.is_tensor(X):
        raise TypeError("operand must be a torch tensor")
    be.require_cuda(X, "operand")
    dt = torch.promote_types(A.dtype, X.dtype)
    if dt != A.dtype:
        raise TypeError(f"operand dtype {X.dtype} does not match operator dtype {A.dtype}")
    X = X.to(dt)
    return X if X.is_contiguous() else X.contiguous()


def lazify(A):
    return A if isinstance(A, LinearOperator) else Dense(A)


# ----------------------------------------------------------------------------------------------------
# operators on the hot path (cola/ops/operators.py)
# ----------------------------------------------------------------------------------------------------
class Dense(LinearOperator):
    """operators.py:12-38"""
    def __init__(self, A):
        self.A = A
        super().__init__(dtype=A.dtype, shape=A.shape)

    def to_dense(self):
        return self.A


class Triangular(Dense):
    """operators.py:41-45"""
    def __init__(self, A, lower=True):
        super().__init__(A)
        self.lower = lower


class Sparse(LinearOperator):
    """operators.py:48-81.  Same COO constructor; the CSR arrays (int32 indices) are built on the device with a
    STABLE row sort, i.e. what the reference constructor means (its own non-stable argsort can misalign
    values and indices).  No scipy / host round trip."""
    def __init__(self, data, row_indices, col_indices, shape):
        super().__init__(dtype=data.dtype, shape=shape)
        row_indices = row_indices.to(data.device)
        col_indices = col_indices.to(data.device)
        order = torch.argsort(row_indices.to(torch.int64), stable=True)
        self.data = data[order].contiguous()
        self.row_indices = row_indices[order]
        self.col_indices = col_indices[order]
        counts = torch.bincount(self.row_indices.to(torch.int64), minlength=shape[0])
        rowptr = torch.zeros(shape[0] + 1, dtype=torch.int64, device=data.device)
        rowptr[1:] = torch.cumsum(counts, 0)
        assert int(self.data.numel()) < 2**31, "int32 CSR indices (operators.py:73-74)"
        self.indptr = rowptr.to(torch.int32).contiguous()
        self.indices = self.col_indices.to(torch.int32).contiguous()
        self.nnz = int(self.data.numel())
        self.max_row_nnz = int(counts.max()) if counts.numel() > 0 else 0

    @classmethod
    def from_csr(cls, indptr, indices, data, shape):
        """Wrap existing CSR arrays (e.g. those of the reference's `Sparse.A`, operators.py:71-75) without
        re-sorting: the operator is exactly the one those arrays describe."""
        self = cls.__new__(cls, data)
        LinearOperator.__init__(self, dtype=data.dtype, shape=shape)
        self.indptr = indptr.to(torch.int32).contiguous()
        self.indices = indices.to(torch.int32).contiguous()
        self.data = data.contiguous()
        counts = (self.indptr[1:] - self.indptr[:-1]).to(torch.int64)
        self.row_indices = torch.repeat_interleave(torch.arange(shape[0], device=data.device), counts)
        self.col_indices = self.indices
        self.nnz = int(self.data.numel())
        self.max_row_nnz = int(counts.max()) if counts.numel() > 0 else 0
        return self

    def _transpose(self):
        return Sparse(self.data, self.col_indices, self.row_indices, (self.shape[1], self.shape[0]))


class ScalarMul(LinearOperator):
    """operators.py:84-101"""
    def __init__(self, c, shape, dtype=None, device=None):
        super().__init__(dtype=dtype or type(c), shape=shape)
        self.c = torch.as_tensor(c, dtype=dtype, device=device)
        self.device = device if device is not None else self.device

    def __str__(self):
        return f"{self.c}"


class Identity(LinearOperator):
    """operators.py:104-127"""
    def __init__(self, shape, dtype):
        super().__init__(dtype=dtype, shape=shape)

    def _infer_annotations(self):
        return {Unitary, PSD}

    def _matmat(self, X):
        return X

    def to(self, device, dtype=None):
        self.device = torch.device(device) if not isinstance(device, torch.device) else device
        return self

    def __str__(self):
        return "I"


def I_like(A):
    op = Identity(dtype=A.dtype, shape=A.shape)
    op.to(A.device)
    return op


class Product(LinearOperator):
    """operators.py:138-164"""
    def __init__(self, *Ms):
        self.Ms = tuple(lazify(M) for M in Ms)
        devices = [M.device for M in self.Ms]
        assert all(x == devices[0] for x in devices), "There is a device mismatch in Product"
        for M1, M2 in zip(Ms[:-1], Ms[1:]):
            if M1.shape[-1] != M2.shape[-2]:
                raise ValueError(f"dimension mismatch {M1.shape} vs {M2.shape}")
        shape = (Ms[0].shape[-2], Ms[-1].shape[-1])
        dtype = reduce(torch.promote_types, (M.dtype for M in self.Ms))
        super().__init__(dtype, shape)
        self.device = devices[0]

    def _infer_annotations(self):  # annotations.py:115-122
        not_commuting = [M for M in self.Ms if not isinstance(M, ScalarMul)]
        if len(not_commuting) == 1:
            return set(not_commuting[0].annotations)
        return _intersect(self.Ms) & {Unitary, Stiefel}

    def __str__(self):
        return "".join(str(M) for M in self.Ms)


class Sum(LinearOperator):
    """operators.py:167-191"""
    def __init__(self, *Ms):
        self.Ms = tuple(lazify(M) for M in Ms)
        devices = [M.device for M in self.Ms]
        assert all(x == devices[0] for x in devices), "There is a device mismatch in Sum"
        shape = Ms[0].shape
        for M in Ms:
            if tuple(M.shape) != tuple(shape):
                raise ValueError(f"dimension mismatch {M.shape} vs {shape}")
        super().__init__(Ms[0].dtype, shape)
        self.device = devices[0]

    def _infer_annotations(self):  # annotations.py:130-132
        return _intersect(self.Ms) - {Unitary, Stiefel}

    def __str__(self):
        return "+".join(str(M) for M in self.Ms)


def _prod(c):
    return reduce(lambda a, b: a * b, c)


class Kronecker(LinearOperator):
    """operators.py:198-230"""
    def __init__(self, *Ms):
        self.Ms = tuple(lazify(M) for M in Ms)
        shape = _prod([Mi.shape[-2] for Mi in Ms]), _prod([Mi.shape[-1] for Mi in Ms])
        dtype = reduce(torch.promote_types, (M.dtype for M in self.Ms))
        super().__init__(dtype, shape)

    def _infer_annotations(self):  # annotations.py:91-93
        return _intersect(self.Ms)

    def to_dense(self):
        return reduce(torch.kron, [M.to_dense() for M in self.Ms])

    def __str__(self):
        return "⊗".join(str(M) for M in self.Ms)


class KronSum(LinearOperator):
    """operators.py:241-275: A (+) B (+) ... = A x I x .. + I x B x .. + ...  (square factors)."""
    def __init__(self, *Ms):
        self.Ms = tuple(lazify(M) for M in Ms)
        shape = _prod([Mi.shape[-2] for Mi in Ms]), _prod([Mi.shape[-1] for Mi in Ms])
        dtype = reduce(torch.promote_types, (M.dtype for M in self.Ms))
        super().__init__(dtype, shape)

    def _infer_annotations(self):  # the reference has no inference rule for KronSum (annotations.py:80-193)
        return set()

    def to_dense(self):
        def ksum(A, B):
            IA = torch.eye(A.shape[-2], dtype=A.dtype, device=A.device)
            IB = torch.eye(B.shape[-2], dtype=B.dtype, device=B.device)
            return torch.kron(A, IB) + torch.kron(IA, B)
        return reduce(ksum, [M.to_dense() for M in self.Ms])

    def __str__(self):
        return "⊕ₖ".join(str(M) for M in self.Ms)


class BlockDiag(LinearOperator):
    """operators.py:277-320"""
    def __init__(self, *Ms, multiplicities=None):
        self.Ms = tuple(lazify(M) for M in Ms)
        self.multiplicities = [1 for _ in Ms] if multiplicities is None else multiplicities
        shape = (sum(Mi.shape[-2] * c for Mi, c in zip(Ms, self.multiplicities)),
                 sum(Mi.shape[-1] * c for Mi, c in zip(Ms, self.multiplicities)))
        dtype = reduce(torch.promote_types, (M.dtype for M in self.Ms))
        super().__init__(dtype, shape)

    def _infer_annotations(self):  # annotations.py:135-137
        return _intersect(self.Ms)

    def to_dense(self):
        blocks = [M.to_dense() for M, c in zip(self.Ms, self.multiplicities) for _ in range(c)]
        return torch.block_diag(*blocks)


class Diagonal(LinearOperator):
    """operators.py:323-348"""
    def __init__(self, diag):
        assert len(diag.shape) == 1, f"diagonal is not a vector, it is of shape {diag.shape=}"
        self.diag = diag
        super().__init__(dtype=diag.dtype, shape=(len(diag), ) * 2)

    def to_dense(self):
        return torch.diag(self.diag)

    def __str__(self):
        return f"diag({self.diag})"


class Tridiagonal(LinearOperator):
    """operators.py:351-372 (alpha lower, beta diagonal, gamma upper band): the container Lanczos returns, and an
    operator in its own right.  With plain vector bands its matmat is the CSR kernel on a 3-entries-per-row matrix
    (`as_sparse`), so shift / Diagonal / dots epilogues fuse as for any Sparse; batched bands (what a batched
    Lanczos returns) keep the elementwise form below."""
    def __init__(self, alpha, beta, gamma):
        def col(v):
            return v.reshape(-1, 1) if v.dim() == 1 else v
        self.alpha, self.beta, self.gamma = col(alpha), col(beta), col(gamma)
        super().__init__(dtype=beta.dtype, shape=(self.beta.shape[0], self.beta.shape[0]))

    def _matmat(self, X):
        if self.beta.dim() == 2 and self.beta.shape[1] == 1:
            return LinearOperator._matmat(self, X)         # compiled plan: CSR core
        out = self.beta * X
        zeros = torch.zeros((1, X.shape[-1]), dtype=X.dtype, device=X.device)
        up = torch.cat([self.gamma * X[1:], zeros], dim=0)
        lo = torch.cat([zeros, self.alpha * X[:-1]], dim=0)
        return out + lo + up

    def as_sparse(self):
        n = self.beta.shape[0]
        dev = self.beta.device
        i = torch.arange(n, device=dev)
        rows = torch.cat([i, i[1:], i[:-1]])
        cols = torch.cat([i, i[:-1], i[1:]])
        data = torch.cat([self.beta[:, 0], self.alpha[:, 0], self.gamma[:, 0]])
        order = torch.argsort(rows * n + cols)
        return Sparse(data[order].contiguous(), rows[order], cols[order], (n, n))

    def to_dense(self):
        m = self.beta.shape[0]
        T = torch.diag(self.beta[:, 0])
        if m > 1:
            T = T + torch.diag(self.alpha[:, 0], -1) + torch.diag(self.gamma[:, 0], 1)
        return T


class Transpose(LinearOperator):
    """operators.py:381-396.  Lazy, like the reference's (`.T` of a Sum / Product / Kronecker / ... is a
    `Transpose[...]` there, and dispatch rules see it as such); what is applied is the transpose pushed down to
    the leaves (`_push_transpose`: transposed Dense / CSR cores, reversed Products), so it runs on the same fused
    kernels as the operator itself."""
    def __init__(self, A):
        super().__init__(dtype=A.dtype, shape=(A.shape[1], A.shape[0]))
        self.A = A
        self.device = A.device

    def _infer_annotations(self):
        return set()

    def pushed(self):
        """The operator tree of A^T, or None when A is an opaque operator without a transpose of its own."""
        token = _version_token(self.A)
        cached = self.__dict__.get("_pushed_cache")
        if cached is None or cached[0] != token:               # transposed leaf copies follow in-place updates
            P = _push_transpose(self.A)
            cached = (token, None if isinstance(P, Transpose) else P)
            self.__dict__["_pushed_cache"] = cached
        return cached[1]

    def _matmat(self, X):
        if self.pushed() is None:
            if type(self.A)._rmatmat is LinearOperator._rmatmat:
                raise NotImplementedError(f"{type(self.A).__name__} has no transposed matmat on this path")
            return self.A._rmatmat(X.T.contiguous()).T.contiguous()
        return LinearOperator._matmat(self, X)                 # plan of the pushed tree

    def _rmatmat(self, X):
        return self.A._matmat(X.T.contiguous()).T

    def __str__(self):
        return f"{str(self.A)}ᵀ"


class Sliced(LinearOperator):
    """operators.py:417-453, column/row slices of a lazy operator (what eig() returns, eigs.py:111)."""
    def __init__(self, A, slices):
        self.A, self.slices = A, slices
        rows = range(A.shape[0])[slices[0]]
        cols = range(A.shape[1])[slices[1]]
        super().__init__(dtype=A.dtype, shape=(len(rows), len(cols)))

    def to_dense(self):
        return self.A.to_dense()[self.slices[0], :][:, self.slices[1]]

    def _matmat(self, X):
        full = torch.zeros((self.A.shape[1], X.shape[1]), dtype=X.dtype, device=X.device)
        full[self.slices[1]] = X
        return (self.A @ full)[self.slices[0]]


def _getitem(self, ids):
    if isinstance(ids, tuple) and len(ids) == 2 and all(isinstance(s, slice) for s in ids):
        return Sliced(self, ids)
    raise NotImplementedError(f"__getitem__ not implemented for {type(ids)}")


LinearOperator.__getitem__ = _getitem


# ----------------------------------------------------------------------------------------------------
# algebra dispatch (cola/fns.py)
# ----------------------------------------------------------------------------------------------------
def dot(A, B):
    if isinstance(B, Identity):
        return A
    if isinstance(A, Identity):
        return B
    a = A.Ms if isinstance(A, Product) else (A, )
    b = B.Ms if isinstance(B, Product) else (B, )
    return Product(*(a + b))


def add(A, B):
    A, B = lazify(A), lazify(B)
    a = A.Ms if type(A) is Sum else (A, )
    b = B.Ms if type(B) is Sum else (B, )
    return Sum(*(a + b))


def mul(A, c):
    if isinstance(A, ScalarMul):
        if isinstance(c, ScalarMul):
            return ScalarMul(A.c * c.c, A.shape, A.dtype, A.device)
        return ScalarMul(A.c * c, A.shape, A.dtype, A.device)
    S = ScalarMul(c, (A.shape[-2], A.shape[-2]), A.dtype, A.device)
    return Product(S, A)


def transpose(A):
    """cola/fns.py:140-168: SelfAdjoint -> itself, Transpose -> its operand, Dense / Triangular / Sparse -> the
    transposed leaf, everything else -> a lazy Transpose."""
    if A.isa(SelfAdjoint):
        return A
    if isinstance(A, Transpose):
        return A.A
    if isinstance(A, Triangular):
        return Triangular(A.A.T.contiguous(), lower=not A.lower)
    if isinstance(A, Dense):
        return Dense(A.A.T.contiguous())
    if isinstance(A, Sparse):
        return A._transpose()
    return Transpose(A)


def _push_transpose(A):
    """A^T as a tree of hot-path operators: the transpose distributed over Sum / Product / Kronecker / KronSum /
    BlockDiag / Tridiagonal down to transposed leaves.  Returns Transpose(A) for operators it cannot open."""
    if A.isa(SelfAdjoint):
        return A
    if isinstance(A, Transpose):
        return A.A
    if isinstance(A, Triangular):
        return Triangular(A.A.T.contiguous(), lower=not A.lower)
    if isinstance(A, Dense):
        return Dense(A.A.T.contiguous())
    if isinstance(A, Sparse):
        return A._transpose()
    if isinstance(A, (Diagonal, Identity, ScalarMul)):
        return A
    if isinstance(A, Kronecker):
        return Kronecker(*[_push_transpose(M) for M in A.Ms])
    if isinstance(A, KronSum):
        return KronSum(*[_push_transpose(M) for M in A.Ms])
    if isinstance(A, Tridiagonal):
        return Tridiagonal(A.gamma, A.beta, A.alpha)
    if isinstance(A, BlockDiag):
        return BlockDiag(*[_push_transpose(M) for M in A.Ms], multiplicities=A.multiplicities)
    if type(A) is Sum:
        return Sum(*[_push_transpose(M) for M in A.Ms])
    if type(A) is Product:
        return Product(*[_push_transpose(M) for M in reversed(A.Ms)])
    return Transpose(A)


def kron(A, B):
    a = A.Ms if isinstance(A, Kronecker) else (lazify(A), )
    b = B.Ms if isinstance(B, Kronecker) else (lazify(B), )
    return Kronecker(*(a + b))


def kronsum(A, B):
    """cola/fns.py:224-242"""
    a = A.Ms if isinstance(A, KronSum) else (lazify(A), )
    b = B.Ms if isinstance(B, KronSum) else (lazify(B), )
    return KronSum(*(a + b))


def densify(A):
    """cola/fns.py:41-47"""
    return A.to_dense() if isinstance(A, LinearOperator) else A


def block_diag(*ops, multiplicities=None):
    return BlockDiag(*ops, multiplicities=multiplicities)


# ----------------------------------------------------------------------------------------------------
# operator compiler: tree -> flat plan -> fused kernels
# ----------------------------------------------------------------------------------------------------
class _DenseCore:
    def __init__(self, M):
        self.M = M if M.is_contiguous() else M.contiguous()
        self.shape = tuple(M.shape)

    def apply(self, X, Y, epi):
        k = X.shape[1]
        be.mode_contract(self.M, self.shape[0], self.shape[1], 1, k, X, Y, epi_x=X if epi.needs_x() else None,
                         **epi.kw())


class _CsrCore:
    """CSR SpMM.  SpMV-shaped applications (k <= 4) of a matrix whose column indices have no locality (a random graph,
    BASELINE config 5: every non-zero pulls a 32-byte DRAM sector for 8 useful bytes once the vector outgrows L2) run
    COLUMN-BLOCKED: the pattern is cut once into vertical strips whose slice of X (SPMV_BLOCK_BYTES) stays resident in
    L2, and the strips are applied one after the other, each accumulating into Y; the CSR arrays still stream through
    once in total, Y is re-read per strip.  Row sums are then taken strip by strip (fp rounding order only)."""
    TILE_STAGE_BYTES = int(os.environ.get("COLA_SPMM_TILE_KB", "106")) << 10    # X rows one ring stage of the staged kernel holds (0: off)
    TILE_STRIP_ROWS = int(os.environ.get("COLA_SPMM_TILE_R", "0"))               # 0: from the stage size
    TILE_SMEM_BYTES = 227 << 10                                                  # opt-in shared memory of one sm_100 CTA
    SPMV_BLOCK_BYTES = int(os.environ.get("COLA_SPMV_BLOCK_MB", "45")) << 20   # measured on cfg5 (2^24 nodes, fp64): 32 MB 2.53 ms, 45 MB 2.18 ms, 68 MB 2.89 ms; plain 3.99; cuSPARSE 2.97

    def __init__(self, S):
        self.S = S
        self.shape = tuple(S.shape)
        self._strips = {}
        self._tiled = {}

    def _tiles(self, X, Y):
        """The tile-local form for the staged kernel (csrc/csr_tiled.cu), or None: rows of X of 128 B .. 4 KB, contiguous
        blocks, a pattern whose tiles gather long column runs (stencil / banded: BASELINE config 2)."""
        row_bytes = X.shape[1] * X.element_size()
        if (self.TILE_STAGE_BYTES <= 0 or row_bytes % 16 or not 128 <= row_bytes <= 4096 or not X.is_contiguous()
                or not Y.is_contiguous() or self.S.nnz == 0):
            return None
        if row_bytes not in self._tiled:
            from .csr_tiles import CsrTiles, STRIPS, banded_like
            if not banded_like(self.S.indices, self.S.row_indices, self.S.nnz):
                self._tiled[row_bytes] = None
                return None
            # one ring stage holds, per tile row: ~1.33 staged rows of X (5-point stencil halo), its non-zeros' offsets
            # and values, two row-pointer words; if the halo turns out wider, halve the strips once or twice
            nz_row = 1.05 * self.S.nnz / max(self.S.shape[0], 1) * (4 + X.element_size()) + 8
            R = self.TILE_STRIP_ROWS
            if R <= 0:
                R = 8
                while 2 * R * STRIPS * (1.33 * row_bytes + nz_row) <= self.TILE_STAGE_BYTES and R < 64:
                    R *= 2
            found = None
            while found is None and R >= 8:
                cap_rows = int(self.TILE_STAGE_BYTES - R * STRIPS * nz_row) // row_bytes
                T = CsrTiles(self.S, R, cap_rows, row_bytes)
                fits = 2 * T.stage_bytes(X.element_size()) + X.shape[1] * 8 + 1024 <= self.TILE_SMEM_BYTES   # a ring of two stages
                if T.worthwhile() and fits:
                    found = T
                R //= 2
            self._tiled[row_bytes] = found
        return self._tiled[row_bytes]

    def _column_strips(self, k, itemsize):
        """None (plain kernel), or [(indptr, indices, data, nnz)] per vertical strip, built on first use."""
        key = (k, itemsize)
        if key in self._strips:
            return self._strips[key]
        S = self.S
        strips = None
        x_bytes = S.shape[1] * k * itemsize
        if self.SPMV_BLOCK_BYTES > 0 and x_bytes > 1.5 * self.SPMV_BLOCK_BYTES and S.nnz > 0:
            rows = S.row_indices.to(torch.int64)
            span = float((S.indices.to(torch.int64) - rows).abs().double().mean()) * k * itemsize
            if span > self.SPMV_BLOCK_BYTES / 4:                # the gather window of a row does not fit L2 anyway
                n_strips = -(-x_bytes // self.SPMV_BLOCK_BYTES)
                width = -(-S.shape[1] // n_strips)
                strip_of = torch.div(S.indices, width, rounding_mode="floor")
                strips = []
                for b in range(n_strips):
                    idx = (strip_of == b).nonzero().reshape(-1)   # ascending: row-major order is kept
                    counts = torch.bincount(rows[idx], minlength=S.shape[0])
                    indptr = torch.zeros(S.shape[0] + 1, dtype=torch.int64, device=idx.device)
                    indptr[1:] = torch.cumsum(counts, 0)
                    strips.append((indptr.to(torch.int32).contiguous(), S.indices[idx].contiguous(), S.data[idx].contiguous(),
                                   int(idx.numel())))
                del strip_of
            del rows
        self._strips[key] = strips
        return strips

    def apply(self, X, Y, epi):
        S = self.S
        strips = self._column_strips(X.shape[1], X.element_size()) if X.shape[1] <= 4 else None
        if strips is None:
            T = self._tiles(X, Y)
            if T is not None:
                return be.csr_spmm_tiled(T, T.values(S.data), S.shape, X, Y, **epi.kw())
            return be.csr_spmm(S.indptr, S.indices, S.data, S.shape, S.nnz, S.max_row_nnz, X, Y, **epi.kw())
        last = len(strips) - 1
        for b, (indptr, indices, data, nnz) in enumerate(strips):
            if b < last:       # partial row sums: Y = alpha * A_b X (+ Y)
                be.csr_spmm(indptr, indices, data, S.shape, nnz, S.max_row_nnz, X, Y, alpha=epi.alpha,
                            accumulate=epi.accumulate or b > 0, gate=epi.gate)
            else:              # the last strip carries the epilogue: shift / diag / dots act on the complete row sum
                kw = epi.kw()
                kw["accumulate"] = epi.accumulate or b > 0
                be.csr_spmm(indptr, indices, data, S.shape, nnz, S.max_row_nnz, X, Y, **kw)


class _KronCore:
    """Chain of mode contractions; factor i sees X as (pre_i, d_i, post_i) with no transposes
    (replaces operators.py:216-223)."""
    def __init__(self, factors):
        self.Fs = [(f if f.is_contiguous() else f.contiguous()) for f in factors]
        self.shape = (_prod([f.shape[0] for f in self.Fs]), _prod([f.shape[1] for f in self.Fs]))
        self._ws = {}
        self.use_tensor_cores = True   # set False to force the exact-fp32 SIMT contraction (A/B checks)

    def _workspace(self, numel, dtype, device, slot):
        key = (slot, dtype, device)
        buf = self._ws.get(key)
        if buf is None or buf.numel() < numel:
            buf = torch.empty(numel, dtype=dtype, device=device)
            self._ws[key] = buf
        return buf

    def _tc_ok(self, X):
        """fp32, every factor 64x64, k % 32 == 0: the tcgen05 / TMA path (csrc/kron_tc.cu)."""
        if X.dtype != torch.float32 or len(self.Fs) < 2:
            return False
        if any(tuple(f.shape) != (64, 64) for f in self.Fs):
            return False
        dims = (ctypes.c_int64 * len(self.Fs))(*[64] * len(self.Fs))
        return bool(be.lib().cdll.cola_kron_tc_supported(len(self.Fs), dims, X.shape[1]))

    def _apply_tc(self, X, Y, epi):
        D = len(self.Fs)
        n, k = X.shape
        ws_bytes = int(be.lib().cdll.cola_kron_tc_workspace_bytes(n, D))
        ws = self._workspace(ws_bytes // 4, X.dtype, X.device, "tc")
        facs = (ctypes.c_void_p * D)(*[f.data_ptr() for f in self.Fs])
        ldf = (ctypes.c_int64 * D)(*[f.stride(0) for f in self.Fs])
        be.lib().call("cola_kron_matmat_tc_f32", D, facs, ldf, be.ptr(X), be.ptr(Y), k, be.ptr(ws),
                      ctypes.c_float(epi.alpha), ctypes.c_float(epi.shift),
                      be.ptr(epi.diag) if epi.diag is not None else None, int(epi.accumulate), be.ptr(epi.dots),
                      be.ptr(epi.dots_row), be.ptr(epi.gate), be.stream_ptr())

    def apply(self, X, Y, epi):
        if self.use_tensor_cores and self._tc_ok(X):
            return self._apply_tc(X, Y, epi)
        k = X.shape[1]
        D = len(self.Fs)
        d_in = [f.shape[1] for f in self.Fs]
        d_out = [f.shape[0] for f in self.Fs]
        src = X
        for i, F in enumerate(self.Fs):
            pre = _prod(d_out[:i]) if i > 0 else 1
            L = _prod(d_in[i + 1:]) if i + 1 < D else 1
            post = L * k
            last = i == D - 1
            # a square 64- / 128-wide fp32 factor with k % 32 == 0: the per-mode tcgen05 kernel (csrc/kron_tc.cu,
            # mode_tc_kernel; BASELINE config 4's Kronecker(128, 128, 64)); anything else: exact SIMT tiles
            tc = self.use_tensor_cores and be.mode_contract_tc_ok(F, pre, L, k, X)
            if last:
                if tc:
                    be.mode_contract_tc(F, pre, L, k, src, Y, epi_x=X if epi.needs_x() else None, **epi.kw())
                else:
                    be.mode_contract(F, d_out[i], d_in[i], pre, post, src, Y, epi_x=X if epi.needs_x() else None,
                                     **epi.kw())
            else:
                numel = pre * d_out[i] * post
                dst = self._workspace(numel, X.dtype, X.device, i % 2)
                if tc:
                    be.mode_contract_tc(F, pre, L, k, src, dst, gate=epi.gate)
                else:
                    be.mode_contract(F, d_out[i], d_in[i], pre, post, src, dst, gate=epi.gate)
                src = dst


class _KronSumCore:
    """sum_i I x .. x F_i x .. x I as D mode contractions that accumulate into Y, in factor order like
    operators.py:261-268 (no `0 * ev` initialisation pass, no moveaxis copies).  The last factor sees X as
    (n / d_D, d_D, k), i.e. its flattened rows are the operator's rows, so it carries the fused epilogue."""
    def __init__(self, factors):
        self.Fs = [(f if f.is_contiguous() else f.contiguous()) for f in factors]
        assert all(f.shape[0] == f.shape[1] for f in self.Fs), "KronSum factors are square"
        self.shape = (_prod([f.shape[0] for f in self.Fs]), ) * 2
        self.use_tensor_cores = True   # set False to force the exact-fp32 SIMT contractions (A/B checks)

    def _tc_ok(self, X):
        """True when every mode of this application takes the per-mode tcgen05 kernel."""
        dims = [f.shape[0] for f in self.Fs]
        return all(be.mode_contract_tc_ok(F, _prod(dims[:i]) if i else 1, _prod(dims[i + 1:]) if i + 1 < len(dims) else 1,
                                          X.shape[1], X) for i, F in enumerate(self.Fs))

    def apply(self, X, Y, epi):
        """Every mode contracts X itself and accumulates into Y; the last one carries the operator epilogue.  Square 64- /
        128-wide fp32 factors with k % 32 == 0 run on the per-mode tcgen05 kernel (csrc/kron_tc.cu, mode_tc_kernel)."""
        k = X.shape[1]
        dims = [f.shape[0] for f in self.Fs]
        D = len(dims)
        for i, F in enumerate(self.Fs):
            pre = _prod(dims[:i]) if i > 0 else 1
            L = _prod(dims[i + 1:]) if i + 1 < D else 1
            acc = epi.accumulate or i > 0
            tc = self.use_tensor_cores and be.mode_contract_tc_ok(F, pre, L, k, X)
            if i == D - 1:
                kw = epi.kw()
                kw["accumulate"] = acc
                if tc:
                    be.mode_contract_tc(F, pre, L, k, X, Y, epi_x=X if epi.needs_x() else None, **kw)
                else:
                    be.mode_contract(F, dims[i], dims[i], pre, L * k, X, Y, epi_x=X if epi.needs_x() else None, **kw)
            elif tc:
                be.mode_contract_tc(F, pre, L, k, X, Y, alpha=epi.alpha, accumulate=acc, gate=epi.gate)
            else:
                be.mode_contract(F, dims[i], dims[i], pre, L * k, X, Y, alpha=epi.alpha, accumulate=acc, gate=epi.gate)


class _BlockDiagCore:
    """One batched contraction per distinct block: kron(I_mult, M) on its row range (operators.py:299-310)."""
    def __init__(self, blocks, mults):
        self.blocks = [(b if b.is_contiguous() else b.contiguous()) for b in blocks]
        self.mults = list(mults)
        self.shape = (sum(b.shape[0] * c for b, c in zip(self.blocks, self.mults)),
                      sum(b.shape[1] * c for b, c in zip(self.blocks, self.mults)))

    def apply(self, X, Y, epi):
        k = X.shape[1]
        ri = ro = 0
        for M, c in zip(self.blocks, self.mults):
            kw = epi.kw()
            if kw.get("diag") is not None:
                kw["diag"] = be.off_ptr(epi.diag, ro)
            if be.mode_contract_tc_ok(M, c, 1, k, X):          # square 64- / 128-wide fp32 blocks: tcgen05 per-mode kernel
                be.mode_contract_tc(M, c, 1, k, be.off_ptr(X, ri * k), be.off_ptr(Y, ro * k),
                                    epi_x=be.off_ptr(X, ri * k) if epi.needs_x() else None, **kw)
            else:
                be.mode_contract(M, M.shape[0], M.shape[1], c, k, be.off_ptr(X, ri * k), be.off_ptr(Y, ro * k),
                                 epi_x=be.off_ptr(X, ri * k) if epi.needs_x() else None, **kw)
            ri += c * M.shape[1]
            ro += c * M.shape[0]


class _OpaqueCore:
    """Any other LinearOperator: call its own `_matmat` and add the result in (not a fused path)."""
    def __init__(self, op):
        self.op = op
        self.shape = tuple(op.shape)

    def apply(self, X, Y, epi):
        Z = self.op._matmat(X)
        Z = Z if Z.is_contiguous() else Z.contiguous()
        be.axpby(Z, Y, epi.alpha, 1.0 if epi.accumulate else 0.0, gate=epi.gate)
        if epi.shift != 0.0 or epi.diag is not None or epi.dots is not None:
            # epilogue as a separate sweep: Y += (shift+diag) X ; dots += <X, Y>
            be.diag_matmat(X, Y, epi.shift, epi.diag, True, epi.dots, epi.dots_row, epi.gate)


class _Epilogue:
    def __init__(self, alpha=1.0, shift=0.0, diag=None, accumulate=False, dots=None, dots_row=None, gate=None):
        self.alpha, self.shift, self.diag, self.accumulate = alpha, shift, diag, accumulate
        self.dots, self.dots_row, self.gate = dots, dots_row, gate

    def needs_x(self):
        return self.shift != 0.0 or self.diag is not None or self.dots is not None

    def kw(self):
        return dict(alpha=self.alpha, shift=self.shift, diag=self.diag, accumulate=self.accumulate, dots=self.dots,
                    dots_row=self.dots_row, gate=self.gate)


class Plan:
    """A X = sum_t scale_t * chain_t X + (shift + diag) o X.  chain_t is a list of cores applied right to left."""
    def __init__(self, shape, dtype):
        self.shape, self.dtype = shape, dtype
        self.terms = []      # (scale, [cores left-to-right])
        self.shift = 0.0
        self.diag = None

    def graph_safe(self):
        """True when every core launches only library kernels on the current stream (no opaque `_matmat`, which may
        allocate or synchronise): the condition for capturing the apply in a CUDA graph."""
        return all(not isinstance(c, _OpaqueCore) and len(ch) == 1 for _, ch in self.terms for c in ch)

    def graph_token(self):
        """Identity of every scratch buffer the apply touches (Kronecker / BlockDiag workspaces grow on demand and a
        captured graph keeps the old pointers): a cached graph is only replayed while this is unchanged."""
        tok = []
        for _, ch in self.terms:
            for c in ch:
                for key, buf in sorted(getattr(c, "_ws", {}).items(), key=lambda kv: str(kv[0])):
                    tok.append((str(key), buf.data_ptr(), buf.numel()))
        return tuple(tok)

    def describe(self):
        parts = [f"{s:g}*" + "@".join(type(c).__name__.strip("_") for c in ch) for s, ch in self.terms]
        if self.shift != 0.0:
            parts.append(f"{self.shift:g}*I")
        if self.diag is not None:
            parts.append("Diag")
        return " + ".join(parts)

    def apply(self, X, Y, dots=None, dots_row=None, gate=None):
        be.require_cuda(X, "operand")
        be.require_cuda(Y, "result buffer")
        if X.dtype != self.dtype or Y.dtype != self.dtype:
            raise TypeError(f"operand dtype {X.dtype}/{Y.dtype} does not match operator dtype {self.dtype}")
        assert X.shape[0] == self.shape[1] and Y.shape[0] == self.shape[0] and X.shape[1] == Y.shape[1]
        if X.shape[1] == 0 or X.shape[0] == 0 or Y.shape[0] == 0:
            return                                             # empty block: nothing to launch (the reference returns (n, 0))
        square = self.shape[0] == self.shape[1]
        n_terms = len(self.terms)
        if n_terms == 0:
            assert square
            be.diag_matmat(X, Y, self.shift, self.diag, False, dots, dots_row, gate)
            return
        for t, (scale, chain) in enumerate(self.terms):
            last_term = t == n_terms - 1
            src = X
            for ci in range(len(chain) - 1, -1, -1):  # right to left
                core = chain[ci]
                final = ci == 0
                if final:
                    if last_term and len(chain) > 1 and (self.shift != 0.0 or self.diag is not None or dots is not None):
                        # The fused epilogue of a core reads the core's own input; at the head of a Product chain
                        # that is the intermediate, not X.  Apply the chain plainly, then one diagonal sweep adds
                        # (shift + diag) o X and takes the <X, Y> column dots.
                        core.apply(src, Y, _Epilogue(scale, 0.0, None, t > 0, None, None, gate))
                        be.diag_matmat(X, Y, self.shift, self.diag, True, dots, dots_row, gate)
                    elif last_term:
                        core.apply(src, Y, _Epilogue(scale, self.shift, self.diag, n_terms > 1, dots, dots_row, gate))
                    else:
                        core.apply(src, Y, _Epilogue(scale, 0.0, None, t > 0, None, None, gate))
                else:
                    tmp = torch.empty((core.shape[0], X.shape[1]), dtype=X.dtype, device=X.device)
                    core.apply(src, tmp, _Epilogue(gate=gate))
                    src = tmp


def _core_of(op):
    if isinstance(op, Dense):
        return _DenseCore(op.A)
    if isinstance(op, Sparse):
        return _CsrCore(op)
    if isinstance(op, Kronecker):
        return _KronCore([_dense_of(M) for M in op.Ms])
    if isinstance(op, KronSum):
        return _KronSumCore([_dense_of(M) for M in op.Ms])
    if isinstance(op, Tridiagonal) and op.beta.dim() == 2 and op.beta.shape[1] == 1:
        return _CsrCore(op.as_sparse())
    if isinstance(op, BlockDiag):
        return _BlockDiagCore([_dense_of(M) for M in op.Ms], op.multiplicities)
    return _OpaqueCore(op)


def _dense_of(op):
    if isinstance(op, Dense):
        return op.A
    return op.to_dense()


def compile_plan(A):
    plan = Plan(tuple(A.shape), A.dtype)

    def add_diag(d):
        plan.diag = d if plan.diag is None else plan.diag + d

    def visit(op, scale):
        if isinstance(op, Identity):
            plan.shift += scale
        elif isinstance(op, ScalarMul):
            plan.shift += scale * float(op.c)
        elif isinstance(op, Diagonal):
            add_diag(op.diag if scale == 1.0 else scale * op.diag)
        elif isinstance(op, Transpose) and op.pushed() is not None:
            visit(op.pushed(), scale)
        elif type(op) is Sum:
            for M in op.Ms:
                visit(M, scale)
        elif type(op) is Product:
            rest = []
            for M in op.Ms:
                if isinstance(M, ScalarMul):
                    scale = scale * float(M.c)
                elif isinstance(M, Identity):
                    continue
                elif isinstance(M, Transpose) and M.pushed() is not None:
                    rest.append(M.pushed())                    # a transposed Dense / CSR leaf is a core like any other
                else:
                    rest.append(M)
            if not rest:
                plan.shift += scale
            elif len(rest) == 1:
                visit(rest[0], scale)
            else:
                plan.terms.append((scale, [_core_of(M) for M in rest]))
        else:
            plan.terms.append((scale, [_core_of(op)]))

    visit(A, 1.0)
    if plan.diag is not None:
        plan.diag = plan.diag.to(A.dtype).contiguous()
    return plan
