"""Drop-in hook for wilson-labs/cola.

The reference resolves its Krylov loops by module attribute at call time (`cg.py:91`, `lanczos.py:221`,
`arnoldi.py:199`, `slq.py:75`, `diagonal_estimation.py:57-58`) and applies operators through `_matmat`, so one
`install()` is the whole integration: it rebinds

    cola.linalg.inverse.cg.run_batched_cg                          (cg.py:94-119)
    cola.linalg.decompositions.lanczos.lanczos_fact                (lanczos.py:235-272)
    cola.linalg.decompositions.arnoldi.arnoldi_fact                (arnoldi.py:289-324)
    cola.linalg.trace.diagonal_estimation.hutchinson_diag_estimate (diagonal_estimation.py:158-210)
    cola.linalg.tbd.slq.slq_fwd                                    (slq.py:37-52)
    Dense/Sparse/Kronecker/KronSum/BlockDiag/Diagonal/Sum/Product/ScalarMul._matmat  (operators.py)

to adapters that take the B200 path when the operator lives on a CUDA device in float32/float64 and its tree
converts (`from_cola`), and call the saved reference function otherwise (CPU, complex, JAX/NumPy backends,
preconditioners that do not convert, exotic operators).  The adapters translate between the
reference's state layouts and this package's: the reference keeps the Krylov basis as (b, n, m+2) with the
vector index fastest, the kernels as (m+2, n, b); what is handed back is a strided view with the reference's
logical shape.  `uninstall()` restores everything.

The reference is not importable on the GPU box, so the adapters are exercised on CPU in
tests/test_plugin_reference.py (FORCE_FAST_PATH): once with oracle-backed stand-ins for the loops, and once with
this package's own host loops running on the test-only kernel stand-ins of tests/host_harness.py, driven through
the reference's public API; the CUDA kernels themselves are checked against the oracle in tests/test_gpu_parity*.py.
"""
import importlib

import torch

from . import ops as bops

# cola_b200.linalg re-exports functions named cg / lanczos / arnoldi, which shadow the submodules as attributes
b_cg = importlib.import_module(__package__ + ".linalg.cg")
b_lanczos = importlib.import_module(__package__ + ".linalg.lanczos")
b_arnoldi = importlib.import_module(__package__ + ".linalg.arnoldi")
b_stoch = importlib.import_module(__package__ + ".linalg.stochastic")
b_unary = importlib.import_module(__package__ + ".linalg.unary")

FORCE_FAST_PATH = False      # tests only: route CPU operators through the adapters as well
_SAVED = {}
_ANNOTATIONS = ("PSD", "SelfAdjoint", "Unitary", "Stiefel")


class NotConvertible(Exception):
    pass


def _is_fast_dtype(dtype):
    return dtype in (torch.float32, torch.float64)


def _on_fast_device(A):
    if FORCE_FAST_PATH:
        return True
    dev = getattr(A, "device", None)
    try:
        return dev is not None and torch.device(dev).type == "cuda"
    except (TypeError, RuntimeError):
        return False


def from_cola(A, cola):
    """Reference operator tree -> mirror classes, one to one (same constructor arguments); annotations are
    carried over by name.  Leaves that are not on the hot path raise NotConvertible (the caller then stays on
    the reference path), except plain `LinearOperator`s with a matmat closure, which are wrapped as opaque."""
    token = _leaf_token(A)
    cached = getattr(A, "_b200_mirror", None)
    if cached is not None and cached[0] == token:
        return cached[1]
    R = cola.ops
    if not _is_fast_dtype(A.dtype):
        raise NotConvertible(f"dtype {A.dtype}")
    unary = getattr(getattr(cola.linalg, "unary", None), "unary", None)
    if isinstance(A, R.Sparse):
        csr = A.A
        M = bops.Sparse.from_csr(csr.crow_indices(), csr.col_indices(), csr.values(), tuple(A.shape))
    elif isinstance(A, R.Triangular):
        M = bops.Triangular(A.A, lower=A.lower)
    elif isinstance(A, R.Dense) and type(A).__name__.startswith("Dense"):
        M = bops.Dense(A.A)
    elif type(A).__name__.startswith("TriangularInv") and hasattr(A, "lower"):
        b_dispatch = importlib.import_module(__package__ + ".linalg.dispatch")
        M = b_dispatch.TriangularInv(bops.Triangular(A.A, lower=A.lower))
    elif isinstance(A, R.Identity):
        M = bops.Identity(tuple(A.shape), A.dtype)
        M.device = torch.device(A.device) if getattr(A, "device", None) is not None else M.device
    elif isinstance(A, R.ScalarMul):
        M = bops.ScalarMul(float(A.c), tuple(A.shape), A.dtype, getattr(A, "device", None))
    elif isinstance(A, R.Diagonal):
        M = bops.Diagonal(A.diag)
    elif isinstance(A, R.Tridiagonal):
        M = bops.Tridiagonal(A.alpha, A.beta, A.gamma)
    elif isinstance(A, R.Kronecker):
        M = bops.Kronecker(*[from_cola(m, cola) for m in A.Ms])
    elif isinstance(A, R.KronSum):
        M = bops.KronSum(*[from_cola(m, cola) for m in A.Ms])
    elif isinstance(A, R.BlockDiag):
        M = bops.BlockDiag(*[from_cola(m, cola) for m in A.Ms], multiplicities=list(A.multiplicities))
    elif isinstance(A, R.Sum):
        M = bops.Sum(*[from_cola(m, cola) for m in A.Ms])
    elif isinstance(A, R.Product):
        M = bops.Product(*[from_cola(m, cola) for m in A.Ms])
    elif isinstance(A, R.Transpose):
        M = from_cola(A.A, cola).T
    elif type(A).__name__ in ("NystromPrecond", "NystromPrecondLazy", "AdaNysPrecond") and hasattr(A, "subspace_scaling"):
        # preconditioners.py:128-130: U diag(s) U^T + I as a two-core chain plus the identity
        U = A.U.contiguous()
        M = bops.Sum(bops.Product(bops.Dense(U), bops.Dense((A.subspace_scaling * U.T).contiguous())),
                     bops.Identity(tuple(A.shape), A.dtype).to(U.device))
    elif unary is not None and isinstance(A, unary.LanczosUnary):
        M = b_stoch.LanczosUnary(from_cola(A.A, cola), A.f, **{k: v for k, v in getattr(A, "kwargs", {}).items()
                                                                if k in ("max_iters", "tol", "pbar")})
    elif unary is not None and isinstance(A, unary.ArnoldiUnary):
        M = b_unary.ArnoldiUnary(from_cola(A.A, cola), A.f, **{k: v for k, v in getattr(A, "kwargs", {}).items()
                                                                if k in ("max_iters", "tol", "pbar")})
    else:
        raise NotConvertible(type(A).__name__)
    ref_names = {getattr(a, "__name__", str(a)) for a in getattr(A, "annotations", ())}
    for name in _ANNOTATIONS:
        if name in ref_names:
            M.annotations = set(M.annotations) | {getattr(bops, name)}
    try:
        A._b200_mirror = (token, M)
    except (AttributeError, TypeError):
        pass
    return M


def _leaf_token(A):
    """(id, in-place version) of the operator's parameter tensors: a cached mirror is reused only while none of them
    was replaced or written in place (mirrors and their plans hold derived copies: folded scalars, CSR index arrays)."""
    try:
        leaves = A.flatten()[0]
    except Exception:
        return None
    return tuple((id(t), t._version) if torch.is_tensor(t) else repr(t) for t in leaves)


def _needs_autograd(A, *tensors):
    """True when autograd is recording and an operand or a parameter of the operator requires grad.  The kernels are
    not differentiable, so those calls stay on the reference's eager code: that is what runs when the reference
    differentiates `A(theta) @ v` inside its custom backward rules (cg.py:72-86, slq.py:10-31).  The solver loops
    themselves are invoked inside `torch.autograd.Function.forward` (custom_autodiff.py:44-49), where recording is off,
    and keep the fast path."""
    if not torch.is_grad_enabled():
        return False
    if any(torch.is_tensor(t) and t.requires_grad for t in tensors):
        return True
    try:
        leaves = A.flatten()[0]
    except Exception:
        return False
    return any(torch.is_tensor(t) and t.requires_grad for t in leaves)


def _mirror_or_none(A, cola, *tensors):
    if not (_on_fast_device(A) and _is_fast_dtype(A.dtype)) or _needs_autograd(A, *tensors):
        return None
    try:
        return from_cola(A, cola)
    except NotConvertible:
        return None


# ------------------------------------------------------------------------------------------------ adapters
def _make_run_batched_cg(cola, ref):
    def run_batched_cg(A, b, x0, max_iters, tol, preconditioner, pbar):
        M = _mirror_or_none(A, cola, b, x0)
        if M is None:
            return ref(A, b, x0, max_iters, tol, preconditioner, pbar)
        if isinstance(preconditioner, cola.ops.Identity):
            P = bops.I_like(M)
        else:
            P = _mirror_or_none(preconditioner, cola)       # e.g. NystromPrecond -> U diag(s) U^T + I in the plan
            if P is None:
                return ref(A, b, x0, max_iters, tol, preconditioner, pbar)
        return b_cg.run_batched_cg(M, b, x0, max_iters, tol, P, pbar)
    return run_batched_cg


def _make_lanczos_fact(cola, ref):
    def lanczos_fact(A, init_val, max_iters=100, tol=1e-7, pbar=False):
        # a state that is not a fresh start is the implicitly restarted variant (lanczos.py:111): reference path
        M = _mirror_or_none(A, cola, init_val[0]) if int(init_val[3]) == 1 else None
        if M is None:
            return ref(A, init_val, max_iters, tol, pbar)
        V0, diag0, _, _ = init_val                               # init_lanczos (lanczos.py:275-284): V0[..., 1] = rhs/|rhs|
        rhs = V0[..., 1].T.contiguous()                          # (n, b)
        st = b_lanczos.lanczos_fact(M, rhs, max_iters, tol, pbar)
        dt = diag0.dtype
        m = diag0.shape[-1]
        V = st.V.permute(2, 1, 0)                                # (b, n, m+2) view of the (m+2, n, b) basis
        diag = st.alpha_acc[:m].to(dt).T.contiguous()            # (b, m)
        subdiag = torch.sqrt(st.sub_sq[:m + 1]).to(dt).T.contiguous()   # (b, m+1)
        i = torch.tensor(st.i, dtype=torch.int32, device=V.device)
        return V, diag, subdiag, i, st.info
    return lanczos_fact


def _make_arnoldi_fact(cola, ref):
    def arnoldi_fact(A, init_val, max_iters, tol, pbar):
        # idx > 0: a restart of implicitly restarted Arnoldi on a partly filled basis (arnoldi.py:99): reference path
        M = _mirror_or_none(A, cola, init_val[0]) if int(init_val[2]) == 0 else None
        if M is None:
            return ref(A, init_val, max_iters, tol, pbar)
        Q0 = init_val[0]                                         # init_arnoldi (arnoldi.py:327-335): Q0[..., 0] = rhs/|rhs|
        rhs = Q0[..., 0].T.contiguous()
        Q, H, idx, info = b_arnoldi.arnoldi_fact(M, rhs, max_iters, tol, pbar)
        return Q.permute(2, 1, 0), H, torch.tensor(int(idx), dtype=torch.int32, device=H.device), info
    return arnoldi_fact


def _make_hutch(cola, ref):
    def hutchinson_diag_estimate(A, k=0, bs=100, tol=3e-2, max_iters=10000, pbar=False, rand='normal', key=None):
        M = _mirror_or_none(A, cola)
        if M is None:
            return ref(A, k, bs, tol, max_iters, pbar, rand, key)
        return b_stoch.hutchinson_diag_estimate(M, k, bs, tol, max_iters, pbar, rand, key)
    return hutchinson_diag_estimate


def _make_slq_fwd(cola, ref):
    """`slq_fwd` carries the reference's custom autograd rule (`@iterative_autograd(slq_bwd)`, slq.py:37): the
    replacement is wrapped the same way, so the forward runs here (inside Function.forward, recording off) and the
    backward stays the reference's slq_bwd -- whose CG solve comes back through the rebound run_batched_cg."""
    inner_ref = getattr(ref, "__wrapped__", None)

    def slq_fwd(A, fun, num_samples, max_iters, tol, pbar, key):
        M = _mirror_or_none(A, cola)
        if M is None:
            return (inner_ref or ref)(A, fun, num_samples, max_iters, tol, pbar, key)
        return b_stoch.slq_fwd(M, fun, num_samples, max_iters, tol, pbar, key)

    if inner_ref is None:
        return slq_fwd
    autodiff = importlib.import_module("cola.utils.custom_autodiff")
    slq_mod = importlib.import_module("cola.linalg.tbd.slq")
    return autodiff.iterative_autograd(slq_mod.slq_bwd)(slq_fwd)


def _make_matmat(cola, ref):
    def _matmat(self, X):
        if torch.is_tensor(X) and (X.is_cuda or FORCE_FAST_PATH) and _is_fast_dtype(X.dtype) and X.dtype == self.dtype \
                and not torch._C._functorch.is_batchedtensor(X):   # under vmap there is no pointer to hand over
            M = _mirror_or_none(self, cola, X)
            if M is not None:
                return M._matmat(X.contiguous())
        return ref(self, X)
    return _matmat


_LOOPS = (
    ("cola.linalg.inverse.cg", "run_batched_cg", _make_run_batched_cg),
    ("cola.linalg.decompositions.lanczos", "lanczos_fact", _make_lanczos_fact),
    ("cola.linalg.decompositions.arnoldi", "arnoldi_fact", _make_arnoldi_fact),
    ("cola.linalg.trace.diagonal_estimation", "hutchinson_diag_estimate", _make_hutch),
    ("cola.linalg.tbd.slq", "slq_fwd", _make_slq_fwd),
)
# Tridiagonal is deliberately absent: the reference vmaps `Tridiagonal.to_dense` over the batched Lanczos output
# (unary.py:52), and functorch-batched leaves cannot be handed to kernels by pointer.  Inside a tree handed to one
# of the loops it still converts (from_cola) and runs on the CSR kernel.
_MATMAT_CLASSES = ("Dense", "Sparse", "Kronecker", "KronSum", "BlockDiag", "Diagonal", "Sum", "Product", "ScalarMul")


def install(cola=None, matmats=True):
    """Rebind the reference's hot-path functions (idempotent).  `cola` defaults to `import cola`."""
    if _SAVED:
        return
    if cola is None:
        cola = importlib.import_module("cola")
    for modname, attr, make in _LOOPS:
        mod = importlib.import_module(modname)
        ref = getattr(mod, attr)
        _SAVED[(mod, attr)] = ref
        setattr(mod, attr, make(cola, ref))
    if matmats:
        for name in _MATMAT_CLASSES:
            klass = getattr(cola.ops, name)
            if "_matmat" in vars(klass):
                ref = vars(klass)["_matmat"]
                _SAVED[(klass, "_matmat")] = ref
                setattr(klass, "_matmat", _make_matmat(cola, ref))


def uninstall():
    for (owner, attr), ref in list(_SAVED.items()):
        setattr(owner, attr, ref)
    _SAVED.clear()
