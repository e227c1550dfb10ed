"""Backward passes of the native API.

The reference wraps its iterative solvers in `iterative_autograd` (cola/utils/custom_autodiff.py:4-86) with two
hand-written rules, restated here on the kernels:

  cg_bwd   (cola/linalg/inverse/cg.py:72-86)   db = run_batched_cg(A, dy, x0, max_iters, tol, P)  -- a second solve with
           the forward call's settings -- and dA = vjp of  theta -> A(theta) @ soln  with cotangent -db;
  slq_bwd  (cola/linalg/tbd/slq.py:10-31)      the probes are re-drawn from the key, solves = cg(A, probes, tol=1e-6,
           max_iters=100), and dA = vjp of  theta -> A(theta) @ probes  with cotangent  g / num_samples * solves.

Both end in `xnp.vjp_derivs` (cola/backends/torch_fns.py:244-260), which the reference leaves to torch autograd over
its eager matmat.  `param_vjp` computes that vjp leaf by leaf with the kernels of csrc/param_grad.cu (SDDMM on the
CSR pattern for Sparse values, row dots for Diagonal / Tridiagonal bands, G V^T for Dense, the mode Gram for
Kronecker / KronSum factors) and the operator algebra by the chain rule (Sum, Product, ScalarMul, Transpose,
BlockDiag), so a training step through `solve` or `logdet` never leaves the device kernels.
"""
import torch

from . import backend as be
from . import ops as O
from . import rng


# ----------------------------------------------------------------------------------------------------------
# parameters of an operator tree
# ----------------------------------------------------------------------------------------------------------
def parameters(A):
    """Floating-point tensors of the operator tree, in attribute order (the leaves `LinearOperator.flatten`,
    cola/ops/operator_base.py:89-95, would return; index arrays are not parameters)."""
    out, seen = [], set()

    def walk(op):
        for val in vars(op).values():
            if torch.is_tensor(val):
                if val.is_floating_point() and id(val) not in seen:
                    seen.add(id(val))
                    out.append(val)
            elif isinstance(val, O.LinearOperator):
                walk(val)
            elif isinstance(val, (tuple, list)) and val and all(isinstance(v, O.LinearOperator) for v in val):
                for v in val:
                    walk(v)

    walk(A)
    return out


def needs_grad(A, *tensors):
    """True when autograd is recording and the operator or one of `tensors` takes part in the graph."""
    if not torch.is_grad_enabled():
        return False
    if any(t is not None and torch.is_tensor(t) and t.requires_grad for t in tensors):
        return True
    return any(p.requires_grad for p in parameters(A))


def _wants(op, wanted):
    return any(id(p) in wanted for p in parameters(op))


# ----------------------------------------------------------------------------------------------------------
# kernel wrappers
# ----------------------------------------------------------------------------------------------------------
def _sddmm(S, G, V, alpha):
    out = torch.empty_like(S.data)
    be.sddmm_csr(S.indptr, S.indices, S.shape[0], G, V, alpha, out)
    return out


def _row_dots(G, g_row0, V, v_row0, n, alpha):
    """out[i] = alpha * sum_c G[g_row0 + i, c] V[v_row0 + i, c]."""
    out = torch.empty(n, dtype=V.dtype, device=V.device)
    be.row_dots(G, g_row0, V, v_row0, n, alpha, out)
    return out


def _gram(G, g_off, Z, z_off, d_g, d_z, pre, post, alpha):
    """C[a, j] = alpha * sum_{p, t} G[p, a, t] Z[p, j, t] with G / Z read as (pre, d, post) blocks from the given
    element offsets."""
    C = torch.zeros((d_g, d_z), dtype=torch.float64, device=Z.device)
    be.gram_nt(G, g_off, Z, z_off, d_g, d_z, pre, post, alpha, C)
    return C.to(Z.dtype)


def _rows(X, r0, r1):
    """Row block of a contiguous (n, k) matrix (contiguous itself)."""
    return X[r0:r1]


# ----------------------------------------------------------------------------------------------------------
# vjp of theta -> A(theta) @ V with cotangent G
# ----------------------------------------------------------------------------------------------------------
def param_vjp(A, V, G, scale=1.0, wanted=None):
    """{id(parameter): gradient} of  scale * <G, A(theta) V>  for every parameter in `wanted` (default: those that
    require grad).  V (n_in, k) and G (n_out, k) are CUDA blocks of the operator's dtype."""
    if wanted is None:
        wanted = [p for p in parameters(A) if p.requires_grad]
    ids = {id(p) for p in wanted}
    grads = {}
    if not ids:
        return grads
    V = V.to(A.dtype).contiguous()
    G = G.to(A.dtype).contiguous()
    if V.dim() == 1:
        V, G = V[:, None].contiguous(), G[:, None].contiguous()
    with torch.no_grad():
        _vjp(A, V, G, float(scale), ids, grads)
    return grads


def _add(grads, p, g):
    g = g.reshape(p.shape).to(p.dtype)
    if id(p) in grads:
        grads[id(p)] = grads[id(p)] + g
    else:
        grads[id(p)] = g


def _vjp(op, V, G, scale, ids, grads):
    if not _wants(op, ids):
        return
    k = V.shape[1]
    dt, dev = V.dtype, V.device
    if isinstance(op, O.Transpose):
        # <G, A^T V> = <V, A G>: the vjp of A with input G and cotangent V
        return _vjp(op.A, G, V, scale, ids, grads)
    if type(op) is O.Dense:
        return _add(grads, op.A, _gram(G, 0, V, 0, op.shape[0], op.shape[1], 1, k, scale))
    if isinstance(op, O.Sparse):
        return _add(grads, op.data, _sddmm(op, G, V, scale))
    if isinstance(op, O.Diagonal):
        return _add(grads, op.diag, _row_dots(G, 0, V, 0, op.shape[0], scale))
    if isinstance(op, O.ScalarMul):
        d = torch.zeros(k, dtype=torch.float64, device=dev)
        be.col_dots(G, V, d)
        return _add(grads, op.c, scale * d.sum())
    if isinstance(op, O.Identity):
        return
    if isinstance(op, O.Tridiagonal) and op.beta.dim() == 2 and op.beta.shape[1] == 1:
        n = op.shape[0]
        if id(op.beta) in ids:
            _add(grads, op.beta, _row_dots(G, 0, V, 0, n, scale))
        if id(op.alpha) in ids:                                 # y[i+1] += alpha[i] x[i]   (operators.py:362-367)
            _add(grads, op.alpha, _row_dots(G, 1, V, 0, n - 1, scale))
        if id(op.gamma) in ids:                                 # y[i]   += gamma[i] x[i+1]
            _add(grads, op.gamma, _row_dots(G, 0, V, 1, n - 1, scale))
        return
    if type(op) is O.Sum:
        for M in op.Ms:
            _vjp(M, V, G, scale, ids, grads)
        return
    if type(op) is O.Product:
        Ms = op.Ms
        # inputs of every factor, right to left; cotangents left to right
        Z = [None] * len(Ms)
        Z[-1] = V
        for i in range(len(Ms) - 1, 0, -1):
            Z[i - 1] = (Ms[i] @ Z[i]).contiguous()
        Gc = G
        for i, M in enumerate(Ms):
            _vjp(M, Z[i], Gc, scale, ids, grads)
            if i + 1 < len(Ms) and any(_wants(Mj, ids) for Mj in Ms[i + 1:]):
                Gc = (M.T @ Gc).contiguous()
        return
    if isinstance(op, O.Kronecker):
        return _vjp_kron(op, V, G, scale, ids, grads)
    if isinstance(op, O.KronSum):
        d = [M.shape[0] for M in op.Ms]
        for i, M in enumerate(op.Ms):
            if not _wants(M, ids):
                continue
            F = _dense_leaf(M)
            pre, post = O._prod(d[:i]) if i else 1, (O._prod(d[i + 1:]) if i + 1 < len(d) else 1) * k
            _add(grads, F, _gram(G, 0, V, 0, d[i], d[i], pre, post, scale))
        return
    if isinstance(op, O.BlockDiag):
        r_in = r_out = 0
        for M, c in zip(op.Ms, op.multiplicities):
            do, di = M.shape
            if _wants(M, ids):
                if type(M) is O.Dense:                          # all copies at once: a sum of c products G_p V_p^T
                    _add(grads, M.A, _gram(G, r_out * k, V, r_in * k, do, di, c, k, scale))
                else:
                    for j in range(c):
                        _vjp(M, _rows(V, r_in + j * di, r_in + (j + 1) * di), _rows(G, r_out + j * do, r_out + (j + 1) * do),
                             scale, ids, grads)
            r_in += c * di
            r_out += c * do
        return
    raise NotImplementedError(f"cola_b200: no parameter gradient rule for {type(op).__name__} on the native path "
                              "(keep the reference operator and use cola_b200.install(), which runs the reference's vjp)")


def _dense_leaf(M):
    if type(M) is not O.Dense:
        raise NotImplementedError("cola_b200: Kronecker / KronSum factor gradients need Dense factors on the native path")
    return M.A


def _vjp_kron(op, V, G, scale, ids, grads):
    """Y = (F_1 x ... x F_D) V as D mode contractions (operators.py:216-223).  For factor i the input of its contraction
    is Z_i = (modes > i already contracted) and its cotangent is G_i = (modes < i contracted with the transposed
    factors); dF_i[a, j] = sum_{p, q} G_i[p, a, q] Z_i[p, j, q]."""
    Fs = [_dense_leaf(M) if _wants(M, ids) else O._dense_of(M) for M in op.Ms]
    Fs = [F if F.is_contiguous() else F.contiguous() for F in Fs]
    D = len(Fs)
    k = V.shape[1]
    dt, dev = V.dtype, V.device
    d_out = [F.shape[0] for F in Fs]
    d_in = [F.shape[1] for F in Fs]
    want = [id(F) in ids for F in Fs]
    first, last = min(i for i in range(D) if want[i]), max(i for i in range(D) if want[i])
    # Z[i]: modes i+1 .. D-1 contracted (right to left), needed down to `first`
    Z = [None] * D
    Z[D - 1] = V
    for i in range(D - 1, first, -1):
        pre = O._prod(d_in[:i]) if i else 1
        post = (O._prod(d_out[i + 1:]) if i + 1 < D else 1) * k
        out = torch.empty(pre * d_out[i] * post, dtype=dt, device=dev)
        be.mode_contract(Fs[i], d_out[i], d_in[i], pre, post, Z[i], out)
        Z[i - 1] = out
    Gc = G
    for i in range(0, last + 1):
        pre = O._prod(d_in[:i]) if i else 1
        post = (O._prod(d_out[i + 1:]) if i + 1 < D else 1) * k
        if want[i]:
            _add(grads, Fs[i], _gram(Gc, 0, Z[i], 0, d_out[i], d_in[i], pre, post, scale))
        if i < last:
            Ft = Fs[i].T.contiguous()
            out = torch.empty(pre * d_in[i] * post, dtype=dt, device=dev)
            be.mode_contract(Ft, d_in[i], d_out[i], pre, post, Gc, out)
            Gc = out


# ----------------------------------------------------------------------------------------------------------
# autograd rules
# ----------------------------------------------------------------------------------------------------------
class _CgSolve(torch.autograd.Function):
    """soln = run(A, rhs)[0] with cg_bwd as its backward.  `run` is the forward call with every setting bound
    (x0, max_iters, tol, P, pbar), exactly what cg_bwd re-applies to the output cotangent."""

    @staticmethod
    def forward(ctx, A, run, holder, rhs, *params):
        soln, *rest = run(A, rhs.detach())
        holder.extend(rest)
        ctx.A, ctx.run = A, run
        ctx.params = params
        ctx.save_for_backward(soln)
        return soln

    @staticmethod
    def backward(ctx, dy):
        (soln, ) = ctx.saved_tensors
        A = ctx.A
        with torch.no_grad():
            db = ctx.run(A, dy.contiguous())[0]                 # cg.py:78
            wanted = [p for p, need in zip(ctx.params, ctx.needs_input_grad[4:]) if need]
            grads = param_vjp(A, soln, db, scale=-1.0, wanted=wanted)     # cg.py:80-85
        return (None, None, None, db if ctx.needs_input_grad[3] else None,
                *[grads.get(id(p)) if need else None for p, need in zip(ctx.params, ctx.needs_input_grad[4:])])


def cg_with_grad(A, rhs, run):
    """Differentiable `run(A, rhs) -> (soln, *rest)`; returns (soln, rest)."""
    holder = []
    params = parameters(A)
    soln = _CgSolve.apply(A, run, holder, rhs, *params)
    return soln, holder


class _Slq(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A, fwd, kwargs, *params):
        out = fwd()
        ctx.A, ctx.kwargs, ctx.params = A, kwargs, params
        return out

    @staticmethod
    def backward(ctx, g):
        from .linalg.cg import cg
        from .linalg.stochastic import friendly_chunks, probe_chunk
        A, kw = ctx.A, ctx.kwargs
        num = kw["num_samples"]
        key = kw.get("key")
        key = rng.PRNGKey(0) if key is None else key             # slq.py:16-17
        wanted = [p for p, need in zip(ctx.params, ctx.needs_input_grad[3:]) if need]
        total = {}
        with torch.no_grad():
            probes = rng.randn(A.shape[1], num, dtype=A.dtype, key=key, device=A.device)      # slq.py:18
            coef = 1.0 / num                                     # slq.py:21
            cb = probe_chunk(A.shape[1], 100, A.dtype, A.device, kw.get("probe_chunk_size"))
            for c0, c1 in friendly_chunks(num, cb, A.dtype):
                Z = probes[:, c0:c1].contiguous()
                solves, _ = cg(A, Z, tol=1e-6, max_iters=100)    # slq.py:19
                d_solves = (coef * g.to(A.dtype)) * solves       # slq.py:22-23
                part = param_vjp(A, Z, d_solves, scale=1.0, wanted=wanted)                   # slq.py:25-29
                for pid, val in part.items():
                    total[pid] = total[pid] + val if pid in total else val
        return (None, None, None, *[total.get(id(p)) if need else None for p, need in zip(ctx.params, ctx.needs_input_grad[3:])])


def slq_with_grad(A, fwd, kwargs):
    return _Slq.apply(A, fwd, kwargs, *parameters(A))
