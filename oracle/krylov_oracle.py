"""CPU restatement (torch, eager, CPU tensors) of CoLA's Krylov hot path.

TEST INFRASTRUCTURE.  This module is the *checker* for the CUDA path in
cola_b200/: it restates, operation by operation and in the same order, what the
reference (wilson-labs/cola, read-only at /root/reference in the build
container) does on its torch backend, so that on CPU tensors its outputs are
bit-identical to the reference's.  That claim is pinned by tests/test_oracle_golden.py
against fixtures generated from the real reference by tests/golden/make_golden.py
(parity PINNED: see tests/golden/MANIFEST.json).

Nothing in cola_b200/ imports this file; bench.py uses it only for the
`cpu_baseline` / `--impl reference` legs.

Each function cites the reference lines it follows (paths relative to
/root/reference).
"""
import hashlib
import time

import numpy as np
import torch

_TINY = 1e-40  # cola/linalg/inverse/cg.py:8


# ----------------------------------------------------------------------------------
# keyed RNG  (cola/backends/torch_fns.py:154-155, 222-241)
# ----------------------------------------------------------------------------------
def sha_key(n):
    raw = n.to_bytes((n.bit_length() + 7) // 8, "big")
    return int(int.from_bytes(hashlib.sha256(raw).digest(), "big") % (2**32 - 1))


PRNGKey = sha_key
next_key = sha_key


def keyed_randn(*shape, dtype, key, device="cpu"):
    saved = torch.random.get_rng_state()
    torch.random.manual_seed(key)
    z = torch.randn(*shape, dtype=dtype, device=device)
    torch.random.set_rng_state(saved)
    return z


# ----------------------------------------------------------------------------------
# operators: only `matmat`, `shape`, `dtype`  (cola/ops/operators.py)
# ----------------------------------------------------------------------------------
class Op:
    def __init__(self, shape, dtype):
        self.shape, self.dtype = tuple(shape), dtype

    def __matmul__(self, X):
        assert X.shape[0] == self.shape[-1], f"dimension mismatch {self.shape} vs {X.shape}"
        if X.dim() == 1:
            return self.matmat(X.reshape(-1, 1)).reshape(-1)
        return self.matmat(X)

    def to_dense(self):
        return self.matmat(torch.eye(self.shape[-1], dtype=self.dtype))


class DenseOp(Op):  # operators.py:12-38
    def __init__(self, M):
        super().__init__(M.shape, M.dtype)
        self.M = M

    def matmat(self, X):
        dt = torch.promote_types(self.dtype, X.dtype)
        return self.M.to(dt) @ X.to(dt)


class SparseOp(Op):
    """operators.py:48-81: COO in, sorted by row, scipy COO->CSR for indptr/indices (int32), torch CSR SpMM.

    Deviation (reference defect, see DESIGN.md "Sparse constructor"): the reference sorts with
    `argsort(row_indices)` which torch does not guarantee to be stable; when it reorders entries
    inside a row the reference pairs its permuted `data` with scipy's canonically ordered `indices`
    and silently builds a different matrix.  The oracle (and cola_b200) use a STABLE row sort, which is
    what the constructor means; the two agree whenever the reference's argsort is order-preserving,
    and tests/golden/make_golden.py repairs the reference instance otherwise."""
    def __init__(self, data, rows, cols, shape):
        from scipy.sparse import coo_array
        super().__init__(shape, data.dtype)
        order = torch.argsort(rows, stable=True)
        self.data, rows, cols = data[order], rows[order], cols[order]
        csr = coo_array((self.data.detach().numpy(), (rows.numpy(), cols.numpy())), shape=shape).tocsr()
        self.indptr = torch.tensor(csr.indptr, dtype=torch.int32)
        self.indices = torch.tensor(csr.indices, dtype=torch.int32)
        self.csr = torch.sparse_csr_tensor(crow_indices=self.indptr, col_indices=self.indices,
                                           values=self.data.detach(), size=shape)

    def matmat(self, X):
        return self.csr @ X


class ScaledIdentityOp(Op):  # Product[ScalarMul, Identity]  operators.py:84-127,153-156
    def __init__(self, c, n, dtype):
        super().__init__((n, n), dtype)
        self.c = torch.tensor(c, dtype=dtype)

    def matmat(self, X):
        return self.c * X


class IdentityOp(Op):  # operators.py:104-127: returns the same tensor
    def __init__(self, n, dtype):
        super().__init__((n, n), dtype)

    def matmat(self, X):
        return X


class ScaledOp(Op):  # c * A  ==  Product[ScalarMul, A]   (cola/fns.py mul, operators.py:153-156)
    def __init__(self, c, A):
        super().__init__(A.shape, A.dtype)
        self.c, self.A = torch.tensor(c, dtype=A.dtype), A

    def matmat(self, X):
        return self.c * self.A.matmat(X)


class DiagonalOp(Op):  # operators.py:323-348
    def __init__(self, d):
        super().__init__((d.shape[0], d.shape[0]), d.dtype)
        self.d = d

    def matmat(self, X):
        return self.d[:, None] * X


class SumOp(Op):  # operators.py:167-191: python sum(), i.e. ((0 + t0) + t1) + ...
    def __init__(self, *terms):
        super().__init__(terms[0].shape, terms[0].dtype)
        self.terms = terms

    def matmat(self, X):
        acc = 0
        for t in self.terms:
            acc = acc + t.matmat(X)
        return acc


class ProductOp(Op):  # operators.py:138-164: right to left
    def __init__(self, *factors):
        super().__init__((factors[0].shape[0], factors[-1].shape[1]), factors[0].dtype)
        self.factors = factors

    def matmat(self, X):
        for f in reversed(self.factors):
            X = f.matmat(X)
        return X


class KroneckerOp(Op):  # operators.py:198-230
    def __init__(self, *factors):
        r = int(np.prod([f.shape[0] for f in factors]))
        c = int(np.prod([f.shape[1] for f in factors]))
        super().__init__((r, c), factors[0].dtype)
        self.factors = factors

    def matmat(self, X):
        E = X.reshape(*[f.shape[-1] for f in self.factors], -1)
        for axis, f in enumerate(self.factors):
            front = torch.moveaxis(E, axis, 0)
            out_shape = (f.shape[0], *front.shape[1:])
            prod = f.matmat(front.reshape(f.shape[-1], -1)).reshape(out_shape)
            E = torch.moveaxis(prod, 0, axis)
        return E.reshape(self.shape[-2], E.shape[-1])


class BlockDiagOp(Op):  # operators.py:277-320
    def __init__(self, *blocks, multiplicities=None):
        self.mult = [1] * len(blocks) if multiplicities is None else list(multiplicities)
        r = sum(b.shape[0] * c for b, c in zip(blocks, self.mult))
        c_ = sum(b.shape[1] * c for b, c in zip(blocks, self.mult))
        super().__init__((r, c_), blocks[0].dtype)
        self.blocks = blocks

    def matmat(self, X):
        k = X.shape[1]
        lo, pieces = 0, []
        for blk, c in zip(self.blocks, self.mult):
            hi = lo + c * blk.shape[-1]
            inner = blk.matmat(X[lo:hi].T.reshape(k * c, blk.shape[-1]).T)
            pieces.append(inner.T.reshape(k, c * blk.shape[0]).T)
            lo = hi
        return torch.cat(pieces, dim=0)


class KronSumOp(Op):  # operators.py:241-275: out = sum_i (I x ... x M_i x ... x I) X, accumulated in factor order
    def __init__(self, *Ms):
        self.Ms = Ms
        self.dtype = Ms[0].dtype
        self.shape = (int(np.prod([M.shape[0] for M in Ms])), int(np.prod([M.shape[1] for M in Ms])))

    def matmat(self, X):
        ev = X.reshape(*[M.shape[1] for M in self.Ms], -1)
        out = 0 * ev
        for i, M in enumerate(self.Ms):
            front = torch.moveaxis(ev, i, 0)
            Mf = M.matmat(front.reshape(M.shape[1], -1)).reshape(M.shape[0], *front.shape[1:])
            out += torch.moveaxis(Mf, 0, i)
        return out.reshape(self.shape[0], ev.shape[-1])


class TridiagonalOp(Op):  # operators.py:351-372 (alpha lower band, beta diagonal, gamma upper band)
    def __init__(self, alpha, beta, gamma):
        self.alpha, self.beta, self.gamma = alpha.reshape(-1, 1), beta.reshape(-1, 1), gamma.reshape(-1, 1)
        self.dtype = beta.dtype
        self.shape = (beta.shape[0], beta.shape[0])

    def matmat(self, X):
        out = self.beta * X
        zeros = torch.zeros((1, X.shape[-1]), dtype=X.dtype)
        up = torch.concat([self.gamma * X[1:], zeros], dim=0)
        lo = torch.concat([zeros, self.alpha * X[:-1]], dim=0)
        return out + lo + up


# ----------------------------------------------------------------------------------
# loop driver with the `info` contract  (cola/utils/torch_tqdm.py:7-71, 86-90)
# ----------------------------------------------------------------------------------
def _tracked_while(error_of, cond, body, state):
    samples, evals = [], 0
    t0 = time.time()
    while True:
        samples.append(float(error_of(state)))
        evals += 1
        if not bool(cond(state)):
            break
        state = body(state)
    per_iter = (time.time() - t0) / evals
    samples.append(float(error_of(state)))
    info = {"iterations": evals, "errors": np.array(samples[2:]), "iteration_time": per_iter}
    return state, info


def _safe_div(num, den):  # cg.py:173-178
    small = torch.tensor(_TINY, dtype=num.real.dtype)
    den = torch.where(torch.abs(den) < small, _TINY, den)
    return num / den


def _colnorm(M):
    return torch.linalg.norm(M, dim=-2, keepdim=True)


# ----------------------------------------------------------------------------------
# CG  (cola/linalg/inverse/cg.py:94-178)
# ----------------------------------------------------------------------------------
def cg(A, B, x0=None, tol=1e-6, max_iters=1000, P=None):
    """Multi-RHS CG.  B (n,k).  Returns x (n,k), r (n,k), k_iters, info."""
    vec = B.dim() == 1
    if vec:
        B = B[:, None]
    x = torch.zeros_like(B) if x0 is None else (x0[:, None] if x0.dim() == 1 else x0)
    P = IdentityOp(A.shape[0], A.dtype) if P is None else P
    scale = _colnorm(B)                                   # cg.py:96
    Bn = _safe_div(B, scale)                              # cg.py:97
    r = Bn - A.matmat(x)                                  # cg.py:123
    z = P.matmat(r)
    p = z
    gamma = torch.sum(torch.conj(r) * z, dim=-2, keepdim=True)
    tol_eff = tol * _colnorm(r) + tol                     # cg.py:101
    zero = torch.tensor(0.0, dtype=B.dtype)
    eps = torch.tensor(_TINY, dtype=B.real.dtype)

    def cond(s):                                          # cg.py:133-138
        return torch.any(_colnorm(s[2]) > tol_eff) & (s[1] < max_iters)

    def body(s):                                          # cg.py:141-170
        x, k, r, p, gamma = s
        done = _colnorm(r) < eps
        Ap = A.matmat(p)
        alpha = _safe_div(gamma, torch.sum(torch.conj(p) * Ap, dim=-2, keepdim=True))
        alpha = torch.where(done, zero, alpha)
        x = x + alpha * p
        r = r - alpha * Ap
        z = P.matmat(r)
        gamma1 = torch.sum(torch.conj(r) * z, dim=-2, keepdim=True)
        beta = torch.where(done, zero, _safe_div(gamma1, gamma))
        p = z + beta * p
        return (x, k + 1, r, p, gamma1)

    def err(s):                                           # cg.py:113-115
        return torch.linalg.norm(s[2], dim=-2).mean()

    (x, k, r, p, gamma), info = _tracked_while(err, cond, body, (x, 0, r, p, gamma))
    x, r = x * scale, r * scale                           # cg.py:119
    if vec:
        x, r = x.reshape(-1), r.reshape(-1)
    return x, r, k, info


class NystromPrecondOp(Op):
    """cola/linalg/preconditioning/preconditioners.py:97-157: P = U diag(s) U^T + I from the rank-r Nystrom sketch."""
    def __init__(self, A, rank, mu=1e-7, eps=1e-8, adjust_mu=True, key=None):
        super().__init__(A.shape, A.dtype)
        key = PRNGKey(42) if key is None else key
        Omega = keyed_randn(A.shape[0], rank, dtype=A.dtype, key=key)
        Omega, _ = torch.linalg.qr(Omega, mode="reduced")              # get_nys_approx :145-157
        Y = A.matmat(Omega)
        nu = eps * torch.linalg.norm(Y)
        Y = Y + nu * Omega
        C = torch.linalg.cholesky(Omega.T @ Y)
        B = torch.linalg.solve_triangular(C, Y.T, upper=False).T
        U, Sigma, _ = torch.linalg.svd(B, full_matrices=False)
        self.Lambda, self.U = torch.clip(Sigma**2.0 - nu, min=0.0), U
        amu = mu * torch.max(self.Lambda) if adjust_mu else mu        # _create_approx :118-126
        self.scaling = ((torch.min(self.Lambda) + amu) / (self.Lambda + amu) - 1)[:, None]

    def matmat(self, V):                                               # :128-130
        return self.U @ (self.scaling * (self.U.T @ V)) + V


# ----------------------------------------------------------------------------------
# Lanczos with CGS2 full reorthogonalisation  (cola/linalg/decompositions/lanczos.py:185-296)
# ----------------------------------------------------------------------------------
def _cgs_pass(V, w):                                      # lanczos.py:293-296
    c = torch.sum(torch.conj(V) * w.unsqueeze(-1), dim=-2, keepdim=True)
    w -= torch.sum(V * c, dim=-1)
    return w


def lanczos_fact(A, rhs, max_iters=100, tol=1e-7):
    """rhs (n,b).  Returns V (b,n,m+2), diag (b,m), sub (b,m+1), i, info  (lanczos.py:235-284)."""
    n, b = rhs.shape
    m = max_iters
    dtype = A.dtype
    diag = torch.zeros(b, m, dtype=dtype)
    sub = torch.zeros(b, m + 1, dtype=dtype)
    V = torch.zeros(b, n, m + 2, dtype=dtype)
    V[..., 1] = torch.clone((rhs / _colnorm(rhs)).T)

    def body(s):
        V, diag, sub, i = s
        V[..., i] = V[..., i] / torch.linalg.norm(V[..., i], dim=-1, keepdim=True)
        w = A.matmat(torch.permute(V, [1, 0, 2])[..., i]).T
        diag[..., i - 1] = torch.sum(torch.conj(w) * V[..., i], dim=-1)
        w -= diag[..., [i - 1]] * V[..., i] + sub[..., [i - 1]] * V[..., i - 1]
        w = _cgs_pass(V, w)
        w = _cgs_pass(V, w)
        V[..., i + 1] = w
        sub[..., i] = torch.linalg.norm(V[..., i + 1], dim=-1)
        return V, diag, sub, i + 1

    def err(s):
        _, _, sub, i = s
        floor = torch.tensor(1e-30, dtype=sub.real.dtype)
        rel = sub[..., i - 1].real / torch.maximum(sub[..., 1].real, floor)
        return torch.max(rel, dim=0)[0] + (i <= 1) * 1.

    def cond(s):
        _, _, sub, i = s
        large = (sub[..., i - 1].real > tol * sub[..., 1].real) | (i <= 1)
        return (i <= m) & torch.any(large)

    (V, diag, sub, i), info = _tracked_while(err, cond, body, (V, diag, sub, 1))
    return V, diag, sub, i, info


def lanczos(A, start, max_iters=100, tol=1e-7):
    """start (n,) or (n,b).  Returns Q (b,n,iters), alpha (b,iters-1) off-diagonal,
    beta (b,iters) diagonal, info; batch dim dropped for a 1-D start (lanczos.py:185-232)."""
    max_iters = min(max_iters, A.shape[0])
    rhs = start[:, None] if start.dim() == 1 else start
    V, diag, sub, i, info = lanczos_fact(A, rhs, max_iters, tol)
    iters = i - 1
    alpha = sub[..., 1:-1][..., :iters - 1]
    beta = diag[..., :iters]
    Q = V[..., 1:-1][..., :iters]
    if start.dim() == 1:
        return Q[0], alpha[0], beta[0], info
    return Q, alpha, beta, info


def tridiag_dense(alpha, beta):
    """Dense T from off-diagonal alpha (..., m-1) and diagonal beta (..., m)
    (operators.py:351-372 Tridiagonal.to_dense semantics)."""
    T = torch.diag_embed(beta)
    if alpha.shape[-1] > 0:
        T = T + torch.diag_embed(alpha, offset=1) + torch.diag_embed(alpha, offset=-1)
    return T


def lanczos_eigs(A, start, max_iters=100, tol=1e-7):      # lanczos.py:34-61
    Q, alpha, beta, info = lanczos(A, start, max_iters, tol)
    lam, S = torch.linalg.eigh(tridiag_dense(alpha, beta))
    order = torch.argsort(lam, dim=-1)
    return lam[..., order], Q @ S[:, order], info


# ----------------------------------------------------------------------------------
# Arnoldi with modified Gram-Schmidt  (cola/linalg/decompositions/arnoldi.py:166-205, 289-335)
# ----------------------------------------------------------------------------------
def arnoldi_fact(A, rhs, max_iters=100, tol=1e-7):
    n, b = rhs.shape
    m = max_iters
    dtype = A.dtype
    H = torch.zeros(b, m + 1, m, dtype=dtype)
    Q = torch.zeros(b, n, m + 1, dtype=dtype)
    nrm = torch.linalg.norm(rhs, dim=-2)
    Q[..., 0] = torch.clone((rhs / nrm).T)
    m_eff = min(m, A.shape[0])

    def cond(s):
        _, H, j, nrm = s
        return (j < m_eff) & torch.any((nrm > tol * H[:, 1, 0].real) | (j <= 0))

    def body(s):
        Q, H, j, _ = s
        w = A.matmat(Q[..., j].T).T
        h = torch.zeros(H.shape[0], H.shape[1], dtype=w.dtype)
        for t in range(0, j + 1):
            h[..., t] = torch.sum(torch.conj(Q[..., t]) * w, dim=-1)
            w = w - h[..., [t]] * Q[..., t]
        nrm = torch.linalg.norm(w, dim=-1, keepdim=True)
        w /= torch.clip(nrm, min=tol / 2.)
        h[..., j + 1] = nrm[:, 0]
        H[..., j] = h
        Q[..., j + 1] = w
        return Q, H, j + 1, nrm[:, 0]

    (Q, H, j, _), info = _tracked_while(lambda s: s[-1][0], cond, body, (Q, H, 0, nrm))
    return Q, H, j, info


def arnoldi(A, start, max_iters=100, tol=1e-7):
    rhs = start[:, None] if start.dim() == 1 else start
    Q, H, _, info = arnoldi_fact(A, rhs, max_iters, tol)
    if start.dim() == 1:
        return Q[0], H[0], info
    return Q, H, info


def gmres(A, rhs, x0=None, max_iters=100, tol=1e-7):
    """cola/linalg/inverse/gmres.py:41-124 (use_householder=False, use_triangular=False; P is accepted by the
    reference but never used, gmres.py:92-124).  rhs (n,) or (n,b) -> (soln, info).  Note the reference's own
    formulation: it drops the LAST ROW of the (m+1, m) Hessenberg and solves the normal equations of the
    square part, `(H^H H + D) y = H^H[:, 0] * beta`, D = identity on the rows that stayed zero."""
    is_vec = rhs.dim() == 1
    if x0 is None:
        x0 = torch.zeros_like(rhs)
    if is_vec:
        rhs, x0 = rhs[..., None], x0[..., None]
    res = rhs - A.matmat(x0)
    m = max_iters
    Q, H, _, info = arnoldi_fact(A, res, m, tol)                  # arnoldi() hands back the untrimmed arrays
    Q, H = Q[:, :, :-1], H[:, :-1, :]
    beta = torch.linalg.norm(res, dim=-2)
    HT = torch.conj(torch.permute(H, [0, 2, 1]))
    largest = torch.max(torch.abs(H), -1)[0]
    overall = torch.max(largest.reshape(largest.shape[0], -1), -1)[0]
    thresh = 10 * tol * overall[:, None]
    padding = torch.where(largest < thresh, torch.ones_like(largest), torch.zeros_like(largest))
    y = torch.linalg.solve(HT @ H + torch.diag_embed(padding), HT[..., 0, None]).squeeze(-1) * beta[:, None]
    y = torch.where(largest < thresh, torch.zeros_like(y), y)
    pred = torch.permute(Q @ y[..., None], [1, 0, 2])[:, :, 0]
    soln = x0 + pred
    return (soln[:, 0] if is_vec else soln), info


def power_iteration(A, tol=1e-6, max_iter=1000, key=None):
    """cola/linalg/eig/power_iteration.py:35-81 (momentum=None) -> (v, eigmax, info)."""
    key = PRNGKey(42) if key is None else key
    v = keyed_randn(A.shape[-1], dtype=A.dtype, key=key)

    def body(s):
        i, v, vprev, eig, eigprev = s
        p = A.matmat(v.reshape(-1, 1)).reshape(-1)
        eig, eigprev = v @ p, eig
        return i + 1, p / torch.linalg.norm(p), v, eig, eigprev

    def err(s):
        *_, eig, eigprev = s
        return abs(eigprev - eig) / eig

    def cond(s):
        return (s[0] < max_iter) & (err(s) > tol)

    eig0, eigprev0 = torch.tensor(10., dtype=A.dtype), torch.tensor(1., dtype=A.dtype)
    (_, v, _, emax, _), info = _tracked_while(err, cond, body, (0, v, v, eig0, eigprev0))
    return v, emax, info


def arnoldi_eigs(A, start, max_iters=100, tol=1e-7):      # arnoldi.py:35-62
    Q, H, info = arnoldi(A, start, max_iters, tol)
    Q, H = Q[:, :-1], H[:-1]
    lam, S = torch.linalg.eig(H)
    return lam, Q.to(S.dtype) @ S, info


# ----------------------------------------------------------------------------------
# f(A) V through Lanczos  (cola/linalg/unary/unary.py:37-60)
# ----------------------------------------------------------------------------------
def lanczos_unary_matmat(A, f, Vin, max_iters=100, tol=1e-7):
    Q, alpha, beta, info = lanczos(A, Vin, max_iters, tol)
    lam, P = torch.linalg.eigh(tridiag_dense(alpha, beta))
    norms = torch.linalg.norm(Vin, dim=0)
    thresh = 10 * torch.finfo(A.dtype).eps * torch.max(torch.abs(lam), dim=1, keepdim=True)[0]
    flam = torch.where(torch.abs(lam) > thresh, f(lam), torch.zeros_like(lam))
    coef = torch.conj(P)[:, 0, :] * norms[:, None]
    out = (Q @ P @ (flam * coef)[..., None])[..., 0]
    return out.T, info

# ----------------------------------------------------------------------------------
# f(A) V through Arnoldi  (cola/linalg/unary/unary.py:63-91); the result is complex
# ----------------------------------------------------------------------------------
def arnoldi_unary_matmat(A, f, Vin, max_iters=100, tol=1e-7):
    Q, H, info = arnoldi(A, Vin, max_iters, tol)          # batched: Q (b, n, m+1), H (b, m+1, m)
    Q, H = Q[:, :, :-1], H[:, :-1]
    lam, P = torch.linalg.eig(H)
    norms = torch.linalg.norm(Vin, dim=0)
    e0 = torch.zeros(P.shape[1], Vin.shape[-1], dtype=P.dtype)
    e0[0] = 1.0
    Pinv0 = torch.linalg.solve(P, e0.T[..., None]).squeeze(-1)
    coef = Pinv0 * norms[:, None]
    thresh = 10 * torch.finfo(A.dtype).eps * torch.max(torch.abs(lam), dim=1, keepdim=True)[0]
    flam = torch.where(torch.abs(lam) > thresh, f(lam), torch.zeros_like(lam))
    out = (Q.to(P.dtype) @ P @ (flam * coef)[..., None])[..., 0]
    return out.T, info


# ----------------------------------------------------------------------------------
# exp / log / sqrt / isqrt as lazy operators  (cola/linalg/unary/unary.py:37-91, 229-335)
# ----------------------------------------------------------------------------------
UNARY_FUNS = {"exp": torch.exp, "log": torch.log, "sqrt": lambda x: x**0.5, "isqrt": lambda x: x**-0.5}


class LanczosUnaryOp(Op):                                  # unary.py:37-60
    def __init__(self, A, f, max_iters, tol):
        super().__init__(A.shape, A.dtype)
        self.A, self.f, self.max_iters, self.tol = A, f, max_iters, tol

    def matmat(self, X):
        return lanczos_unary_matmat(self.A, self.f, X, self.max_iters, self.tol)[0]


class ArnoldiUnaryOp(LanczosUnaryOp):                      # unary.py:63-91
    def matmat(self, X):
        return arnoldi_unary_matmat(self.A, self.f, X, self.max_iters, self.tol)[0]


def unary_operator(fn, A, alg, max_iters, tol):
    """The dispatch rules the fixtures exercise: exp(KronSum) = Kronecker of exp(factor) (unary.py:244-246),
    pow(Kronecker, a) = Kronecker of pow(factor, a) (:303-305), otherwise LanczosUnary / ArnoldiUnary (:134-142)."""
    if fn == "exp" and isinstance(A, KronSumOp):
        return KroneckerOp(*[unary_operator(fn, M, alg, max_iters, tol) for M in A.Ms])
    if fn in ("sqrt", "isqrt") and isinstance(A, KroneckerOp):
        return KroneckerOp(*[unary_operator(fn, M, alg, max_iters, tol) for M in A.factors])
    cls = LanczosUnaryOp if alg == "lanczos" else ArnoldiUnaryOp
    return cls(A, UNARY_FUNS[fn], max_iters, tol)


# ----------------------------------------------------------------------------------
# Hutchinson / exact diagonals  (cola/linalg/trace/diagonal_estimation.py:84-128, 158-210)
# ----------------------------------------------------------------------------------
def hutchinson_diag(matmat, n, dtype, tol=3e-2, max_iters=10000, rand="normal", key=None, k=0):
    bs = min(100, n)
    assert tol > 1e-3, "tolerance chosen too high for stochastic diagonal estimation"
    assert rand in ["normal", "rademacher"], "rand must be 'normal' or 'rademacher'"
    key = PRNGKey(42) if key is None else key

    def body(s):
        i, s1, s2, key = s
        key = next_key(key)
        z = keyed_randn(n, bs, dtype=dtype, key=key)
        if rand == "rademacher":
            z = torch.sign(z)
        z2 = torch.roll(z, -k, 0)                                         # diagonal_estimation.py:190-193
        z2[slice(0, abs(k)) if k <= 0 else slice(-abs(k), None)] = 0
        slc = slice(abs(k), None) if -k > 0 else slice(None, -abs(k) or None)
        est = (matmat(z) * z2)[slc]
        return i + 1, s1 + est.sum(-1), s2 + (est**2).sum(-1), key

    def err(s):
        i, s1, s2, _ = s
        mean = s1 / (i * bs)
        se = torch.sqrt((s2 / (i * bs) - mean**2) / (i * bs))
        return torch.mean(se / torch.maximum(torch.abs(mean), .1 * torch.ones_like(mean)))

    def cond(s):
        return (s[0] == 0) | ((s[0] < max_iters) & (err(s) > tol))

    zeros = torch.zeros(n - abs(k), dtype=dtype)
    (i, s1, _, _), info = _tracked_while(err, cond, body, (0, zeros, zeros, key))
    return s1 / (i * bs), info


def exact_diag(matmat, n, dtype, k=0):
    """cola/linalg/trace/diagonal_estimation.py:84-128: blocks of 100 identity columns, the k-th diagonal picked out
    with a shifted copy of the block (get_I_chunk_like)."""
    bs = min(100, n)
    eye = torch.eye(n, dtype=dtype)
    total = 0.
    for i in range(0, n, bs):
        if k == 0:
            chunk = shifted = eye[:, i:i + bs]
        elif k <= 0:
            kk = abs(k)
            I_chunk = eye[:, i:i + bs + kk]
            padded = torch.zeros(n, bs + kk, dtype=dtype)
            padded[:, :I_chunk.shape[-1]] = I_chunk
            chunk, shifted = I_chunk[:, :bs], padded[:, kk:kk + bs]
        else:
            I_chunk = eye[:, max(i - k, 0):i + bs]
            padded = torch.zeros(n, bs + k, dtype=dtype)
            padded[:, -I_chunk.shape[-1]:] = I_chunk
            chunk, shifted = I_chunk[:, -bs:], padded[:, :bs]
        total = total + (matmat(chunk) * shifted).sum(-1)
    return total[abs(k):] if k <= 0 else total[:(-k or None)]


def diag(A, k=0, alg="hutch", **hutch_kw):
    """cola/linalg/trace/diag_trace.py:56-119: structure rules first -- a Sum is taken term by term, so Dense /
    Diagonal / Kronecker-of-those terms give exact diagonals whatever the algorithm -- then Hutch or Exact on what is
    left (here: CSR, Product, scaled operators; `c * Identity` is a Product[ScalarMul, Identity] in the reference and
    has no rule of its own)."""
    if isinstance(A, DenseOp):
        return torch.diag(A.M, diagonal=k)
    if isinstance(A, (IdentityOp, DiagonalOp)):
        if k == 0:
            return A.d if isinstance(A, DiagonalOp) else torch.ones(A.shape[0], dtype=A.dtype)
        return torch.zeros(A.shape[0] - abs(k), dtype=A.dtype)
    if isinstance(A, SumOp):
        return sum(diag(M, k, alg, **hutch_kw) for M in A.terms)
    if isinstance(A, BlockDiagOp):
        assert k == 0
        return torch.concat([d for M, m in zip(A.blocks, A.mult) for d in [diag(M, k, alg, **hutch_kw)] * m])
    if isinstance(A, (KroneckerOp, KronSumOp)):
        assert k == 0
        Ms = A.factors if isinstance(A, KroneckerOp) else A.Ms
        ds = [diag(M, k, alg, **hutch_kw) for M in Ms]
        slices = [[None] * i + [slice(None)] + [None] * (len(ds) - i - 1) for i in range(len(ds))]
        parts = [d[tuple(sl)] for d, sl in zip(ds, slices)]
        out = parts[0]
        for p in parts[1:]:
            out = out * p if isinstance(A, KroneckerOp) else out + p
        return out.reshape(-1)
    if alg == "exact":
        return exact_diag(A.matmat, A.shape[0], A.dtype, k)
    return hutchinson_diag(A.matmat, A.shape[0], A.dtype, k=k, **hutch_kw)[0]


# ----------------------------------------------------------------------------------
# SLQ  (cola/linalg/tbd/slq.py:37-75)
# ----------------------------------------------------------------------------------
def slq(A, f, max_iters=100, tol=1e-5, vtol=0.1, key=None, probes=None):
    num = max(int(1 / vtol**2), 1)
    eps = torch.finfo(A.dtype).eps
    Z = keyed_randn(A.shape[1], num, dtype=A.dtype, key=key) if probes is None else probes
    _, alpha, beta, _ = lanczos(A, Z, max_iters, tol)
    lam, S = torch.linalg.eigh(tridiag_dense(alpha, beta))
    tau = S[..., 0, :]
    cut = 10 * eps * torch.max(lam, dim=1, keepdim=True)[0]
    flam = torch.where(torch.abs(lam) > cut, f(lam), torch.zeros_like(lam))
    est = A.shape[-2] * torch.sum(tau**2 * flam, dim=-1)
    return torch.mean(est, dim=0)


def slq_per_probe(A, f, Z, max_iters=100, tol=1e-5):
    """Per-probe quadrature values n * sum_j tau_j^2 f(lambda_j) for probe block Z (n,b):
    what slq() averages.  Used to time / check probe chunks (work is linear in probes)."""
    eps = torch.finfo(A.dtype).eps
    _, alpha, beta, _ = lanczos(A, Z, max_iters, tol)
    lam, S = torch.linalg.eigh(tridiag_dense(alpha, beta))
    tau = S[..., 0, :]
    cut = 10 * eps * torch.max(lam, dim=1, keepdim=True)[0]
    flam = torch.where(torch.abs(lam) > cut, f(lam), torch.zeros_like(lam))
    return A.shape[-2] * torch.sum(tau**2 * flam, dim=-1)


# ----------------------------------------------------------------------------------
# backward passes  (cola/linalg/inverse/cg.py:72-86, cola/linalg/tbd/slq.py:10-31)
# ----------------------------------------------------------------------------------
def _param_vjp(make_op, params, V, G):
    """xnp.vjp_derivs(fun = theta -> A(theta) @ V, primals = theta, duals = G) (cola/backends/torch_fns.py:244-260, real
    dtypes: the conjugations are no-ops): torch autograd through the operator's eager matmat.  `make_op(params)` builds
    the operator from the parameter tensors.  A SparseOp is rebuilt from its dense equivalent here: torch's CSR SpMM has
    no autograd (the reference's own backward raises on Sparse for that reason, tests/golden/make_golden_bwd.py), the
    gradient of the dense equivalent restricted to the pattern is what the rule means."""
    ps = [p.detach().clone().requires_grad_(True) for p in params]
    out = _differentiable(make_op(ps)).matmat(V)
    return torch.autograd.grad(out, ps, grad_outputs=G, allow_unused=True)


def _differentiable(A):
    """The same operator with every SparseOp applied through index_add on its COO triplets (differentiable in the
    values), recursively through the composite operators."""
    if isinstance(A, SparseOp):
        rows = torch.repeat_interleave(torch.arange(A.shape[0]), (A.indptr[1:] - A.indptr[:-1]).to(torch.int64))
        cols = A.indices.to(torch.int64)
        data = A.data

        class _Coo(Op):
            def matmat(self, X):
                return torch.zeros((A.shape[0], X.shape[1]), dtype=X.dtype).index_add(0, rows, data[:, None] * X[cols])

        return _Coo(A.shape, A.dtype)
    for attr in ("terms", "factors", "blocks", "Ms"):
        if hasattr(A, attr):
            setattr(A, attr, [_differentiable(t) for t in getattr(A, attr)])
    if isinstance(A, ScaledOp):
        A.A = _differentiable(A.A)
    return A


def cg_bwd(make_op, params, soln, dy, x0=None, tol=1e-6, max_iters=1000, P=None):
    """cg.py:72-86: db = run_batched_cg(A, dy, x0, max_iters, tol, P); dA = vjp(theta -> A(theta) @ soln, -db).
    Returns (d_params, db)."""
    A = make_op([p.detach() for p in params])
    db, *_ = cg(A, dy, x0=x0, tol=tol, max_iters=max_iters, P=P)
    return _param_vjp(make_op, params, soln, -db), db


def slq_bwd(make_op, params, g, num_samples, key=None):
    """slq.py:10-31: probes re-drawn from the key, solves = cg(A, probes, tol=1e-6, max_iters=100), dA = vjp(theta ->
    A(theta) @ probes, g / num_samples * solves).  (As the reference notes, this assumes f = log.)"""
    A = make_op([p.detach() for p in params])
    key = sha_key(0) if key is None else key
    probes = keyed_randn(A.shape[1], num_samples, dtype=A.dtype, key=key)
    solves, *_ = cg(A, probes, tol=1e-6, max_iters=100)
    return _param_vjp(make_op, params, probes, (1.0 / num_samples) * g * solves)
