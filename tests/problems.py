"""Seeded synthetic problems shared by the golden generator, the oracle tests and the GPU
parity tests.  Inputs come from numpy's legacy RandomState (stable across versions), so
the three sides see bit-identical matrices and right-hand sides.

A problem's operator is a small nested *spec* (tuples), turned into
  * reference operators by tests/golden/make_golden.py   (build container only),
  * oracle operators by `to_oracle`                        (CPU checker),
  * cola_b200 operators by `to_b200`                       (the CUDA product path).
"""
import numpy as np
import torch

NP = {torch.float32: np.float32, torch.float64: np.float64}


def rs(seed):
    return np.random.RandomState(seed)


def t(a, dtype):
    return torch.tensor(np.ascontiguousarray(a), dtype=dtype)


# ---------------------------------------------------------------------------- generators
def spd_dense(n, dtype, seed, lo=1e-2, coeff=0.75):
    """Random orthogonal Q, geometric-ish spectrum in (lo, 1+lo]; symmetrised.  Same construction idea as
    the reference's generate_pd_from_diag(generate_spectrum(...)) (cola/utils/utils_for_tests.py:164-196)."""
    g = rs(seed)
    Q, _ = np.linalg.qr(g.normal(size=(n, n)))
    spec = np.sort(coeff**(np.arange(n) * 40.0 / n))[::-1] + lo
    A = (Q * spec) @ Q.T
    A = 0.5 * (A + A.T)
    return t(A, dtype)


def nonsym_dense(n, dtype, seed):
    g = rs(seed)
    A = g.normal(size=(n, n)) / np.sqrt(n) + np.diag(np.linspace(1.0, 3.0, n))
    return t(A, dtype)


def laplacian_2d_coo(g, dtype, shift=0.0):
    """5-point Laplacian on a g x g grid: kron(I,T)+kron(T,I), T=tridiag(-1,2,-1),
    entries sorted by (row, col).  Returns data, rows, cols (int64), shape."""
    n = g * g
    idx = np.arange(n, dtype=np.int64)
    ix, iy = idx // g, idx % g
    rows = [idx, idx[iy > 0], idx[iy < g - 1], idx[ix > 0], idx[ix < g - 1]]
    cols = [idx, idx[iy > 0] - 1, idx[iy < g - 1] + 1, idx[ix > 0] - g, idx[ix < g - 1] + g]
    vals = [np.full(n, 4.0 + shift)] + [np.full(len(r), -1.0) for r in rows[1:]]
    rows, cols, vals = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
    order = np.lexsort((cols, rows))
    return (t(vals[order], dtype), torch.tensor(rows[order]), torch.tensor(cols[order]), (n, n))


def graph_laplacian_coo(n, deg_pairs, dtype, seed):
    """Random undirected graph Laplacian L = D - W with deg_pairs*n random pairs, symmetrised
    (SURVEY 8d cfg5).  Duplicate pairs are merged (weights summed)."""
    g = rs(seed)
    a = g.randint(0, n, size=deg_pairs * n).astype(np.int64)
    b = g.randint(0, n, size=deg_pairs * n).astype(np.int64)
    keep = a != b
    a, b = a[keep], b[keep]
    r = np.concatenate([a, b])
    c = np.concatenate([b, a])
    key = r * n + c
    key, counts = np.unique(key, return_counts=True)
    r, c, w = key // n, key % n, counts.astype(np.float64)
    degree = np.bincount(r, weights=w, minlength=n)
    rows = np.concatenate([r, np.arange(n, dtype=np.int64)])
    cols = np.concatenate([c, np.arange(n, dtype=np.int64)])
    vals = np.concatenate([-w, degree])
    order = np.lexsort((cols, rows))
    return (t(vals[order], dtype), torch.tensor(rows[order]), torch.tensor(cols[order]), (n, n))


def kron_factor(d, dtype, seed):
    """K = G G^T / d + 0.5 I  (SURVEY 8d cfg3/cfg4)."""
    G = rs(seed).normal(size=(d, d))
    return t(G @ G.T / d + 0.5 * np.eye(d), dtype)


def randn_np(shape, dtype, seed):
    return t(rs(seed).normal(size=shape), dtype)


# ---------------------------------------------------------------------------- named problems
def problem(name):
    """Returns dict(spec=..., ann='psd'|'sa'|None, dtype=..., plus case-specific inputs)."""
    f32, f64 = torch.float32, torch.float64
    if name == "cfg1_dense1024":      # BASELINE config 1
        return dict(spec=("dense", spd_dense(1024, f32, 21)), ann="psd", dtype=f32, B=randn_np((1024, ), f32, 1))
    if name == "dense96_f32":
        return dict(spec=("dense", spd_dense(96, f32, 3)), ann="psd", dtype=f32, B=randn_np((96, 5), f32, 4))
    if name == "dense96_f64":
        return dict(spec=("dense", spd_dense(96, f64, 3)), ann="psd", dtype=f64, B=randn_np((96, 5), f64, 4))
    if name in ("lap24_f32", "lap24_f64"):
        dt = f32 if name.endswith("f32") else f64
        return dict(spec=("csr", *laplacian_2d_coo(24, dt)), ann="psd", dtype=dt, B=randn_np((576, 8), dt, 0))
    if name == "lap16_shift_f32":     # CSR + c*I + Diagonal composition
        d = t(rs(5).uniform(0.5, 1.5, size=256), f32)
        spec = ("sum", [("csr", *laplacian_2d_coo(16, f32)), ("scaled_identity", 0.25, 256), ("diag", d)])
        return dict(spec=spec, ann="psd", dtype=f32, B=randn_np((256, 3), f32, 6))
    if name == "kron888_f32":         # BASELINE config 3, scaled down
        fs = [("dense", kron_factor(8, f32, i)) for i in range(3)]
        spec = ("sum", [("kron", fs), ("scaled_identity", 0.1, 512)])
        return dict(spec=spec, ann="psd", dtype=f32, B=randn_np((512, 16), f32, 0))
    if name == "kron465_diag_f64":    # BASELINE config 4 shape (Kronecker + Diagonal), ragged dims
        fs = [("dense", kron_factor(d, f64, i)) for i, d in enumerate((4, 6, 5))]
        dg = t(rs(3).uniform(size=120) + 0.5, f64)
        spec = ("sum", [("kron", fs), ("diag", dg)])
        return dict(spec=spec, ann="psd", dtype=f64, B=randn_np((120, 7), f64, 2))
    if name == "kron884_diag_f32":
        fs = [("dense", kron_factor(d, f32, i)) for i, d in enumerate((8, 8, 4))]
        dg = t(rs(3).uniform(size=256) + 0.5, f32)
        spec = ("sum", [("kron", fs), ("diag", dg)])
        return dict(spec=spec, ann="psd", dtype=f32, B=randn_np((256, 16), f32, 2))
    if name == "blockdiag_f32":
        blocks = [("dense", spd_dense(6, f32, 11)), ("dense", spd_dense(4, f32, 12)), ("dense", spd_dense(9, f32, 13))]
        spec = ("blockdiag", blocks, [2, 3, 1])
        return dict(spec=spec, ann="psd", dtype=f32, B=randn_np((33, 4), f32, 8))
    if name == "product_f64":         # Product[Dense, Dense] + scaled core
        M = nonsym_dense(40, f64, 14)
        spec = ("sum", [("product", [("dense", M.T.contiguous()), ("dense", M)]), ("scale", 0.5, ("dense", spd_dense(40, f64, 15)))])
        return dict(spec=spec, ann="psd", dtype=f64, B=randn_np((40, 3), f64, 9))
    if name == "nonsym48_f32":
        return dict(spec=("dense", nonsym_dense(48, f32, 17)), ann=None, dtype=f32, B=randn_np((48, 3), f32, 18))
    if name == "nonsym48_f64":
        return dict(spec=("dense", nonsym_dense(48, f64, 17)), ann=None, dtype=f64, B=randn_np((48, 3), f64, 18))
    if name == "graph2k_f64":         # BASELINE config 5, scaled down
        return dict(spec=("csr", *graph_laplacian_coo(2048, 8, f64, 7)), ann="sa", dtype=f64,
                    B=randn_np((2048, ), f64, 19))
    if name == "kron888_pure_f32":    # a bare Kronecker product: the structure rules of pow / inv / logdet apply
        fs = [("psd", ("dense", kron_factor(8, f32, i))) for i in range(3)]
        return dict(spec=("kron", fs), ann="psd", dtype=f32, B=randn_np((512, 16), f32, 0))
    if name in ("kronsum465_f64", "kronsum884_f32"):   # KronSum (operators.py:241-275)
        dt = f64 if name.endswith("f64") else f32
        dims = (4, 6, 5) if dt == f64 else (8, 8, 4)
        fs = [("psd", ("dense", kron_factor(d, dt, 20 + i))) for i, d in enumerate(dims)]   # annotated factors
        n = int(np.prod(dims))
        return dict(spec=("kronsum", fs), ann="psd", dtype=dt, B=randn_np((n, 7 if dt == f64 else 16), dt, 21))
    if name == "tridiag200_f64":      # non-symmetric Tridiagonal (operators.py:351-372)
        g = rs(22)
        spec = ("tridiag", t(g.normal(size=199), f64), t(g.normal(size=200) + 3.0, f64), t(g.normal(size=199), f64))
        return dict(spec=spec, ann=None, dtype=f64, B=randn_np((200, 5), f64, 23))
    if name == "tridiag200_shift_f32":   # symmetric Tridiagonal + c*I + Diagonal: the CSR core with a fused epilogue
        g = rs(24)
        off = t(g.uniform(-1.0, 1.0, size=199), f32)
        spec = ("sum", [("tridiag", off, t(g.uniform(2.5, 3.5, size=200), f32), off), ("scaled_identity", 0.5, 200),
                        ("diag", t(g.uniform(0.0, 1.0, size=200), f32))])
        return dict(spec=spec, ann="psd", dtype=f32, B=randn_np((200, 6), f32, 25))
    raise KeyError(name)


def upcast(spec):
    """Same spec with every floating tensor promoted to float64 (oracle sensitivity probes)."""
    if torch.is_tensor(spec):
        return spec.double() if spec.is_floating_point() else spec
    if isinstance(spec, (list, tuple)):
        return type(spec)(upcast(v) for v in spec)
    return spec


def stable_window(name, tol, max_iters, rtol):
    """Number of leading entries of info['errors'] over which the CG residual trace of problem `name` is
    numerically well-posed at tolerance rtol: the oracle is re-run on a perturbed copy of the problem
    (fp32 problems: in float64; fp64 problems: right-hand side perturbed by 1e-15 relative) and the window
    ends where the two oracle traces stop agreeing to rtol/4.  Beyond it CG's own rounding sensitivity, not
    the implementation, decides the digits (e.g. tiny block-diagonal systems after Krylov exhaustion)."""
    from oracle import krylov_oracle as ko
    P = problem(name)
    _, _, _, base = ko.cg(to_oracle(P["spec"]), P["B"], tol=tol, max_iters=max_iters)
    if P["dtype"] == torch.float32:
        _, _, _, other = ko.cg(to_oracle(upcast(P["spec"])), P["B"].double(), tol=tol, max_iters=max_iters)
    else:
        noise = 1.0 + 1e-15 * t(rs(999).normal(size=tuple(P["B"].shape)), torch.float64)
        _, _, _, other = ko.cg(to_oracle(P["spec"]), P["B"] * noise, tol=tol, max_iters=max_iters)
    a, b = base["errors"], other["errors"]
    m = min(len(a), len(b))
    bad = np.nonzero(np.abs(a[:m] - b[:m]) > 0.25 * rtol * np.abs(b[:m]))[0]
    return int(bad[0]) if len(bad) else m


# ---------------------------------------------------------------------------- spec -> operators
def to_oracle(spec):
    from oracle import krylov_oracle as ko
    kind = spec[0]
    if kind == "dense":
        return ko.DenseOp(spec[1])
    if kind == "csr":
        return ko.SparseOp(*spec[1:])
    if kind == "diag":
        return ko.DiagonalOp(spec[1])
    if kind == "scaled_identity":
        dtype = None
        return ("scaled_identity", spec[1], spec[2])
    if kind == "scale":
        return ko.ScaledOp(spec[1], to_oracle(spec[2]))
    if kind == "kron":
        return ko.KroneckerOp(*[to_oracle(s) for s in spec[1]])
    if kind == "blockdiag":
        return ko.BlockDiagOp(*[to_oracle(s) for s in spec[1]], multiplicities=spec[2])
    if kind == "product":
        return ko.ProductOp(*[to_oracle(s) for s in spec[1]])
    if kind == "psd":
        return to_oracle(spec[1])
    if kind == "kronsum":
        return ko.KronSumOp(*[to_oracle(s) for s in spec[1]])
    if kind == "tridiag":
        return ko.TridiagonalOp(*spec[1:])
    if kind == "sum":
        terms = [to_oracle(s) for s in spec[1]]
        dtype = next(tm.dtype for tm in terms if not isinstance(tm, tuple))
        terms = [ko.ScaledIdentityOp(tm[1], tm[2], dtype) if isinstance(tm, tuple) else tm for tm in terms]
        return ko.SumOp(*terms)
    raise KeyError(kind)


def to_b200(spec, device, ann=None):
    import cola_b200 as cb
    ops = cb.ops
    kind = spec[0]

    def rec(s, dtype_hint=None):
        k = s[0]
        if k == "dense":
            return ops.Dense(s[1].to(device))
        if k == "csr":
            return ops.Sparse(s[1].to(device), s[2].to(device), s[3].to(device), s[4])
        if k == "diag":
            return ops.Diagonal(s[1].to(device))
        if k == "scaled_identity":
            I = ops.Identity((s[2], s[2]), dtype_hint)
            I.to(device)
            return s[1] * I
        if k == "scale":
            return s[1] * rec(s[2])
        if k == "kron":
            return ops.Kronecker(*[rec(x) for x in s[1]])
        if k == "blockdiag":
            return ops.BlockDiag(*[rec(x) for x in s[1]], multiplicities=s[2])
        if k == "product":
            return ops.Product(*[rec(x) for x in s[1]])
        if k == "psd":
            return cb.PSD(rec(s[1]))
        if k == "kronsum":
            return ops.KronSum(*[rec(x) for x in s[1]])
        if k == "tridiag":
            return ops.Tridiagonal(*[v.to(device) for v in s[1:]])
        if k == "sum":
            first = rec(s[1][0])
            out = first
            for x in s[1][1:]:
                out = out + rec(x, first.dtype)
            return out
        raise KeyError(k)

    A = rec(spec)
    if ann == "psd":
        A = cb.PSD(A)
    elif ann == "sa":
        A = cb.SelfAdjoint(A)
    return A
