"""GPU parity tests for the SURVEY 8f rows built after the main path: f(A)v through Lanczos / Arnoldi with the
exp / log / sqrt / isqrt / pow dispatch surface, exact and off-diagonal estimators, KronSum and Tridiagonal
matmats (and CG on them).  Same bar as tests/test_gpu_parity.py: the CUDA path through the C ABI against the CPU
oracle on identical inputs and against fixtures the real reference produced; 1e-5 (fp32) / 1e-10 (fp64)."""
import numpy as np
import pytest
import torch

from tests import problems as pb
from tests import test_gpu_parity as gp
from tests.golden_cases import DIAG_CASES, NEXT_CG_CASES, NEXT_MATMAT_PROBLEMS, UNARY_CASES
from tests.test_gpu_parity import rel, tol_of

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def cb():
    import cola_b200
    assert torch.cuda.is_available()
    cola_b200.backend.lib()  # fails loudly if the extension is missing
    cola_b200.rng.PROBE_DEVICE = "cpu"
    return cola_b200


def crel(a, b):
    """Relative distance that also takes complex arrays."""
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


# ------------------------------------------------------------------------------------------- KronSum / Tridiagonal
@pytest.mark.parametrize("name", NEXT_MATMAT_PROBLEMS)
def test_matmat_kronsum_tridiagonal(name, golden, cb):
    P = pb.problem(name)
    A = pb.to_b200(P["spec"], DEV, P["ann"])
    O = pb.to_oracle(P["spec"])
    t = tol_of(P["dtype"])
    X = pb.randn_np((A.shape[1], 6), P["dtype"], 100)
    g = golden("matmat_" + name)
    assert rel(A @ X.to(DEV), g["Y"]) < t
    assert rel(A @ X[:, 0].contiguous().to(DEV), g["y"]) < t
    X2 = pb.randn_np((A.shape[1], 33), P["dtype"], 101)          # ragged RHS count
    assert rel(A @ X2.to(DEV), O.matmat(X2)) < t
    # fused <x, y> dots, and the transpose (KronSum of transposed factors / swapped bands)
    Xd = X2.to(DEV)
    Y = torch.empty_like(Xd)
    dots = torch.zeros(33, dtype=torch.float64, device=DEV)
    A.matmat_into(Xd, Y, dots=dots)
    assert rel(Y, O.matmat(X2)) < t and rel(dots, (X2.double() * O.matmat(X2).double()).sum(0)) < t
    dense = O.matmat(torch.eye(A.shape[1], dtype=P["dtype"]))
    assert rel(A.T @ Xd, dense.T @ X2) < 10 * t
    assert rel(A.to_dense(), dense) < t


@pytest.mark.parametrize("case", sorted(NEXT_CG_CASES))
def test_cg_on_kronsum_and_tridiagonal(case, golden, cb, monkeypatch):
    monkeypatch.setitem(gp.CG_CASES, case, NEXT_CG_CASES[case])
    gp.test_cg_vs_oracle_and_golden(case, golden, cb)


# ------------------------------------------------------------------------------------------- f(A) v
@pytest.mark.parametrize("case", sorted(UNARY_CASES))
def test_unary_functions(case, golden, cb):
    from oracle import krylov_oracle as ko
    name, fn, alg, m, tol = UNARY_CASES[case]
    P = pb.problem(name)
    A = pb.to_b200(P["spec"], DEV, P["ann"])
    L = cb.linalg
    algo = (L.Arnoldi if alg == "arnoldi" else L.Lanczos)(max_iters=m, tol=tol)
    F = getattr(L, fn)(A, algo)
    g = golden(case)
    assert type(F).__name__ == str(g["kind"])
    Y = F @ P["B"].to(DEV)
    assert tuple(Y.shape) == tuple(g["Y"].shape) and Y.cpu().numpy().dtype == g["Y"].dtype
    Yo = ko.unary_operator(fn, pb.to_oracle(P["spec"]), alg, m, tol).matmat(P["B"])
    # f(A)v inherits the conditioning of the small eigenproblem (eig(H) / eigh(T)): 100x the strict bar, as for
    # the log(A) V fixtures of the main path
    t = 100 * tol_of(P["dtype"])
    assert crel(Y, Yo) < t, crel(Y, Yo)
    assert crel(Y, g["Y"]) < t, crel(Y, g["Y"])
    if alg == "arnoldi":
        # identity that does not depend on any implementation: exp/log/sqrt of the dense matrix
        import scipy.linalg as sl
        Ad = pb.to_oracle(P["spec"]).matmat(torch.eye(A.shape[0], dtype=P["dtype"])).double().numpy()
        dense = {"exp": sl.expm, "log": sl.logm, "sqrt": sl.sqrtm}[fn](Ad) @ P["B"].double().numpy()
        assert crel(Y.to(torch.complex128), dense) < (1e-4 if P["dtype"] == torch.float32 else 1e-6)


def test_unary_structure_rules(cb):
    """unary.py:181-205 and the integer cases of pow (:265-300) that need no solver."""
    L, ops = cb.linalg, cb.ops
    d = torch.linspace(0.5, 2.0, 12, dtype=torch.float64, device=DEV)
    alg = L.Lanczos(max_iters=12, tol=1e-12)
    assert rel(L.exp(ops.Diagonal(d), alg).diag, torch.exp(d)) == 0.0
    assert rel(L.isqrt(ops.Diagonal(d), alg).diag, d**-0.5) < 1e-15
    I = ops.I_like(ops.Diagonal(d))
    X = torch.ones(12, 2, dtype=torch.float64, device=DEV)
    assert rel(L.exp(I, alg) @ X, np.e * np.ones((12, 2))) < 1e-15
    assert rel(L.log(3.0 * I, alg) @ X, np.log(3.0) * np.ones((12, 2))) < 1e-15
    M = pb.spd_dense(6, torch.float64, 31).to(DEV)
    B = ops.BlockDiag(cb.PSD(ops.Dense(M)), ops.Diagonal(d[:4]), multiplicities=[2, 1])
    FB = L.sqrt(B, L.Lanczos(max_iters=6, tol=1e-12))
    assert isinstance(FB, ops.BlockDiag) and FB.multiplicities == [2, 1]
    w, V = torch.linalg.eigh(M)
    sq = (V * w.sqrt()) @ V.T
    ref = torch.block_diag(sq, sq, torch.diag(d[:4].sqrt()))
    Xb = torch.randn(16, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(0)).to(DEV)
    assert rel(FB @ Xb, ref @ Xb) < 1e-9
    A = cb.PSD(ops.Dense(M))
    assert rel(L.pow(A, 2, alg) @ Xb[:6], M @ (M @ Xb[:6])) < 1e-13
    assert isinstance(L.pow(A, 0, alg), ops.Identity)
    with pytest.raises(AssertionError, match="SelfAdjoint"):
        L.sqrt(ops.Dense(M), alg)
    # small PSD operators with Auto take the dense eigh (unary.py:113-131,160-168)
    assert rel(L.sqrt(A) @ Xb[:6], sq @ Xb[:6]) < 1e-12


def test_pow_minus_one_is_a_solve(cb):
    """pow(A, -1, Lanczos) -> inv(A, CG) and pow(A, -1, Arnoldi) -> inv(A, GMRES)  (unary.py:283-297)."""
    L = cb.linalg
    P = pb.problem("dense96_f64")
    A = pb.to_b200(P["spec"], DEV, P["ann"])
    B = P["B"].to(DEV)
    X = L.pow(A, -1, L.Lanczos(max_iters=500, tol=1e-11)) @ B
    assert rel(A @ X, B) < 1e-8
    P = pb.problem("nonsym48_f64")
    A = pb.to_b200(P["spec"], DEV, P["ann"])
    B = P["B"].to(DEV)
    # 30 steps: the reference's GMRES squares the Hessenberg (normal equations), so the full 48-step run on n = 48
    # loses accuracy again (2e-5) -- see the note on the gmres fixtures in tests/golden/make_golden.py
    X = L.pow(A, -1, L.Arnoldi(max_iters=30, tol=1e-12)) @ B
    assert rel(A @ X, B) < 1e-7


# ------------------------------------------------------------------------------------------- diagonals
@pytest.mark.parametrize("case", sorted(DIAG_CASES))
def test_exact_and_offset_diagonals(case, golden, cb):
    name, k, alg = DIAG_CASES[case]
    P = pb.problem(name)
    A = pb.to_b200(P["spec"], DEV, P["ann"])
    L = cb.linalg
    g = golden(case)
    t = tol_of(P["dtype"])
    if alg == "exact":
        d = L.diag(A, k, L.Exact())
        assert rel(d, g["dense_diag"]) < t
    else:
        d = L.diag(A, k, L.Hutch(tol=2e-2, max_iters=4, key=cb.rng.PRNGKey(9)))
    assert tuple(d.shape) == tuple(g["diag"].shape) and rel(d, g["diag"]) < t


def test_exact_diag_ragged_blocks_and_trace(cb):
    """n = 250 is not a multiple of the 100-column block: every offset still comes out exact; trace = sum diag."""
    L, ops = cb.linalg, cb.ops
    M = pb.randn_np((250, 250), torch.float64, 41)
    A = ops.Product(ops.Dense(M.to(DEV)), ops.Dense(torch.eye(250, dtype=torch.float64, device=DEV)))
    for k in (0, 1, -1, 7, -120):
        assert rel(L.diag(A, k, L.Exact()), torch.diagonal(M, offset=k)) < 1e-13
    assert abs(float(L.trace(A, L.Exact())) - float(torch.trace(M))) < 1e-10


def test_slogdet_lanczos_rule(golden, cb):
    """logdet.py:111-117 returns (tr/|tr|, |tr log A|); restated as is (det < 1 here, so the sign is -1)."""
    L = cb.linalg
    P = pb.problem("dense96_f64")
    A = pb.to_b200(P["spec"], DEV, P["ann"])
    g = golden("slogdet_lanczos_dense96_f64")
    sign, mag = L.slogdet(A, L.Lanczos(max_iters=40, tol=1e-12), L.Hutch(tol=2e-2, max_iters=2, key=cb.rng.PRNGKey(42)))
    assert float(sign) == float(g["sign"]) == -1.0
    assert abs(float(mag) - float(g["logdet"])) < 1e-8 * float(g["logdet"])


# ------------------------------------------------------------------------------------------- tensor-core paths
def test_kronsum_tensor_core_path(cb):
    """KronSum with 64x64 fp32 factors runs on the tcgen05 per-mode kernel (3xTF32), every mode contracting X and
    accumulating into Y: fp32-grade accuracy against an fp64 reference, agreement with the exact SIMT contractions,
    fused shift / diagonal / dots, and accumulation behind another core of a Sum."""
    ops = cb.ops
    for D, k in [(2, 32), (2, 96), (3, 64)]:
        Fs = [pb.kron_factor(64, torch.float32, 60 + i) for i in range(D)]
        n = 64**D
        dg = pb.t(pb.rs(10).uniform(size=n) + 0.5, torch.float32)
        KS = ops.KronSum(*[ops.Dense(F.to(DEV)) for F in Fs])
        A = cb.PSD(KS + 0.1 * ops.I_like(KS) + ops.Diagonal(dg.to(DEV)))
        core = A.plan().terms[0][1][0]
        X = pb.randn_np((n, k), torch.float32, 4).to(DEV)
        assert type(core).__name__ == "_KronSumCore" and core._tc_ok(X)
        Y = torch.empty_like(X)
        dots = torch.zeros(k, dtype=torch.float64, device=DEV)
        A.matmat_into(X, Y, dots=dots)
        E = X.double().reshape(*([64] * D), k)
        ref = torch.zeros_like(E)
        for i, F in enumerate(Fs):
            ref = ref + torch.moveaxis(torch.tensordot(F.double().to(DEV), torch.moveaxis(E, i, 0), dims=1), 0, i)
        ref = ref.reshape(n, k) + (0.1 + dg.double().to(DEV))[:, None] * X.double()
        assert rel(Y, ref) < 2e-6, (D, k, rel(Y, ref))
        assert rel(dots, (X.double() * Y.double()).sum(0)) < 1e-6    # fp32 partials over 32 values, fp64 across tiles
        core.use_tensor_cores = False
        Y2 = torch.empty_like(X)
        A.matmat_into(X, Y2)
        core.use_tensor_cores = True
        assert rel(Y, Y2) < 2e-6
        Y3 = torch.full_like(X, 7.0)                             # closed gate: no launch touches Y
        A.matmat_into(X, Y3, gate=torch.tensor([1], dtype=torch.int32, device=DEV))
        assert bool((Y3 == 7.0).all())


def test_tensor_core_epilogue_dots_cover_earlier_terms(cb):
    """Sum[CSR core, Kronecker or KronSum on the tensor cores]: the tensor-core term is applied last with
    accumulate, and its fused <x, y> must be that of the whole operator (what CG's p^T A p needs)."""
    ops = cb.ops
    n, k = 4096, 64
    g = pb.rs(12)
    off = pb.t(g.uniform(-1.0, 1.0, size=n - 1), torch.float32).to(DEV)
    Tri = ops.Tridiagonal(off, pb.t(g.uniform(2.5, 3.5, size=n), torch.float32).to(DEV), off)   # CSR core
    Fs = [ops.Dense(pb.kron_factor(64, torch.float32, 70 + i).to(DEV)) for i in range(2)]
    X = pb.randn_np((n, k), torch.float32, 5).to(DEV)
    for last in (ops.Kronecker(*Fs), ops.KronSum(*Fs)):
        A = Tri + last + 0.25 * ops.I_like(last)
        plan = A.plan()
        assert len(plan.terms) == 2 and plan.terms[1][1][0]._tc_ok(X)
        Y = torch.empty_like(X)
        dots = torch.zeros(k, dtype=torch.float64, device=DEV)
        A.matmat_into(X, Y, dots=dots)
        ref = (Tri.to_dense().double() + last.to_dense().double()) @ X.double() + 0.25 * X.double()
        assert rel(Y, ref) < 2e-6
        assert rel(dots, (X.double() * Y.double()).sum(0)) < 1e-6    # fp32 partials over 32 values, fp64 across tiles


def test_triangular_inverse(cb):
    """inv(Triangular) -> TriangularInv (inv.py:149-165): a triangular solve per application, its transpose, and its
    use as an opaque core inside a fused plan (shift + dots epilogue as a separate sweep)."""
    L, ops = cb.linalg, cb.ops
    g = torch.Generator().manual_seed(0)
    M = (torch.tril(torch.randn(40, 40, dtype=torch.float64, generator=g)) + 4 * torch.eye(40, dtype=torch.float64)).to(DEV)
    X = torch.randn(40, 5, dtype=torch.float64, generator=g).to(DEV)
    T = ops.Triangular(M, lower=True)
    Ti = L.inv(T)
    assert type(Ti).__name__ == "TriangularInv"
    assert rel(M @ (Ti @ X), X) < 1e-13 and rel((X.T @ Ti) @ M, X.T) < 1e-13
    assert T.T.lower is False and rel(M.T @ (L.inv(T.T) @ X), X) < 1e-13
    A = Ti + 0.5 * ops.I_like(Ti)
    ref = torch.linalg.solve_triangular(M, X, upper=False) + 0.5 * X
    dots = torch.zeros(5, dtype=torch.float64, device=DEV)
    Y = torch.empty_like(X)
    A.matmat_into(X, Y, dots=dots)
    assert rel(Y, ref) < 1e-13 and rel(dots, (X * ref).sum(0)) < 1e-13
    # inv(L L^T) = L^-T L^-1, the Cholesky-style inverse of the reference (decompositions.py:147-175)
    S = M @ M.T
    Sinv = L.inv(T.T) @ Ti
    assert rel(S @ (Sinv @ X), X) < 1e-11


def test_pinv_normal_equations(cb):
    """pinv(A, CG) (pinv.py:65-71): CG on A^T A as a two-core Product chain of the plan (transposed Dense core, Dense
    core), applied to A^T b; rectangular operator, checked against the dense least-squares solution."""
    L, ops = cb.linalg, cb.ops
    M = pb.randn_np((300, 40), torch.float64, 93).to(DEV)
    B = pb.randn_np((300, 5), torch.float64, 94).to(DEV)
    A = ops.Dense(M)
    P = L.pinv(A, L.CG(tol=1e-12, max_iters=200))
    assert tuple(P.shape) == (40, 300)
    assert P.Ms[0].Ms[0].A.plan().describe() == "1*DenseCore@DenseCore"
    x = P @ B
    assert rel(x, torch.linalg.lstsq(M, B).solution) < 1e-8
    assert rel(M.T @ (M @ x), M.T @ B) < 1e-9                    # normal equations
    assert float(L.eigmin(cb.PSD(ops.Dense((M.T @ M).contiguous())), L.Lanczos(max_iters=40, tol=1e-12))) > 0


def test_spmv_column_strips(cb, monkeypatch):
    """SpMV-shaped applications of a pattern without locality run column-blocked (ops._CsrCore: vertical strips whose
    slice of X stays in L2, BASELINE config 5).  Forced here by a small strip size: same result as the plain kernel
    and as a dense fp64 product, with the fused epilogue (shift, Diagonal, <x, y> dots), inside a Sum (accumulate)
    and through a Lanczos run."""
    ops = cb.ops
    g = torch.Generator().manual_seed(5)
    n = 3000
    for dt, tol in [(torch.float64, 1e-13), (torch.float32, 2e-6)]:
        rows = torch.randint(0, n, (12 * n, ), generator=g)
        cols = torch.randint(0, n, (12 * n, ), generator=g)
        key = torch.unique(torch.cat([rows * n + cols, cols * n + rows, torch.arange(n) * (n + 1)]))
        rows, cols = key // n, key % n
        vals = torch.randn(rows.numel(), dtype=dt, generator=g)
        vals = torch.where(rows == cols, torch.full_like(vals, 30.0), 0.5 * (vals + torch.zeros_like(vals)))
        Ad = torch.zeros(n, n, dtype=torch.float64)
        Ad[rows, cols] = vals.double()
        Ad = 0.5 * (Ad + Ad.T)
        vals = Ad[rows, cols].to(dt)
        dg = torch.rand(n, dtype=dt, generator=g)
        D2 = torch.randn(n, n, dtype=dt, generator=g) / n
        for k in (1, 2, 4):
            X = torch.randn(n, k, dtype=dt, generator=g).to(DEV)
            outs = []
            for strip_bytes in (1 << 40, 4096):
                monkeypatch.setattr(ops._CsrCore, "SPMV_BLOCK_BYTES", strip_bytes)
                S = ops.Sparse(vals.to(DEV), rows.to(DEV), cols.to(DEV), (n, n))
                A = ops.Dense(D2.to(DEV)) + S + 0.25 * ops.I_like(S) + ops.Diagonal(dg.to(DEV))
                core = A.plan().terms[-1][1][0]
                Y = torch.empty_like(X)
                dots = torch.zeros(k, dtype=torch.float64, device=DEV)
                A.matmat_into(X, Y, dots=dots)
                assert (core._column_strips(k, X.element_size()) is not None) == (strip_bytes == 4096)
                ref = (D2.double() + Ad + torch.diag(0.25 + dg.double())) @ X.double().cpu()
                assert rel(Y, ref) < tol, (dt, k, strip_bytes, rel(Y, ref))
                assert rel(dots, (X.double().cpu() * ref).sum(0)) < tol
                outs.append(Y)
            assert rel(outs[0], outs[1]) < tol
    # a Lanczos run on the strips: same Ritz values as on the plain kernel
    evs = []
    for strip_bytes in (1 << 40, 4096):
        monkeypatch.setattr(ops._CsrCore, "SPMV_BLOCK_BYTES", strip_bytes)
        S = cb.SelfAdjoint(ops.Sparse(vals.double().to(DEV), rows.to(DEV), cols.to(DEV), (n, n)))
        ev, _ = cb.linalg.eig(S, 4, "LM", cb.linalg.Lanczos(max_iters=40, tol=1e-12, key=3))
        evs.append(ev)
    assert rel(evs[0], evs[1]) < 1e-10


@pytest.mark.parametrize("g", [96, 100, 160, 201])
def test_spmm_pipelined_kernel_ragged_grids(g, cb):
    """The software-pipelined SpMM (wide right-hand-side blocks) on grids whose row count is not a multiple of the tile,
    fp32 and fp64, with the fused epilogue -- against a dense-equivalent fp64 product."""
    for dt, k, tol in [(torch.float32, 64, 2e-6), (torch.float64, 8, 1e-13), (torch.float32, 32, 2e-6)]:
        data, rows, cols, shape = pb.laplacian_2d_coo(g, dt)
        n = shape[0]
        S = cb.ops.Sparse(data.to(DEV), rows.to(DEV), cols.to(DEV), shape)
        dg = torch.rand(n, dtype=dt, generator=torch.Generator().manual_seed(g))
        A = S + 0.5 * cb.ops.I_like(S) + cb.ops.Diagonal(dg.to(DEV))
        X = pb.randn_np((n, k), dt, 7).to(DEV)
        Y = torch.empty_like(X)
        dots = torch.zeros(k, dtype=torch.float64, device=DEV)
        A.matmat_into(X, Y, dots=dots)
        ref = torch.sparse_coo_tensor(torch.stack([rows, cols]), data.double(), shape).to(DEV) @ X.double() \
            + (0.5 + dg.double().to(DEV))[:, None] * X.double()
        assert rel(Y, ref) < tol, (g, dt, k, rel(Y, ref))
        assert rel(dots, (X.double() * ref).sum(0)) < tol


def test_mode_contract_tensor_core_path(cb):
    """cola_mode_contract_tc_f32: one mode of a Kronecker chain on tcgen05 (3xTF32) for square 64- / 128-wide factors --
    the modes the fused 64^D kernel does not take (BASELINE config 4: Kronecker(128, 128, 64)).  Against fp64, with the
    operator epilogue on the last mode, and through a Kronecker(128, 64, 128) + Diagonal operator vs the exact SIMT path."""
    be = cb.backend
    g = torch.Generator().manual_seed(11)
    for d, pre, L, k in [(64, 4, 1, 32), (64, 3, 8, 64), (128, 1, 4, 32), (128, 8, 1, 64), (128, 3, 4, 96), (128, 5, 12, 32)]:
        F = (torch.randn(d, d, generator=g) / d**0.5 + 0.5 * torch.eye(d)).to(DEV)
        X = torch.randn(pre * d * L, k, generator=g).to(DEV)
        assert be.mode_contract_tc_ok(F, pre, L, k, X)
        out = torch.full_like(X, float("nan"))
        be.mode_contract_tc(F, pre, L, k, X, out, alpha=1.5)
        ref = 1.5 * torch.einsum("aj,pjlr->palr", F.double(), X.double().reshape(pre, d, L, k)).reshape(pre * d * L, k)
        assert rel(out, ref) < 3e-6, (d, pre, L, k, rel(out, ref))
        if L == 1:
            dg = torch.rand(pre * d, generator=g).to(DEV)
            dots = torch.zeros(k, dtype=torch.float64, device=DEV)
            Y0 = torch.randn(pre * d, k, generator=g).to(DEV)
            Y = Y0.clone()
            be.mode_contract_tc(F, pre, L, k, X, Y, alpha=1.5, shift=0.25, diag=dg, epi_x=X, accumulate=True, dots=dots)
            ref2 = ref + (0.25 + dg.double())[:, None] * X.double() + Y0.double()
            assert rel(Y, ref2) < 3e-6
            assert rel(dots, (X.double() * Y.double()).sum(0)) < 1e-6
    assert not be.mode_contract_tc_ok(torch.eye(96, device=DEV), 4, 1, 32, X)        # other sizes: SIMT tiles
    assert not be.mode_contract_tc_ok(torch.eye(64, device=DEV), 4, 1, 24, X)
    Fs = [(torch.randn(d, d, generator=g) / d**0.5 + 0.5 * torch.eye(d)).to(DEV) for d in (128, 64, 128)]
    n = 128 * 64 * 128
    dg = torch.rand(n, generator=g).to(DEV)
    K = cb.ops.Kronecker(*[cb.ops.Dense(F) for F in Fs])
    A = K + cb.ops.Diagonal(dg)
    X = torch.randn(n, 32, generator=g).to(DEV)
    core = A.plan().terms[0][1][0]
    Y1, Y2 = torch.empty_like(X), torch.empty_like(X)
    d1, d2 = torch.zeros(32, dtype=torch.float64, device=DEV), torch.zeros(32, dtype=torch.float64, device=DEV)
    A.matmat_into(X, Y1, dots=d1)
    core.use_tensor_cores = False
    A.matmat_into(X, Y2, dots=d2)
    core.use_tensor_cores = True
    assert rel(Y1, Y2) < 5e-6 and rel(d1, d2) < 5e-6
    # BlockDiag with 64- / 128-wide blocks: one per-mode launch per distinct block (pre = multiplicity), epilogue offsets
    B1, B2, B3 = Fs[1], Fs[0], (torch.randn(40, 40, generator=g) / 6 + torch.eye(40)).to(DEV)
    BD = cb.ops.BlockDiag(cb.ops.Dense(B1), cb.ops.Dense(B2), cb.ops.Dense(B3), multiplicities=[4, 8, 3])
    nb = 4 * 64 + 8 * 128 + 3 * 40
    dgb = torch.rand(nb, generator=g).to(DEV)
    Ab = BD + 0.5 * cb.ops.I_like(BD) + cb.ops.Diagonal(dgb)
    Xb = torch.randn(nb, 64, generator=g).to(DEV)
    Yb = torch.empty_like(Xb)
    db = torch.zeros(64, dtype=torch.float64, device=DEV)
    Ab.matmat_into(Xb, Yb, dots=db)
    dense = torch.block_diag(*([B1.double()] * 4 + [B2.double()] * 8 + [B3.double()] * 3)) + torch.diag(0.5 + dgb.double())
    refb = dense @ Xb.double()
    assert rel(Yb, refb) < 3e-6 and rel(db, (Xb.double() * refb).sum(0)) < 3e-6



def _banded(n, offsets, dt, g):
    """COO of a banded matrix with the given diagonals (random values, dominant main diagonal)."""
    rows, cols, vals = [], [], []
    for off in offsets:
        r = torch.arange(max(0, -off), min(n, n - off))
        rows.append(r)
        cols.append(r + off)
        vals.append(torch.full((r.numel(), ), 8.0, dtype=dt) if off == 0 else -torch.rand(r.numel(), dtype=dt, generator=g))
    return torch.cat(vals), torch.cat(rows), torch.cat(cols)


def test_spmm_staged_tiles(cb, monkeypatch):
    """Wide blocks on stencil / banded patterns take the staged kernel (csrc/csr_tiled.cu: the X rows a tile of 8 strips
    gathers arrive in shared memory once, as bulk-copied runs; cola_b200/csr_tiles.py builds the tile-local form).  Same
    result as the register-gather kernel and as a dense fp64 product: 2-D tiles with a ragged tail, boundary strips, the
    fused epilogue and accumulate, irregular tiles (gather from global) inside the same launch, in-place value updates,
    and a CG solve (BASELINE config 2's operator at test size)."""
    from cola_b200.csr_tiles import CsrTiles
    ops = cb.ops
    be = cb.backend
    g = torch.Generator().manual_seed(11)
    for dt, tol, ks in [(torch.float64, 1e-13, (16, 32, 64)), (torch.float32, 3e-6, (32, 64))]:
        for n, offsets in [(96 * 100, (-96, -1, 0, 1, 96)), (64 * 130 + 5, (-64, -1, 0, 1, 64)), (9000, (-128, -2, -1, 0, 1, 2, 128))]:
            vals, rows, cols = _banded(n, offsets, dt, g)
            Ad = torch.zeros(n, n, dtype=torch.float64)
            Ad[rows, cols] = vals.double()
            dg = torch.rand(n, dtype=dt, generator=g)
            for k in ks:
                X = torch.randn(n, k, dtype=dt, generator=g).to(DEV)
                outs = []
                for stage in (0, 106 << 10):
                    monkeypatch.setattr(ops._CsrCore, "TILE_STAGE_BYTES", stage)
                    S = ops.Sparse(vals.to(DEV), rows.to(DEV), cols.to(DEV), (n, n))
                    A = S + 0.25 * ops.I_like(S) + ops.Diagonal(dg.to(DEV))
                    core = A.plan().terms[-1][1][0]
                    Y = torch.empty_like(X)
                    dots = torch.zeros(k, dtype=torch.float64, device=DEV)
                    A.matmat_into(X, Y, dots=dots)
                    T = core._tiles(X, Y)
                    assert (T is not None) == (stage > 0), (n, k, stage)
                    if T is not None:
                        assert T.n_tiles2d > 0 and T.n_tiles > T.n_tiles2d       # strips a far diagonal apart + a consecutive-row tail
                    ref = (Ad + torch.diag(0.25 + dg.double())) @ X.double().cpu()
                    assert rel(Y, ref) < tol, (dt, n, k, stage, rel(Y, ref))
                    assert rel(dots, (X.double().cpu() * ref).sum(0)) < max(tol, 1e-6 if stage else tol)
                    S2 = ops.Sparse(vals.to(DEV), rows.to(DEV), cols.to(DEV), (n, n))
                    assert rel((S + 2.0 * S2) @ X, 3.0 * (Ad @ X.double().cpu())) < tol     # second term accumulates into Y
                    outs.append(Y)
                assert rel(outs[0], outs[1]) < tol
    # irregular tiles: a pattern without runs forced through the staged kernel (every tile gathers from global memory),
    # and a banded pattern whose staged rows exceed a deliberately small stage (mixed regular / irregular tiles)
    n = 3000
    r = torch.randint(0, n, (9 * n, ), generator=g)
    c = torch.randint(0, n, (9 * n, ), generator=g)
    key = torch.unique(r * n + c)
    r, c = key // n, key % n
    v = torch.randn(r.numel(), dtype=torch.float64, generator=g)
    vb, rb, cbnd = _banded(n, (-64, -1, 0, 1, 64), torch.float64, g)
    for (vv, rr, cc), R, cap, all_irregular in (((v, r, c), 16, 64, True), ((vb, rb, cbnd), 16, 170, False)):
        S = ops.Sparse(vv.to(DEV), rr.to(DEV), cc.to(DEV), (n, n))
        T = CsrTiles(S, R, cap, 16 * 8)
        assert (T.n_regular == 0) if all_irregular else (0 < T.n_regular < T.n_tiles), (T.n_regular, T.n_tiles)
        X = torch.randn(n, 16, dtype=torch.float64, generator=g).to(DEV)
        Y = torch.empty_like(X)
        dots = torch.zeros(16, dtype=torch.float64, device=DEV)
        be.csr_spmm_tiled(T, T.values(S.data), S.shape, X, Y, alpha=2.0, shift=0.5, dots=dots)
        Ad = torch.zeros(n, n, dtype=torch.float64)
        Ad[rr, cc] = vv
        ref = 2.0 * (Ad @ X.cpu()) + 0.5 * X.cpu()
        assert rel(Y, ref) < 1e-13 and rel(dots, (X.cpu() * ref).sum(0)) < 1e-6
    # a pattern without diagonals is screened out before the tile form is built (the build sorts every non-zero)
    S = ops.Sparse(v.to(DEV), r.to(DEV), c.to(DEV), (n, n))
    Xr = torch.randn(n, 16, dtype=torch.float64, generator=g).to(DEV)
    import cola_b200.csr_tiles as ct
    with monkeypatch.context() as mp:
        mp.setattr(ct, "CsrTiles", lambda *a, **k: (_ for _ in ()).throw(AssertionError("tile form built for a random pattern")))
        assert S.plan().terms[-1][1][0]._tiles(Xr, torch.empty_like(Xr)) is None
    # values written in place are picked up (the padded copy is refreshed), and a CG solve runs on the staged kernel
    monkeypatch.setattr(ops._CsrCore, "TILE_STAGE_BYTES", 106 << 10)
    data, rows, cols, shape = pb.laplacian_2d_coo(96, torch.float64)
    S = ops.Sparse(data.to(DEV), rows.to(DEV), cols.to(DEV), shape)
    X = torch.randn(shape[0], 16, dtype=torch.float64, generator=g).to(DEV)
    Y1 = S @ X
    S.data.mul_(3.0)
    assert rel(S @ X, 3.0 * Y1) < 1e-14
    S.data.div_(3.0)
    A = cb.PSD(S + 0.05 * ops.I_like(S))
    assert A.plan().terms[-1][1][0]._tiles(X, Y1) is not None
    sol, info = cb.linalg.cg(A, X, tol=1e-10, max_iters=2000)
    Ad = torch.zeros(shape, dtype=torch.float64)
    Ad[rows, cols] = data
    Ad += 0.05 * torch.eye(shape[0], dtype=torch.float64)
    assert rel(Ad @ sol.cpu(), X.cpu()) < 1e-8
