"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs and
against the golden vectors the real reference produced (tests/golden/).

Tolerances are the north-star bar: relative 1e-5 for fp32, 1e-10 for fp64, on per-iteration residual norms,
Lanczos coefficients, solutions, eigenvalues and log-determinants."""
import numpy as np
import pytest
import torch

from tests import problems as pb
from tests.golden_cases import GMRES_CASES, PCG_CASES, POWER_CASES, ARNOLDI_CASES, CG_CASES, LANCZOS_CASES, MATMAT_PROBLEMS

pytestmark = pytest.mark.gpu

F32_TOL, F64_TOL = 1e-5, 1e-10


def rel(a, b):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    a, b = a.astype(np.float64), b.astype(np.float64)
    den = 0.5 * (np.linalg.norm(a) + np.linalg.norm(b))
    return 0.0 if den == 0 else float(np.linalg.norm(a - b) / den)


def tol_of(dtype):
    return F32_TOL if dtype == torch.float32 else F64_TOL


@pytest.fixture(scope="module")
def cb():
    import cola_b200
    assert torch.cuda.is_available()
    cola_b200.backend.lib()  # fails loudly if the extension is missing
    cola_b200.rng.PROBE_DEVICE = "cpu"  # draw probes on the CPU generator: same stream as the CPU oracle
    return cola_b200


DEV = "cuda:0"


# ------------------------------------------------------------------------------------------- matmats
@pytest.mark.parametrize("name", MATMAT_PROBLEMS)
def test_matmat(name, golden, cb):
    P = pb.problem(name)
    A = pb.to_b200(P["spec"], DEV, P["ann"])
    X = pb.randn_np((A.shape[1], 6), P["dtype"], 100)
    g = golden("matmat_" + name)
    t = tol_of(P["dtype"])
    assert rel(A @ X.to(DEV), g["Y"]) < t
    assert rel(A @ X[:, 0].contiguous().to(DEV), g["y"]) < t
    O = pb.to_oracle(P["spec"])
    X2 = pb.randn_np((A.shape[1], 33), P["dtype"], 101)   # ragged RHS count: scalar-width path
    assert rel(A @ X2.to(DEV), O.matmat(X2)) < t


def test_matmat_fused_dots_and_gate(cb):
    """The epilogue's column dots equal sum(X * (A X)) and a closed gate makes the launch a no-op."""
    for name in ["lap24_f32", "kron884_diag_f32", "dense96_f64", "blockdiag_f32", "lap16_shift_f32"]:
        P = pb.problem(name)
        A = pb.to_b200(P["spec"], DEV, P["ann"])
        X = pb.randn_np((A.shape[1], 8), P["dtype"], 5).to(DEV)
        Y = torch.empty_like(X)
        dots = torch.zeros((3, 8), dtype=torch.float64, device=DEV)
        row = torch.tensor([2], dtype=torch.int32, device=DEV)
        A.matmat_into(X, Y, dots=dots, dots_row=row)
        ref = (X.double() * Y.double()).sum(0)
        assert rel(dots[2], ref) < 1e-12 and float(dots[:2].abs().sum()) == 0.0
        Y2 = torch.full_like(X, 7.0)
        closed = torch.tensor([1], dtype=torch.int32, device=DEV)
        A.matmat_into(X, Y2, dots=dots, dots_row=row, gate=closed)
        assert bool((Y2 == 7.0).all()) and rel(dots[2], ref) < 1e-12


# ------------------------------------------------------------------------------------------- CG
@pytest.mark.parametrize("case", sorted(CG_CASES))
def test_cg_vs_oracle_and_golden(case, golden, cb):
    from oracle import krylov_oracle as ko
    name, tol, iters = CG_CASES[case]
    P = pb.problem(name)
    A = pb.to_b200(P["spec"], DEV, P["ann"])
    x, info = cb.linalg.CG(tol=tol, max_iters=iters)(A, P["B"].to(DEV))
    xo, ro, ko_iters, info_o = ko.cg(pb.to_oracle(P["spec"]), P["B"], tol=tol, max_iters=iters)
    g = golden(case)
    t = tol_of(P["dtype"])
    # The trace is compared at the north-star tolerance over the window in which CG itself is numerically
    # well-posed (tests/problems.py:stable_window); iteration counts of convergence-limited runs may differ by the
    # rounding at the threshold crossing.
    assert info_o["iterations"] == int(g["iterations"])
    assert abs(info["iterations"] - info_o["iterations"]) <= 2
    window = min(pb.stable_window(name, tol, iters, t), len(info["errors"]), len(info_o["errors"]))
    assert window >= 3, window
    np.testing.assert_allclose(info["errors"][:window], info_o["errors"][:window], rtol=t)
    np.testing.assert_allclose(info["errors"][:window], g["errors"][:window], rtol=t)
    if tol > 1e-20:   # converged solves: both solutions are within cond*tol of the exact one
        assert rel(x, xo) < 200 * max(t, tol)
        assert rel(x, g["x"]) < 200 * max(t, tol)
    else:
        assert rel(x, xo) < 1e3 * t
    print(f"{case}: window {window} of {len(info_o['errors'])}, iterations {info['iterations']} vs {info_o['iterations']}")


def test_cg_first_iterations_tight(cb):
    """Per-iteration residual norms over the first iterations at the strict north-star tolerance."""
    from oracle import krylov_oracle as ko
    for name, t in [("lap24_f32", F32_TOL), ("lap24_f64", F64_TOL)]:
        P = pb.problem(name)
        A = pb.to_b200(P["spec"], DEV, P["ann"])
        x, info = cb.linalg.CG(tol=1e-30, max_iters=25)(A, P["B"].to(DEV))
        xo, _, _, info_o = ko.cg(pb.to_oracle(P["spec"]), P["B"], tol=1e-30, max_iters=25)
        assert info["iterations"] == info_o["iterations"] == 26
        np.testing.assert_allclose(info["errors"], info_o["errors"], rtol=t)
        assert rel(x, xo) < 10 * t


def test_cg_vector_x0_and_solve_surface(golden, cb):
    P = pb.problem("dense96_f32")
    A = pb.to_b200(P["spec"], DEV, P["ann"])
    x0 = pb.randn_np(tuple(P["B"].shape), P["dtype"], 77)
    x, info = cb.linalg.CG(tol=1e-6, max_iters=500, x0=x0.to(DEV))(A, P["B"].to(DEV))
    g = golden("cg_dense96_f32_x0")
    assert abs(info["iterations"] - int(g["iterations"])) <= 1 and rel(x, g["x"]) < 1e-4
    xs = cb.linalg.solve(A, P["B"].to(DEV), cb.linalg.CG(tol=1e-6, max_iters=500))
    assert rel(xs, golden("solve_dense96_f32")["x"]) < 1e-4
    Ainv = cb.linalg.inv(A, cb.linalg.CG(tol=1e-6, max_iters=500))
    xv = Ainv @ P["B"][:, 0].contiguous().to(DEV)
    assert xv.shape == (96, ) and "iterations" in Ainv.info and rel(xv, golden("solve_dense96_f32")["x"][:, 0]) < 1e-4


def test_cg_edge_cases(cb):
    P = pb.problem("dense96_f64")
    A = pb.to_b200(P["spec"], DEV, P["ann"])
    # zero right-hand side: cond_fun false at k=0, no iterations, x = 0
    x, info = cb.linalg.CG(tol=1e-8, max_iters=50)(A, torch.zeros(96, 2, dtype=torch.float64, device=DEV))
    assert info["iterations"] == 1 and float(x.abs().sum()) == 0.0
    # max_iters = 0
    x, info = cb.linalg.CG(tol=1e-8, max_iters=0)(A, P["B"].to(DEV))
    assert info["iterations"] == 1 and float(x.abs().sum()) == 0.0
    # one converged column next to live ones keeps iterating until all meet the tolerance (any(), cg.py:136)
    B = P["B"].clone()
    B[:, 1] = 0.0
    from oracle import krylov_oracle as ko
    x, info = cb.linalg.CG(tol=1e-9, max_iters=300)(A, B.to(DEV))
    xo, _, _, info_o = ko.cg(pb.to_oracle(P["spec"]), B, tol=1e-9, max_iters=300)
    assert abs(info["iterations"] - info_o["iterations"]) <= 1 and rel(x, xo) < 1e-7
    assert float(x[:, 1].abs().sum()) == 0.0


def test_cg_workspace_is_not_shared_while_in_use(cb):
    """A graph-eligible solve keeps its state and captured batch on the operator.  A solve that finds that workspace in
    use (another thread / stream) takes fresh buffers instead of writing into it, and
    release_cg_workspace() drops it."""
    P = pb.problem("dense96_f64")
    A = pb.to_b200(P["spec"], DEV, P["ann"])
    B = P["B"].to(DEV)
    alg = cb.linalg.CG(tol=1e-10, max_iters=200)
    x1, info1 = alg(A, B)
    ws = A.__dict__["_cg_workspace"]
    assert ws["graph"] is not None and not ws["busy"]
    ws["busy"] = True                                    # as if a concurrent solve held it
    marker = ws["x"].clone()
    x2, info2 = alg(A, B)
    assert torch.equal(ws["x"], marker)                  # untouched
    assert info2["iterations"] == info1["iterations"] and rel(x2, x1) < 1e-9    # atomics order: last-bit differences
    ws["busy"] = False
    x3, _ = alg(A, B)                                    # replays the cached batch again
    assert rel(x3, x1) < 1e-9
    cb.linalg.release_cg_workspace(A)
    assert "_cg_workspace" not in A.__dict__


@pytest.mark.parametrize("case", sorted(PCG_CASES))
def test_pcg_nystrom_vs_oracle_and_golden(case, golden, cb):
    """Preconditioned CG: NystromPrecond built on the device (library QR/Cholesky/SVD of the
    n x r sketch, as in the reference) and applied through the operator plan inside the CG loop."""
    from oracle import krylov_oracle as ko
    name, rank, tol, iters = PCG_CASES[case]
    P = pb.problem(name)
    A = pb.to_b200(P["spec"], DEV, P["ann"])
    Ao = pb.to_oracle(P["spec"])
    saved = cb.rng.PROBE_DEVICE
    cb.rng.PROBE_DEVICE = "cpu"                      # draw the sketch on the CPU generator, like the oracle / golden run
    try:
        Nys = cb.linalg.NystromPrecond(A, rank=rank, key=cb.rng.PRNGKey(3))
    finally:
        cb.rng.PROBE_DEVICE = saved
    Nys_o = ko.NystromPrecondOp(Ao, rank, key=ko.PRNGKey(3))
    g = golden(case)
    t = tol_of(P["dtype"])
    B = P["B"].to(DEV)
    assert rel(Nys.Lambda, Nys_o.Lambda) < 50 * t and rel(Nys.Lambda, g["Lambda"]) < 50 * t
    assert rel(Nys @ B, Nys_o.matmat(P["B"])) < 50 * t and rel(Nys @ B, g["PB"]) < 50 * t
    dots = torch.zeros(B.shape[1], dtype=torch.float64, device=DEV)
    Z = torch.empty_like(B)
    Nys.matmat_into(B, Z, dots=dots)                # <b, P b> fused into the apply of a Product chain + shift
    assert rel(dots, (B.double() * Z.double()).sum(0)) < (1e-6 if P["dtype"] == torch.float32 else 1e-12)
    x, info = cb.linalg.CG(tol=tol, max_iters=iters, P=Nys)(A, B)
    xo, _, _, info_o = ko.cg(Ao, P["B"], tol=tol, max_iters=iters, P=Nys_o)
    assert abs(info["iterations"] - info_o["iterations"]) <= 1 and abs(info["iterations"] - int(g["iterations"])) <= 1
    tol_x = 2e-5 if P["dtype"] == torch.float32 else 1e-8
    assert rel(x, xo) < tol_x and rel(x, g["x"]) < tol_x
    # leading residual norms (the sketch factorizations run in cuSOLVER here and LAPACK there; CG amplifies the ulps)
    m = min(len(info["errors"]), len(info_o["errors"]), 8)
    assert rel(info["errors"][:m], info_o["errors"][:m]) < (5e-3 if P["dtype"] == torch.float32 else 1e-8)
    # preconditioning must pay for itself on these problems
    _, info_plain = cb.linalg.CG(tol=tol, max_iters=iters)(A, B)
    assert info["iterations"] <= info_plain["iterations"]


def test_mode_contract_split_k(cb):
    """Short-and-wide factors (the U^T r product of a Nystrom preconditioner): K split over the grid with atomics."""
    be = cb.backend
    g = torch.Generator().manual_seed(21)
    for dt, t in [(torch.float32, 2e-6), (torch.float64, 1e-13)]:
        for (d_out, d_in, post) in [(24, 20000, 40), (64, 9001, 256), (1, 8192, 1), (33, 70001, 7)]:
            M = (torch.randn(d_out, d_in, dtype=dt, generator=g) / d_in**0.5).to(DEV)
            X = torch.randn(d_in, post, dtype=dt, generator=g).to(DEV)
            Y = torch.full((d_out, post), 7.0, dtype=dt, device=DEV)          # must be overwritten, not accumulated
            be.mode_contract(M, d_out, d_in, 1, post, X, Y, alpha=0.5)
            assert rel(Y, 0.5 * (M.double() @ X.double())) < t, (dt, d_out, d_in, post)
    gate = torch.ones(1, dtype=torch.int32, device=DEV)
    Y = torch.full((24, 40), 7.0, device=DEV)
    M = torch.randn(24, 20000, generator=g).to(DEV); X = torch.randn(20000, 40, generator=g).to(DEV)
    be.mode_contract(M, 24, 20000, 1, 40, X, Y, gate=gate)
    assert float((Y - 7.0).abs().max()) == 0.0                                # gated launch leaves the output alone


def test_pcg_nystrom_larger_grid(cb):
    """PCG on a 128x128 grid Laplacian + 0.05 I (n = 16384): the U^T r product takes the split-K path."""
    from oracle import krylov_oracle as ko
    vals, rows, cols, shape = pb.laplacian_2d_coo(128, torch.float64)
    n = shape[0]
    S = cb.ops.Sparse(vals.to(DEV), rows.to(DEV), cols.to(DEV), shape)
    A = cb.PSD(S + 0.05 * cb.ops.I_like(S))
    Ao = ko.SumOp(ko.SparseOp(vals, rows, cols, shape), ko.ScaledIdentityOp(0.05, n, torch.float64))
    B = pb.randn_np((n, 4), torch.float64, 12)
    saved = cb.rng.PROBE_DEVICE
    cb.rng.PROBE_DEVICE = "cpu"
    try:
        Nys = cb.linalg.NystromPrecond(A, rank=40, mu=1e-4, key=cb.rng.PRNGKey(9))
    finally:
        cb.rng.PROBE_DEVICE = saved
    Nys_o = ko.NystromPrecondOp(Ao, 40, mu=1e-4, key=ko.PRNGKey(9))
    assert rel(Nys @ B.to(DEV), Nys_o.matmat(B)) < 1e-9
    x, info = cb.linalg.CG(tol=1e-8, max_iters=2000, P=Nys)(A, B.to(DEV))
    xo, _, _, info_o = ko.cg(Ao, B, tol=1e-8, max_iters=2000, P=Nys_o)
    assert abs(info["iterations"] - info_o["iterations"]) <= 2
    assert rel(x, xo) < 1e-6
    assert rel(A @ x, B) < 1e-6


def test_product_chain_epilogue(cb):
    """shift / diagonal / fused <x, y> dots on an operator whose only core is a Product chain: the epilogue must see
    the operator's input, not the chain's intermediate."""
    g = torch.Generator().manual_seed(2)
    M = torch.randn(40, 24, dtype=torch.float64, generator=g).to(DEV)
    d = (torch.rand(40, dtype=torch.float64, generator=g) + 0.5).to(DEV)
    A = cb.PSD(cb.ops.Product(cb.ops.Dense(M), cb.ops.Dense(M.T.contiguous())) + cb.ops.Diagonal(d)
               + 0.3 * cb.ops.I_like(cb.ops.Dense(M @ M.T)))
    X = torch.randn(40, 5, dtype=torch.float64, generator=g).to(DEV)
    ref = M @ (M.T @ X) + d[:, None] * X + 0.3 * X
    assert rel(A @ X, ref) < 1e-13
    dots = torch.zeros(5, dtype=torch.float64, device=DEV)
    Y = torch.empty_like(X)
    A.matmat_into(X, Y, dots=dots)
    assert rel(Y, ref) < 1e-13 and rel(dots, (X * ref).sum(0)) < 1e-13
    x, info = cb.linalg.CG(tol=1e-12, max_iters=200)(A, X)
    assert rel(A @ x, X) < 1e-10


def test_cg_graph_workspace_reuse(cb):
    """Graph-eligible solves keep their state (and the captured iteration batch) on the operator: a second solve
    with the same shape replays the first one's graph.  Results must not depend on that history, earlier outputs
    must stay intact, and a change of shape / max_iters / scratch buffers must recapture."""
    from oracle import krylov_oracle as ko
    P = pb.problem("kron888_f32")
    A = pb.to_b200(P["spec"], DEV, P["ann"])
    Ao = pb.to_oracle(P["spec"])
    B1 = P["B"]
    B2 = pb.randn_np(tuple(B1.shape), P["dtype"], 321)
    alg = cb.linalg.CG(tol=1e-6, max_iters=200)
    x1, i1 = alg(A, B1.to(DEV))
    x1_copy = x1.clone()
    assert A.__dict__["_cg_workspace"]["graph"] is not None          # 3 batches of 16 iterations at least
    g_first = A.__dict__["_cg_workspace"]["graph"]
    x2, i2 = alg(A, B2.to(DEV))
    assert A.__dict__["_cg_workspace"]["graph"] is g_first           # replayed, not recaptured
    assert torch.equal(x1, x1_copy)                                  # outputs are copies of the workspace
    x1b, i1b = alg(A, B1.to(DEV))
    assert rel(x1b, x1) < 1e-6 and i1b["iterations"] == i1["iterations"]       # same inputs, replayed graph
    for B, x, info in ((B1, x1, i1), (B2, x2, i2)):
        xo, _, _, info_o = ko.cg(Ao, B, tol=1e-6, max_iters=200)
        assert abs(info["iterations"] - info_o["iterations"]) <= 1     # the stop test can flip on the last ulp
        assert rel(x, xo) < 2e-5
    x0 = pb.randn_np(tuple(B1.shape), P["dtype"], 5)
    x3, _ = cb.linalg.CG(tol=1e-6, max_iters=200, x0=x0.to(DEV))(A, B1.to(DEV))
    assert rel(x3, x1) < 5e-4                                        # other start, same tolerance: cond * tol apart
    # different number of right-hand sides -> new workspace; then back
    x4, _ = alg(A, B1[:, :3].contiguous().to(DEV))
    assert A.__dict__["_cg_workspace"]["key"][1] == 3 and rel(x4, x1[:, :3]) < 5e-4   # `any` stop rule over 3 columns
    x5, _ = alg(A, B1.to(DEV))
    assert rel(x5, x1) < 1e-6


def test_cg_full_size_properties(cb):
    """BASELINE config 2 at a GPU-friendly slice of full size (1024^2 grid, 64 RHS, fp32): properties that
    do not need the oracle: the true residual b - A x matches the recurrence residual and decreases."""
    g = 1024
    data, rows, cols, shape = pb.laplacian_2d_coo(g, torch.float32)
    A = cb.PSD(cb.ops.Sparse(data.to(DEV), rows.to(DEV), cols.to(DEV), shape))
    torch.manual_seed(0)
    B = torch.randn(shape[0], 64, device=DEV)
    x, info = cb.linalg.CG(tol=1e-30, max_iters=40)(A, B)
    r_true = B - A @ x
    rn = torch.linalg.norm(r_true, dim=0) / torch.linalg.norm(B, dim=0)
    assert abs(float(rn.mean()) - info["errors"][-1]) < 1e-4 * info["errors"][-1] + 1e-6
    assert info["errors"][-1] < info["errors"][0] and info["iterations"] == 41
    # linearity: solving for 2B gives 2x (normalised system => same iterates)
    x2, _ = cb.linalg.CG(tol=1e-30, max_iters=40)(A, 2 * B)
    assert rel(x2, 2 * x) < 1e-6


# ------------------------------------------------------------------------------------------- Lanczos
@pytest.mark.parametrize("case", sorted(LANCZOS_CASES))
def test_lanczos_vs_oracle_and_golden(case, golden, cb):
    from oracle import krylov_oracle as ko
    name, m, tol, batched = LANCZOS_CASES[case]
    P = pb.problem(name)
    A = pb.to_b200(P["spec"], DEV, P["ann"])
    start = P["B"] if (batched or P["B"].dim() == 1) else P["B"][:, 0].contiguous()
    Q, T, info = cb.linalg.Lanczos(start_vector=start.to(DEV), max_iters=m, tol=tol)(A)
    Qo, ao, bo, info_o = ko.lanczos(pb.to_oracle(P["spec"]), start, m, tol)
    g = golden(case)
    t = tol_of(P["dtype"])
    assert info["iterations"] == info_o["iterations"] == int(g["iterations"])
    alpha, beta = T.alpha[..., 0], T.beta[..., 0]
    assert tuple(alpha.shape) == tuple(g["alpha"].shape) and tuple(beta.shape) == tuple(g["beta"].shape)
    # the first 12 coefficients at the strict bar; the whole run a little looser (Lanczos amplifies rounding)
    w = 12
    assert rel(alpha[..., :w], ao[..., :w]) < t and rel(beta[..., :w], bo[..., :w]) < t
    assert rel(alpha[..., :w], g["alpha"][..., :w]) < t and rel(beta[..., :w], g["beta"][..., :w]) < t
    assert rel(alpha, ao) < 100 * t and rel(beta, bo) < 100 * t
    Qd = Q.to_dense()
    assert tuple(Qd.shape) == tuple(Qo.shape)
    Qn = Qd.cpu().numpy()
    if Qn.ndim == 2 and Qn.shape[0] > 1000:
        Qn = Qn[::16]
    assert rel(Qn[..., :w], g["Q"][..., :w]) < 10 * t
    # orthonormality of the basis (full reorthogonalisation does its job)
    Qb = Qd if Qd.dim() == 3 else Qd[None]
    G = Qb.transpose(1, 2).double() @ Qb.double()
    eye = torch.eye(G.shape[-1], dtype=torch.float64, device=G.device)
    assert float((G - eye).abs().max()) < (1e-4 if P["dtype"] == torch.float32 else 1e-10)


def test_lanczos_known_answers_and_early_stop(golden, cb):
    """Hand-derived cases the reference's own tests pin (tests/algorithms/test_lanczos.py:268-300)."""
    L = cb.linalg
    A = cb.SelfAdjoint(cb.ops.Dense(torch.diag(torch.tensor([4., 2., 1.])).to(DEV)))
    Q, T, info = L.Lanczos(start_vector=torch.tensor([[1.0, 0.0, 0.0]]).T.to(DEV), max_iters=3, tol=1e-7)(A)
    g = golden("lanczos_case_early")
    assert info["iterations"] == int(g["iterations"]) == 2
    assert tuple(T.beta[..., 0].shape) == tuple(g["beta"].shape) and float(T.beta[0, 0, 0]) == 4.0
    assert tuple(T.alpha[..., 0].shape) == tuple(g["alpha"].shape)
    beta, alpha = [1., 3., 7.], [0.1, 1.0]
    M = torch.tensor([[beta[2], 0, alpha[1]], [0, beta[0], alpha[0]], [alpha[1], alpha[0], beta[1]]])
    Q, T, info = L.Lanczos(start_vector=torch.tensor([[0.0, 1.0, 0.]]).T.to(DEV), max_iters=3, tol=1e-7)(
        cb.SelfAdjoint(cb.ops.Dense(M.to(DEV))))
    assert info["iterations"] - 1 == 3
    assert rel(T.beta[0, :, 0], beta) < 1e-6 and rel(T.alpha[0, :, 0], alpha) < 1e-6


def test_eig_lanczos_default_start(golden, cb):
    P = pb.problem("graph2k_f64")
    A = pb.to_b200(P["spec"], DEV, P["ann"])
    vals, vecs = cb.linalg.eig(A, 6, "LM", cb.linalg.Lanczos(max_iters=48, tol=1e-12))
    g = golden("eig_graph2k_f64_default_start")
    assert rel(vals, g["eigvals"]) < F64_TOL
    V = vecs.to_dense()
    assert tuple(V.shape) == (2048, 6)
    assert rel(V.abs()[::16], np.abs(g["eigvecs"])) < 1e-7
    # Ritz pairs satisfy A v = lambda v to the accuracy of the top of the spectrum
    res = torch.linalg.norm(A @ V[:, -1].contiguous() - vals[-1] * V[:, -1]) / vals[-1]
    assert float(res) < 1e-6


# ------------------------------------------------------------------------------------------- Arnoldi
@pytest.mark.parametrize("case", sorted(ARNOLDI_CASES))
def test_arnoldi_vs_oracle_and_golden(case, golden, cb):
    from oracle import krylov_oracle as ko
    name, m, tol, batched = ARNOLDI_CASES[case]
    P = pb.problem(name)
    A = pb.to_b200(P["spec"], DEV, P["ann"])
    start = P["B"] if batched else P["B"][:, 0].contiguous()
    Q, H, info = cb.linalg.Arnoldi(start_vector=start.to(DEV), max_iters=m, tol=tol)(A)
    Qo, Ho, info_o = ko.arnoldi(pb.to_oracle(P["spec"]), start, m, tol)
    g = golden(case)
    t = tol_of(P["dtype"])
    assert info["iterations"] == info_o["iterations"] == int(g["iterations"])
    Hd, Qd = H.to_dense(), Q.to_dense()
    assert tuple(Hd.shape) == tuple(g["H"].shape) and tuple(Qd.shape) == tuple(g["Q"].shape)
    # Leading columns at the strict bar.  Later columns of an Arnoldi factorisation are forward-unstable: the
    # window is where the oracle itself is stable under a 1-ulp perturbation of the start vector; over the whole
    # run the backward-stable invariant of MGS-Arnoldi is checked instead: A Q_m = Q_{m+1} H to rounding.  (MGS
    # does not keep Q orthonormal -- the fp32 oracle loses it to 2e-2 after 20 steps -- so that is not asserted.)
    ulp = 1e-7 if P["dtype"] == torch.float32 else 1e-15
    noise = 1.0 + ulp * pb.t(pb.rs(5).normal(size=tuple(start.shape)), P["dtype"])
    _, Hp, _ = ko.arnoldi(pb.to_oracle(P["spec"]), start * noise, m, tol)
    w = 1
    while w < m and rel(Hp[..., :w + 2, :w + 1], Ho[..., :w + 2, :w + 1]) < t / 4:
        w += 1
    assert w >= 3, w
    assert rel(Hd[..., :w + 1, :w], Ho[..., :w + 1, :w]) < t and rel(Hd[..., :w + 1, :w], g["H"][..., :w + 1, :w]) < t
    assert rel(Qd[..., :w], Qo[..., :w]) < 10 * t and rel(Qd[..., :w], g["Q"][..., :w]) < 10 * t
    Qb = (Qd if Qd.dim() == 3 else Qd[None]).double()
    Hb = (Hd if Hd.dim() == 3 else Hd[None]).double()
    Ad = P["spec"][1].double().to(DEV)
    steps = info["iterations"] - 1
    eps = 1e-6 if P["dtype"] == torch.float32 else 1e-14
    for q, h in zip(Qb, Hb):
        resid = Ad @ q[:, :steps] - q[:, :steps + 1] @ h[:steps + 1, :steps]
        assert float(resid.abs().max()) < 20 * eps * float(Ad.abs().max())
    print(f"{case}: strict window {w} of {m} columns")


def test_eig_arnoldi(golden, cb):
    P = pb.problem("nonsym48_f64")
    A = pb.to_b200(P["spec"], DEV, P["ann"])
    vals, vecs = cb.linalg.eig(A, 48, "LM", cb.linalg.Arnoldi(start_vector=P["B"][:, 0].contiguous().to(DEV),
                                                                max_iters=48, tol=1e-12))
    mags = np.sort(np.abs(vals.cpu().numpy()))
    assert rel(mags, golden("eig_arnoldi_nonsym48_f64")["eigvals_sorted_abs"]) < 1e-8


# ------------------------------------------------------------------------------------------- power iteration (8f item 3)
@pytest.mark.parametrize("case", sorted(POWER_CASES))
def test_power_iteration_vs_oracle_and_golden(case, golden, cb):
    from oracle import krylov_oracle as ko
    name, tol, iters = POWER_CASES[case]
    P = pb.problem(name)
    A = pb.to_b200(P["spec"], DEV, P["ann"])
    saved = cb.rng.PROBE_DEVICE
    cb.rng.PROBE_DEVICE = "cpu"                      # same start vector as the oracle / golden run
    try:
        v, emax, info = cb.linalg.PowerIteration(tol=tol, max_iter=iters, key=cb.rng.PRNGKey(11))(A)
    finally:
        cb.rng.PROBE_DEVICE = saved
    vo, emax_o, info_o = ko.power_iteration(pb.to_oracle(P["spec"]), tol=tol, max_iter=iters, key=ko.PRNGKey(11))
    g = golden(case)
    t = tol_of(P["dtype"])
    # the stop test |eig_prev - eig| / eig > tol can flip on the last ulp near convergence
    assert abs(info["iterations"] - info_o["iterations"]) <= 1 and abs(info["iterations"] - int(g["iterations"])) <= 1
    assert abs(float(emax) - float(emax_o)) <= 10 * t * abs(float(emax_o))
    assert abs(float(emax) - float(g["eigmax"])) <= 10 * t * abs(float(g["eigmax"]))
    if info["iterations"] == info_o["iterations"]:
        assert rel(v, vo) < 1e3 * t and rel(v, g["v"]) < 1e3 * t
    m = min(len(info["errors"]), len(info_o["errors"])) - 3          # the last entries sit at the rounding floor
    assert rel(info["errors"][:m], info_o["errors"][:m]) < (1e-2 if P["dtype"] == torch.float32 else 1e-5)


def test_eigmax_dispatch(golden, cb):
    P = pb.problem("dense96_f64")
    A = pb.to_b200(P["spec"], DEV, P["ann"])
    saved = cb.rng.PROBE_DEVICE
    cb.rng.PROBE_DEVICE = "cpu"
    try:
        e = cb.linalg.eigmax(A, cb.linalg.PowerIteration(tol=1e-9, max_iter=400, key=cb.rng.PRNGKey(11)))
        vals, vecs = cb.linalg.eig(A, 1, "LM", cb.linalg.PowerIteration(tol=1e-9, max_iter=400, key=cb.rng.PRNGKey(11)))
    finally:
        cb.rng.PROBE_DEVICE = saved
    assert abs(float(e) - float(golden("eigmax_dense96_f64")["eigmax"])) < 1e-9
    assert vals.shape == (1,) and vecs.shape == (96, 1)
    with pytest.raises(AssertionError):
        cb.linalg.eig(A, 2, "LM", cb.linalg.PowerIteration())


# ------------------------------------------------------------------------------------------- GMRES (8f item 2)
@pytest.mark.parametrize("case", sorted(GMRES_CASES))
def test_gmres_vs_oracle_and_golden(case, golden, cb):
    """GMRES = device Arnoldi + the reference's normal-equation solve of the square Hessenberg (gmres.py:110-118).
    That solve squares cond(H), so the bar on x is set by the oracle's own sensitivity to a 1-ulp perturbation of
    b; the residual of the returned solution must be as small as the oracle's."""
    from oracle import krylov_oracle as ko
    name, m, tol, vec = GMRES_CASES[case]
    P = pb.problem(name)
    A = pb.to_b200(P["spec"], DEV, P["ann"])
    Ao = pb.to_oracle(P["spec"])
    b = P["B"][:, 0].contiguous() if vec else P["B"]
    x, info = cb.linalg.GMRES(tol=tol, max_iters=m)(A, b.to(DEV))
    xo, info_o = ko.gmres(Ao, b, max_iters=m, tol=tol)
    g = golden(case)
    assert info["iterations"] == info_o["iterations"] == int(g["iterations"])
    assert tuple(x.shape) == tuple(g["x"].shape)
    t = tol_of(P["dtype"])
    ulp = 1e-7 if P["dtype"] == torch.float32 else 1e-15
    Ad = pb.to_oracle(pb.upcast(P["spec"]))
    b64 = b.double()

    def resid(sol):
        sol = torch.as_tensor(np.asarray(sol.cpu() if torch.is_tensor(sol) else sol)).double()
        return float(torch.linalg.norm(b64 - Ad @ sol) / torch.linalg.norm(b64))
    sens, worst_resid = t, resid(xo)
    for seed in (6, 7, 8, 9):            # the oracle under 1-ulp perturbations of b (full-space runs are ill-posed)
        noise = 1.0 + ulp * pb.t(pb.rs(seed).normal(size=tuple(b.shape)), P["dtype"])
        xp, _ = ko.gmres(Ao, b * noise, max_iters=m, tol=tol)
        sens, worst_resid = max(sens, rel(xp, xo)), max(worst_resid, resid(xp))
    bar = 20 * sens
    assert rel(x, xo) < bar and rel(x, g["x"]) < bar, (rel(x, xo), rel(x, g["x"]), bar)
    assert resid(x) <= 2 * worst_resid + 100 * t, (resid(x), worst_resid)
    print(f"{case}: rel(x, oracle) {rel(x, xo):.2e} (bar {bar:.2e}), residual {resid(x):.2e} vs oracle {resid(xo):.2e}")


def test_gmres_x0_solve_and_auto(golden, cb):
    P = pb.problem("nonsym48_f64")
    A = pb.to_b200(P["spec"], DEV, P["ann"])
    B = P["B"].to(DEV)
    x0 = pb.randn_np(tuple(P["B"].shape), P["dtype"], 78)
    x, info = cb.linalg.GMRES(tol=1e-12, max_iters=20, x0=x0.to(DEV))(A, B)
    g = golden("gmres_nonsym48_f64_x0")
    assert info["iterations"] == int(g["iterations"]) and rel(x, g["x"]) < 1e-8
    x = cb.linalg.solve(A, B, cb.linalg.GMRES(tol=1e-12, max_iters=20))       # inv.py:23-39, 60-62
    assert rel(x, golden("solve_gmres_nonsym48_f64")["x"]) < 1e-8
    Ainv = cb.linalg.inv(A, cb.linalg.GMRES(tol=1e-12, max_iters=20))
    y = Ainv @ B[:, 0].contiguous()
    assert y.shape == (48,) and Ainv.info["iterations"] > 0
    with pytest.raises(RuntimeError):
        cb.linalg.GMRES()(A, P["B"])                                          # CPU right-hand side: no fallback
    with pytest.raises(NotImplementedError):
        cb.linalg.gmres(A, B, use_householder=True)


# ------------------------------------------------------------------------------------------- SLQ / Hutch
@pytest.mark.parametrize("name,m,vtol", [("kron884_diag_f32", 25, 0.25), ("kron465_diag_f64", 30, 0.2),
                                          ("lap24_f64", 40, 0.25)])
def test_slq_logdet_identical_probes(name, m, vtol, golden, cb):
    P = pb.problem(name)
    A = pb.to_b200(P["spec"], DEV, P["ann"])
    g = golden("slq_" + name)
    val = cb.linalg.stochastic_lanczos_quad(A, torch.log, max_iters=m, tol=1e-7, vtol=vtol, key=int(g["key"]))
    t = tol_of(P["dtype"])
    assert abs(float(val) - float(g["logdet"])) <= 10 * t * abs(float(g["logdet"]))
    # chunked probes give the same estimate (per-probe work is independent)
    val2 = cb.linalg.stochastic_lanczos_quad(A, torch.log, max_iters=m, tol=1e-7, vtol=vtol, key=int(g["key"]),
                                             probe_chunk_size=5)
    assert abs(float(val2) - float(val)) <= 10 * t * abs(float(val))


@pytest.mark.parametrize("name,m", [("kron884_diag_f32", 25), ("kron465_diag_f64", 30)])
def test_log_matmat_and_hutch_logdet(name, m, golden, cb):
    P = pb.problem(name)
    A = pb.to_b200(P["spec"], DEV, P["ann"])
    L = cb.linalg
    t = tol_of(P["dtype"])
    F = L.LanczosUnary(A, torch.log, max_iters=m, tol=1e-7)
    assert rel(F @ P["B"].to(DEV), golden("logA_matmat_" + name)["Y"]) < 100 * t
    g = golden("hutch_logdet_" + name)
    val = L.logdet(A, L.Lanczos(max_iters=m, tol=1e-7), L.Hutch(tol=2e-2, max_iters=3, key=int(g["key"])))
    assert abs(float(val) - float(g["logdet"])) <= 100 * t * abs(float(g["logdet"]))
    dg = L.Hutch(tol=2e-2, max_iters=3, key=int(g["key"]))(L.LanczosUnary(A, torch.log, max_iters=m, tol=1e-7), 0)
    assert rel(dg, g["diag"]) < 100 * t


def test_hutch_rademacher_and_errors(golden, cb):
    P = pb.problem("dense96_f64")
    A = pb.to_b200(P["spec"], DEV, P["ann"])
    dg = cb.linalg.Hutch(tol=5e-2, max_iters=4, rand="rademacher", key=cb.rng.PRNGKey(7))(A, 0)
    assert rel(dg, golden("hutch_diag_dense96_f64")["diag"]) < 1e-10
    with pytest.raises(AssertionError, match="tolerance chosen too high"):
        cb.linalg.Hutch(tol=1e-4)(A, 0)
    ex = cb.linalg.diag(A, 0, cb.linalg.Exact())
    assert rel(ex, torch.diagonal(P["spec"][1])) < 1e-12


# ------------------------------------------------------------------------------------------- tensor-core Kronecker
def test_kronecker_tensor_core_path(cb):
    """64x64 factors take the tcgen05 / TMA path (3xTF32): fp32-grade accuracy against an fp64 reference, the
    fused epilogue (shift, diagonal, pAp dots), agreement with the exact SIMT contraction, and CG parity with the
    CPU oracle on a BASELINE-config-3-shaped operator (scaled to two factors so the oracle stays fast)."""
    from oracle import krylov_oracle as ko
    torch.manual_seed(0)
    for D, k in [(2, 32), (2, 96), (3, 64)]:
        Fs = [pb.kron_factor(64, torch.float32, 40 + i) for i in range(D)]
        n = 64**D
        dg = pb.t(pb.rs(9).uniform(size=n) + 0.5, torch.float32)
        K = cb.ops.Kronecker(*[cb.ops.Dense(F.to(DEV)) for F in Fs])
        A = cb.PSD(K + 0.1 * cb.ops.I_like(K) + cb.ops.Diagonal(dg.to(DEV)))
        core = A.plan().terms[0][1][0]
        X = pb.randn_np((n, k), torch.float32, 3).to(DEV)
        assert core._tc_ok(X)
        Y = torch.empty_like(X)
        dots = torch.zeros(k, dtype=torch.float64, device=DEV)
        A.matmat_into(X, Y, dots=dots)
        E = X.double().reshape(*([64] * D), k)
        for i, F in enumerate(Fs):
            E = torch.moveaxis(torch.tensordot(F.double().to(DEV), torch.moveaxis(E, i, 0), dims=1), 0, i)
        ref = E.reshape(n, k) + (0.1 + dg.double().to(DEV))[:, None] * X.double()
        assert rel(Y, ref) < 2e-6, (D, k, rel(Y, ref))
        assert rel(dots, (X.double() * Y.double()).sum(0)) < 1e-6    # fp32 partials over 32 values, fp64 across tiles
        core.use_tensor_cores = False
        Y2 = torch.empty_like(X)
        A.matmat_into(X, Y2)
        core.use_tensor_cores = True
        assert rel(Y, Y2) < 2e-6
    # CG through the tensor-core matmat vs the CPU oracle
    Fs = [pb.kron_factor(64, torch.float32, 50 + i) for i in range(2)]
    B = pb.randn_np((4096, 32), torch.float32, 8)
    K = cb.ops.Kronecker(*[cb.ops.Dense(F.to(DEV)) for F in Fs])
    A = cb.PSD(K + 0.1 * cb.ops.I_like(K))
    x, info = cb.linalg.CG(tol=1e-30, max_iters=30)(A, B.to(DEV))
    Ao = ko.SumOp(ko.KroneckerOp(*[ko.DenseOp(F) for F in Fs]), ko.ScaledIdentityOp(0.1, 4096, torch.float32))
    xo, _, _, info_o = ko.cg(Ao, B, tol=1e-30, max_iters=30)
    A64 = ko.SumOp(ko.KroneckerOp(*[ko.DenseOp(F.double()) for F in Fs]), ko.ScaledIdentityOp(0.1, 4096, torch.float64))
    _, _, _, info_64 = ko.cg(A64, B.double(), tol=1e-30, max_iters=30)
    bad = np.nonzero(np.abs(info_o["errors"] - info_64["errors"]) > 0.25 * F32_TOL * info_64["errors"])[0]
    window = int(bad[0]) if len(bad) else len(info_o["errors"])
    assert window >= 5
    np.testing.assert_allclose(info["errors"][:window], info_o["errors"][:window], rtol=F32_TOL)
    assert info["iterations"] == info_o["iterations"] == 31
    print(f"tensor-core CG: strict window {window} of {len(info_o['errors'])}")


# ------------------------------------------------------------------------------------------- kernels directly
def test_tridiag_ql_first_row(cb):
    """Eigenvalues and first eigenvector components from the QL kernel against fp64 torch.linalg.eigh, including
    matrices with (near-)zero couplings (early-terminated Lanczos) and m = 1, 2."""
    be = cb.backend
    g = torch.Generator().manual_seed(5)
    for m, b in [(1, 3), (2, 5), (7, 33), (40, 64), (100, 70), (129, 2)]:
        d = torch.randn(m, b, dtype=torch.float64, generator=g) + 3.0
        e = torch.rand(m, b, dtype=torch.float64, generator=g) + 0.05
        if m > 4:
            e[2, 0] = 0.0            # exactly decoupled block
            e[3, 1] = 1e-30          # numerically decoupled
        e[m - 1] = 0.0
        T = torch.diag_embed(d.T) + torch.diag_embed(e[:m - 1].T, offset=1) + torch.diag_embed(e[:m - 1].T, offset=-1)
        lam_ref, Q = torch.linalg.eigh(T)
        w_ref = Q[:, 0, :]**2
        dd, ee, zz = d.to(DEV).contiguous(), e.to(DEV).contiguous(), torch.empty(m, b, dtype=torch.float64, device=DEV)
        status = torch.ones(b, dtype=torch.int32, device=DEV)
        be.tridiag_eig_first_row(dd, ee, zz, status)
        assert int(status.abs().sum()) == 0
        lam, order = torch.sort(dd.T.cpu(), dim=1)
        w = torch.gather((zz.T.cpu())**2, 1, order)
        assert float((lam - lam_ref).abs().max()) < 1e-12 * float(lam_ref.abs().max()), (m, b)
        assert float((w.sum(1) - 1).abs().max()) < 1e-12
        # weights of (numerically) equal eigenvalues are only defined as a sum: compare the quadrature itself
        for f in (lambda x: torch.log(x * x + 1.0), lambda x: x**3, lambda x: torch.exp(-x * x)):
            q, q_ref = (w * f(lam)).sum(1), (w_ref * f(lam_ref)).sum(1)
            assert float((q - q_ref).abs().max()) < 1e-11 * float(q_ref.abs().max()), (m, b)


def test_mode_contract_big_tile_path(cb):
    """The 128x128x8 register-tiled fp32 path (factors with d_out % 128 == 0) against fp64 torch, including a
    ragged last column tile, a rectangular factor, the shift/diag epilogue and accumulate; and that shapes it
    does not take (dots requested, fp64) still agree."""
    be = cb.backend
    g = torch.Generator().manual_seed(11)
    for (d_out, d_in, pre, post) in [(128, 128, 1, 4096), (128, 128, 3, 200), (256, 64, 2, 132), (128, 136, 1, 64)]:
        M = torch.randn(d_out, d_in, generator=g).to(DEV)
        X = torch.randn(pre, d_in, post, generator=g).to(DEV)
        Y = torch.empty(pre, d_out, post, device=DEV)
        be.mode_contract(M, d_out, d_in, pre, post, X, Y)
        ref = torch.einsum("aj,pjq->paq", M.double(), X.double())
        assert rel(Y, ref) < 2e-6, (d_out, d_in, pre, post)
    d, pre, post = 128, 2, 260
    M = torch.randn(d, d, generator=g).to(DEV)
    X = torch.randn(pre, d, post, generator=g).to(DEV)
    dg = torch.rand(pre * d, generator=g).to(DEV)
    Y0 = torch.randn(pre, d, post, generator=g).to(DEV)
    Y = Y0.clone()
    be.mode_contract(M, d, d, pre, post, X, Y, alpha=0.5, shift=0.25, diag=dg, epi_x=X, accumulate=True)
    ref = 0.5 * torch.einsum("aj,pjq->paq", M.double(), X.double()) + (0.25 + dg.double().reshape(pre, d, 1)) * X.double() \
        + Y0.double()
    assert rel(Y, ref) < 2e-6
    dots = torch.zeros(post, dtype=torch.float64, device=DEV)
    Y2 = torch.empty_like(Y)
    be.mode_contract(M, d, d, pre, post, X, Y2, epi_x=X, dots=dots)             # dots -> 64x64 tile kernel
    assert rel(Y2, torch.einsum("aj,pjq->paq", M.double(), X.double())) < 2e-6
    assert rel(dots, (X.double() * Y2.double()).sum((0, 1))) < 1e-6


def test_reorth_kernels_shapes(cb):
    """C = V^T W and W -= V C for ragged / wide / single-column probe blocks in both dtypes."""
    be = cb.backend
    for dt, t in [(torch.float32, 2e-6), (torch.float64, 1e-13)]:
        for n, b, nv in [(1000, 1, 7), (777, 3, 5), (512, 8, 33), (300, 64, 40), (257, 100, 9), (129, 160, 4),
                         (64, 256, 3)]:
            torch.manual_seed(n + b)
            V = torch.randn(nv, n, b, dtype=dt, device=DEV)
            W = torch.randn(n, b, dtype=dt, device=DEV)
            C = torch.zeros(nv, b, dtype=torch.float64, device=DEV)
            be.reorth_dots(V, 1, nv, W, C)
            ref = torch.einsum("jnb,nb->jb", V.double(), W.double())
            assert float(C[0].abs().sum()) == 0.0
            assert rel(C[1:], ref[1:]) < t, (dt, n, b, nv)
            W2 = W.clone()
            nrm = torch.zeros(b, dtype=torch.float64, device=DEV)
            be.reorth_update(V, 1, nv, W2, C, sign=-1.0, wnorm2=nrm)
            refW = W.double() - torch.einsum("jnb,jb->nb", V[1:].double(), C[1:].to(dt).double())
            assert rel(W2, refW) < 10 * t, (dt, n, b, nv)
            assert rel(nrm, (W2.double()**2).sum(0)) < 1e-12


def test_reorth_fused_update_dots(cb):
    """The one-sweep middle step of CGS2 (W -= V C1, C2 = V^T W_new) against the two separate kernels and an fp64
    restatement: ragged row counts (partial last chunk), folded single column, tall 3-D TMA chunks, both dtypes,
    every vector count from 1 up; shapes outside the fused envelope report 'not launched' (False)."""
    be = cb.backend
    launched = 0
    for dt, t in [(torch.float32, 3e-6), (torch.float64, 1e-13)]:
        for n, b, nv in [(4096, 64, 2), (4099, 64, 13), (5003, 64, 51), (2050, 64, 101), (3001, 128, 30), (9001, 8, 33),
                         (1 << 15, 1, 65), (40000, 1, 9), (1022, 16, 120), (777, 3, 5), (2048, 256, 6)]:
            torch.manual_seed(n + b + nv)
            V = torch.randn(nv, n, b, dtype=dt, device=DEV) / n**0.5
            W = torch.randn(n, b, dtype=dt, device=DEV)
            C1 = torch.zeros(nv, b, dtype=torch.float64, device=DEV)
            be.reorth_dots(V, 1, nv, W, C1)
            W_f = W.clone()
            C2_f = torch.zeros_like(C1)
            ok = be.reorth_update_dots(V, 1, nv, W_f, C1, C2_f, sign=-1.0)
            if not ok:
                assert b == 3 or (b == 1 and n % (16 // W.element_size()) != 0), (dt, n, b, nv)
                assert torch.equal(W_f, W) and float(C2_f.abs().sum()) == 0.0   # nothing was launched
                continue
            launched += 1
            W_s = W.clone()
            C2_s = torch.zeros_like(C1)
            be.reorth_update(V, 1, nv, W_s, C1, sign=-1.0)
            be.reorth_dots(V, 1, nv, W_s, C2_s)
            refW = W.double() - torch.einsum("jnb,jb->nb", V[1:].double(), C1[1:].to(dt).double())
            assert rel(W_f, refW) < 10 * t, (dt, n, b, nv)
            assert rel(W_f, W_s) < 10 * t, (dt, n, b, nv)
            # pass-2 coefficients are O(eps) cancellations: compare on the scale of the pass-1 coefficients
            refC2 = torch.einsum("jnb,nb->jb", V[1:].double(), W_f.double())
            scale = float(C1[1:].abs().max())
            assert float((C2_f[1:] - refC2).abs().max()) < 20 * t * scale, (dt, n, b, nv)
            assert float(C2_f[0].abs().sum()) == 0.0
    assert launched >= 18
    # the device-side gate skips the launch's work
    gate = torch.ones(1, dtype=torch.int32, device=DEV)
    V = torch.randn(5, 4096, 64, device=DEV); W = torch.randn(4096, 64, device=DEV)
    C1 = torch.ones(5, 64, dtype=torch.float64, device=DEV); C2 = torch.zeros_like(C1)
    W0 = W.clone()
    assert be.reorth_update_dots(V, 1, 5, W, C1, C2, gate=gate)
    assert torch.equal(W, W0) and float(C2.abs().sum()) == 0.0


def test_vector_sweeps_fold_and_ragged(cb):
    be = cb.backend
    for dt in (torch.float32, torch.float64):
        for n, k in [(1001, 1), (4096, 1), (333, 3), (200, 64), (50, 100), (17, 1500)]:
            torch.manual_seed(n * k)
            X = torch.randn(n, k, dtype=dt, device=DEV)
            Y = torch.randn(n, k, dtype=dt, device=DEV)
            d = torch.zeros(k, dtype=torch.float64, device=DEV)
            be.col_dots(X, Y, d)
            assert rel(d, (X.double() * Y.double()).sum(0)) < 1e-12
            Z = Y.clone()
            be.axpby(X, Z, 2.0, -3.0)
            assert rel(Z, 2 * X - 3 * Y) < (1e-6 if dt == torch.float32 else 1e-14)
            sq = (X.double()**2).sum(0)
            O = torch.empty_like(X)
            be.col_scale(X, O, sq, take_sqrt=True, mode=2)
            assert rel(O, X / torch.sqrt(sq).to(dt)) < (1e-6 if dt == torch.float32 else 1e-14)


def test_vector_sweeps_cluster_reduction(cb):
    """Sweeps that would end in many same-address fp64 atomics (grid x columns >= 48K) run as 8-CTA thread-block
    clusters whose column sums are folded through distributed shared memory first (csrc/sweep.cuh).  Same sums as the
    plain launch (up to the fp64 summation order) and as a double-precision torch reduction: the cluster path, the
    plain path just below the threshold, a grid that is not a multiple of the cluster size before rounding, several
    column slabs (k > 1024), and a gated launch (every CTA of every cluster leaves before the first cluster sync)."""
    be = cb.backend
    for dt in (torch.float32, torch.float64):
        for n, k in [(40000, 128), (40000, 256), (9001, 512), (300000, 32), (3000, 1500), (1 << 18, 64)]:
            torch.manual_seed(n + k)
            X = torch.randn(n, k, dtype=dt, device=DEV)
            Y = torch.randn(n, k, dtype=dt, device=DEV)
            d = torch.zeros(k, dtype=torch.float64, device=DEV)
            be.col_dots(X, Y, d)
            assert rel(d, (X.double() * Y.double()).sum(0)) < 1e-12, (dt, n, k)
            d2 = torch.zeros(k, dtype=torch.float64, device=DEV)
            be.col_dots(X, Y, d2)
            assert rel(d, d2) < 1e-14                            # run to run: only the order of the fp64 atomics moves
    # a CG solve whose sweeps take the cluster path (128 columns), against a dense solve
    n, k = 20000, 128
    g = torch.Generator().manual_seed(3)
    dg = (1.0 + torch.rand(n, dtype=torch.float64, generator=g)).to(DEV)
    lo = (0.3 * torch.rand(n - 1, dtype=torch.float64, generator=g)).to(DEV)
    A = cb.PSD(cb.ops.Tridiagonal(lo, dg, lo))
    B = torch.randn(n, k, dtype=torch.float64, generator=g).to(DEV)
    x, info = cb.linalg.cg(A, B, tol=1e-11, max_iters=500)
    R = dg[:, None] * x
    R[1:] += lo[:, None] * x[:-1]
    R[:-1] += lo[:, None] * x[1:]
    assert rel(R, B) < 1e-9
