"""GPU parity tests at BASELINE.json's shapes (VERDICT r1 "next round" item 1): the CUDA path against the CPU oracle on
identical seeded inputs at the sizes the bench reports, not toy sizes.

  cfg2  CSR 5-point Laplacian, 2048^2 grid (n = 4,194,304), 64 RHS, fp32: one matmat, the fused pAp dots and the first
        10 CG iterations (`info['errors']` and x), i.e. exactly the kernel variants bench.py times;
  cfg3  Kronecker(64x64 x3) + 0.1 I, n = 262,144, 128 RHS, fp32, 20 CG iterations, tensor-core and exact SIMT paths;
  cfg4  Kronecker(128,128,64) + Diagonal, n = 2^20, fp32: Lanczos alpha / beta over 20 steps for 2 probes and the SLQ
        log-determinant value of those probes;
  cfg5  graph Laplacian reduced to 2^20 nodes (avg degree 16), fp64, Lanczos m = 32: alpha / beta and the top Ritz values.

Tolerances are the north-star bar (1e-5 fp32, 1e-10 fp64).  At these sizes the fp32 ORACLE (= the reference's fp32
arithmetic on the CPU) is itself 1e-4 away from exact arithmetic from the first iteration on: its column reductions
over n = 2^18 .. 2^22 rows accumulate sequentially in fp32 (first B200 run of this file: cfg2 errors[0] 0.372484 in the
fp32 oracle, 0.3725923 in its own fp64 run, 0.372592 on the GPU, whose reductions accumulate in fp64).  The fp32
configs are therefore held (a) to the bar against the oracle run in fp64 on the same fp32 inputs, and (b) against the
fp32 oracle to within that oracle's own measured distance from its fp64 run.  fp64 (cfg5) is held to 1e-10 over the
window in which the oracle is stable under a 1e-15 perturbation.  The oracle needs ~2-3 minutes of host time for the
whole file."""
import numpy as np
import pytest
import torch

from tests import problems as pb

pytestmark = pytest.mark.gpu

F32_TOL, F64_TOL = 1e-5, 1e-10
DEV = "cuda:0"
# BASELINE shapes (module constants so that a CPU dry run of the test bodies can shrink them)
CFG3_D, CFG3_K = 64, 128
CFG4_DIMS = (128, 128, 64)
CFG5_LOG2N = 20


def rel(a, b):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    a, b = a.astype(np.float64), b.astype(np.float64)
    den = 0.5 * (np.linalg.norm(a) + np.linalg.norm(b))
    return 0.0 if den == 0 else float(np.linalg.norm(a - b) / den)


def window_of(a, b, rtol):
    """Leading entries over which two ORACLE traces (base and perturbed) agree to rtol / 4."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    m = min(len(a), len(b))
    bad = np.nonzero(np.abs(a[:m] - b[:m]) > 0.25 * rtol * np.abs(b[:m]))[0]
    return int(bad[0]) if len(bad) else m


def assert_trace(got, o64, o32, rtol, what):
    """got vs the fp64 oracle at rtol, and vs the fp32 oracle within that oracle's own distance from fp64."""
    got, o64, o32 = (np.asarray(a, dtype=np.float64) for a in (got, o64, o32))
    assert got.shape == o64.shape == o32.shape, (what, got.shape, o64.shape, o32.shape)
    np.testing.assert_allclose(got, o64, rtol=rtol, err_msg=f"{what}: vs the fp64 oracle")
    floor = np.abs(o32 - o64)
    bad = np.abs(got - o32) > 2.0 * floor + rtol * np.abs(o64)
    assert not bad.any(), (what, "vs the fp32 oracle beyond its own fp32 noise", got[bad], o32[bad], o64[bad])
    return float((floor / np.abs(o64)).max())


@pytest.fixture(scope="module")
def cb():
    import cola_b200
    assert torch.cuda.is_available()
    cola_b200.backend.lib()
    cola_b200.rng.PROBE_DEVICE = "cpu"
    return cola_b200


@pytest.fixture(scope="module")
def ko():
    from oracle import krylov_oracle
    torch.set_num_threads(max(1, torch.get_num_threads()))
    return krylov_oracle


# ------------------------------------------------------------------------------------------------ cfg2
@pytest.mark.parametrize("g", [256, 2048])
def test_cfg2_spmm_64rhs_vs_oracle(g, cb, ko):
    """The SpMM variant CG runs on cfg2 (64 fp32 right-hand sides: 16 lanes x float4 per row) against the oracle's
    torch.sparse_csr @ dense, plain and with the fused shift / diagonal / p^T A p epilogue."""
    from bench import laplacian_coo, rhs_block
    data, rows, cols, shape = laplacian_coo(g, torch.float32, "cpu")
    n = shape[0]
    X = rhs_block(n, 64, seed=11)
    Ao = ko.SparseOp(data, rows, cols, shape)
    Yo = Ao.matmat(X)
    S = cb.ops.Sparse(data.to(DEV), rows.to(DEV), cols.to(DEV), shape)
    Xd = X.to(DEV)
    Y = S @ Xd
    assert rel(Y, Yo) < 1e-6, rel(Y, Yo)
    # element-wise too: a 5-term fp32 sum of O(1) values
    assert float((Y.cpu() - Yo).abs().max()) < 2e-5
    dg = torch.rand(n, generator=torch.Generator().manual_seed(5)) + 0.5
    A = cb.PSD(S + 0.25 * cb.ops.I_like(S) + cb.ops.Diagonal(dg.to(DEV)))
    Y2 = torch.empty_like(Xd)
    dots = torch.zeros(64, dtype=torch.float64, device=DEV)
    A.matmat_into(Xd, Y2, dots=dots)
    ref = Yo.double() + (0.25 + dg.double())[:, None] * X.double()
    assert rel(Y2, ref) < 1e-6
    assert rel(dots, (X.double() * ref).sum(0)) < 1e-6
    # fp32 blocks: the products of a tile's rows (<= 8 per thread) are summed in fp32 and folded into the fp64
    # accumulators once per tile (csr_spmm_pipe_kernel), so the fused dots agree with an all-fp64 sum to ~1e-7
    assert rel(dots, (Xd.double() * Y2.double()).sum(0)) < 1e-6
    # ragged widths of the same operator: 48 (12 lanes), 16, 8 columns
    for k in (48, 16, 8):
        assert rel(S @ Xd[:, :k].contiguous(), Yo[:, :k]) < 1e-6, k


def test_cfg2_full_scale_cg_trace(cb, ko):
    """BASELINE config 2 at full size: first 10 CG iterations against the oracle on the same operator and RHS block."""
    from bench import laplacian_coo, rhs_block
    iters = 10
    data, rows, cols, shape = laplacian_coo(2048, torch.float32, "cpu")
    B = rhs_block(shape[0], 64, seed=0)
    xo, _, its_o, info_o = ko.cg(ko.SparseOp(data, rows, cols, shape), B, tol=1e-30, max_iters=iters)
    x64, _, _, info_64 = ko.cg(ko.SparseOp(data.double(), rows, cols, shape), B.double(), tol=1e-30, max_iters=iters)
    A = cb.PSD(cb.ops.Sparse(data.to(DEV), rows.to(DEV), cols.to(DEV), shape))
    x, info = cb.linalg.CG(tol=1e-30, max_iters=iters)(A, B.to(DEV))
    assert info["iterations"] == info_o["iterations"] == info_64["iterations"] == iters + 1
    floor = assert_trace(info["errors"], info_64["errors"], info_o["errors"], F32_TOL, "cfg2 info['errors']")
    assert rel(x, x64) < F32_TOL, rel(x, x64)
    assert rel(x, xo) < F32_TOL + 2 * rel(xo, x64), (rel(x, xo), rel(xo, x64))
    print(f"cfg2: trace rel vs fp64 oracle {rel(info['errors'], info_64['errors']):.2e}, x rel {rel(x, x64):.2e}; "
          f"fp32 oracle's own distance from fp64: trace {floor:.2e}, x {rel(xo, x64):.2e}")


# ------------------------------------------------------------------------------------------------ cfg3
def _cfg3(ko, dtype=torch.float32):
    Fs = [pb.kron_factor(CFG3_D, dtype, 60 + i) for i in range(3)]
    n = CFG3_D**3
    Ao = ko.SumOp(ko.KroneckerOp(*[ko.DenseOp(F) for F in Fs]), ko.ScaledIdentityOp(0.1, n, dtype))
    return Fs, n, Ao


def test_cfg3_full_scale_cg_trace(cb, ko):
    """BASELINE config 3 at full size (64^3, 128 RHS, fp32): 20 CG iterations, tensor-core (3xTF32) and exact SIMT
    contraction paths, against the oracle in fp64 (bar) and in fp32 (within its own noise)."""
    iters = 20
    Fs, n, Ao = _cfg3(ko)
    B = torch.randn(n, CFG3_K, generator=torch.Generator().manual_seed(0))
    xo, _, _, info_o = ko.cg(Ao, B, tol=1e-30, max_iters=iters)
    A64 = ko.SumOp(ko.KroneckerOp(*[ko.DenseOp(F.double()) for F in Fs]), ko.ScaledIdentityOp(0.1, n, torch.float64))
    x64, _, _, info_64 = ko.cg(A64, B.double(), tol=1e-30, max_iters=iters)
    # window: where the fp64 oracle is insensitive (at the bar) to an fp32-rounding-sized perturbation of the factors
    noise = [1.0 + 6e-8 * torch.randn(F.shape, dtype=torch.float64, generator=torch.Generator().manual_seed(9 + i))
             for i, F in enumerate(Fs)]
    Ap = ko.SumOp(ko.KroneckerOp(*[ko.DenseOp(F.double() * z) for F, z in zip(Fs, noise)]),
                  ko.ScaledIdentityOp(0.1, n, torch.float64))
    _, _, _, info_p = ko.cg(Ap, B.double(), tol=1e-30, max_iters=iters)
    window = window_of(info_p["errors"], info_64["errors"], F32_TOL)
    assert window >= 8, window
    K = cb.ops.Kronecker(*[cb.PSD(cb.ops.Dense(F.to(DEV))) for F in Fs])
    A = cb.PSD(K + 0.1 * cb.ops.I_like(K))
    core = A.plan().terms[0][1][0]
    for tc in (True, False):
        core.use_tensor_cores = tc
        if tc:
            assert core._tc_ok(B.to(DEV)) == (CFG3_D == 64)
        x, info = cb.linalg.CG(tol=1e-30, max_iters=iters)(A, B.to(DEV))
        assert info["iterations"] == info_o["iterations"] == info_64["iterations"]
        floor = assert_trace(info["errors"][:window], info_64["errors"][:window], info_o["errors"][:window], F32_TOL,
                             f"cfg3 info['errors'], tensor cores: {tc}")
        assert rel(x, x64) < max(F32_TOL, 4 * rel(xo, x64)), (tc, rel(x, x64), rel(xo, x64))
        print(f"cfg3 tc={tc}: window {window}/{len(info_o['errors'])}, trace rel {rel(info['errors'][:window], info_64['errors'][:window]):.2e}, "
              f"x rel {rel(x, x64):.2e} (fp32 oracle vs fp64: trace {floor:.2e}, x {rel(xo, x64):.2e})")
    core.use_tensor_cores = True


# ------------------------------------------------------------------------------------------------ cfg4
def test_cfg4_shape_lanczos_and_slq(cb, ko):
    """BASELINE config 4's operator (Kronecker(128,128,64) + Diagonal, n = 2^20, fp32): 20 Lanczos steps for 2 probes
    -- alpha / beta -- and the SLQ log-determinant value of the same probes."""
    dims, m = CFG4_DIMS, 20
    Fs = [pb.kron_factor(d, torch.float32, 70 + i) for i, d in enumerate(dims)]
    n = dims[0] * dims[1] * dims[2]
    dg = torch.rand(n, generator=torch.Generator().manual_seed(3)) + 0.5
    Z = torch.randn(n, 2, generator=torch.Generator().manual_seed(42))
    Ao = ko.SumOp(ko.KroneckerOp(*[ko.DenseOp(F) for F in Fs]), ko.DiagonalOp(dg))
    _, ao, bo, info_o = ko.lanczos(Ao, Z, m, 1e-7)
    A64 = ko.SumOp(ko.KroneckerOp(*[ko.DenseOp(F.double()) for F in Fs]), ko.DiagonalOp(dg.double()))
    _, a64, b64, _ = ko.lanczos(A64, Z.double(), m, 1e-7)
    # window: leading coefficients on which the fp64 oracle is insensitive (at the bar) to an fp32-rounding-sized
    # perturbation of the diagonal
    zn = 1.0 + 6e-8 * torch.randn(n, dtype=torch.float64, generator=torch.Generator().manual_seed(4))
    Ap = ko.SumOp(ko.KroneckerOp(*[ko.DenseOp(F.double()) for F in Fs]), ko.DiagonalOp(dg.double() * zn))
    _, ap, bp, _ = ko.lanczos(Ap, Z.double(), m, 1e-7)
    wa = min(window_of(ap[p].numpy(), a64[p].numpy(), F32_TOL) for p in range(2))
    wb = min(window_of(bp[p].numpy(), b64[p].numpy(), F32_TOL) for p in range(2))
    w = min(wa, wb)
    assert w >= 6, (wa, wb)
    K = cb.ops.Kronecker(*[cb.PSD(cb.ops.Dense(F.to(DEV))) for F in Fs])
    A = cb.PSD(K + cb.ops.Diagonal(dg.to(DEV)))
    Q, T, info = cb.linalg.Lanczos(start_vector=Z.to(DEV), max_iters=m, tol=1e-7)(A)
    alpha, beta = T.alpha[..., 0].cpu(), T.beta[..., 0].cpu()
    assert info["iterations"] == info_o["iterations"]
    assert tuple(alpha.shape) == tuple(ao.shape) and tuple(beta.shape) == tuple(bo.shape)
    assert_trace(alpha[:, :w].numpy(), a64[:, :w].numpy(), ao[:, :w].numpy(), F32_TOL, "cfg4 alpha")
    assert_trace(beta[:, :w].numpy(), b64[:, :w].numpy(), bo[:, :w].numpy(), F32_TOL, "cfg4 beta")
    assert rel(alpha, a64) < 100 * F32_TOL and rel(beta, b64) < 100 * F32_TOL
    # SLQ value of the same two probes (quadrature is a smooth function of T: tight even past the window)
    from cola_b200.linalg import stochastic
    est = stochastic.slq_per_probe(A, torch.log, Z.to(DEV), m, 1e-7)
    est_o = ko.slq_per_probe(Ao, torch.log, Z, m, 1e-7)
    est_64 = ko.slq_per_probe(A64, torch.log, Z.double(), m, 1e-7)
    floor = float((est_o.double() - est_64).abs().max() / est_64.abs().max())     # the oracle's own fp32 noise
    assert rel(est, est_64) < F32_TOL, (rel(est, est_64), floor)
    assert rel(est, est_o) < F32_TOL + 2 * floor, (rel(est, est_o), floor)
    print(f"cfg4: coefficient window {w}/{m}, slq rel vs fp64 oracle {rel(est, est_64):.2e} (fp32 oracle vs fp64 {floor:.2e})")


# ------------------------------------------------------------------------------------------------ cfg5
def test_cfg5_shape_lanczos_eig(cb, ko):
    """BASELINE config 5 reduced to 2^20 nodes (avg degree 16, fp64): Lanczos m = 32 with full reorthogonalisation,
    single start vector: alpha / beta against the oracle and the top Ritz values."""
    n, m = 1 << CFG5_LOG2N, 32
    data, rows, cols, shape = pb.graph_laplacian_coo(n, 8, torch.float64, 77)
    v0 = torch.randn(n, dtype=torch.float64, generator=torch.Generator().manual_seed(7))
    Ao = ko.SparseOp(data, rows, cols, shape)
    _, ao, bo, info_o = ko.lanczos(Ao, v0, m, 1e-12)
    ao, bo = ao.reshape(1, -1), bo.reshape(1, -1)              # 1-D start: the oracle drops the batch dimension
    noise = 1.0 + 1e-15 * torch.randn(n, dtype=torch.float64, generator=torch.Generator().manual_seed(8))
    _, a2, b2, _ = ko.lanczos(Ao, v0 * noise, m, 1e-12)
    a2, b2 = a2.reshape(1, -1), b2.reshape(1, -1)
    w = min(window_of(ao[0].numpy(), a2[0].numpy(), F64_TOL), window_of(bo[0].numpy(), b2[0].numpy(), F64_TOL))
    assert w >= 8, w
    L = cb.SelfAdjoint(cb.ops.Sparse(data.to(DEV), rows.to(DEV), cols.to(DEV), shape))
    Q, T, info = cb.linalg.Lanczos(start_vector=v0.to(DEV), max_iters=m, tol=1e-12)(L)
    alpha, beta = T.alpha[..., 0].cpu().reshape(1, -1), T.beta[..., 0].cpu().reshape(1, -1)
    assert info["iterations"] == info_o["iterations"]
    np.testing.assert_allclose(alpha[0, :w].numpy(), ao[0, :w].numpy(), rtol=F64_TOL)
    np.testing.assert_allclose(beta[0, :w].numpy(), bo[0, :w].numpy(), rtol=F64_TOL)
    assert rel(alpha, ao) < 1e-8 and rel(beta, bo) < 1e-8
    # top Ritz values: eigenvalues of T are a stable function of (alpha, beta)
    ev = torch.linalg.eigvalsh(ko.tridiag_dense(ao, bo))[0]
    vals, _ = cb.linalg.eig(L, 8, "LM", cb.linalg.Lanczos(start_vector=v0.to(DEV), max_iters=m, tol=1e-12))
    assert rel(vals, ev[-8:]) < F64_TOL, rel(vals, ev[-8:])
    print(f"cfg5: coefficient window {w}/{m}")
