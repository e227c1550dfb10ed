"""cola_b200.install() against the REAL reference on the B200.

The unmodified reference package is imported from baseline/_ref (installed there by baseline/install_ref.py in the
build container; git-ignored, shipped with the repo snapshot) -- or from /root/reference where that tree exists.  Every
case builds the reference's own operators twice: on the CPU, evaluated by the unpatched reference, and on `cuda:0`,
evaluated through the reference's public API with the plugin installed (no FORCE_FAST_PATH, no stand-ins: real CUDA
tensors, real kernels, the `(b, n, m+2)` permuted views and `Sparse.from_csr` of a CUDA `torch.sparse_csr` included).
The kernel-launch counter proves the fast path ran."""
import numpy as np
import pytest
import torch

from tests import problems as pb

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def cola():
    from baseline.install_ref import import_reference
    try:
        return import_reference()
    except ImportError as exc:                                  # pragma: no cover
        pytest.fail(f"the reference must travel to the GPU box in baseline/_ref: {exc}")


@pytest.fixture(scope="module")
def cb():
    import cola_b200
    assert torch.cuda.is_available()
    cola_b200.backend.lib()
    cola_b200.rng.PROBE_DEVICE = "cpu"     # probes drawn on the host generator: the CPU reference's draws
    return cola_b200


@pytest.fixture
def plugin(cola, cb):
    from cola_b200 import plugin as pl
    pl.uninstall()
    yield pl
    pl.FORCE_FAST_PATH = False
    pl.uninstall()


def aligned_sparse(cola, data, rows, cols, shape):
    """The reference's Sparse constructor sorts with a non-stable argsort and can misalign values and indices
    (DESIGN.md); inputs here are (row, col)-sorted, so the intended CSR values are `data` itself."""
    S = cola.ops.Sparse(data, rows, cols, shape)
    if not torch.equal(S.col_indices.to(torch.int32), S.A.col_indices()):
        S.data, S.row_indices, S.col_indices = data, rows, cols
        S.A = torch.sparse_csr_tensor(S.A.crow_indices(), S.A.col_indices(), data, size=shape)
    return S


def build(cola, device, dtype=torch.float64):
    """Reference operators of every hot-path type, from seeded host data, with their leaves on `device`."""
    R = cola.ops
    t = lambda x: x.to(device)                                   # noqa: E731
    K1, K2 = pb.kron_factor(12, dtype, 1), pb.kron_factor(10, dtype, 2)
    d = pb.t(pb.rs(3).uniform(size=120) + 0.5, dtype)
    kron = cola.PSD(R.Kronecker(cola.PSD(R.Dense(t(K1))), cola.PSD(R.Dense(t(K2)))) + R.Diagonal(t(d)))
    data, rows, cols, shape = pb.laplacian_2d_coo(24, dtype, shift=0.3)
    lap = cola.PSD(aligned_sparse(cola, t(data), t(rows), t(cols), shape))
    N = R.Dense(t(pb.nonsym_dense(48, dtype, 4)))
    blk = cola.PSD(R.BlockDiag(cola.PSD(R.Dense(t(pb.kron_factor(8, dtype, 5)))), cola.PSD(R.Dense(t(pb.kron_factor(6, dtype, 6)))),
                               multiplicities=[3, 2]))
    prod = R.Dense(t(pb.kron_factor(36, dtype, 7))) @ R.Diagonal(t(pb.t(pb.rs(8).uniform(size=36) + 1.0, dtype)))
    ksum = cola.PSD(R.KronSum(cola.PSD(R.Dense(t(K1))), cola.PSD(R.Dense(t(K2)))))
    return dict(kron=kron, lap=lap, N=N, blk=blk, prod=prod, ksum=ksum)


def surface(cola, ops, device, dtype=torch.float64):
    """The reference's public API on the hot path.  Returns {name: tensor | int | array}."""
    from importlib import import_module as im
    from cola.linalg.decompositions.decompositions import Arnoldi, Lanczos
    from cola.linalg.inverse.cg import CG
    from cola.linalg.inverse.gmres import GMRES
    from cola.linalg.tbd.slq import stochastic_lanczos_quad
    from cola.linalg.trace.diagonal_estimation import Hutch
    ref_lanczos = im("cola.linalg.decompositions.lanczos").lanczos
    ref_arnoldi = im("cola.linalg.decompositions.arnoldi").arnoldi
    out = {}
    g = lambda shape, seed: pb.randn_np(shape, dtype, seed).to(device)   # noqa: E731
    for name, A in ops.items():
        X = g((A.shape[1], 5), 20)
        out[f"matmat_{name}"] = A @ X
        out[f"matvec_{name}"] = A @ X[:, 0].contiguous()
    kron, lap, N, blk = ops["kron"], ops["lap"], ops["N"], ops["blk"]
    tight = 1e-10 if dtype == torch.float64 else 1e-6
    for name, A in (("kron", kron), ("lap", lap), ("blk", blk), ("ksum", ops["ksum"])):
        B = g((A.shape[0], 4), 21)
        x, info = CG(tol=tight, max_iters=300)(A, B)
        out[f"cg_x_{name}"], out[f"cg_it_{name}"], out[f"cg_errors_{name}"] = x, info["iterations"], info["errors"][:12]
    out["solve_lap"] = cola.linalg.solve(lap, g((576, 3), 22), CG(tol=tight, max_iters=400))
    Nys = im("cola.linalg.preconditioning.preconditioners").NystromPrecond(kron, rank=8, key=kron.xnp.PRNGKey(1))
    xp, infop = CG(tol=tight, max_iters=100, P=Nys)(kron, g((120, 4), 23))
    out["pcg_x"], out["pcg_it"] = xp, infop["iterations"]
    Q, T, info = ref_lanczos(kron, g((120, 3), 24), 10, 1e-12)           # batched start block: (b, n, m) view
    out["lanczos_Q"], out["lanczos_beta"], out["lanczos_alpha"], out["lanczos_it"] = Q.A, T.beta, T.alpha, info["iterations"]
    Q, T, info = ref_lanczos(lap, g((576,), 25), 12, 1e-12)
    out["lanczos1_Q"], out["lanczos1_T"] = Q.to_dense(), T.to_dense()
    Q, H, info = ref_arnoldi(N, g((48,), 26), 9, 1e-12)
    out["arnoldi_Q"], out["arnoldi_H"], out["arnoldi_it"] = Q.to_dense(), H.to_dense(), info["iterations"]
    out["gmres_x"] = cola.linalg.solve(N, g((48, 2), 27), GMRES(tol=1e-10, max_iters=20))
    vals, vecs = cola.linalg.eig(lap, 4, "LM", Lanczos(start_vector=g((576,), 28), max_iters=40, tol=1e-12))
    out["eigvals"] = vals
    key = kron.xnp.PRNGKey(5)
    out["slq"] = stochastic_lanczos_quad(kron, torch.log, max_iters=20, tol=1e-9, vtol=0.2, key=key)
    out["logdet"] = cola.linalg.logdet(kron, Lanczos(max_iters=20, tol=1e-10), Hutch(tol=2e-2, max_iters=2, key=key))
    out["hutch_diag"] = Hutch(tol=2e-2, max_iters=2, key=key)(lap, 0)
    out["sqrtA"] = cola.linalg.sqrt(kron, Lanczos(max_iters=20, tol=1e-12)) @ g((120, 3), 29)
    out["expN"] = cola.linalg.exp(N, Arnoldi(max_iters=16, tol=1e-12)) @ g((48, 2), 30)
    return out


def compare(ref, got, tol):
    for name, r in ref.items():
        gval = got[name]
        if isinstance(r, (int, np.integer)):
            assert gval == r, (name, gval, r)
            continue
        r = torch.as_tensor(r).detach().cpu()
        gval = torch.as_tensor(gval).detach().cpu()
        assert tuple(gval.shape) == tuple(r.shape) and gval.dtype == r.dtype, (name, gval.shape, r.shape, gval.dtype, r.dtype)
        if r.is_complex():
            r, gval = torch.view_as_real(r), torch.view_as_real(gval)
        err = float((gval.double() - r.double()).abs().max() / max(float(r.double().abs().max()), 1e-300))
        assert err < tol, (name, err)


def test_reference_api_on_cuda_through_install(cola, cb, plugin):
    """fp64: the reference on the CPU (unpatched) vs the reference's API on CUDA operators with install()."""
    ref = surface(cola, build(cola, "cpu"), "cpu")
    plugin.install(cola)
    before = cb.backend.lib().launch_count()
    got = surface(cola, build(cola, DEV), DEV)
    launched = cb.backend.lib().launch_count() - before
    assert launched > 500, launched                              # the loops and matmats really ran on the kernels
    for k, v in got.items():
        if torch.is_tensor(v):
            assert v.is_cuda, k
    compare(ref, got, 2e-8)
    # after uninstall() the same CUDA operators run the reference's eager code again (no kernel launches)
    plugin.uninstall()
    before = cb.backend.lib().launch_count()
    ops = build(cola, DEV)
    y = ops["kron"] @ pb.randn_np((120, 5), torch.float64, 20).to(DEV)
    assert cb.backend.lib().launch_count() == before
    compare({"y": ref["matmat_kron"]}, {"y": y}, 1e-12)


def test_reference_api_on_cuda_fp32(cola, cb, plugin):
    ref = surface(cola, build(cola, "cpu", torch.float32), "cpu", torch.float32)
    plugin.install(cola)
    got = surface(cola, build(cola, DEV, torch.float32), DEV, torch.float32)
    # iteration counts of tolerance-limited fp32 solves may flip at the threshold; everything else at fp32 accuracy
    for k in [k for k in ref if k.endswith("_it") or "_it_" in k]:
        assert abs(ref.pop(k) - got.pop(k)) <= 2, k
    for k in [k for k in ref if k.startswith("cg_errors_")]:
        r, g_ = np.asarray(ref.pop(k)), np.asarray(got.pop(k))
        np.testing.assert_allclose(g_[:6], r[:6], rtol=1e-4)
    compare(ref, got, 5e-3)                                      # converged solutions: cond * eps apart


def test_from_cola_on_cuda_leaves(cola, cb, plugin):
    """from_cola on CUDA operators: leaves are taken by reference (no copies), the CSR arrays of the reference's CUDA
    torch.sparse_csr are wrapped as they are, and the Lanczos basis handed back is a strided view."""
    from importlib import import_module as im
    ops = build(cola, DEV)
    M = plugin.from_cola(ops["lap"], cola)
    A = ops["lap"].A
    assert M.indptr.is_cuda and M.indices.data_ptr() == A.col_indices().data_ptr() or M.indices.dtype == torch.int32
    assert M.data.data_ptr() == A.values().data_ptr()
    K = plugin.from_cola(ops["kron"], cola)
    assert K.Ms[0].Ms[0].A.data_ptr() == ops["kron"].Ms[0].Ms[0].A.data_ptr()
    plugin.install(cola)
    lz = im("cola.linalg.decompositions.lanczos")
    V0 = pb.randn_np((120, 3), torch.float64, 31).to(DEV)
    Q, T, info = lz.lanczos(ops["kron"], V0, 8, 1e-12)
    assert tuple(Q.A.shape) == (3, 120, 8) and not Q.A.is_contiguous() and Q.A.is_cuda     # (b, n, m) view of (m, n, b)
    G = Q.A.transpose(1, 2) @ Q.A
    assert float((G - torch.eye(8, dtype=torch.float64, device=DEV)).abs().max()) < 1e-12


def test_reference_backward_rules_with_solves_on_kernels(cola, cb, plugin):
    """Under install() the reference's custom backward rules (cg_bwd cg.py:72-86, slq_bwd slq.py:10-31) keep working:
    forward solve, backward solve and SLQ forward on the kernels; gradients equal the CPU reference's."""
    from cola.linalg.inverse.cg import CG
    R = cola.ops
    dtype = torch.float64

    def loss(device):
        K1 = pb.kron_factor(12, dtype, 1).to(device).requires_grad_(True)
        d = pb.t(pb.rs(3).uniform(size=120) + 0.5, dtype).to(device).requires_grad_(True)
        A = cola.PSD(R.Kronecker(cola.PSD(R.Dense(K1)), cola.PSD(R.Dense(pb.kron_factor(10, dtype, 2).to(device)))) + R.Diagonal(d))
        b = pb.randn_np((120, 2), dtype, 40).to(device).requires_grad_(True)
        x = cola.linalg.solve(A, b, CG(tol=1e-12, max_iters=300))
        val = (x**2).sum()
        val.backward()
        return val.detach().cpu(), K1.grad.cpu(), d.grad.cpu()

    ref = loss("cpu")
    plugin.install(cola)
    before = cb.backend.lib().launch_count()
    got = loss(DEV)
    assert cb.backend.lib().launch_count() - before > 50
    for r, g_ in zip(ref, got):
        assert float((r - g_).abs().max() / r.abs().max()) < 1e-7
