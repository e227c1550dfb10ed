"""cola_b200.plugin against the REAL reference on CPU tensors (the live /root/reference tree in the build container,
else the copy installed in baseline/_ref; the CUDA counterpart is tests/test_plugin_gpu.py).

Three things are pinned here, all on CPU:
  1. install()/uninstall() rebind exactly the documented symbols, and with CPU operators every call falls through
     to the reference code (results unchanged);
  2. from_cola maps reference operator trees to the mirror classes one to one (classes, shapes, leaves, annotations);
  3. the adapters' layout glue: with FORCE_FAST_PATH and the CUDA loops replaced by oracle-backed stand-ins that
     speak this package's layouts ((m+2, n, b) basis, fp64 accumulators), the reference's own public API
     (CG, Lanczos, Arnoldi, stochastic_lanczos_quad) returns what the unpatched reference returns.
The CUDA loops themselves are compared with the same oracle in tests/test_gpu_parity.py.
"""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from baseline.install_ref import import_reference, reference_sys_path  # noqa: E402

if not reference_sys_path():
    pytest.skip("the reference is neither at /root/reference nor installed in baseline/_ref", allow_module_level=True)
cola = import_reference()  # noqa: E402  (the reference)
REF = "/root/reference"
from cola.linalg.inverse.cg import CG  # noqa: E402
from cola.linalg.tbd.slq import stochastic_lanczos_quad  # noqa: E402

from cola_b200 import ops as bops  # noqa: E402
from cola_b200 import plugin  # noqa: E402
b_lanczos = plugin.b_lanczos  # the submodule (cola_b200.linalg.lanczos the attribute is the function)
from oracle import krylov_oracle as ko  # noqa: E402

R = cola.ops


def make_tree(dtype=torch.float64):
    g = torch.Generator().manual_seed(3)
    K1 = torch.randn(4, 4, dtype=dtype, generator=g); K1 = K1 @ K1.T / 4 + 0.5 * torch.eye(4, dtype=dtype)
    K2 = torch.randn(3, 3, dtype=dtype, generator=g); K2 = K2 @ K2.T / 3 + 0.5 * torch.eye(3, dtype=dtype)
    d = torch.rand(12, dtype=dtype, generator=g) + 0.5
    A = R.Kronecker(cola.PSD(R.Dense(K1)), cola.PSD(R.Dense(K2))) + R.Diagonal(d)
    return cola.PSD(A), (K1, K2, d)


def mirror_to_oracle(M):
    if isinstance(M, bops.Sparse):
        return ko.SparseOp(M.data, M.row_indices, M.col_indices, M.shape)
    if isinstance(M, bops.Dense):
        return ko.DenseOp(M.A)
    if isinstance(M, bops.Identity):
        return ko.IdentityOp(M.shape[0], M.dtype)
    if isinstance(M, bops.ScalarMul):
        return ko.ScaledIdentityOp(float(M.c), M.shape[0], M.dtype)
    if isinstance(M, bops.Diagonal):
        return ko.DiagonalOp(M.diag)
    if isinstance(M, bops.Kronecker):
        return ko.KroneckerOp(*[mirror_to_oracle(m) for m in M.Ms])
    if isinstance(M, bops.BlockDiag):
        return ko.BlockDiagOp(*[mirror_to_oracle(m) for m in M.Ms], multiplicities=M.multiplicities)
    if isinstance(M, bops.Sum):
        return ko.SumOp(*[mirror_to_oracle(m) for m in M.Ms])
    if isinstance(M, bops.Product):
        return ko.ProductOp(*[mirror_to_oracle(m) for m in M.Ms])
    raise TypeError(type(M))


@pytest.fixture
def installed():
    plugin.install(cola)
    yield
    plugin.FORCE_FAST_PATH = False
    plugin.uninstall()


def test_install_rebinds_and_cpu_falls_through(installed):
    from importlib import import_module as im   # `cola.linalg.trace` the attribute is a function, not the package
    ra, rl = im("cola.linalg.decompositions.arnoldi"), im("cola.linalg.decompositions.lanczos")
    rcg, rslq = im("cola.linalg.inverse.cg"), im("cola.linalg.tbd.slq")
    rde = im("cola.linalg.trace.diagonal_estimation")
    for mod, name in [(rcg, "run_batched_cg"), (rl, "lanczos_fact"), (ra, "arnoldi_fact"),
                      (rde, "hutchinson_diag_estimate"), (rslq, "slq_fwd")]:
        assert getattr(mod, name).__module__ == "cola_b200.plugin", name
    assert R.Dense._matmat.__module__ == "cola_b200.plugin"
    A, _ = make_tree()
    b = torch.randn(12, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    x_in, _ = CG(tol=1e-10, max_iters=50)(A, b)
    plugin.uninstall()
    assert rcg.run_batched_cg.__module__ == "cola.linalg.inverse.cg"
    assert R.Dense._matmat.__module__ == "cola.ops.operators"
    x_ref, _ = CG(tol=1e-10, max_iters=50)(A, b)
    assert torch.equal(x_in, x_ref)          # CPU operators never leave the reference code


def test_from_cola_tree_mapping():
    A, (K1, K2, d) = make_tree()
    M = plugin.from_cola(A, cola)
    assert isinstance(M, bops.Sum) and M.shape == (12, 12) and M.isa(bops.PSD)
    kron, diag = M.Ms
    assert isinstance(kron, bops.Kronecker) and [type(f) for f in kron.Ms] == [bops.Dense, bops.Dense]
    assert kron.Ms[0].A is K1 and kron.Ms[1].A is K2 and all(f.isa(bops.PSD) for f in kron.Ms)
    assert isinstance(diag, bops.Diagonal) and diag.diag is d
    assert plugin.from_cola(A, cola) is M                         # cached on the reference instance
    # scalar * identity, product, block diagonal, transpose, sparse (CSR arrays taken as they are)
    I = R.I_like(A)
    P = plugin.from_cola(2.5 * I, cola)
    assert isinstance(P, bops.Product) and isinstance(P.Ms[0], bops.ScalarMul) and float(P.Ms[0].c) == 2.5
    B = plugin.from_cola(R.BlockDiag(R.Dense(K1), R.Dense(K2), multiplicities=[2, 1]), cola)
    assert isinstance(B, bops.BlockDiag) and B.multiplicities == [2, 1] and B.shape == (11, 11)
    rows = torch.tensor([0, 0, 1, 2, 2]); cols = torch.tensor([0, 2, 1, 0, 2])
    vals = torch.tensor([1., 2., 3., 4., 5.], dtype=torch.float64)
    S_ref = R.Sparse(vals, rows, cols, (3, 3))
    S = plugin.from_cola(S_ref, cola)
    assert isinstance(S, bops.Sparse) and S.nnz == 5 and S.max_row_nnz == 2
    assert torch.equal(S.indptr, S_ref.A.crow_indices().to(torch.int32))
    assert torch.equal(S.data, S_ref.A.values())
    from importlib import import_module as im
    Nys = im("cola.linalg.preconditioning.preconditioners").NystromPrecond(A, rank=4, key=A.xnp.PRNGKey(1))
    Pm = plugin.from_cola(Nys, cola)
    assert isinstance(Pm, bops.Sum) and isinstance(Pm.Ms[0], bops.Product) and isinstance(Pm.Ms[1], bops.Identity)
    Um, SUt = Pm.Ms[0].Ms[0].A, Pm.Ms[0].Ms[1].A
    X = torch.randn(12, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(4))
    assert torch.allclose(Um @ (SUt @ X) + X, Nys @ X, atol=1e-12)
    with pytest.raises(plugin.NotConvertible):
        plugin.from_cola(R.Dense(K1.to(torch.complex64)), cola)
    with pytest.raises(plugin.NotConvertible):
        plugin.from_cola(R.Householder(d[:, None]), cola)


def _standin_cg(M, b, x0, max_iters, tol, P, pbar=False):
    Po = None if isinstance(P, bops.Identity) else mirror_to_oracle(P)
    x, r, k, info = ko.cg(mirror_to_oracle(M), b, x0=x0, tol=tol, max_iters=max_iters, P=Po)
    return x, r, k, info


def _standin_lanczos_fact(M, rhs, max_iters=100, tol=1e-7, pbar=False):
    V, diag, sub, i, info = ko.lanczos_fact(mirror_to_oracle(M), rhs, max_iters, tol)
    Vm = V.permute(2, 1, 0).contiguous()                            # this package's (m+2, n, b) layout
    return b_lanczos.LanczosState(Vm, diag.T.double().contiguous(), (sub.T.double()**2).contiguous(), int(i), info)


def _standin_arnoldi_fact(M, rhs, max_iters, tol, pbar=False):
    Q, H, j, info = ko.arnoldi_fact(mirror_to_oracle(M), rhs, max_iters, tol)
    return Q.permute(2, 1, 0).contiguous(), H, int(j), info


def test_adapter_layouts_against_reference(installed, monkeypatch):
    b_arnoldi, b_cg = plugin.b_arnoldi, plugin.b_cg
    A, _ = make_tree()
    g = torch.Generator().manual_seed(1)
    B = torch.randn(12, 4, dtype=torch.float64, generator=g)
    v = torch.randn(12, dtype=torch.float64, generator=g)

    plugin.uninstall()
    x_ref, info_ref = CG(tol=1e-9, max_iters=40)(A, B)
    from importlib import import_module as im
    ref_lanczos = im("cola.linalg.decompositions.lanczos").lanczos
    ref_arnoldi = im("cola.linalg.decompositions.arnoldi").arnoldi
    Ql_ref, Tl_ref, il_ref = ref_lanczos(A, v, 8, 1e-12)
    Qb_ref, Tb_ref, _ = ref_lanczos(A, B, 6, 1e-12)
    Qa_ref, Ha_ref, ia_ref = ref_arnoldi(A, v, 7, 1e-12)
    key = A.xnp.PRNGKey(5)
    slq_ref = stochastic_lanczos_quad(A, torch.log, max_iters=10, tol=1e-9, vtol=0.25, key=key)
    Nys = im("cola.linalg.preconditioning.preconditioners").NystromPrecond(A, rank=4, key=A.xnp.PRNGKey(1))
    xp_ref, infop_ref = CG(tol=1e-9, max_iters=40, P=Nys)(A, B)

    plugin.install(cola)
    plugin.FORCE_FAST_PATH = True
    monkeypatch.setattr(b_cg, "run_batched_cg", _standin_cg)
    monkeypatch.setattr(b_lanczos, "lanczos_fact", _standin_lanczos_fact)
    monkeypatch.setattr(b_arnoldi, "arnoldi_fact", _standin_arnoldi_fact)
    monkeypatch.setattr(plugin.b_stoch, "lanczos_fact", _standin_lanczos_fact)      # name imported into the module
    monkeypatch.setattr(plugin.b_stoch, "probe_chunk", lambda *a, **k: 5)           # 16 probes in chunks of 5
    monkeypatch.setattr(plugin.b_stoch, "USE_TRIDIAG_QL", False)                    # the QL kernel is CUDA-only
    # operator applications inside the stand-ins must not recurse into the (CUDA-only) mirror matmats
    plugin.uninstall(); plugin.install(cola, matmats=False); plugin.FORCE_FAST_PATH = True

    x, info = CG(tol=1e-9, max_iters=40)(A, B)
    assert torch.allclose(x, x_ref, rtol=1e-10, atol=1e-12)
    assert info["iterations"] == info_ref["iterations"]
    xp, infop = CG(tol=1e-9, max_iters=40, P=Nys)(A, B)             # preconditioner converted by from_cola
    assert torch.allclose(xp, xp_ref, rtol=1e-9, atol=1e-11) and infop["iterations"] == infop_ref["iterations"]
    Ql, Tl, il = ref_lanczos(A, v, 8, 1e-12)
    assert Ql.to_dense().shape == Ql_ref.to_dense().shape
    assert torch.allclose(Ql.to_dense(), Ql_ref.to_dense(), atol=1e-10)
    assert torch.allclose(Tl.to_dense(), Tl_ref.to_dense(), atol=1e-10)
    assert il["iterations"] == il_ref["iterations"]
    Qb, Tb, _ = ref_lanczos(A, B, 6, 1e-12)
    # batched start block: the leaves of the vmapped Dense / Tridiagonal carry the batch dimension
    assert Qb.A.shape == Qb_ref.A.shape == (4, 12, 6)
    assert torch.allclose(Qb.A, Qb_ref.A, atol=1e-10)
    assert torch.allclose(Tb.beta, Tb_ref.beta, atol=1e-10) and torch.allclose(Tb.alpha, Tb_ref.alpha, atol=1e-10)
    Qa, Ha, ia = ref_arnoldi(A, v, 7, 1e-12)
    assert torch.allclose(Qa.to_dense(), Qa_ref.to_dense(), atol=1e-10)
    assert torch.allclose(Ha.to_dense(), Ha_ref.to_dense(), atol=1e-10)
    assert ia["iterations"] == ia_ref["iterations"]
    slq = stochastic_lanczos_quad(A, torch.log, max_iters=10, tol=1e-9, vtol=0.25, key=key)
    assert abs(float(slq) - float(slq_ref)) < 1e-8 * abs(float(slq_ref))


def test_reference_api_over_host_loops_on_emulated_kernels(installed):
    """The whole drop-in, minus the silicon: the reference's public API with install(), this package's own host
    loops (not stand-ins) and the test-only kernel statements of tests/host_harness.py: CG and Nystrom-preconditioned
    CG through solve(), structured matmats incl. KronSum / Tridiagonal, Lanczos, Arnoldi,
    eig, SLQ, logdet(Lanczos, Hutch), Hutch off-diagonals, exp / sqrt through the reference's own LanczosUnary /
    ArnoldiUnary on the rebound factorisations."""
    from importlib import import_module as im

    from cola.linalg.decompositions.decompositions import Arnoldi, Lanczos
    from cola.linalg.trace.diagonal_estimation import Hutch
    from tests.host_harness import emulated_kernels
    ref_lanczos = im("cola.linalg.decompositions.lanczos").lanczos
    ref_arnoldi = im("cola.linalg.decompositions.arnoldi").arnoldi
    A, (K1, K2, d) = make_tree()
    g = torch.Generator().manual_seed(11)
    B = torch.randn(12, 4, dtype=torch.float64, generator=g)
    v = torch.randn(12, dtype=torch.float64, generator=g)
    N = R.Dense(torch.randn(12, 12, dtype=torch.float64, generator=g) / 4 + 2 * torch.eye(12, dtype=torch.float64))
    KS = cola.PSD(R.KronSum(cola.PSD(R.Dense(K1)), cola.PSD(R.Dense(K2))))
    Tr = R.Tridiagonal(torch.randn(11, dtype=torch.float64, generator=g), torch.randn(12, dtype=torch.float64, generator=g),
                       torch.randn(11, dtype=torch.float64, generator=g))
    key = A.xnp.PRNGKey(5)

    def run():
        out = {}
        out["matmat"] = A @ B
        x, info = CG(tol=1e-9, max_iters=40)(A, B)
        out["cg_x"], out["cg_it"], out["cg_errors"] = x, info["iterations"], info["errors"]
        out["solve_x"] = cola.linalg.solve(KS, B, CG(tol=1e-9, max_iters=60))
        Nys = im("cola.linalg.preconditioning.preconditioners").NystromPrecond(A, rank=4, key=A.xnp.PRNGKey(1))
        xp, infop = CG(tol=1e-9, max_iters=40, P=Nys)(A, B)
        out["pcg_x"], out["pcg_it"] = xp, infop["iterations"]
        out["kronsum"] = KS @ B
        out["tridiag"] = Tr @ B
        Q, T, info = ref_lanczos(A, B, 6, 1e-12)
        out["lanczos_Q"], out["lanczos_beta"], out["lanczos_alpha"], out["lanczos_it"] = Q.A, T.beta, T.alpha, info["iterations"]
        Q, H, info = ref_arnoldi(N, v, 7, 1e-12)
        out["arnoldi_Q"], out["arnoldi_H"], out["arnoldi_it"] = Q.to_dense(), H.to_dense(), info["iterations"]
        Q, H, info = ref_arnoldi(Tr, v, 7, 1e-12)              # Tridiagonal inside a loop: converted, CSR core
        out["arnoldi_tridiag_H"] = H.to_dense()
        vals, vecs = cola.linalg.eig(A, 3, "LM", Lanczos(start_vector=v, max_iters=12, tol=1e-12))
        out["eigvals"] = vals
        out["slq"] = stochastic_lanczos_quad(A, torch.log, max_iters=10, tol=1e-9, vtol=0.25, key=key)
        out["logdet"] = cola.linalg.logdet(A, Lanczos(max_iters=12, tol=1e-10), Hutch(tol=2e-2, max_iters=2, key=key))
        out["hutch_k1"] = Hutch(tol=2e-2, max_iters=2, key=key)(A, 1)
        out["sqrtA"] = cola.linalg.sqrt(A, Lanczos(max_iters=12, tol=1e-12)) @ B
        Lc = R.Triangular(torch.linalg.cholesky(A.to_dense()), lower=True)       # Cholesky-style inverse as an operator
        Ainv = cola.linalg.inv(Lc.T) @ cola.linalg.inv(Lc)
        Q, T, info = ref_lanczos(cola.SelfAdjoint(Ainv), v, 6, 1e-12)             # TriangularInv cores inside a loop
        out["lanczos_trinv_beta"] = T.beta
        out["expN"] = cola.linalg.exp(N, Arnoldi(max_iters=12, tol=1e-12)) @ B
        return out

    plugin.uninstall()
    ref = run()
    plugin.install(cola)
    plugin.FORCE_FAST_PATH = True
    launched = []
    import cola_b200.backend as be
    with emulated_kernels():
        for name in ("mode_contract", "csr_spmm", "reorth_update", "mgs_chain", "tridiag_eig_first_row"):
            fn = getattr(be, name)
            setattr(be, name, (lambda f, n: (lambda *a, **k: (launched.append(n), f(*a, **k))[1]))(fn, name))
        cg_lib = be.lib()
        cg_call = cg_lib.call
        cg_lib.call = lambda name, *a: (launched.append(name), cg_call(name, *a))[1]
        got = run()
    # the B200 host loops really ran (not the reference's): their kernels were "launched"
    assert {"mode_contract", "csr_spmm", "reorth_update", "mgs_chain", "tridiag_eig_first_row", "cola_cg_update_xp_f64",
            "cola_cg_advance_f64"} <= set(launched)
    for name, r in ref.items():
        gval = got[name]
        if isinstance(r, int):
            assert gval == r, name
            continue
        r, gval = torch.as_tensor(r), torch.as_tensor(gval)
        assert gval.shape == r.shape and gval.dtype == r.dtype, name
        err = float((gval - r).abs().max() / r.abs().max())
        assert err < 1e-9, (name, err)


def test_reference_own_suite_over_host_loops(tmp_path):
    """The reference's OWN test suite (/root/reference/tests, its non-JAX default selection: 290 tests) run with
    install() active, CPU operators routed through the adapters and the kernels replaced by the test-only statements
    (tests/ref_suite_plugin.py): every torch fp32/fp64 case that reaches CG, Lanczos, Arnoldi, SLQ, Hutchinson or a
    structured matmat executes this package's host loops, everything else (complex, NumPy backend, autograd
    recordings, implicitly restarted variants) must fall through to the reference code.  All of them have to pass."""
    import json
    import subprocess
    if not os.path.isdir(os.path.join(REF, "tests")):
        pytest.skip("the reference's own test-suite is only in the build container's /root/reference tree")
    stats = tmp_path / "stats.json"
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1", COLA_B200_REF_SUITE_STATS=str(stats),
               PYTHONPATH=os.pathsep.join([os.path.join(HERE, "golden", "refshim"), REF, os.path.dirname(HERE)]))
    cmd = [sys.executable, "-m", "pytest", os.path.join(REF, "tests"), "-q", "-p", "no:cacheprovider", "-p",
           "tests.ref_suite_plugin", "-m", "not big and not tricky and not market and not jax", "--tb=line"]
    out = subprocess.run(cmd, cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=900)
    tail = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-2000:]
    assert out.returncode == 0, out.stdout[-4000:]
    assert " passed" in tail and "failed" not in tail and int(tail.split(" passed")[0].split()[-1]) >= 290, tail
    calls = json.load(open(stats))
    # the fast path was really taken: every loop family launched its kernels
    for name in ("cola_cg_*", "mode_contract", "lanczos_three_term", "reorth_update", "mgs_chain", "tridiag_eig_first_row",
                 "diag_matmat", "col_scale"):
        assert calls.get(name, 0) > 0, (name, calls)


def test_random_operator_trees_through_the_plugin():
    """Random expression trees (tests/test_differential_reference._random_tree) evaluated with the unpatched
    reference and again with install() over this package's host loops: matmat, transposed matmat, CG, logdet
    (Lanczos + Hutch), eig, sqrt, GMRES, Hutchinson diag and exp (Arnoldi) agree (113 trees were run this way without
    a mismatch; 25 are kept here)."""
    import importlib
    import random

    from cola.linalg.decompositions.decompositions import Arnoldi, Lanczos
    from cola.linalg.inverse.gmres import GMRES
    from cola.linalg.trace.diagonal_estimation import Hutch
    from tests.host_harness import emulated_kernels
    import cola_b200 as cb
    tree = importlib.import_module("tests.test_differential_reference")._random_tree
    f64 = torch.float64

    def surface(A, T, B, v, n, key):
        return dict(mm=T @ B, mmT=T.T @ B, cg=CG(tol=1e-10, max_iters=100)(A, B)[0],
                    ld=cola.linalg.logdet(A, Lanczos(max_iters=n, tol=1e-10), Hutch(tol=2e-2, max_iters=2, key=key)),
                    ev=cola.linalg.eig(A, 2, "LM", Lanczos(start_vector=v, max_iters=n, tol=1e-12))[0],
                    sq=cola.linalg.sqrt(A, Lanczos(max_iters=n, tol=1e-12)) @ B,
                    gm=GMRES(tol=1e-12, max_iters=n)(T, B)[0],
                    dg=cola.linalg.diag(T, 0, Hutch(tol=2e-2, max_iters=2, key=key)),
                    ex=cola.linalg.exp(T, Arnoldi(max_iters=n, tol=1e-12)) @ B)

    rng = random.Random(1)
    compared = 0
    for t in range(25):
        n = rng.choice([6, 8, 12])
        build = tree(rng, rng.choice([1, 2]), n)
        I = R.Identity((n, n), f64)
        B = torch.randn(n, 2, dtype=f64, generator=torch.Generator().manual_seed(t))
        v, key = B[:, 0].contiguous(), cb.rng.PRNGKey(4)
        try:
            T = build(R, cola.PSD, cola.SelfAdjoint)
            ref = surface(cola.PSD(T.T @ T + 0.5 * I), T, B, v, n, key)
        except Exception:        # a tree the reference itself cannot evaluate
            continue
        plugin.install(cola)
        plugin.FORCE_FAST_PATH = True
        try:
            with emulated_kernels():
                T2 = build(R, cola.PSD, cola.SelfAdjoint)
                got = surface(cola.PSD(T2.T @ T2 + 0.5 * I), T2, B, v, n, key)
        finally:
            plugin.FORCE_FAST_PATH = False
            plugin.uninstall()
        for name, r in ref.items():
            r, g = torch.as_tensor(r), torch.as_tensor(got[name])
            tol = 1e-5 if name in ("gm", "ex") else 1e-7
            assert g.shape == r.shape and float((g - r).abs().max()) <= tol * max(float(r.abs().max()), 1e-300), \
                (t, type(T).__name__, name)
        compared += 1
    assert compared >= 18
