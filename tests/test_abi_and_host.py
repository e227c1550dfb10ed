"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol include/cola_b200.h declares,
the operator compiler flattens trees correctly, and the product path refuses CPU tensors loudly."""
import ctypes
import os

import pytest
import torch

import cola_b200 as cb
from cola_b200 import backend as be
from cola_b200 import ops


def test_library_exports_every_declared_symbol():
    decls = be.parse_header()
    assert len(decls) >= 30
    cdll = ctypes.CDLL(be.LIB_PATH)
    for name in decls:
        assert hasattr(cdll, name), f"{name} declared in include/cola_b200.h but not exported"
    L = be.lib()
    assert L.cdll.cola_version() >= 1
    for family in ("csr_spmm", "mode_contract", "diag_matmat", "col_dots", "col_scale", "axpby", "cg_update_r",
                   "cg_update_xp", "cg_tol", "cg_advance", "reorth_dots", "reorth_update", "lanczos_three_term",
                   "mgs_link"):
        for sfx in ("f32", "f64"):
            assert f"cola_{family}_{sfx}" in decls


def test_no_gpu_reports_status_not_crash():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = be.lib()
    rc = L.cdll.cola_device_info(None, None, None)
    assert rc == -3  # COLA_E_NOGPU
    assert b"no CUDA device" in L.cdll.cola_last_error()


def test_bad_arguments_are_rejected_without_touching_the_gpu():
    L = be.lib()
    rc = L.cdll.cola_col_dots_f32(None, None, 4, 4, 4, None, None, None)
    assert rc == -1 and b"null pointer" in L.cdll.cola_last_error()
    with pytest.raises(RuntimeError, match="status -1"):
        L.call("cola_reorth_dots_f64", None, 0, 0, 1, None, 4, 4, None, None, None)


def test_cpu_tensors_raise_no_fallback():
    A = cb.PSD(ops.Dense(torch.eye(4)))
    with pytest.raises(RuntimeError, match="CUDA-only"):
        A @ torch.ones(4, 2)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        cb.linalg.CG()(A, torch.ones(4, 2))
    with pytest.raises(RuntimeError, match="CUDA-only"):
        cb.linalg.Lanczos(start_vector=torch.ones(4))(A)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(be, "LIB_PATH", os.path.join(os.path.dirname(be.LIB_PATH), "nope.so"))
    monkeypatch.setattr(be, "_LIB", None)
    with pytest.raises(RuntimeError, match="no CPU or eager-torch fallback"):
        be.lib()


def test_plan_compiler_flattens_compositions():
    K = ops.Kronecker(ops.Dense(torch.eye(2)), ops.Dense(torch.eye(3)))
    A = K + 0.1 * ops.I_like(K) + ops.Diagonal(torch.ones(6))
    plan = A.plan()
    assert len(plan.terms) == 1 and type(plan.terms[0][1][0]).__name__ == "_KronCore"
    assert abs(plan.shift - 0.1) < 1e-7 and plan.diag is not None  # c is stored in the operator dtype, like the reference
    B = 2.0 * ops.Dense(torch.eye(6)) + ops.Dense(torch.ones(6, 6)) @ ops.Dense(torch.eye(6))
    pb = B.plan()
    assert [s for s, _ in pb.terms] == [2.0, 1.0] and len(pb.terms[1][1]) == 2
    assert "DenseCore" in pb.describe()
    C = 3.0 * ops.I_like(K)
    pc = C.plan()
    assert pc.terms == [] and pc.shift == 3.0


def test_annotations_follow_reference_rules():
    D = cb.PSD(ops.Dense(torch.eye(3)))
    assert D.isa(cb.PSD) and D.isa(cb.SelfAdjoint) and not ops.Dense(torch.eye(3)).isa(cb.PSD)
    K = ops.Kronecker(D, D)
    assert K.isa(cb.PSD)                                   # annotations.py:91-93
    S = K + 0.1 * ops.I_like(K)
    assert S.isa(cb.PSD)                                   # Sum intersect, Product with one non-scalar factor
    assert not (K + ops.Dense(torch.ones(9, 9))).isa(cb.SelfAdjoint)
    with pytest.raises(AssertionError, match="CG only valid for PSD"):
        cb.linalg.inv(ops.Dense(torch.eye(3)), cb.linalg.CG())
    with pytest.raises(ValueError, match="dimension mismatch"):
        ops.Sum(ops.Dense(torch.eye(3)), ops.Dense(torch.eye(4)))
    with pytest.raises(AssertionError, match="dimension mismatch"):
        D @ torch.ones(4, 1)


def test_sparse_constructor_builds_aligned_csr_on_cpu_tensors():
    data = torch.tensor([1., 2., 3., 4., 5., 6.])
    rows = torch.tensor([2, 0, 1, 2, 0, 2])
    cols = torch.tensor([0, 1, 3, 1, 3, 2])
    S = ops.Sparse(data, rows, cols, (3, 4))
    assert S.indptr.tolist() == [0, 2, 3, 6] and S.indptr.dtype == torch.int32
    assert S.indices.tolist() == [1, 3, 3, 0, 1, 2] and S.data.tolist() == [2., 5., 3., 1., 4., 6.]


def test_rng_matches_reference_key_chain():
    from oracle import krylov_oracle as ko
    assert cb.rng.PRNGKey(42) == ko.PRNGKey(42)
    z1 = cb.rng.randn(5, 3, dtype=torch.float32, device="cpu", key=cb.rng.PRNGKey(7))
    z2 = ko.keyed_randn(5, 3, dtype=torch.float32, key=ko.PRNGKey(7))
    assert torch.equal(z1, z2)
