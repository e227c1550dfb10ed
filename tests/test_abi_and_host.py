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
                   "mgs_link", "mgs_chain", "csr_spmm_tiled"):
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


def test_probe_blocks_are_split_into_vector_friendly_widths():
    """100 Hutchinson probes -> 64 + 32 + 4 (stochastic.friendly_chunks): every block a power of two no wider than
    the chunk limit, tails narrower than one 16-byte vector as a single block, ranges contiguous from `start`."""
    from cola_b200.linalg.stochastic import friendly_chunks
    assert friendly_chunks(100, 96, torch.float32) == [(0, 64), (64, 96), (96, 100)]
    assert friendly_chunks(1024, 64, torch.float32) == [(64 * i, 64 * (i + 1)) for i in range(16)]
    assert friendly_chunks(24, 64, torch.float32, start=12) == [(12, 28), (28, 36)]
    assert friendly_chunks(3, 64, torch.float32) == [(0, 3)]
    assert friendly_chunks(7, 64, torch.float64) == [(0, 4), (4, 6), (6, 7)]
    for k, cb_ in [(1, 1), (37, 8), (513, 128), (130, 64)]:
        ch = friendly_chunks(k, cb_, torch.float32)
        assert ch[0][0] == 0 and ch[-1][1] == k and all(a[1] == b[0] for a, b in zip(ch, ch[1:]))
        assert all(c1 - c0 <= max(cb_, 3) for c0, c1 in ch)


def test_sparse_from_csr_wraps_arrays_as_they_are():
    indptr = torch.tensor([0, 2, 3, 6], dtype=torch.int64)
    indices = torch.tensor([3, 1, 3, 2, 0, 1], dtype=torch.int64)          # deliberately unsorted inside rows
    data = torch.tensor([5., 2., 3., 6., 1., 4.])
    S = ops.Sparse.from_csr(indptr, indices, data, (3, 4))
    assert S.indptr.dtype == torch.int32 and S.indices.tolist() == [3, 1, 3, 2, 0, 1] and S.data is not None
    assert S.nnz == 6 and S.max_row_nnz == 3 and S.row_indices.tolist() == [0, 0, 1, 2, 2, 2]
    assert S.shape == (3, 4) and S.dtype == torch.float32


def test_preconditioned_cg_and_gmres_are_cuda_only():
    A = cb.PSD(ops.Dense(torch.eye(4)))
    with pytest.raises(RuntimeError, match="CUDA-only"):
        cb.linalg.GMRES(max_iters=3)(A, torch.ones(4, 2))
    with pytest.raises(RuntimeError, match="CUDA-only"):
        cb.linalg.CG(P=ops.Diagonal(torch.ones(4)))(A, torch.ones(4, 2))


def test_operand_on_another_device_is_refused(monkeypatch):
    """Kernels launch on the current device's stream; an operand on a different GPU of the same process must raise
    instead of being dereferenced by the wrong device (checked with a stand-in tensor: there is no GPU here)."""
    import types

    import torch

    from cola_b200 import backend as be
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    other = types.SimpleNamespace(is_cuda=True, device=types.SimpleNamespace(index=1))
    with pytest.raises(RuntimeError, match="current CUDA device"):
        be.require_cuda(other, "operand")
    be.require_cuda(types.SimpleNamespace(is_cuda=True, device=types.SimpleNamespace(index=0)), "operand")
    # an operand that requires grad is refused while autograd records (no silent cut of the graph), accepted otherwise
    wants_grad = types.SimpleNamespace(is_cuda=True, device=types.SimpleNamespace(index=0), requires_grad=True)
    with pytest.raises(RuntimeError, match="not differentiable"):
        be.require_cuda(wants_grad, "operand")
    with torch.no_grad():
        be.require_cuda(wants_grad, "operand")
