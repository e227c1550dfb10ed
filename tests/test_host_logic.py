"""Host-logic tests without a GPU: the bodies of the GPU parity tests (tests/test_gpu_parity*.py) for everything
that is orchestrated in Python -- plan compiler + matmat epilogues, CG / preconditioned CG (device-side stop rule,
16-iteration batches, trace reconstruction), Lanczos, Arnoldi, GMRES, power iteration, SLQ, Hutchinson, f(A)v --
are run on CPU tensors with the kernels replaced by tests/host_harness.py (statements of the header's semantics).
They check loop logic against the oracle and the reference's golden vectors; the kernels themselves, and the
CUDA-graph replay of CG batches, are checked only on the GPU."""
import pytest
import torch

import cola_b200
from tests import test_gpu_parity as gp
from tests import test_gpu_parity_next as gn
from tests.golden_cases import (ARNOLDI_CASES, CG_CASES, DIAG_CASES, GMRES_CASES, LANCZOS_CASES, MATMAT_PROBLEMS,
                                NEXT_CG_CASES, NEXT_MATMAT_PROBLEMS, PCG_CASES, POWER_CASES, UNARY_CASES)
from tests.host_harness import emulated_kernels


@pytest.fixture(params=[False, True], ids=["two-kernel-cgs2", "fused-cgs2"])
def emu_both(request, monkeypatch):
    monkeypatch.setattr(gp, "DEV", "cpu")
    monkeypatch.setattr(gn, "DEV", "cpu")
    with emulated_kernels(fused=request.param):
        yield cola_b200


@pytest.fixture
def emu(monkeypatch):
    monkeypatch.setattr(gp, "DEV", "cpu")
    monkeypatch.setattr(gn, "DEV", "cpu")
    with emulated_kernels():
        yield cola_b200


def test_gate_is_closed_outside_the_harness():
    A = cola_b200.ops.Dense(torch.eye(4))
    with emulated_kernels():
        assert torch.equal(A @ torch.ones(4, 2), torch.ones(4, 2))
    with pytest.raises(RuntimeError, match="CUDA-only"):
        A @ torch.ones(4, 2)


@pytest.mark.parametrize("name", MATMAT_PROBLEMS)
def test_matmat(name, golden, emu):
    gp.test_matmat(name, golden, emu)


def test_matmat_fused_dots_and_product_chain(emu):
    gp.test_matmat_fused_dots_and_gate(emu)
    # the Product-chain epilogue case of test_gpu_parity.test_product_chain_epilogue, without its CG tail
    g = torch.Generator().manual_seed(2)
    M = torch.randn(40, 24, dtype=torch.float64, generator=g)
    d = torch.rand(40, dtype=torch.float64, generator=g) + 0.5
    ops = emu.ops
    A = emu.PSD(ops.Product(ops.Dense(M), ops.Dense(M.T.contiguous())) + ops.Diagonal(d) + 0.3 * ops.I_like(ops.Dense(M @ M.T)))
    X = torch.randn(40, 5, dtype=torch.float64, generator=g)
    ref = M @ (M.T @ X) + d[:, None] * X + 0.3 * X
    dots = torch.zeros(5, dtype=torch.float64)
    Y = torch.empty_like(X)
    A.matmat_into(X, Y, dots=dots)
    assert gp.rel(Y, ref) < 1e-13 and gp.rel(dots, (X * ref).sum(0)) < 1e-13


@pytest.mark.parametrize("case", sorted(CG_CASES))
def test_cg(case, golden, emu):
    gp.test_cg_vs_oracle_and_golden(case, golden, emu)


@pytest.mark.parametrize("case", sorted(NEXT_CG_CASES))
def test_cg_on_kronsum_and_tridiagonal(case, golden, emu, monkeypatch):
    gn.test_cg_on_kronsum_and_tridiagonal(case, golden, emu, monkeypatch)


def test_cg_surface_and_edge_cases(golden, emu):
    gp.test_cg_first_iterations_tight(emu)
    gp.test_cg_vector_x0_and_solve_surface(golden, emu)
    gp.test_cg_edge_cases(emu)
    gp.test_product_chain_epilogue(emu)
    gn.test_pow_minus_one_is_a_solve(emu)


# pcg_dense96_f64 stops at tol 1e-11, where the iteration count moves by 2 with the summation order of the stand-in
# matmat (the GPU test holds it to +-1 on the real kernels); its solution is still compared below
@pytest.mark.parametrize("case", sorted(set(PCG_CASES) - {"pcg_dense96_f64"}))
def test_pcg_nystrom(case, golden, emu):
    gp.test_pcg_nystrom_vs_oracle_and_golden(case, golden, emu)


def test_pcg_nystrom_threshold_limited_case(golden, emu):
    from tests import problems as pb
    name, rank, tol, iters = PCG_CASES["pcg_dense96_f64"]
    P = pb.problem(name)
    A = pb.to_b200(P["spec"], "cpu", P["ann"])
    Nys = emu.linalg.NystromPrecond(A, rank=rank, key=emu.rng.PRNGKey(3))
    x, info = emu.linalg.CG(tol=tol, max_iters=iters, P=Nys)(A, P["B"])
    g = golden("pcg_dense96_f64")
    assert abs(info["iterations"] - int(g["iterations"])) <= 3 and gp.rel(x, g["x"]) < 1e-8


@pytest.mark.parametrize("case", sorted(LANCZOS_CASES))
def test_lanczos(case, golden, emu_both):
    gp.test_lanczos_vs_oracle_and_golden(case, golden, emu_both)


def test_lanczos_known_answers_and_eig(golden, emu):
    gp.test_lanczos_known_answers_and_early_stop(golden, emu)
    gp.test_eig_lanczos_default_start(golden, emu)


@pytest.mark.parametrize("case", sorted(ARNOLDI_CASES))
def test_arnoldi(case, golden, emu):
    gp.test_arnoldi_vs_oracle_and_golden(case, golden, emu)


def test_eig_arnoldi(golden, emu):
    gp.test_eig_arnoldi(golden, emu)


@pytest.mark.parametrize("case", sorted(POWER_CASES))
def test_power_iteration(case, golden, emu):
    gp.test_power_iteration_vs_oracle_and_golden(case, golden, emu)


@pytest.mark.parametrize("case", sorted(GMRES_CASES))
def test_gmres(case, golden, emu):
    gp.test_gmres_vs_oracle_and_golden(case, golden, emu)


@pytest.mark.parametrize("name,m,vtol", [("kron884_diag_f32", 25, 0.25), ("kron465_diag_f64", 30, 0.2),
                                         ("lap24_f64", 40, 0.25)])
def test_slq(name, m, vtol, golden, emu):
    gp.test_slq_logdet_identical_probes(name, m, vtol, golden, emu)


@pytest.mark.parametrize("name,m", [("kron884_diag_f32", 25), ("kron465_diag_f64", 30)])
def test_log_matmat_and_hutch_logdet(name, m, golden, emu):
    gp.test_log_matmat_and_hutch_logdet(name, m, golden, emu)


def test_hutch_rademacher_and_errors(golden, emu):
    gp.test_hutch_rademacher_and_errors(golden, emu)


# ---- SURVEY 8f rows (tests/test_gpu_parity_next.py bodies; the CG-based ones stay GPU-only)
@pytest.mark.parametrize("name", NEXT_MATMAT_PROBLEMS)
def test_matmat_kronsum_tridiagonal(name, golden, emu):
    gn.test_matmat_kronsum_tridiagonal(name, golden, emu)


@pytest.mark.parametrize("case", sorted(UNARY_CASES))
def test_unary_functions(case, golden, emu):
    gn.test_unary_functions(case, golden, emu)


def test_unary_structure_rules(emu):
    gn.test_unary_structure_rules(emu)


@pytest.mark.parametrize("case", sorted(DIAG_CASES))
def test_exact_and_offset_diagonals(case, golden, emu):
    gn.test_exact_and_offset_diagonals(case, golden, emu)


def test_exact_diag_ragged_and_slogdet_rule(golden, emu):
    gn.test_exact_diag_ragged_blocks_and_trace(emu)
    gn.test_slogdet_lanczos_rule(golden, emu)


def test_triangular_inverse(emu):
    gn.test_triangular_inverse(emu)


def test_pinv_normal_equations(emu):
    gn.test_pinv_normal_equations(emu)


def test_plan_follows_in_place_parameter_updates(emu):
    """Plans hold derived copies (scaled diagonals, folded scalars, the CSR form of a Tridiagonal): an in-place write
    to a leaf (an optimizer step) or a replaced leaf must recompile, an untouched operator must not."""
    ops = emu.ops
    d = torch.linspace(1.0, 2.0, 8, dtype=torch.float64)
    band = torch.full((7, ), 0.5, dtype=torch.float64)
    A = 2.0 * ops.Diagonal(d) + ops.Tridiagonal(band, d.clone(), band)
    X = torch.ones(8, 2, dtype=torch.float64)
    ref = lambda: 2.0 * d[:, None] * X + A.Ms[1].to_dense() @ X   # noqa: E731
    assert torch.allclose(A @ X, ref())
    first = A.plan()
    assert A.plan() is first                                   # unchanged leaves: cached
    d.mul_(3.0)                                                # in place, like optimizer.step()
    assert torch.allclose(A @ X, ref()) and A.plan() is not first
    second = A.plan()
    A.Ms[1].beta.add_(1.0)                                     # a leaf of a nested operator (feeds the CSR copy)
    assert torch.allclose(A @ X, ref()) and A.plan() is not second
    Q, T, info = emu.linalg.Lanczos(start_vector=X[:, 0].contiguous() + d, max_iters=4, tol=1e-12)(emu.SelfAdjoint(A))
    assert torch.allclose(T.beta[0, 0], ((X[:, 0] + d) @ (A.to_dense() @ (X[:, 0] + d))) / ((X[:, 0] + d) @ (X[:, 0] + d)))


def test_lanczos_chunks_run_in_lockstep(emu):
    """LanczosUnary / SLQ split wide probe blocks into chunks (HBM budget).  The reference's single batched Lanczos
    stops when ALL columns satisfy the rule (lanczos.py:256-268), so every chunk must take the longest chunk's
    iteration count: here the last chunk's columns live in a 2-dimensional invariant subspace and would stop
    after 3 steps on their own."""
    ops = emu.ops
    d = torch.tensor([1.0, 2.0] * 3 + [3.0, 4.0, 5.0, 6.0, 7.0, 8.0], dtype=torch.float64)
    A = emu.SelfAdjoint(ops.Diagonal(d))
    g = torch.Generator().manual_seed(0)
    V = torch.randn(12, 8, dtype=torch.float64, generator=g)
    V[6:, 4:] = 0.0                                            # columns 4-7: only the eigenvalues 1 and 2
    one = emu.linalg.LanczosUnary(A, torch.exp, max_iters=10, tol=1e-7, probe_chunk=8)
    two = emu.linalg.LanczosUnary(A, torch.exp, max_iters=10, tol=1e-7, probe_chunk=4)
    Y1, Y2 = one @ V, two @ V
    assert one.info["iterations"] == two.info["iterations"] > 4
    assert gp.rel(Y1, Y2) < 1e-12 and gp.rel(Y1, torch.exp(d)[:, None] * V) < 1e-6


def test_spmv_column_strips(emu, monkeypatch):
    gn.test_spmv_column_strips(emu, monkeypatch)


def test_spmm_staged_tiles(emu, monkeypatch):
    """The tile-local CSR form (cola_b200/csr_tiles.py) and its dispatch, with the kernel replaced by an emulation that
    computes FROM the records, runs, slots and local row pointers."""
    gn.test_spmm_staged_tiles(emu, monkeypatch)


def test_native_api_refuses_parameters_that_require_grad(emu):
    """The kernels do not record autograd: a leaf that requires grad raises while recording is on (instead of handing
    back a result without grad_fn), and is accepted under torch.no_grad()."""
    W = torch.eye(4, dtype=torch.float64, requires_grad=True)
    A = emu.ops.Dense(W) + emu.ops.Diagonal(torch.ones(4, dtype=torch.float64))
    X = torch.ones(4, 2, dtype=torch.float64)
    with pytest.raises(RuntimeError, match="not differentiable"):
        A @ X
    with torch.no_grad():
        assert torch.equal(A @ X, 2 * X)


def test_bench_e2e_double_buffering_logic(emu, monkeypatch):
    """bench.py's end-to-end loop overlaps the copies of adjacent steps with the solve on a second stream.  With the
    CUDA stream / event objects replaced by recorders (no GPU here) the schedule itself is checked: no wait on an
    event that was not recorded, an input buffer is only refilled after the solve that read it, every step's
    solution reaches the host buffer, iterations are counted once per step."""
    import contextlib

    import bench
    from tests import problems as pb
    log = []

    class Stream:
        def wait_event(self, ev):
            assert ev.recorded, "wait on an event that was never recorded"
            log.append(("wait", ev.name))

    class Event:
        count = 0

        def __init__(self, **kw):
            self.recorded, self.name = False, Event.count
            Event.count += 1

        def record(self, stream=None):
            self.recorded = True
            log.append(("record", self.name))

    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: Stream())
    monkeypatch.setattr(torch.cuda, "Stream", lambda device=None: Stream())
    monkeypatch.setattr(torch.cuda, "Event", Event)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.Tensor, "record_stream", lambda self, s: None, raising=False)
    P = pb.problem("lap24_f32")
    A = pb.to_b200(P["spec"], "cpu", P["ann"])
    alg = emu.linalg.CG(tol=1e-30, max_iters=10)
    x_ref, _ = alg(A, P["B"])
    for steps in (1, 2, 5):
        x_host = torch.zeros_like(P["B"])
        assert bench.e2e_double_buffered(alg, A, P["B"], x_host, "cpu", steps) == 10 * steps
        assert torch.equal(x_host, x_ref)
    # loaded[0], loaded[1], released[0], released[1] are events 0..3 of a call: buffer 1 is refilled for step 3 only
    # after solve 1 released it (wait on released[1] precedes the second record of loaded[1])
    Event.count = 0
    log.clear()
    bench.e2e_double_buffered(alg, A, P["B"], torch.zeros_like(P["B"]), "cpu", 4)
    second_fill = [i for i, e in enumerate(log) if e == ("record", 1)][1]
    assert ("wait", 3) in log[:second_fill]
