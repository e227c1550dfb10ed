"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: column ranges, the single SLQ all-reduce and the
per-block Hutchinson all-reduce.  The per-probe arithmetic is stubbed with the CPU oracle (the CUDA kernels cannot
run here); what is under test is that the sharded estimate equals the unsharded one on identical probes."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, results):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cola_b200 as cb
    from cola_b200.linalg import stochastic
    from oracle import krylov_oracle as ko
    from tests import problems as pb
    cb.rng.PROBE_DEVICE = "cpu"
    P = pb.problem("kron465_diag_f64")
    Ao = pb.to_oracle(P["spec"])

    class Stub:  # just enough of a LinearOperator for slq_fwd / Hutch host logic
        shape, dtype, device = Ao.shape, Ao.dtype, torch.device("cpu")

        def __matmul__(self, Z):
            return ko.lanczos_unary_matmat(Ao, torch.log, Z, 20, 1e-12)[0]

    # the per-chunk factorisation + quadrature (lanczos_fact on the kernels) stated with the oracle
    stochastic.lanczos_chunks_lockstep = (lambda A, blocks, m, tol, pbar, process, group=None:
                                          [ko.slq_per_probe(Ao, torch.log, get(), m, tol) for get in blocks])
    from tests import host_harness
    cb.backend.row_dots = host_harness.row_dots      # the estimator's two running sums (cola_row_dots_*), stated on CPU
    key = cb.rng.PRNGKey(42)
    num = max(int(1 / 0.2**2), 1)   # 24: same rounding as stochastic_lanczos_quad (slq.py:74)
    val = stochastic.slq_fwd(Stub(), torch.log, num_samples=num, max_iters=20, tol=1e-12, pbar=False, key=key,
                             probe_chunk_size=4, group=dist.group.WORLD)
    ref = ko.slq(Ao, torch.log, max_iters=20, tol=1e-12, vtol=0.2, key=key)
    dg, info = stochastic.hutchinson_diag_estimate(Stub(), 0, tol=2e-2, max_iters=2, key=key, group=dist.group.WORLD)
    dref, iref = ko.hutchinson_diag(lambda Z: ko.lanczos_unary_matmat(Ao, torch.log, Z, 20, 1e-12)[0], Ao.shape[0],
                                    Ao.dtype, tol=2e-2, max_iters=2, key=key)
    lo, hi = cb.sharding.column_range(num, rank, world)
    results[rank] = (float(val), float(ref), float((dg - dref).abs().max()), info["iterations"], iref["iterations"],
                     lo, hi)
    dist.destroy_process_group()


def test_sharded_slq_and_hutch_match_unsharded():
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
    assert len(results) == world
    spans = sorted((results[r][5], results[r][6]) for r in range(world))
    assert spans[0][0] == 0 and spans[-1][1] == 24 and spans[0][1] == spans[1][0]
    for r in range(world):
        val, ref, derr, it, itref, _, _ = results[r]
        assert abs(val - ref) <= 1e-10 * abs(ref)      # one all-reduce of (sum, count) == the mean over all probes
        assert derr < 1e-10 and it == itref            # per-block all-reduce reproduces the global stopping rule
    assert results[0][0] == results[1][0]


def _worker_host_loops(rank, world, port, results):
    """Same checks with the package's own operators and host loops (kernels replaced by tests/host_harness.py), plus
    the RHS-sharded CG solve of cola_b200.sharding.solve_sharded."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cola_b200 as cb
    from cola_b200.linalg import stochastic
    from tests import problems as pb
    from tests.host_harness import emulated_kernels
    P = pb.problem("kron465_diag_f64")
    with emulated_kernels():
        A = pb.to_b200(P["spec"], "cpu", P["ann"])
        key = cb.rng.PRNGKey(42)
        kw = dict(num_samples=25, max_iters=20, tol=1e-12, pbar=False, key=key, probe_chunk_size=4)
        val = stochastic.slq_fwd(A, torch.log, group=dist.group.WORLD, **kw)
        ref = stochastic.slq_fwd(A, torch.log, **kw)
        F = cb.linalg.LanczosUnary(A, torch.log, max_iters=20, tol=1e-12)
        dg, info = stochastic.hutchinson_diag_estimate(F, 0, tol=2e-2, max_iters=2, key=key, group=dist.group.WORLD)
        dref, iref = stochastic.hutchinson_diag_estimate(F, 0, tol=2e-2, max_iters=2, key=key)
        alg = cb.linalg.CG(tol=1e-30, max_iters=15)             # fixed budget: iterates identical to the unsharded solve
        xs, _ = cb.sharding.solve_sharded(A, P["B"], alg, group=dist.group.WORLD)
        xl, _ = cb.sharding.solve_sharded(A, P["B"], alg, group=dist.group.WORLD, gather=False)
        xref, _ = alg(A, P["B"])
        # tolerance-limited solve whose ranks stop at very different iterations: rank 0 gets two eigenvectors
        # (converged after one step), rank 1 three generic columns.  The global rule must give the unsharded count,
        # trace and solution (rank 0 is advanced with its rule switched off, cg.py:_global_stop_rule)
        Pd = pb.problem("dense96_f64")
        Ad = pb.to_b200(Pd["spec"], "cpu", Pd["ann"])
        Bd = Pd["B"].clone()
        Bd[:, :2] = torch.linalg.eigh(Pd["spec"][1])[1][:, [3, 40]]
        tol_alg = cb.linalg.CG(tol=1e-9, max_iters=500)
        xt, info_t = cb.sharding.solve_sharded(Ad, Bd, tol_alg, group=dist.group.WORLD)
        xt_ref, info_ref = tol_alg(Ad, Bd)
        xloc, info_loc = tol_alg(Ad, cb.sharding.shard_columns(Bd, dist.group.WORLD)[0])   # what a local rule would do
        # exact diagonals: 100-column identity blocks dealt round-robin, one all-reduce (n = 576: 6 blocks, ragged)
        Pl = pb.problem("lap24_f64")
        Al = pb.to_b200(Pl["spec"], "cpu", Pl["ann"])
        dense = pb.to_oracle(Pl["spec"]).matmat(torch.eye(576, dtype=torch.float64))
        diag_err = max(float((cb.linalg.exact_diag(Al, kk, 100, group=dist.group.WORLD)
                              - torch.diagonal(dense, offset=kk)).abs().max()) for kk in (0, 1, -24))
    lo, hi = cb.sharding.column_range(P["B"].shape[1], rank, world)
    results[rank] = (float(val), float(ref), float((dg - dref).abs().max() / dref.abs().max()), info["iterations"],
                     iref["iterations"], float((xs - xref).abs().max() / xref.abs().max()),
                     float((xl - xref[:, lo:hi]).abs().max()), tuple(xl.shape),
                     dict(it=info_t["iterations"], it_ref=info_ref["iterations"], it_local=info_loc["iterations"],
                          xerr=float((xt - xt_ref).abs().max() / xt_ref.abs().max()),
                          trace_err=float(abs(info_t["errors"] - info_ref["errors"]).max() / info_ref["errors"].max()),
                          n_err=(len(info_t["errors"]), len(info_ref["errors"])), diag_err=diag_err))
    dist.destroy_process_group()


def test_sharded_host_loops_match_unsharded():
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    port = 31500 + os.getpid() % 2000
    mp.spawn(_worker_host_loops, args=(world, port, results), nprocs=world, join=True)
    assert len(results) == world
    for r in range(world):
        val, ref, derr, it, itref, xerr, xlerr, shape, tl = results[r]
        assert abs(val - ref) <= 1e-12 * abs(ref)      # 25 probes, 13 + 12 by rank, chunks of 4: one all-reduce
        assert derr < 1e-12 and it == itref
        assert xerr < 1e-12 and xlerr < 1e-12 and shape[1] in (3, 4)
        assert tl["it"] == tl["it_ref"] and tl["n_err"][0] == tl["n_err"][1], tl
        assert tl["xerr"] < 1e-10 and tl["trace_err"] < 1e-10, tl
        assert tl["diag_err"] < 1e-13, tl
    assert results[0][0] == results[1][0]
    # the scenario really exercises the protocol: left alone, rank 0 would have stopped long before rank 1
    assert results[0][8]["it_local"] < results[1][8]["it_local"] == results[1][8]["it_ref"]


def _worker_fewer_rhs_than_ranks(rank, world, port, results):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cola_b200 as cb
    from tests import problems as pb
    from tests.host_harness import emulated_kernels
    Pd = pb.problem("dense96_f64")
    with emulated_kernels():
        Ad = pb.to_b200(Pd["spec"], "cpu", Pd["ann"])
        B1 = Pd["B"][:, :1].contiguous()                        # ONE right-hand side, two ranks: rank 0's block is (n, 0)
        alg = cb.linalg.CG(tol=1e-9, max_iters=500)
        xs, info = cb.sharding.solve_sharded(Ad, B1, alg, group=dist.group.WORLD)
        xref, info_ref = alg(Ad, B1)
        lo, hi = cb.sharding.column_range(1, rank, world)
    results[rank] = (hi - lo, info["iterations"], info_ref["iterations"],
                     float((xs - xref).abs().max() / xref.abs().max()),
                     float(abs(info["errors"] - info_ref["errors"]).max() / info_ref["errors"].max()))
    dist.destroy_process_group()


def test_fewer_rhs_than_ranks():
    """a rank with an empty column block must neither raise nor leave the others waiting in the
    stop-rule all-reduce; every rank returns the unsharded iteration count, trace and solution."""
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    port = 33500 + os.getpid() % 2000
    mp.spawn(_worker_fewer_rhs_than_ranks, args=(world, port, results), nprocs=world, join=True)
    assert sorted(results[r][0] for r in range(world)) == [0, 1]
    for r in range(world):
        _, it, it_ref, xerr, terr = results[r]
        assert it == it_ref and xerr < 1e-10 and terr < 1e-10, results[r]


def test_column_range_covers_everything():
    from cola_b200.sharding import column_range
    for total in (1, 7, 64, 100, 1024):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                lo, hi = column_range(total, r, world)
                got.extend(range(lo, hi))
            assert got == list(range(total))
