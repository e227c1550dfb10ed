"""Batched (multi-RHS) conjugate gradients on the fused sm_100a kernels.

Same interface and semantics as the reference (cola/linalg/inverse/cg.py): `CG(tol, max_iters, pbar, x0, P)`,
`cg(A, rhs, ...) -> (soln, info)`, `run_batched_cg(A, b, x0, max_iters, tol, preconditioner, pbar)
-> (x, r, k, info)` with `info = {'iterations', 'errors', 'iteration_time'}` as built by
cola/utils/torch_tqdm.py:7-71.

What is different underneath (see DESIGN.md "CG"):
  * one iteration = 3 kernels + a 1-block scalar kernel instead of ~87 eager ATen calls
      matmat(+shift/diag epilogue) with the p^T A p column dots fused      (cg.py:145, 157-158)
      r -= alpha Ap ; gamma' = <r,r>                                        (cg.py:148-150, 165-166)
      x += alpha p ; p = r + beta p   (p read once for both)                (cg.py:147, 151-153)
  * alpha, beta, the has_converged mask, the iteration counter and the stopping rule live in a
    device-resident control block; the host only polls it every `check_every` iterations, so the two
    device->host syncs per iteration of torch_tqdm.py:42,88 are gone;
  * the per-iteration residual norms are kept on the device (gamma trace) and `info['errors']` is rebuilt
    from them once at the end.
"""
import gc
import time
from dataclasses import dataclass
from typing import Any

import numpy as np
import torch

from .. import backend as be
from ..ops import I_like, Identity, LinearOperator
from .algorithm_base import Algorithm

_small_value = 1e-40
CHECK_EVERY = 16  # iterations enqueued between polls of the device-side stop flag
USE_CUDA_GRAPH = True  # replay batches of iterations as a CUDA graph once the loop is warm
GRAPH_MAX_ELEMS = 64 * 1024 * 1024
# Set by cola_b200.sharding.solve_sharded for the duration of a RHS-sharded solve: the process group over which the
# stopping rule `any(||r|| > tol_eff)` (cg.py:133-138) and the error trace are global, as in the unsharded solve.
STOP_RULE_GROUP = None


def _world(group):
    if group is None:
        return 1
    import torch.distributed as dist
    return dist.get_world_size(group) if dist.is_initialized() else 1


def _global_stop_rule(group, it, drive, reevaluate, ctl, tol_eff, max_iters):
    """Reproduces the reference's stopping rule over ALL right-hand sides when they are sharded over ranks.

    The reference stops at the first iteration K at which every column satisfies ||r|| <= tol_eff (or at
    max_iters) and all columns are iterated up to K.  Each rank's device-side rule stops at the first iteration
    K_r at which ITS columns are satisfied, so K >= max_r K_r.  Protocol (one 16-byte all-reduce per round, usually
    one or two rounds, nothing inside the iterations): agree on K* = max_r K_r; ranks that stopped earlier advance
    to K* with the rule switched off (tol_eff = -1, iteration cap K*), then re-evaluate the real rule at K*; ranks
    whose residuals rose above tolerance again continue to their next stop; repeat until every rank stands at the
    same iteration."""
    import torch.distributed as dist
    dev = ctl.device
    while True:
        ks = torch.tensor([it, -it], dtype=torch.int64, device=dev)
        dist.all_reduce(ks, op=dist.ReduceOp.MAX, group=group)
        k_max, k_min = int(ks[0]), -int(ks[1])
        if k_max == k_min:
            return it
        if it < k_max:
            saved = tol_eff.clone()
            tol_eff.fill_(-1.0)                                 # rs > -1 for every finite residual: rule off
            ctl[2:3].fill_(k_max)
            ctl[1:2].zero_()
            advanced = drive(k_max)
            tol_eff.copy_(saved)
            ctl[2:3].fill_(max_iters)
            ctl[1:2].zero_()
            reevaluate()                                        # the real cond_fun at the current iteration
            if advanced == it:                                  # no progress (all residuals NaN): give up agreeing
                return drive(max_iters)
            it = advanced
        it = drive(max_iters)                                   # no-op when the rule is satisfied at K*


def _global_error_trace(group, col_norms):
    """info['errors'] tracks mean_columns ||r|| (cg.py:103-105): over all ranks' columns it is the all-reduced column
    sum divided by the all-reduced column count.  col_norms: (iterations + 1, k_local)."""
    import torch.distributed as dist
    packed = torch.cat([col_norms.sum(dim=1), col_norms.new_tensor([col_norms.shape[1]])])
    dist.all_reduce(packed, group=group)
    return packed[:-1] / packed[-1]


def _empty_block_cg(b, max_iters):
    """A rank whose column block is empty (fewer right-hand sides than ranks, sharding.solve_sharded): no kernel runs,
    but the rank still takes part in the collectives of `_global_stop_rule` / `_global_error_trace` so that the
    other ranks do not wait for it; its `info` is the global one.  It has no residual of its own, so in every
    round of the agreement it simply stands at the largest iteration count any rank reports."""
    t0 = time.time()
    it = 0
    group = STOP_RULE_GROUP if _world(STOP_RULE_GROUP) > 1 else None
    col_norms = torch.zeros((1, 0), dtype=torch.float64, device=b.device)
    if group is not None:
        import torch.distributed as dist
        while True:
            ks = torch.tensor([it, -it], dtype=torch.int64, device=b.device)
            dist.all_reduce(ks, op=dist.ReduceOp.MAX, group=group)
            k_max, k_min = int(ks[0]), -int(ks[1])
            if k_max == k_min:
                break
            it = k_max
        col_norms = torch.zeros((it + 1, 0), dtype=torch.float64, device=b.device)
        trace = _global_error_trace(group, col_norms).cpu().numpy()
    else:
        trace = np.full(1, np.nan)
    samples = np.concatenate([trace, trace[-1:]])
    info = {"iterations": it + 1, "errors": samples[2:].astype(np.float64), "iteration_time": (time.time() - t0) / (it + 1)}
    return torch.empty_like(b), torch.empty_like(b), it, info


@dataclass
class CG(Algorithm):
    """cola/linalg/inverse/cg.py:13-36"""
    tol: float = 1e-6
    max_iters: int = 1000
    pbar: bool = False
    x0: Any = None
    P: Any = None

    def __call__(self, A, b):
        return cg(A, b, **self.__dict__)


def cg(A: LinearOperator, rhs, x0=None, P=None, tol=1e-6, max_iters=5000, pbar=False):
    """cola/linalg/inverse/cg.py:39-69"""
    is_vector = len(rhs.shape) == 1
    if x0 is None:
        x0 = None  # zero initial guess: r0 = b exactly, the A @ x0 matmat is skipped
    if is_vector:
        rhs = rhs[..., None]
        x0 = x0[..., None] if x0 is not None else None
    if P is None:
        P = I_like(A)
    from .. import autograd as ag
    if ag.needs_grad(A, rhs):
        # run_cg carries a custom backward (cg.py:72-91): cg_bwd re-applies the forward call's settings to the
        # output cotangent and takes the parameter vjp at the solution
        def run(A_, b_):
            return run_batched_cg(A_, b_, x0, max_iters, tol, P, pbar=pbar)
        soln, rest = ag.cg_with_grad(A, rhs, run)
        infodict = rest[-1]
    else:
        soln, *_, infodict = run_batched_cg(A, rhs, x0, max_iters, tol, P, pbar=pbar)
    soln = soln.reshape(-1) if is_vector else soln
    return soln, infodict


def run_batched_cg(A, b, x0, max_iters, tol, preconditioner, pbar=False):
    """cola/linalg/inverse/cg.py:94-119 on the device.  b (n,k) -> (x (n,k), r (n,k), k_iters, info)."""
    if not isinstance(preconditioner, Identity):
        return _run_batched_pcg(A, b, x0, max_iters, tol, preconditioner, pbar)
    be.require_cuda(b, "right-hand side")
    dt = A.dtype
    b = b.to(dt).contiguous()
    n, k = b.shape
    if k == 0:
        return _empty_block_cg(b, int(max_iters))
    dev = b.device
    max_iters = int(max_iters)
    lib = be.lib()
    sx = be.sfx(dt)
    st = be.stream_ptr

    # graphs pay off where the loop is launch-bound (vector blocks up to ~256 MB); larger problems spend
    # milliseconds per kernel and the capture would cost more than it saves
    graph_ok = USE_CUDA_GRAPH and n * k <= GRAPH_MAX_ELEMS and A.plan().graph_safe()
    # A captured batch of iterations is tied to the buffers it was captured on.  For graph-eligible sizes the
    # state lives in a workspace kept on the operator (one per operator, latest shape), so that a second solve with
    # the same shape replays the graph of the first instead of capturing (and later destroying) its own: capture,
    # instantiation and teardown cost more than a whole 100-iteration solve at these sizes.
    ws = None
    if graph_ok:
        key = (n, k, dt, max_iters, str(dev))
        cached = A.__dict__.get("_cg_workspace")
        if cached is not None and cached.get("busy"):
            graph_ok = False                                     # in use by a concurrent solve: fresh state, eager loop
        elif cached is not None and cached["key"] == key:
            ws = cached
        else:
            ws = {"key": key, "graph": None,
                  "x": torch.empty_like(b), "r": torch.empty_like(b), "p": torch.empty_like(b), "ap": torch.empty_like(b),
                  "gamma": torch.empty((max_iters + 2, k), dtype=torch.float64, device=dev),
                  "pap": torch.empty((max_iters + 1, k), dtype=torch.float64, device=dev),
                  "tol_eff": torch.empty(k, dtype=dt, device=dev),
                  "ctl": torch.empty(4, dtype=torch.int32, device=dev)}
            A.__dict__["_cg_workspace"] = ws

    def _solve():
        # ---- setup: normalise RHS, residual, gamma0, tolerances (cg.py:96-101, 122-130)
        mult_sq = torch.zeros(k, dtype=torch.float64, device=dev)
        be.col_dots(b, b, mult_sq)
        r = ws["r"] if ws else torch.empty_like(b)
        be.col_scale(b, r, mult_sq, take_sqrt=True, mode=1)          # b / ||b|| (safe)
        x = ws["x"] if ws else torch.empty_like(b)
        if x0 is None:
            x.zero_()
        else:
            x.copy_(x0.to(dt))
            ax = torch.empty_like(b)
            A.matmat_into(x, ax)
            be.axpby(ax, r, -1.0, 1.0)                               # r0 = b - A x0
            del ax
        if ws:
            p, ap, gamma, pap, tol_eff, ctl = ws["p"], ws["ap"], ws["gamma"], ws["pap"], ws["tol_eff"], ws["ctl"]
            p.copy_(r)
            gamma.zero_()
            pap.zero_()
            ctl.copy_(be.small_ints([0, 0, max_iters, k], dev))
        else:
            p = r.clone()
            ap = torch.empty_like(b)
            gamma = torch.zeros((max_iters + 2, k), dtype=torch.float64, device=dev)
            pap = torch.zeros((max_iters + 1, k), dtype=torch.float64, device=dev)
            tol_eff = torch.empty(k, dtype=dt, device=dev)
            ctl = be.small_ints([0, 0, max_iters, k], dev)
        be.col_dots(r, r, gamma[0])
        lib.call(f"cola_cg_tol_{sx}", be.ptr(gamma), be.scalar(dt, tol), be.ptr(tol_eff), k, st())
        it_ptr, done_ptr = ctl[0:1], ctl[1:2]
        lib.call(f"cola_cg_advance_{sx}", be.ptr(ctl), be.ptr(gamma), be.ptr(tol_eff), 0, st())   # initial cond_fun

        def enqueue(n_iters):
            for _ in range(n_iters):
                A.matmat_into(p, ap, dots=pap, dots_row=it_ptr, gate=done_ptr)
                lib.call(f"cola_cg_update_r_{sx}", be.ptr(r), be.ptr(ap), n, k, k, be.ptr(ctl), be.ptr(gamma),
                         be.ptr(pap), be.ptr(gamma), st())
                lib.call(f"cola_cg_update_xp_{sx}", be.ptr(x), be.ptr(r), be.ptr(p), n, k, k, be.ptr(ctl), be.ptr(gamma),
                         be.ptr(pap), st())
                lib.call(f"cola_cg_advance_{sx}", be.ptr(ctl), be.ptr(gamma), be.ptr(tol_eff), 1, st())

        t0 = time.time()
        if ws and ws["graph"] is not None and ws.get("token") != A.plan().graph_token():
            ws["graph"] = None                                       # the operator's scratch buffers moved: recapture
        loop = {"graph": ws["graph"] if ws else None, "batches": 0}

        def drive(limit):
            """Enqueue gated batches until the device-side rule (or the iteration cap `limit`) stops the loop.  The control
            block is snapshotted after every batch (stream-ordered, DMA-free) and read ONE BATCH LATER: the next batch is
            already enqueued when the host waits for a snapshot, so the device does not idle through the host round trip
            (event wake-up, Python, launch).  Iterations enqueued past the stopping point are device-side no-ops."""
            c = be.read_small(ctl)
            it, done = int(c[0]), int(c[1])
            pending = []                                         # snapshots of batches enqueued but not looked at yet
            while True:
                if done:
                    return it
                if len(pending) >= 2 or (pending and it + len(pending) * CHECK_EVERY >= limit):
                    c = be.publish_result(pending.pop(0))
                    it, done = int(c[0]), int(c[1])
                    continue
                remaining = limit - it - len(pending) * CHECK_EVERY
                want_capture = graph_ok and loop["graph"] is None and loop["batches"] >= 1 and remaining >= 2 * CHECK_EVERY
                if want_capture and pending:                     # capturing synchronises anyway: look at what is in flight first
                    c = be.publish_result(pending.pop(0))
                    it, done = int(c[0]), int(c[1])
                    continue
                if want_capture:
                    # Every kernel reads the iteration index and the stop flag from the device control block, so a batch
                    # of iterations is the same launch sequence each time: capture it once, replay it (launch cost -> ~0).
                    # Iterations past the stopping point are device-side no-ops, exactly as in the eager batches.
                    graph = torch.cuda.CUDAGraph()
                    # No garbage collection while the stream is capturing: a cyclic collection can finalise another
                    # operator's cached CUDAGraph (cudaGraphExecDestroy / cudaFree), which invalidates the capture in
                    # progress (seen as cudaErrorStreamCaptureInvalidated, depending on test order).  torch.cuda.graph()
                    # itself runs gc.collect() just before capture begins.
                    gc_on = gc.isenabled()
                    gc.disable()
                    try:
                        with torch.cuda.graph(graph):
                            enqueue(CHECK_EVERY)
                    finally:
                        if gc_on:
                            gc.enable()
                    loop["graph"] = graph
                    ws["graph"], ws["token"] = graph, A.plan().graph_token()
                    # capture does not execute: fall through to the replay below
                if loop["graph"] is not None:
                    loop["graph"].replay()
                else:
                    enqueue(max(1, min(CHECK_EVERY, remaining)))
                loop["batches"] += 1
                pending.append(be.publish_async(ctl))

        it = drive(max_iters)
        group = STOP_RULE_GROUP if _world(STOP_RULE_GROUP) > 1 else None
        if group is not None:
            it = _global_stop_rule(group, it, drive, lambda: lib.call(f"cola_cg_advance_{sx}", be.ptr(ctl), be.ptr(gamma),
                                                                        be.ptr(tol_eff), 0, st()), ctl, tol_eff, max_iters)
        elapsed = time.time() - t0

        # ---- info dict exactly as while_loop_winfo builds it (torch_tqdm.py:35-62): the tracked error is sampled
        # before every cond evaluation (it+1 of them) and once more after the loop; the first two are dropped.
        col_norms = torch.sqrt(gamma[:it + 1])
        trace = be.read_small(col_norms.mean(dim=1) if group is None else _global_error_trace(group, col_norms)).numpy()
        samples = np.concatenate([trace, trace[-1:]])
        info = {"iterations": it + 1, "errors": samples[2:].astype(np.float64), "iteration_time": elapsed / (it + 1)}
        if ws:                                                       # hand back copies: the workspace is reused
            x_out, r_out = torch.empty_like(x), torch.empty_like(r)
        else:
            x_out, r_out = x, r
        be.col_scale(x, x_out, mult_sq, take_sqrt=True, mode=0)      # x * ||b||  (cg.py:119)
        be.col_scale(r, r_out, mult_sq, take_sqrt=True, mode=0)
        return x_out, r_out, it, info

    # A workspace in use (a second solve on the same operator from another thread / stream) is never shared: that solve
    # takes fresh buffers and the eager loop.  `release_cg_workspace(A)` frees the cached one.
    if ws is not None:
        ws["busy"] = True
    try:
        return _solve()
    finally:
        if ws is not None:
            ws["busy"] = False


def release_cg_workspace(A):
    """Drops the solve state and the captured CUDA graph a graph-eligible solve leaves on the operator (four (n, k)
    blocks, two fp64 traces, the graph): call it when no further solve of that shape is coming."""
    A.__dict__.pop("_cg_workspace", None)


def _run_batched_pcg(A, b, x0, max_iters, tol, P, pbar=False):
    """Preconditioned branch of cola/linalg/inverse/cg.py:94-170: z = P r, gamma = <r, z>,
    p = z + beta p, while the stopping rule, the `has_converged` mask's trace and `info['errors']` keep using ||r||.
    Built from the same kernels as the plain loop:
        matmat(p -> Ap, <p, Ap> fused)  ->  r-sweep (r -= alpha Ap, writes ||r||^2 into its own trace)
        ->  P applied to r with <r, z> fused into gamma  ->  xp-sweep with z in the role of r  ->  advance on ||r||^2.
    P is any operator of this package (its plan is applied like A's); device-side gating and the 16-iteration
    batches between host polls are unchanged.  One deviation: the per-column `has_converged` mask inside the sweeps
    tests gamma = <r, P r> instead of ||r||^2 against 1e-80; both vanish together for a positive definite P, and
    the mask only matters for residuals that are exactly zero."""
    be.require_cuda(b, "right-hand side")
    assert tuple(P.shape) == tuple(A.shape), "preconditioner shape mismatch"
    A.plan(), P.plan()                                         # validate / compile once; the loop uses matmat_into
    dt = A.dtype
    b = b.to(dt).contiguous()
    n, k = b.shape
    if k == 0:
        return _empty_block_cg(b, int(max_iters))
    dev = b.device
    max_iters = int(max_iters)
    lib = be.lib()
    sx = be.sfx(dt)
    st = be.stream_ptr

    mult_sq = torch.zeros(k, dtype=torch.float64, device=dev)
    be.col_dots(b, b, mult_sq)
    r = torch.empty_like(b)
    be.col_scale(b, r, mult_sq, take_sqrt=True, mode=1)          # b / ||b|| (safe)
    if x0 is None:
        x = torch.zeros_like(b)
    else:
        x = x0.to(dt).contiguous().clone()
        ax = torch.empty_like(b)
        A.matmat_into(x, ax)
        be.axpby(ax, r, -1.0, 1.0)                               # r0 = b - A x0
        del ax
    gamma = torch.zeros((max_iters + 2, k), dtype=torch.float64, device=dev)    # <r, z>
    rnorm2 = torch.zeros((max_iters + 2, k), dtype=torch.float64, device=dev)   # <r, r>
    pap = torch.zeros((max_iters + 1, k), dtype=torch.float64, device=dev)
    z = torch.empty_like(b)
    P.matmat_into(r, z, dots=gamma[0])                           # z0 = P r0, gamma0 = <r0, z0>  (cg.py:124-127)
    p = z.clone()
    ap = torch.empty_like(b)
    be.col_dots(r, r, rnorm2[0])
    tol_eff = torch.empty(k, dtype=dt, device=dev)
    lib.call(f"cola_cg_tol_{sx}", be.ptr(rnorm2), be.scalar(dt, tol), be.ptr(tol_eff), k, st())
    ctl = be.small_ints([0, 0, max_iters, k], dev)
    it_ptr, done_ptr = ctl[0:1], ctl[1:2]
    lib.call(f"cola_cg_advance_{sx}", be.ptr(ctl), be.ptr(rnorm2), be.ptr(tol_eff), 0, st())
    gamma_next = gamma[1:]                                       # row `it` of this view is gamma[it + 1]

    def enqueue(n_iters):
        for _ in range(n_iters):
            A.matmat_into(p, ap, dots=pap, dots_row=it_ptr, gate=done_ptr)
            lib.call(f"cola_cg_update_r_{sx}", be.ptr(r), be.ptr(ap), n, k, k, be.ptr(ctl), be.ptr(gamma),
                     be.ptr(pap), be.ptr(rnorm2), st())          # ||r_new||^2 -> rnorm2[it + 1]
            P.matmat_into(r, z, dots=gamma_next, dots_row=it_ptr, gate=done_ptr)   # <r_new, z_new> -> gamma[it + 1]
            lib.call(f"cola_cg_update_xp_{sx}", be.ptr(x), be.ptr(z), be.ptr(p), n, k, k, be.ptr(ctl), be.ptr(gamma),
                     be.ptr(pap), st())
            lib.call(f"cola_cg_advance_{sx}", be.ptr(ctl), be.ptr(rnorm2), be.ptr(tol_eff), 1, st())

    t0 = time.time()

    def drive(limit):
        while True:
            c = be.read_small(ctl)
            it, done = int(c[0]), int(c[1])
            if done:
                return it
            enqueue(min(CHECK_EVERY, limit - it))

    it = drive(max_iters)
    group = STOP_RULE_GROUP if _world(STOP_RULE_GROUP) > 1 else None
    if group is not None:
        it = _global_stop_rule(group, it, drive, lambda: lib.call(f"cola_cg_advance_{sx}", be.ptr(ctl), be.ptr(rnorm2),
                                                                    be.ptr(tol_eff), 0, st()), ctl, tol_eff, max_iters)
    elapsed = time.time() - t0
    col_norms = torch.sqrt(rnorm2[:it + 1])
    trace = be.read_small(col_norms.mean(dim=1) if group is None else _global_error_trace(group, col_norms)).numpy()
    samples = np.concatenate([trace, trace[-1:]])
    info = {"iterations": it + 1, "errors": samples[2:].astype(np.float64), "iteration_time": elapsed / (it + 1)}
    be.col_scale(x, x, mult_sq, take_sqrt=True, mode=0)          # x * ||b||  (cg.py:119)
    be.col_scale(r, r, mult_sq, take_sqrt=True, mode=0)
    return x, r, it, info
