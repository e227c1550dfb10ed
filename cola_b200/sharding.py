"""Sharding of independent right-hand sides / probe vectors over the GPUs of one box.

One process per GPU (`torch.distributed`); the operator is replicated, columns are split contiguously by rank and
there is no data-path collective inside the Krylov loops: SLQ ends with ONE all-reduce of (sum, count), Hutchinson
all-reduces its two running sums once per 100-probe block (its stopping rule is global), CG agrees on the global
stopping iteration after the local loops end (a 16-byte all-reduce) and all-reduces its error trace once.  The reference has no multi-device code at all; this module is new."""
import torch


def column_range(total, rank, world):
    """Contiguous column slice [lo, hi) of rank `rank` out of `world` (all columns covered, sizes differ by <= 1)."""
    return (rank * total) // world, ((rank + 1) * total) // world


def shard_columns(X, group=None):
    """This rank's column block of an (n, k) right-hand-side matrix."""
    import torch.distributed as dist
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_initialized() else (0, 1)
    lo, hi = column_range(X.shape[1], rank, world)
    return X[:, lo:hi].contiguous(), (lo, hi)


def solve_sharded(A, B, alg, group=None, gather=True):
    """CG solve with the RHS columns sharded over ranks.  There is no collective inside the iterations.  The
    reference's stopping rule `any(||r|| > tol_eff)` runs over ALL columns (cg.py:133-138), so after its own columns
    have converged a rank agrees with the others on the common iteration count (one 16-byte all-reduce per round,
    cola_b200/linalg/cg.py:_global_stop_rule) and `info` -- iterations and the mean-residual trace -- is the unsharded
    solve's; with a fixed iteration budget that exchange is a single all-reduce confirming the same count.  Returns
    the full solution on every rank if `gather`."""
    import importlib

    import torch.distributed as dist
    cg = importlib.import_module(__package__ + ".linalg.cg")
    Bl, (lo, hi) = shard_columns(B, group)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    saved = cg.STOP_RULE_GROUP
    cg.STOP_RULE_GROUP = (group if group is not None else dist.group.WORLD) if world > 1 else None
    try:
        xl, info = alg(A, Bl)
    finally:
        cg.STOP_RULE_GROUP = saved
    if not gather or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return xl, info
    world = dist.get_world_size(group)
    n, k = B.shape
    out = torch.zeros((n, k), dtype=xl.dtype, device=xl.device)
    out[:, lo:hi] = xl
    dist.all_reduce(out, group=group)   # disjoint column blocks: sum == concatenation
    return out, info
