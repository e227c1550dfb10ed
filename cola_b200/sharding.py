"""Sharding of independent right-hand sides / probe vectors over the GPUs of one box.

One process per GPU (`torch.distributed`); the operator is replicated, columns are split contiguously by rank and
there is no data-path collective inside the Krylov loops: SLQ ends with ONE all-reduce of (sum, count), Hutchinson
all-reduces its two running sums once per 100-probe block (its stopping rule is global), fixed-length CG needs
none.  The reference has no multi-device code at all; this module is new."""
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
    """CG solve with the RHS columns sharded over ranks.  Each rank solves its block independently (the stopping
    rule `any(||r|| > tol)` is evaluated per rank: with a fixed iteration budget the iterates are identical to the
    unsharded solve; with a tolerance a rank may stop a few iterations earlier than the slowest column elsewhere
    would force, never later).  Returns the full solution on every rank if `gather`."""
    import torch.distributed as dist
    Bl, (lo, hi) = shard_columns(B, group)
    xl, info = alg(A, Bl)
    if not gather or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return xl, info
    world = dist.get_world_size(group)
    n, k = B.shape
    out = torch.zeros((n, k), dtype=xl.dtype, device=xl.device)
    out[:, lo:hi] = xl
    dist.all_reduce(out, group=group)   # disjoint column blocks: sum == concatenation
    return out, info
