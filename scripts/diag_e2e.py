"""Where does the end-to-end step go at N > 1?  Per-rank copy bandwidths (alone / both directions / under a solve) and the
overlap of the double-buffered schedule.  torchrun --nproc-per-node N scripts/diag_e2e.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import cola_b200 as cb
from bench import laplacian_coo, e2e_double_buffered

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

n, k = 2048 * 2048, 64
B_host = torch.randn(n, k).pin_memory()
x_host = torch.empty(n, k).pin_memory()
Bd = torch.empty(n, k, device=dev)
xd = torch.randn(n, k, device=dev)
s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

def timed_copy(which, reps=3):
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t0 = time.perf_counter()
    for _ in range(reps):
        if which in ("h2d", "both"):
            with torch.cuda.stream(s_in):
                Bd.copy_(B_host, non_blocking=True)
        if which in ("d2h", "both"):
            with torch.cuda.stream(s_out):
                x_host.copy_(xd, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return dt

res = {}
for which in ("h2d", "d2h", "both"):
    res[which] = timed_copy(which)
data, rows, cols, shape = laplacian_coo(2048, torch.float32, dev)
A = cb.PSD(cb.ops.Sparse(data, rows, cols, shape))
alg = cb.linalg.CG(tol=1e-30, max_iters=50)
for _ in range(2):
    alg(A, Bd)
barrier()
t0 = time.perf_counter(); alg(A, Bd); torch.cuda.synchronize(); res["solve"] = time.perf_counter() - t0
# copies in both directions while a solve runs
barrier()
t0 = time.perf_counter()
with torch.cuda.stream(s_in):
    Bd2 = torch.empty_like(Bd); Bd2.copy_(B_host, non_blocking=True)
with torch.cuda.stream(s_out):
    x_host.copy_(xd, non_blocking=True)
alg(A, Bd)
t_solve = time.perf_counter() - t0
torch.cuda.synchronize()
res["solve_with_copies"] = t_solve; res["solve_with_copies_total"] = time.perf_counter() - t0
for steps in (4, 10):
    barrier()
    t0 = time.perf_counter()
    e2e_double_buffered(alg, A, B_host, x_host, dev, steps)
    res[f"e2e_{steps}_per_step"] = (time.perf_counter() - t0) / steps
msg = f"rank {rank}/{world}: " + ", ".join(f"{k_} {v*1e3:.1f} ms" for k_, v in res.items())
gb = 1.073741824
msg += f" | h2d {gb/res['h2d']:.1f} GB/s, d2h {gb/res['d2h']:.1f} GB/s, both {2*gb/res['both']:.1f} GB/s total"
print(msg, flush=True)
if world > 1:
    dist.destroy_process_group()
