"""Nystrom-preconditioned CG at BASELINE cfg2 scale: cost of one preconditioner apply and of one PCG iteration."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cola_b200 as cb
from bench import laplacian_coo, time_kernel
dev = torch.device("cuda:0")
g, k, rank = int(os.environ.get("GRID", 2048)), 64, int(os.environ.get("RANK", 32))
vals, rows, cols, shape = laplacian_coo(g, torch.float32, dev)
A = cb.PSD(cb.ops.Sparse(vals, rows, cols, shape))
n = shape[0]
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
P = cb.linalg.NystromPrecond(A, rank=rank, mu=1e-6)
torch.cuda.synchronize()
print(f"NystromPrecond(rank={rank}) construction: {time.perf_counter() - t0:.2f} s")
X = torch.randn(n, k, device=dev); Z = torch.empty_like(X); d = torch.zeros(k, dtype=torch.float64, device=dev)
print(f"P apply (+<r,z>): {time_kernel(lambda: P.matmat_into(X, Z, dots=d), reps=5):.3f} ms")
B = torch.randn(n, k, generator=torch.Generator().manual_seed(0)).to(dev)
for name, alg in [("CG", cb.linalg.CG(tol=1e-30, max_iters=20)), ("PCG", cb.linalg.CG(tol=1e-30, max_iters=20, P=P))]:
    alg(A, B)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    x, info = alg(A, B)
    torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t0) / 20 * 1e3:.2f} ms/iteration, residual after 20 iterations {info['errors'][-1]:.3e}")

from cola_b200 import backend as be
U = P.U.contiguous(); SUt = (P.subspace_scaling * U.T).contiguous()
C = torch.empty(rank, k, device=dev)
print(f"  s*U^T r (split-K, {rank} x {n} times {n} x {k}): {time_kernel(lambda: be.mode_contract(SUt, rank, n, 1, k, X, C), reps=5):.3f} ms")
print(f"  U c ({n} x {rank} times {rank} x {k}): {time_kernel(lambda: be.mode_contract(U, n, rank, 1, k, C, Z), reps=5):.3f} ms")
print(f"  + r and <r, z>: {time_kernel(lambda: be.diag_matmat(X, Z, 1.0, None, True, d), reps=5):.3f} ms")
