"""Small invocations of the round's new kernels for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cola_b200 as cb
from cola_b200 import backend as be
dev = torch.device("cuda:0")
# cluster-reduced sweep (grid x columns >= 48K) and the plain one
for n, k in ((20000, 128), (5000, 16)):
    X = torch.randn(n, k, device=dev); Y = torch.randn(n, k, device=dev)
    d = torch.zeros(k, dtype=torch.float64, device=dev)
    be.col_dots(X, Y, d)
    print("col_dots", n, k, float((d - (X.double() * Y.double()).sum(0)).abs().max()))
# Arnoldi chain (cooperative launch)
n = 4096
A = cb.ops.Tridiagonal(torch.rand(n - 1, device=dev, dtype=torch.float64), 4 + torch.rand(n, device=dev, dtype=torch.float64), torch.rand(n - 1, device=dev, dtype=torch.float64))
Q, H, info = cb.linalg.arnoldi(A, torch.randn(n, 4, device=dev, dtype=torch.float64), max_iters=6, tol=1e-12)
print("arnoldi ok", info["iterations"])
torch.cuda.synchronize()
# single-column reorth_dots (own kernel): ragged last chunk, odd and even basis sizes
for n, nj, dt in ((100000, 5, torch.float64), (4100, 2, torch.float32), (2048 * 8 + 8, 1, torch.float64)):
    V = torch.randn(nj + 1, n, 1, dtype=dt, device=dev); W = torch.randn(n, 1, dtype=dt, device=dev)
    C = torch.zeros(nj + 1, 1, dtype=torch.float64, device=dev)
    be.reorth_dots(V, 1, nj + 1, W, C)
    ref = V[1:, :, 0].double() @ W[:, 0].double()
    print("reorth_dots1", n, nj, float((C[1:, 0] - ref).abs().max() / ref.abs().max()))
torch.cuda.synchronize()
