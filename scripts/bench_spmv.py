"""SpMV (k=1, fp64) on a random graph Laplacian (cfg5 shape) -- tuning aid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cola_b200 as cb
from bench import time_kernel
dev = torch.device("cuda:0")
log2n = int(os.environ.get("LOG2N", 24)); n = 1 << log2n
g = torch.Generator(device=dev).manual_seed(7)
a = torch.randint(0, n, (8 * n,), device=dev, generator=g); b = torch.randint(0, n, (8 * n,), device=dev, generator=g)
keep = a != b; a, b = a[keep], b[keep]
key = torch.unique(torch.cat([a * n + b, b * n + a])); r, c = key // n, key % n
deg = torch.bincount(r, minlength=n).to(torch.float64); idx = torch.arange(n, device=dev)
L = cb.ops.Sparse(torch.cat([-torch.ones(r.numel(), dtype=torch.float64, device=dev), deg]), torch.cat([r, idx]), torch.cat([c, idx]), (n, n))
del a, b, key, r, c
for k in (1, 2, 4):
    x = torch.randn(n, k, dtype=torch.float64, device=dev); y = torch.empty_like(x); d = torch.zeros(k, dtype=torch.float64, device=dev)
    ms = time_kernel(lambda: L.matmat_into(x, y, dots=d), reps=10)
    by = L.nnz * 12 + 4 * (n + 1) + 2 * n * k * 8
    ref = torch.sparse_csr_tensor(L.indptr, L.indices, L.data, size=(n, n)) @ x
    print(f"k={k}: spmv+dots {ms:.3f} ms, {by/ms*1e-6:.0f} GB/s algorithmic, max err {float((ref-y).abs().max()):.2e}")
ms = time_kernel(lambda: torch.sparse_csr_tensor(L.indptr, L.indices, L.data, size=(n, n)) @ x[:, :1].contiguous(), reps=5)
print(f"cuSPARSE spmv k=1: {ms:.3f} ms")
