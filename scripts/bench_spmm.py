"""Micro-benchmark of the CSR SpMM kernel on BASELINE config 2 (tuning aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cola_b200 as cb
from bench import laplacian_coo, time_kernel

dev = torch.device("cuda:0")
g = int(os.environ.get("GRID", 2048)); k = int(os.environ.get("K", 64))
data, rows, cols, shape = laplacian_coo(g, torch.float32, dev)
A = cb.ops.Sparse(data, rows, cols, shape)
n = shape[0]
p = torch.randn(n, k, device=dev); ap = torch.empty_like(p)
pap = torch.zeros((4, k), dtype=torch.float64, device=dev)
ms = time_kernel(lambda: A.matmat_into(p, ap, dots=pap), reps=30)
by = A.nnz * 8 + 4 * (n + 1) + 2 * n * k * 4
print(f"spmm+dots: {ms:.4f} ms  {by/ms*1e-6:.0f} GB/s algorithmic")
ms = time_kernel(lambda: A.matmat_into(p, ap), reps=30)
print(f"spmm plain: {ms:.4f} ms  {by/ms*1e-6:.0f} GB/s algorithmic")
ref = torch.sparse_csr_tensor(A.indptr, A.indices, A.data, size=shape) @ p
print("max err vs torch.sparse:", float((ref - ap).abs().max()))
if os.environ.get("CUSPARSE"):
    ms = time_kernel(lambda: torch.sparse.mm(torch.sparse_csr_tensor(A.indptr, A.indices, A.data, size=shape), p), reps=10)
    print(f"torch cuSPARSE spmm: {ms:.4f} ms")
