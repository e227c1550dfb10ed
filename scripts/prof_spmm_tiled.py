"""One configuration of the staged SpMM on cfg2 (for ncu): R / cap / warps from the environment."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cola_b200 as cb
if os.environ.get("COLA_LIB"):
    import cola_b200.backend as _be
    _be.LIB_PATH = os.path.abspath(os.environ["COLA_LIB"])
from cola_b200 import backend as be
from cola_b200.csr_tiles import CsrTiles
from bench import laplacian_coo, time_kernel

dev = torch.device("cuda:0")
g = int(os.environ.get("GRID", 2048)); k = int(os.environ.get("K", 64))
R = int(os.environ.get("R", 32)); cap = int(os.environ.get("CAP", 400))
data, rows, cols, shape = laplacian_coo(g, torch.float32, dev)
A = cb.ops.Sparse(data, rows, cols, shape)
n = shape[0]
p = torch.randn(n, k, device=dev); ap = torch.empty_like(p)
pap = torch.zeros((4, k), dtype=torch.float64, device=dev)
T = CsrTiles(A, R, cap, k * A.data.element_size()); vals = T.values(A.data)
by = A.nnz * 8 + 4 * (n + 1) + 2 * n * k * 4
ms = time_kernel(lambda: be.csr_spmm_tiled(T, vals, shape, p, ap, dots=pap), reps=int(os.environ.get("REPS", 20)))
print(f"staged R={R} cap={cap} warps={os.environ.get('COLA_SPMM_TILE_WARPS')}: {ms:.4f} ms  {by/ms*1e-6:.0f} GB/s algorithmic")
