"""One launch of each reorthogonalisation kernel at cfg5 (b = 1 fp64) or cfg4 (b = 64 fp32) shapes, for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cola_b200 as cb
from cola_b200 import backend as be
dev = torch.device("cuda:0")
which = os.environ.get("SHAPE", "cfg5")
n, b, dt, nj = ((1 << 24, 1, torch.float64, 64) if which == "cfg5" else (1 << 20, 64, torch.float32, 100))
V = torch.randn(nj + 1, n, b, dtype=dt, device=dev)
W = torch.randn(n, b, dtype=dt, device=dev)
C = torch.zeros(nj + 1, b, dtype=torch.float64, device=dev); C2 = torch.zeros_like(C)
for _ in range(3):
    be.reorth_dots(V, 1, nj + 1, W, C)
    C.mul_(1e-6)
    be.reorth_update_dots(V, 1, nj + 1, W, C, C2, sign=-1.0)
    be.reorth_update(V, 1, nj + 1, W, C2, sign=-1.0)
torch.cuda.synchronize()
