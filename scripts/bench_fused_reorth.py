"""Tuning aid for the fused update+dots reorthogonalisation kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cola_b200 import backend as be
from bench import time_kernel
dev = torch.device("cuda:0")
n = int(os.environ.get("N", 1 << 20)); b = int(os.environ.get("B", 64)); nj = int(os.environ.get("NJ", 100))
dt = torch.float32
V = torch.randn(nj + 1, n, b, dtype=dt, device=dev)
W = torch.randn(n, b, dtype=dt, device=dev)
C = torch.zeros(nj + 1, b, dtype=torch.float64, device=dev)
C2 = torch.zeros_like(C)
ms = time_kernel(lambda: be.reorth_update_dots(V, 1, nj + 1, W, C, C2, sign=-1.0), reps=5)
print(f"n={n} b={b} nj={nj} R={os.environ.get('COLA_FU_R','auto')}: fused {ms:.3f} ms {(nj + 2) * n * b * 4 / ms * 1e-6:.0f} GB/s")
