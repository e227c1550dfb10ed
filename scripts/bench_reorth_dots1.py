"""reorth_dots on a single fp64 column (cfg5 shape): chunk height / resident CTAs / row parts."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cola_b200 as cb
from cola_b200 import backend as be
from bench import time_kernel
dev = torch.device("cuda:0")
n, b, dt = 1 << 24, 1, torch.float64
for nj in (16, 64, 120):
    V = torch.randn(nj + 1, n, b, dtype=dt, device=dev); W = torch.randn(n, b, dtype=dt, device=dev)
    C = torch.zeros(nj + 1, b, dtype=torch.float64, device=dev)
    ms = time_kernel(lambda: be.reorth_dots(V, 1, nj + 1, W, C), reps=10)
    ref = (V[1:, :, 0] @ W[:, 0])
    C.zero_(); be.reorth_dots(V, 1, nj + 1, W, C); torch.cuda.synchronize()
    err = float((C[1:, 0] - ref).abs().max() / ref.abs().max())
    print(f"WROWS={os.environ.get('COLA_REORTH_WROWS','-')} CTAS={os.environ.get('COLA_REORTH_DOTS_CTAS','-')} Q={os.environ.get('COLA_REORTH_QSPLIT','-')} nj={nj}: {ms:.3f} ms {(nj+1)*n*8/ms*1e-6:.0f} GB/s err {err:.1e}", flush=True)
    del V, W
