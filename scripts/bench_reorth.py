"""Micro-benchmark of the reorthogonalisation kernels (tuning aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cola_b200 as cb
from cola_b200 import backend as be
from bench import time_kernel
dev = torch.device("cuda:0")
for (n, b, dt, nj) in [(1 << 20, 100, torch.float32, 50), (1 << 20, 64, torch.float32, 50), (1 << 20, 64, torch.float32, 100), (1 << 20, 128, torch.float32, 50), (1 << 24, 1, torch.float64, 64), (1 << 22, 8, torch.float32, 32)]:
    s = 4 if dt == torch.float32 else 8
    V = torch.randn(nj + 1, n, b, dtype=dt, device=dev)
    W = torch.randn(n, b, dtype=dt, device=dev)
    C = torch.zeros(nj + 1, b, dtype=torch.float64, device=dev)
    nrm = torch.zeros(b, dtype=torch.float64, device=dev)
    ms = time_kernel(lambda: be.reorth_dots(V, 1, nj + 1, W, C), reps=10)
    by = (nj + 1) * n * b * s
    print(f"n=2^{n.bit_length()-1} b={b} {dt} nj={nj}: dots {ms:.3f} ms {by/ms*1e-6:.0f} GB/s", end="; ")
    ms = time_kernel(lambda: be.reorth_update(V, 1, nj + 1, W, C, sign=-1.0, wnorm2=nrm), reps=10)
    by = (nj + 2) * n * b * s
    print(f"update {ms:.3f} ms {by/ms*1e-6:.0f} GB/s", end="; ")
    # fused update+dots: correctness against the two separate kernels, then timing
    C.zero_(); be.reorth_dots(V, 1, nj + 1, W, C); C.mul_(1e-3)
    W1 = W.clone(); W2 = W.clone()
    Ca = torch.zeros_like(C); Cb = torch.zeros_like(C)
    be.reorth_update(V, 1, nj + 1, W1, C, sign=-1.0); be.reorth_dots(V, 1, nj + 1, W1, Ca)
    ok = be.reorth_update_dots(V, 1, nj + 1, W2, C, Cb, sign=-1.0)
    torch.cuda.synchronize()
    if ok:
        print(f"fused: W err {(W1 - W2).abs().max().item():.2e}  C2 rel err "
              f"{((Ca - Cb).abs().max() / Ca.abs().max()).item():.2e}", end="; ")
        ms = time_kernel(lambda: be.reorth_update_dots(V, 1, nj + 1, W2, C, Cb, sign=-1.0), reps=10)
        print(f"fused {ms:.3f} ms {by/ms*1e-6:.0f} GB/s (one sweep)", end="; ")
    else:
        print("fused: unsupported", end="; ")
    X = torch.randn(n, b, dtype=dt, device=dev); Y = torch.randn(n, b, dtype=dt, device=dev)
    ms = time_kernel(lambda: be.lanczos_three_term(W, X, Y, nrm, nrm), reps=10)
    print(f"three_term {ms:.3f} ms {4*n*b*s/ms*1e-6:.0f} GB/s")
    del V, W
