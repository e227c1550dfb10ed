"""Micro-benchmark of the SIMT mode contraction on BASELINE config 4's modes (tuning aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cola_b200 import backend as be
from bench import time_kernel
dev = torch.device("cuda:0")
k = 64
for (d, pre, post) in [(128, 1, 128 * 64 * k), (128, 128, 64 * k), (64, 128 * 128, k)]:
    g = torch.Generator(device="cpu").manual_seed(d + pre)
    M = torch.randn(d, d, generator=g).to(dev)
    X = torch.randn(pre, d, post, generator=g).to(dev)
    Y = torch.empty_like(X)
    be.mode_contract(M, d, d, pre, post, X, Y)
    ref = torch.einsum("aj,pjq->paq", M.double(), X.double())
    err = float((Y.double() - ref).abs().max() / ref.abs().max())
    ms = time_kernel(lambda: be.mode_contract(M, d, d, pre, post, X, Y), reps=10)
    fl = 2.0 * d * d * pre * post
    print(f"d={d} pre={pre} post={post}: {ms:.3f} ms  {fl / ms * 1e-9:.1f} TFLOP/s  {2 * X.numel() * 4 / ms * 1e-6:.0f} GB/s  rel err {err:.1e}")
