"""Bring-up check of the per-mode tcgen05 contraction (cola_mode_contract_tc_f32) against fp64, then cfg4-shape timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cola_b200 as cb
from cola_b200 import backend as be
from bench import time_kernel

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
for d, pre, L, k in [(64, 4, 1, 32), (64, 1, 8, 64), (128, 1, 4, 32), (128, 8, 1, 64), (128, 3, 4, 96), (64, 128 * 128 // 16, 1, 64),
                     (128, 1, 512, 64), (128, 16, 32, 64)]:
    F = (torch.randn(d, d, generator=g) / d**0.5 + 0.5 * torch.eye(d)).to(dev)
    X = torch.randn(pre * d * L, k, generator=g).to(dev)
    assert be.mode_contract_tc_ok(F, pre, L, k, X)
    out = torch.full_like(X, float("nan"))
    be.mode_contract_tc(F, pre, L, k, X, out, alpha=1.5)
    ref = 1.5 * torch.einsum("aj,pjlr->palr", F.double(), X.double().reshape(pre, d, L, k)).reshape(pre * d * L, k)
    err = float((out.double() - ref).norm() / ref.norm())
    msg = f"d={d} pre={pre} L={L} k={k}: rel err {err:.2e}"
    if L == 1:
        dg = torch.rand(pre * d, generator=g).to(dev)
        dots = torch.zeros(k, dtype=torch.float64, device=dev)
        Y0 = torch.randn(pre * d, k, generator=g).to(dev)
        Y = Y0.clone()
        be.mode_contract_tc(F, pre, L, k, X, Y, alpha=1.5, shift=0.25, diag=dg, epi_x=X, accumulate=True, dots=dots)
        ref2 = ref + (0.25 + dg.double())[:, None] * X.double() + Y0.double()
        msg += f"; epilogue rel err {float((Y.double() - ref2).norm() / ref2.norm()):.2e}, dots {float(((X.double() * ref2).sum(0) - dots).abs().max() / dots.abs().max()):.2e}"
    print(msg, flush=True)
# BASELINE config 4 shapes: Kronecker(128, 128, 64), n = 2^20, 64 probes
n, k = 1 << 20, 64
X = torch.randn(n, k, device=dev)
W = torch.empty_like(X)
for d, pre, L in [(128, 1, 8192), (128, 128, 64), (64, 16384, 1)]:
    F = (torch.randn(d, d, generator=g) / d**0.5 + 0.5 * torch.eye(d)).to(dev)
    ms = time_kernel(lambda: be.mode_contract_tc(F, pre, L, k, X, W), reps=20)
    ms2 = time_kernel(lambda: be.mode_contract(F, d, d, pre, L * k, X, W), reps=5)
    fl = 2 * d * n * k
    print(f"cfg4 mode d={d} pre={pre} L={L}: tc {ms*1e3:.0f} us ({2*n*k*4/ms*1e-6:.0f} GB/s, {fl/ms*1e-9:.0f} TFLOP/s fp32-equivalent), simt {ms2*1e3:.0f} us")
dg = torch.rand(n, device=dev); dots = torch.zeros(k, dtype=torch.float64, device=dev)
F = (torch.randn(64, 64, generator=g) / 8 + 0.5 * torch.eye(64)).to(dev)
ms = time_kernel(lambda: be.mode_contract_tc(F, 16384, 1, k, X, W, diag=dg, epi_x=X, dots=dots), reps=20)
ms2 = time_kernel(lambda: be.mode_contract(F, 64, 64, 16384, k, X, W, diag=dg, epi_x=X, dots=dots), reps=5)
print(f"cfg4 last mode with Diagonal + dots epilogue: tc {ms*1e3:.0f} us, simt {ms2*1e3:.0f} us")
