"""Bring-up check of the tcgen05 Kronecker path against fp64 and the SIMT path (run under `timeout`)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cola_b200 as cb
from bench import time_kernel

dev = torch.device("cuda:0")
torch.manual_seed(0)
for D, k in [(2, 32), (2, 64), (3, 32), (3, 128)]:
    Fs = [torch.randn(64, 64, device=dev) / 8 + 0.5 * torch.eye(64, device=dev) for _ in range(D)]
    K = cb.ops.Kronecker(*[cb.ops.Dense(F) for F in Fs])
    n = 64**D
    X = torch.randn(n, k, device=dev)
    A = K + 0.1 * cb.ops.I_like(K)
    core = A.plan().terms[0][1][0]
    Y = torch.empty_like(X); dots = torch.zeros(k, dtype=torch.float64, device=dev)
    A.matmat_into(X, Y, dots=dots)
    torch.cuda.synchronize()
    # fp64 reference by mode products
    E = X.double().reshape(*([64] * D), k)
    for i, F in enumerate(Fs):
        E = torch.moveaxis(torch.tensordot(F.double(), torch.moveaxis(E, i, 0), dims=1), 0, i)
    ref = E.reshape(n, k) + 0.1 * X.double()
    err = float((Y.double() - ref).norm() / ref.norm())
    derr = float(((X.double() * ref).sum(0) - dots).abs().max() / dots.abs().max())
    core.use_tensor_cores = False
    Y2 = torch.empty_like(X)
    A.matmat_into(X, Y2)
    err2 = float((Y2.double() - ref).norm() / ref.norm())
    core.use_tensor_cores = True
    print(f"D={D} k={k}: tc rel err {err:.2e} (dots {derr:.2e}); simt fp32 rel err {err2:.2e}")
# timing at BASELINE config 3
D, k = 3, 128
Fs = [torch.randn(64, 64, device=dev) / 8 + 0.5 * torch.eye(64, device=dev) for _ in range(D)]
K = cb.ops.Kronecker(*[cb.ops.Dense(F) for F in Fs]); A = K + 0.1 * cb.ops.I_like(K)
X = torch.randn(64**3, k, device=dev); Y = torch.empty_like(X); dots = torch.zeros(k, dtype=torch.float64, device=dev)
ms = time_kernel(lambda: A.matmat_into(X, Y, dots=dots), reps=20)
fl = 2 * 64**3 * k * 3 * 64
print(f"cfg3 matmat tc: {ms*1e3:.1f} us  -> {2*X.numel()*4/ms*1e-6:.0f} GB/s algorithmic, {fl/ms*1e-9:.1f} TFLOP/s useful ({3*fl/ms*1e-9:.1f} issued)")
A.plan().terms[0][1][0].use_tensor_cores = False
ms2 = time_kernel(lambda: A.matmat_into(X, Y, dots=dots), reps=5)
print(f"cfg3 matmat simt: {ms2*1e3:.1f} us")
