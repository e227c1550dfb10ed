import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cola_b200 as cb
from cola_b200.linalg import stochastic
import cola_b200.linalg.lanczos as _unused
import importlib
lz = importlib.import_module("cola_b200.linalg.lanczos")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bench_extra import factor, dev
dims = (128, 128, 64)
Fs = [factor(d, i) for i, d in enumerate(dims)]
n = 1 << 20
dg = (torch.rand(n) + 0.5).to(dev)
K = cb.ops.Kronecker(*[cb.PSD(cb.ops.Dense(F)) for F in Fs]); A = cb.PSD(K + cb.ops.Diagonal(dg))
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for chunk in (64, 64, 128, 64):
    Z = torch.randn(n, chunk, device=dev)
    t0 = T(); st = lz.lanczos_fact(A, Z, 100, 1e-7); t1 = T()
    Tm = stochastic._tridiag_dense(st); t2 = T()
    ev, Q = torch.linalg.eigh(Tm); t3 = T()
    print(f"chunk {chunk}: lanczos {t1-t0:.3f}s  tridiag {t2-t1:.3f}s  eigh {t3-t2:.3f}s  mem {torch.cuda.memory_allocated()/2**30:.1f} GiB reserved {torch.cuda.memory_reserved()/2**30:.1f} GiB")
    del st, Tm, ev, Q

# per-kernel view of one Lanczos step at j = 50 and j = 100 vectors (CUDA events)
from bench import time_kernel
from cola_b200 import backend as be
b = 64
Z = torch.randn(n, b, device=dev)
W = torch.empty_like(Z)
acc = torch.zeros(b, dtype=torch.float64, device=dev)
print(f"matmat (Kron 128,128,64 + Diagonal, dots fused): {time_kernel(lambda: A.matmat_into(Z, W, dots=acc), reps=10):.3f} ms")
print(f"matmat plain: {time_kernel(lambda: A.matmat_into(Z, W), reps=10):.3f} ms")
os.environ["COLA_MC_NO_BIG"] = "1"
print(f"matmat (64x64x16 tiles only): {time_kernel(lambda: A.matmat_into(Z, W, dots=acc), reps=10):.3f} ms")
del os.environ["COLA_MC_NO_BIG"]
for nj in (50, 100):
    V = torch.randn(nj + 2, n, b, device=dev)
    C = torch.zeros(nj + 2, b, dtype=torch.float64, device=dev); C2 = torch.zeros_like(C)
    nrm = torch.zeros(b, dtype=torch.float64, device=dev)
    t_d = time_kernel(lambda: be.reorth_dots(V, 1, nj + 1, W, C), reps=5)
    t_f = time_kernel(lambda: be.reorth_update_dots(V, 1, nj + 1, W, C, C2), reps=5)
    t_u = time_kernel(lambda: be.reorth_update(V, 1, nj + 1, W, C2, wnorm2=nrm), reps=5)
    t_3 = time_kernel(lambda: be.lanczos_three_term(W, V[nj], V[nj - 1], acc, nrm), reps=5)
    t_s = time_kernel(lambda: be.col_scale(V[nj], V[nj], nrm, take_sqrt=True, mode=2), reps=5)
    print(f"nj={nj}: dots {t_d:.3f}  fused {t_f:.3f}  update {t_u:.3f}  three_term {t_3:.3f}  scale {t_s:.3f} ms")
    del V
