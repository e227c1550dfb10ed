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
