"""What does the per-step host poll of the Lanczos stop rule cost?  cfg5-shaped run with and without the synchronising read."""
import os, sys, time, importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cola_b200 as cb
from cola_b200 import backend as be
lz = importlib.import_module("cola_b200.linalg.lanczos")
dev = torch.device("cuda:0")
n = 1 << int(os.environ.get("LOG2N", 24)); m = 128
g = torch.Generator(device=dev).manual_seed(7)
a = torch.randint(0, n, (8 * n,), device=dev, generator=g); b = torch.randint(0, n, (8 * n,), device=dev, generator=g)
keep = a != b; a, b = a[keep], b[keep]
key = torch.unique(torch.cat([a * n + b, b * n + a])); r, c = key // n, key % n
del key, a, b, keep
deg = torch.bincount(r, minlength=n).to(torch.float64); idx = torch.arange(n, device=dev)
L = cb.SelfAdjoint(cb.ops.Sparse(torch.cat([-torch.ones(r.numel(), dtype=torch.float64, device=dev), deg]), torch.cat([r, idx]), torch.cat([c, idx]), (n, n)))
del r, c
v0 = torch.randn(n, 1, dtype=torch.float64, device=dev)
real_read = be.read_small
def fake_read(t):
    return torch.ones(t.shape, dtype=t.dtype)            # no synchronisation: the loop runs all m steps blind
for name, fn in (("polling every step", real_read), ("no poll (blind)", fake_read), ("polling every step", real_read), ("no poll (blind)", fake_read)):
    lz.be.read_small = fn
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st = lz.lanczos_fact(L, v0, m, 1e-12)
    torch.cuda.synchronize(); t = time.perf_counter() - t0
    print(f"{name}: {t*1e3:.1f} ms for {m} steps ({t/m*1e3:.3f} ms/step)", flush=True)
    del st
