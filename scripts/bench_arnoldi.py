"""Arnoldi factorisation: the cooperative MGS chain against the link-by-link launches (timing + same factorisation)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cola_b200 as cb
import importlib
ar = importlib.import_module("cola_b200.linalg.arnoldi")

dev = torch.device("cuda:0")
for n, b, dt, m in ((1 << 20, 16, torch.float32, 30), (1 << 20, 16, torch.float64, 30), (1 << 22, 1, torch.float64, 40), (1 << 18, 8, torch.float32, 30)):
    g = torch.Generator(device="cpu").manual_seed(1)
    lo = torch.randn(n - 1, dtype=dt, generator=g).to(dev); up = torch.randn(n - 1, dtype=dt, generator=g).to(dev)
    d = (4 + torch.rand(n, dtype=dt, generator=g)).to(dev)
    A = cb.ops.Tridiagonal(lo, d, up)
    V = torch.randn(n, b, dtype=dt, generator=g).to(dev)
    res = {}
    for chain in (False, True):
        ar.USE_CHAIN = chain
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            Q, H, idx, info = ar.arnoldi_fact(A, V, m, 1e-12)
            torch.cuda.synchronize(); t = time.perf_counter() - t0
        res[chain] = (Q, H, t)
    links = sum(i + 2 for i in range(m))
    blk = n * b * V.element_size()
    algo = sum((i + 1) * blk + 2 * blk for i in range(m))          # each q_j once + w read / written, per step
    dq = float((res[True][0] - res[False][0]).abs().max()); dh = float((res[True][1] - res[False][1]).abs().max())
    print(f"n=2^{n.bit_length()-1} b={b} {str(dt)[6:]} m={m}: links {res[False][2]*1e3:.1f} ms ({res[False][2]/links*1e6:.0f} us/link), "
          f"chain {res[True][2]*1e3:.1f} ms ({res[True][2]/links*1e6:.0f} us/link, {algo/res[True][2]*1e-9:.0f} GB/s of the q-once model); "
          f"max |dQ| {dq:.2e} |dH| {dh:.2e}")
