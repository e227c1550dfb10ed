"""Staged SpMM (csrc/csr_tiled.cu) against the register-gather kernel and torch.sparse on stencil grids; timing on cfg2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cola_b200 as cb
from cola_b200 import backend as be
from cola_b200.csr_tiles import CsrTiles
from bench import laplacian_coo, time_kernel

dev = torch.device("cuda:0")
bad = 0
for g, k, dt, R, cap in ((96, 16, torch.float64, 32, 400), (160, 32, torch.float32, 32, 400), (256, 64, torch.float32, 32, 400),
                         (256, 32, torch.float64, 16, 200), (100, 32, torch.float32, 32, 400), (512, 64, torch.float32, 16, 200)):
    data, rows, cols, shape = laplacian_coo(g, dt, dev)
    A = cb.ops.Sparse(data, rows, cols, shape)
    T = CsrTiles(A, R, cap, k * A.data.element_size())
    n = shape[0]
    X = torch.randn(n, k, device=dev, dtype=dt)
    ref = (torch.sparse_csr_tensor(A.indptr, A.indices, A.data.double(), size=shape) @ X.double())
    for mode in ("plain", "epi", "dots", "acc"):
        Y = torch.full((n, k), 0.5, device=dev, dtype=dt)
        dots = torch.zeros((2, k), dtype=torch.float64, device=dev)
        diag = torch.rand(n, device=dev, dtype=dt)
        kw = {}
        want = ref
        if mode in ("epi", "dots"):
            kw = dict(alpha=0.5, shift=0.25, diag=diag)
            want = 0.5 * ref + (0.25 + diag.double())[:, None] * X.double()
        if mode == "dots":
            kw["dots"] = dots
        if mode == "acc":
            kw = dict(accumulate=True)
            want = ref + 0.5
        be.csr_spmm_tiled(T, T.values(A.data), shape, X, Y, **kw)
        torch.cuda.synchronize()
        err = float((Y.double() - want).abs().max())
        tol = 1e-12 if dt == torch.float64 else 2e-5
        derr = 0.0
        if mode == "dots":
            wd = (X.double() * Y.double()).sum(0)
            derr = float(((dots[0] - wd).abs() / wd.abs().clamp_min(1)).max())
        ok = err < tol and derr < 1e-5
        bad += not ok
        print(f"g={g} k={k} {str(dt)[6:]} R={R} tiles={T.n_tiles} 2d={T.n_tiles2d} regular={T.n_regular} cap_rows={T.cap_rows} {mode}: err {err:.2e} dots {derr:.1e} {'ok' if ok else 'FAIL'}")
print("FAILURES", bad)
if os.environ.get("CHECK_ONLY"):
    sys.exit(1 if bad else 0)

g = int(os.environ.get("GRID", 2048)); k = int(os.environ.get("K", 64))
data, rows, cols, shape = laplacian_coo(g, torch.float32, dev)
A = cb.ops.Sparse(data, rows, cols, shape)
n = shape[0]
p = torch.randn(n, k, device=dev); ap = torch.empty_like(p)
pap = torch.zeros((4, k), dtype=torch.float64, device=dev)
by = A.nnz * 8 + 4 * (n + 1) + 2 * n * k * 4
ms = time_kernel(lambda: be.csr_spmm(A.indptr, A.indices, A.data, shape, A.nnz, A.max_row_nnz, p, ap, dots=pap), reps=20)
print(f"register-gather spmm+dots: {ms:.4f} ms  {by/ms*1e-6:.0f} GB/s algorithmic")
want = ap.clone()
for R, cap in ((32, 400), (16, 200)):
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(); T = CsrTiles(A, R, cap, k * A.data.element_size()); vals = T.values(A.data); t1.record(); torch.cuda.synchronize()
    for ch, warps, order in ((1, 16, 0), (1, 23, 0), (2, 8, 0), (2, 12, 0), (2, 16, 0), (2, 16, 1), (4, 8, 0), (4, 15, 0)):
        os.environ["COLA_SPMM_TILE_NB"] = str(order)
        os.environ["COLA_SPMM_TILE_WARPS"] = str(warps)
        os.environ["COLA_SPMM_TILE_CH"] = str(ch)
        ap.zero_()
        ms = time_kernel(lambda: be.csr_spmm_tiled(T, vals, shape, p, ap, dots=pap), reps=20)
        print(f"staged R={R} cap={cap} ch={ch} warps={warps} batched={order} (cap_rows {T.cap_rows} cap_nz {T.cap_nz} regular {T.n_regular}/{T.n_tiles}, build {t0.elapsed_time(t1):.0f} ms): "
              f"{ms:.4f} ms  {by/ms*1e-6:.0f} GB/s algorithmic, max diff vs register-gather {float((ap-want).abs().max()):.2e}")
    del T, vals
