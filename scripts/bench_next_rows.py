"""Device timings of the SURVEY 8f rows built after the main path (KronSum / Tridiagonal matmats, f(A)v through
Arnoldi and Lanczos, exact and off-diagonal estimators).  One JSON line per measurement: CUDA events on the
launching stream, 3 warm-ups, operands larger than L2 where the kernel streams (said per line).  The peak is
MEASURED_PEAKS.json's HBM copy bandwidth when present.

    python scripts/bench_next_rows.py > profiles/r1_next_rows.jsonl
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import cola_b200 as cb  # noqa: E402
from bench import laplacian_coo, time_kernel  # noqa: E402

dev = torch.device("cuda:0")
L, ops = cb.linalg, cb.ops
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    HBM = float(PEAK.get("hbm_gbs", 6551.7))
except Exception:
    HBM = 6551.7


def emit(**kw):
    print(json.dumps(kw), flush=True)


def section(fn):
    try:
        fn()
    except Exception as e:  # keep the other measurements
        emit(row=fn.__name__, error=f"{type(e).__name__}: {e}"[:300])


def kron_factor(d, seed):
    G = torch.randn(d, d, generator=torch.Generator().manual_seed(seed))
    return (G @ G.T / d + 0.5 * torch.eye(d)).to(dev)


def kronsum_matmat():
    dims, k = (64, 64, 64), 128
    A = ops.KronSum(*[ops.Dense(kron_factor(d, i)) for i, d in enumerate(dims)])
    n = A.shape[0]
    X = torch.randn(n, k, device=dev)
    Y = torch.empty_like(X)
    ms = time_kernel(lambda: A.matmat_into(X, Y), reps=10)
    alg = sum(d * d for d in dims) * 4 + 2 * n * k * 4
    emit(row="KronSum matmat (operators.py:261-268)", workload="KronSum(64,64,64) fp32, 128 RHS", ms=ms,
         algorithmic_GB=alg / 1e9, achieved_GBps=alg / ms / 1e6, frac_of_hbm=alg / ms / 1e6 / HBM,
         note="tcgen05 per-mode kernel (mode_tc_kernel, 3xTF32), one launch per factor: each mode reads X and accumulates "
              "into Y (moves (3D-1) blocks of 134 MB: 1.07 GB for D = 3); operand 134 MB > L2")
    core = A.plan().terms[0][1][0]
    core.use_tensor_cores = False
    ms = time_kernel(lambda: A.matmat_into(X, Y), reps=10)
    emit(row="KronSum matmat, exact-fp32 SIMT contractions (A/B)", workload="same", ms=ms, achieved_GBps=alg / ms / 1e6,
         frac_of_hbm=alg / ms / 1e6 / HBM, note="3 accumulating mode_contract launches: X read 3x, Y read-modified twice")


def tridiagonal_matmat():
    n, k = 1 << 22, 64
    g = torch.Generator().manual_seed(1)
    A = ops.Tridiagonal(torch.randn(n - 1, generator=g).to(dev), (torch.randn(n, generator=g) + 3).to(dev),
                        torch.randn(n - 1, generator=g).to(dev))
    X = torch.randn(n, k, device=dev)
    Y = torch.empty_like(X)
    ms = time_kernel(lambda: A.matmat_into(X, Y), reps=10)
    nnz = 3 * n - 2
    alg = nnz * 8 + 4 * (n + 1) + 2 * n * k * 4
    emit(row="Tridiagonal matmat (operators.py:365-372)", workload="n=2^22 fp32, 64 RHS, CSR core (3 nnz/row)", ms=ms,
         algorithmic_GB=alg / 1e9, achieved_GBps=alg / ms / 1e6, frac_of_hbm=alg / ms / 1e6 / HBM,
         note="operand 1.07 GB > L2")


def arnoldi_unary():
    n, b, m = 1 << 20, 16, 30
    g = torch.Generator().manual_seed(2)
    A = ops.Tridiagonal((0.3 * torch.randn(n - 1, generator=g)).to(dev), (torch.rand(n, generator=g) + 1).to(dev),
                        (0.3 * torch.randn(n - 1, generator=g)).to(dev))
    V = torch.randn(n, b, device=dev)
    F = L.exp(-1.0 * A, L.Arnoldi(max_iters=m, tol=1e-12))
    F @ V
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    Y = F @ V
    torch.cuda.synchronize()
    s = time.perf_counter() - t0
    # MGS: step j reads w and q_0..q_j once each (+ matmat 2 passes + normalise 2): sum_j (j + 6) vectors
    alg = sum(j + 6 for j in range(m)) * n * b * 4 + 2 * (m + 1) * n * b * 4
    emit(row="ArnoldiUnary exp(-A) V (unary.py:63-91)", workload=f"Tridiagonal n=2^20 fp32, {b} vectors, {m} Arnoldi steps",
         seconds=s, algorithmic_GB=alg / 1e9, achieved_GBps=alg / s / 1e9, frac_of_hbm=alg / s / 1e9 / HBM,
         result_dtype=str(Y.dtype), note="whole call incl. per-step host poll, eig(H) (library) and Q@coef (2 sweeps)")


def lanczos_unary_sqrt():
    dims, b, m = (128, 128, 64), 64, 30
    A = cb.PSD(ops.Kronecker(*[ops.Dense(kron_factor(d, i)) for i, d in enumerate(dims)])
               + ops.Diagonal((torch.rand(1 << 20, generator=torch.Generator().manual_seed(3)) + 0.5).to(dev)))
    V = torch.randn(A.shape[0], b, device=dev)
    F = L.sqrt(A, L.Lanczos(max_iters=m, tol=1e-7))
    F @ V
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    F @ V
    torch.cuda.synchronize()
    s = time.perf_counter() - t0
    n = A.shape[0]
    alg = (2 * m * m + 16 * m) * n * 4 * b
    emit(row="LanczosUnary sqrt(A) V (unary.py:37-60)", workload=f"cfg4 operator (n=2^20) fp32, {b} vectors, {m} Lanczos steps",
         seconds=s, algorithmic_GB=alg / 1e9, achieved_GBps=alg / s / 1e9, frac_of_hbm=alg / s / 1e9 / HBM,
         note="SURVEY 8d Lanczos byte model (4-sweep CGS2)")


def exact_diag():
    dims = (64, 64)
    A = ops.Kronecker(*[ops.Dense(kron_factor(d, i)) for i, d in enumerate(dims)]) + ops.Diagonal(torch.rand(4096).to(dev))
    L.diag(A, 0, L.Exact())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d = L.diag(A, 0, L.Exact())      # (k != 0 on a Sum raises, as in the reference: diagonal_estimation.py asserts there)
    torch.cuda.synchronize()
    emit(row="exact_diag k=0 (diagonal_estimation.py:117-128)", workload="Kronecker(64,64)+Diagonal, n=4096 fp32, 41 blocks of 100",
         seconds=time.perf_counter() - t0, n_out=int(d.numel()), note="launch-bound: 41 fused matmats on (4096, 100) blocks")


def hutch_offdiag():
    vals, rows, cols, shape = laplacian_coo(1024, torch.float32, dev)
    A = cb.PSD(ops.Sparse(vals, rows, cols, shape))
    alg = L.Hutch(tol=2e-2, max_iters=4, key=cb.rng.PRNGKey(9))
    alg(A, 1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d = alg(A, 1)
    torch.cuda.synchronize()
    emit(row="Hutchinson k=1 (diagonal_estimation.py:158-210)", workload="CSR Laplacian 1024^2 fp32, 100-probe blocks, <=4 blocks",
         seconds=time.perf_counter() - t0, n_out=int(d.numel()), mean=float(d.mean()), expect=-1.0,
         note="first off-diagonal of the 5-point Laplacian is -1 except at grid-row boundaries")


if __name__ == "__main__":
    before = cb.backend.lib().launch_count()
    for fn in (kronsum_matmat, tridiagonal_matmat, arnoldi_unary, lanczos_unary_sqrt, exact_diag, hutch_offdiag):
        section(fn)
    emit(row="total", gpu_launches=cb.backend.lib().launch_count() - before, hbm_peak_GBps=HBM)
