"""Secondary BASELINE configs (3, 4, 5) on one B200: prints one JSON line per workload.
    python scripts/bench_extra.py [cfg3] [cfg4] [cfg5] [--probes P] [--nodes LOG2N]
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cola_b200 as cb
from bench import peaks

RANK = int(os.environ.get("RANK", 0)); WORLD = int(os.environ.get("WORLD_SIZE", 1)); LOCAL = int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device(f"cuda:{LOCAL}")
torch.cuda.set_device(LOCAL)
GROUP = None
if WORLD > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
    GROUP = dist.group.WORLD
PEAK, _ = peaks()


def factor(d, seed):
    g = torch.Generator().manual_seed(seed)
    G = torch.randn(d, d, generator=g)
    return (G @ G.T / d + 0.5 * torch.eye(d)).to(dev)


def timed(fn, reps=1):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3, out


def cfg1():
    """BASELINE config 1 (Dense 1024^2 SPD, 1 RHS, fp32) on the GPU path (the config itself is the CPU-runnable one)."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from tests import problems as pb
    P = pb.problem("cfg1_dense1024")
    A = pb.to_b200(P["spec"], str(dev), P["ann"])
    b = P["B"].to(dev)
    alg = cb.linalg.CG(tol=1e-6, max_iters=1000)
    alg(A, b)
    s, (x, info) = timed(lambda: alg(A, b), reps=5)
    its = info["iterations"] - 1
    print(json.dumps({"workload": "cfg1: CG, PSD(Dense 1024x1024), 1 RHS, fp32, tol 1e-6 (GPU path; CUDA-graph batches)",
                      "iterations": its, "iters_per_s": its / s, "solve_ms": s * 1e3,
                      "final_error": float(info["errors"][-1])}))


def refcuda(iters=10):
    """The reference ALGORITHM (oracle port: same eager torch ops, two host syncs per iteration) on CUDA tensors:
    cuSPARSE SpMM + ~45 elementwise launches per iteration -- the 'reference on the same B200' bar of SURVEY 8d."""
    from oracle import krylov_oracle as ko
    from bench import laplacian_coo, rhs_block
    data, rows, cols, shape = laplacian_coo(2048, torch.float32, "cpu")
    A = ko.SparseOp(data, rows, cols, shape)
    A.csr = A.csr.to(dev)
    B = rhs_block(shape[0], 64, 0).to(dev)
    ko.cg(A, B, tol=1e-30, max_iters=2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, _, its, info = ko.cg(A, B, tol=1e-30, max_iters=iters)
    torch.cuda.synchronize()
    s = time.perf_counter() - t0
    print(json.dumps({"workload": "cfg2 with the reference's eager torch algorithm on CUDA tensors (oracle port, cuSPARSE SpMM)",
                      "iters_per_s": its / s, "ms_per_iter": s / its * 1e3}))


def cfg3(iters=100):
    Fs = [factor(64, i) for i in range(3)]
    K = cb.ops.Kronecker(*[cb.PSD(cb.ops.Dense(F)) for F in Fs])
    A = K + 0.1 * cb.ops.I_like(K)
    n, k = 64**3, 128
    B = torch.randn(n, k, generator=torch.Generator().manual_seed(0)).to(dev)
    alg = cb.linalg.CG(tol=1e-30, max_iters=iters)
    for _ in range(5):     # a 100-iteration solve is ~50 ms: let the clocks ramp before timing
        alg(A, B)
    s, (x, info) = timed(lambda: alg(A, B), reps=10)
    by = 3 * 64 * 64 * 4 + 11 * n * k * 4
    r_true = float((torch.linalg.norm(B - A @ x, dim=0) / torch.linalg.norm(B, dim=0)).mean())
    print(json.dumps({"workload": "cfg3: CG, Kronecker(64x64 x3)+0.1 I, n=262144, 128 RHS, fp32 (tcgen05 3xTF32 matmat)",
                      "iters_per_s": iters / s, "ms_per_iter": s / iters * 1e3, "roofline_frac": by * iters / s * 1e-9 / PEAK,
                      "algorithmic_GB_per_iter": by * 1e-9, "final_mean_rel_residual": r_true,
                      "recurrence_residual": float(info["errors"][-1])}))


def cfg4(probes=128, m=int(os.environ.get("LANCZOS_M", 100))):
    dims = (128, 128, 64)
    Fs = [factor(d, i) for i, d in enumerate(dims)]
    n = dims[0] * dims[1] * dims[2]
    dg = (torch.rand(n, generator=torch.Generator().manual_seed(3)) + 0.5).to(dev)
    K = cb.ops.Kronecker(*[cb.PSD(cb.ops.Dense(F)) for F in Fs])
    A = cb.PSD(K + cb.ops.Diagonal(dg))
    chunk = int(os.environ.get("PROBE_CHUNK", 64))
    vtol = 1.0 / (probes ** 0.5)
    f = lambda: cb.linalg.stochastic_lanczos_quad(A, torch.log, max_iters=m, tol=1e-7, vtol=vtol * 0.9999, key=42,
                                                   probe_chunk_size=chunk, group=GROUP)
    cb.linalg.stochastic_lanczos_quad(A, torch.log, max_iters=m, tol=1e-7, vtol=1.0 / (chunk ** 0.5) * 0.9999, key=1,
                                      probe_chunk_size=chunk)   # warm-up (one chunk): kernels loaded, allocator holds the basis block
    if GROUP is not None:
        dist.barrier()
    s, val = timed(f)
    if GROUP is not None:
        t = torch.tensor([s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        s = float(t[0])
        if RANK != 0:
            return
    by_probe = (2 * m * m + 16 * m) * n * 4
    # exact logdet of Kronecker + Diagonal is not separable; report the estimate and the Kronecker-only bound
    print(json.dumps({"workload": f"cfg4: SLQ logdet, Lanczos {m} iters x {probes} probes, Kronecker(128,128,64)+Diagonal, n=2^20, fp32, "
                                  f"probe chunk {chunk}",
                      "n_gpus": WORLD, "sharding": "probe columns by rank, one NCCL all-reduce of (sum, count)",
                      "seconds": s, "seconds_per_probe": s / probes, "extrapolated_1024_probes_s": s / probes * 1024,
                      "roofline_frac_per_gpu": by_probe * probes / WORLD / s * 1e-9 / PEAK, "algorithmic_GB_per_probe": by_probe * 1e-9,
                      "logdet_estimate": float(val)}))


def cfg4api(m=int(os.environ.get("LANCZOS_M", 100))):
    """The API form of BASELINE config 4: logdet(A, Lanczos(max_iters=100), Hutch(max_iters=10, key=42)),
    i.e. Hutchinson blocks of 100 probes over LanczosUnary(A, log) (logdet.py:111-117)."""
    dims = (128, 128, 64)
    Fs = [factor(d, i) for i, d in enumerate(dims)]
    n = dims[0] * dims[1] * dims[2]
    dg = (torch.rand(n, generator=torch.Generator().manual_seed(3)) + 0.5).to(dev)
    K = cb.ops.Kronecker(*[cb.PSD(cb.ops.Dense(F)) for F in Fs])
    A = cb.PSD(K + cb.ops.Diagonal(dg))
    hutch = cb.linalg.Hutch(max_iters=10, tol=0.0031, key=cb.rng.PRNGKey(42))   # tol small enough to run all 10 blocks
    lz = cb.linalg.Lanczos(max_iters=m, tol=1e-7)
    cb.linalg.logdet(A, cb.linalg.Lanczos(max_iters=8, tol=1e-7), cb.linalg.Hutch(max_iters=1, key=cb.rng.PRNGKey(1)))   # warm-up
    s, val = timed(lambda: cb.linalg.logdet(A, lz, hutch))
    info = getattr(hutch, "info", None)
    print(json.dumps({"workload": f"cfg4 (API form): logdet(A, Lanczos(max_iters={m}), Hutch(max_iters=10)), blocks of 100 probes over "
                                  "LanczosUnary(A, log), Kronecker(128,128,64)+Diagonal, n=2^20, fp32",
                      "seconds": s, "seconds_per_100_probe_block": s / 10, "logdet_estimate": float(val)}))


def cfg5(log2n=24, m=int(os.environ.get("LANCZOS_M", 128))):
    n = 1 << log2n
    g = torch.Generator(device=dev).manual_seed(7)
    a = torch.randint(0, n, (8 * n,), device=dev, generator=g)
    b = torch.randint(0, n, (8 * n,), device=dev, generator=g)
    keep = a != b
    a, b = a[keep], b[keep]
    key = torch.unique(torch.cat([a * n + b, b * n + a]))
    r, c = key // n, key % n
    del key, a, b
    deg = torch.bincount(r, minlength=n).to(torch.float64)
    idx = torch.arange(n, device=dev)
    rows = torch.cat([r, idx]); cols = torch.cat([c, idx])
    vals = torch.cat([-torch.ones(r.numel(), dtype=torch.float64, device=dev), deg])
    del r, c
    L = cb.SelfAdjoint(cb.ops.Sparse(vals, rows, cols, (n, n)))
    nnz = L.nnz
    del rows, cols, vals
    torch.cuda.empty_cache()
    alg = cb.linalg.Lanczos(max_iters=m, tol=1e-12, key=7)
    # warm-up: a 4-step run loads the kernels and creates the cuSOLVER handle used by the final (m x m) eigh
    cb.linalg.eig(L, 2, "LM", cb.linalg.Lanczos(max_iters=4, tol=1e-12, key=7))
    s, (ev, V) = timed(lambda: cb.linalg.eig(L, 64, "LM", alg))
    by = (2 * m * m + 16 * m) * n * 8 + m * (nnz * 12 + 4 * (n + 1))
    vtop = V.to_dense()[:, -1].contiguous()
    res = float(torch.linalg.norm(L @ vtop - ev[-1] * vtop) / ev[-1])
    print(json.dumps({"workload": f"cfg5: Lanczos eig top-64, full reorth, graph Laplacian 2^{log2n} nodes (nnz={nnz}), fp64, m={m}",
                      "seconds": s, "iters_per_s": m / s, "roofline_frac": by / s * 1e-9 / PEAK, "algorithmic_TB": by * 1e-12,
                      "lambda_max": float(ev[-1]), "top_ritz_rel_residual": res}))


if __name__ == "__main__":
    which = [a for a in sys.argv[1:] if a.startswith("cfg")] or ["cfg3", "cfg4", "cfg5"]
    probes = int(sys.argv[sys.argv.index("--probes") + 1]) if "--probes" in sys.argv else 128
    log2n = int(sys.argv[sys.argv.index("--nodes") + 1]) if "--nodes" in sys.argv else 24
    for w in which:
        {"cfg1": cfg1, "cfg3": cfg3, "cfg4api": cfg4api, "cfg4": lambda: cfg4(probes), "cfg5": lambda: cfg5(log2n), "cfgref": refcuda}[w]()
    if GROUP is not None:
        dist.barrier()
        dist.destroy_process_group()
