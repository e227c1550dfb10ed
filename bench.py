"""bench.py -- headline benchmark of the Krylov hot path (BASELINE.json: "CG iters/s (4M-row CSR, 64 RHS)").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg2|cfg3|slq] [--small]

One *step* = one CG solve (`cola_b200.linalg.CG(tol=1e-30, max_iters=ITERS)`) on BASELINE config 2:
5-point Laplacian on a 2048x2048 grid as CSR (n = 4,194,304, nnz = 20,963,328, int32 indices), 64 right-hand
sides, fp32, ITERS fixed iterations per solve (tol=1e-30 so every solve does exactly ITERS iterations).
`value` = CG iterations per second (whole job, inputs resident in HBM); at N > 1 every rank solves its own
64-RHS block of the same operator (RHS sharding, no data-path collective) and value counts all ranks' iterations.
`e2e` = same metric with the right-hand sides in pinned HOST memory and the solution copied back, every step's copies
inside the timed region; they run double-buffered on a copy stream under the next step's iterations (checked against
the device-resident solution afterwards; a failure falls back to the serial copy-solve-copy form, see `e2e.mode`).  `roofline` = the dominant kernel's algorithmic bytes / CUDA-event time vs the measured HBM peak.
`cpu_baseline` = the CPU oracle (torch CPU restatement of the reference, oracle/) on the same workload, bounded.

`--impl reference` times that CPU oracle alone (the reference itself is pure Python + packages absent on the GPU
box; DESIGN.md "Reference arm").
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ITERS = 50          # CG iterations per solve (step)
GRID = 2048         # cfg2 grid side
K_RHS = 64


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


def laplacian_coo(g, dtype, device):
    """5-point Laplacian kron(I,T)+kron(T,I), T=tridiag(-1,2,-1), as (row, col)-sorted COO."""
    n = g * g
    idx = torch.arange(n, dtype=torch.int64, device=device)
    ix, iy = idx // g, idx % g
    masks = [iy > 0, iy < g - 1, ix > 0, ix < g - 1]
    offs = [-1, 1, -g, g]
    rows = [idx] + [idx[m] for m in masks]
    cols = [idx] + [idx[m] + o for m, o in zip(masks, offs)]
    vals = [torch.full((n, ), 4.0, dtype=dtype, device=device)] + \
           [torch.full((int(m.sum()), ), -1.0, dtype=dtype, device=device) for m in masks]
    rows, cols, vals = torch.cat(rows), torch.cat(cols), torch.cat(vals)
    order = torch.argsort(rows * n + cols)
    return vals[order], rows[order], cols[order], (n, n)


def rhs_block(n, k, seed, dtype=torch.float32):
    gen = torch.Generator().manual_seed(seed)
    return torch.randn(n, k, dtype=dtype, generator=gen)


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).

    Sampling goes through NVML in-process (nvidia_ml_py): a query costs microseconds.  Spawning `nvidia-smi -lms`
    (the fallback when NVML cannot be imported) stalls kernel launches for tens of milliseconds per sample on some
    boxes, which showed up as a 4-10 % slower timed region."""
    REASONS = (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40))
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.thread, self.stop = index, [], None, None, threading.Event()
        self.source = None

    def __enter__(self):
        if os.environ.get("COLA_BENCH_NO_CLOCKS"):
            return self
        period = float(os.environ.get("COLA_BENCH_CLOCK_MS", "100")) * 1e-3
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def loop():
                while not self.stop.is_set():
                    try:
                        sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        mask = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                        self.rows.append((sm, mx, [n for n, bit in self.REASONS if mask & bit]))
                    except Exception:
                        pass
                    self.stop.wait(period)
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            self.source = "nvml"
            return self
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "500", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            self.source = "nvidia-smi"
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            try:
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                self.rows.append((float(r[1]), float(r[2]), [n for n, v in zip(names, r[4:8]) if v.lower().startswith("active")]))
            except Exception:
                pass

    def __exit__(self, *exc):
        self.stop.set()
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        elif self.thread is not None:
            self.thread.join(timeout=1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": self.source}
        reasons = sorted({n for _, _, rs in self.rows for n in rs})
        return {"sm_mhz": statistics.median(r[0] for r in self.rows), "sm_max_mhz": max(r[1] for r in self.rows),
                "reasons": reasons, "samples": len(self.rows), "source": self.source}


def e2e_double_buffered(alg, A, B_host, x_host, dev, steps):
    """End-to-end steps through the public API with HOST buffers: every step copies its right-hand-side block from
    pinned host memory to the device and its solution back, all inside the timed region.  The copies run on a second
    stream: step s+1's H2D and step s's D2H overlap step s+1's iterations (two device input buffers, the usual
    prefetch of a data loader), instead of idling the GPU ~40 ms per 110 ms solve."""
    main = torch.cuda.current_stream()
    side = torch.cuda.Stream(device=dev)
    n, k = B_host.shape
    Bd = [torch.empty((n, k), dtype=B_host.dtype, device=dev) for _ in range(2)]
    loaded = [torch.cuda.Event() for _ in range(2)]            # H2D into buffer i finished
    released = [torch.cuda.Event() for _ in range(2)]          # the solve that read buffer i finished
    iters = 0
    with torch.cuda.stream(side):
        Bd[0].copy_(B_host, non_blocking=True)
        loaded[0].record(side)
    for s in range(steps):
        cur, nxt = s % 2, (s + 1) % 2
        if s + 1 < steps:
            with torch.cuda.stream(side):
                if s >= 1:
                    side.wait_event(released[nxt])             # buffer nxt was the input of solve s-1
                Bd[nxt].copy_(B_host, non_blocking=True)
                loaded[nxt].record(side)
        main.wait_event(loaded[cur])
        x, info = alg(A, Bd[cur])
        released[cur].record(main)
        x.record_stream(side)                                  # keep x's memory until the side stream has read it
        with torch.cuda.stream(side):
            side.wait_event(released[cur])
            x_host.copy_(x, non_blocking=True)
        iters += info["iterations"] - 1
    torch.cuda.synchronize()
    return iters


def time_kernel(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps  # ms


def cg_kernel_breakdown(A, n, k, dev):
    """Per-kernel device time of one CG iteration (CUDA events on the launching stream)."""
    import cola_b200.backend as be
    lib = be.lib()
    dt = torch.float32
    x = torch.zeros(n, k, dtype=dt, device=dev)
    r = torch.randn(n, k, dtype=dt, device=dev)
    p = r.clone()
    ap = torch.empty_like(r)
    gamma = torch.ones((4, k), dtype=torch.float64, device=dev)
    pap = torch.full((4, k), 1e6, dtype=torch.float64, device=dev)
    ctl = torch.tensor([0, 0, 1000, k], dtype=torch.int32, device=dev)
    it_ptr, done_ptr = ctl[0:1], ctl[1:2]
    s = 4
    out = {}
    out["csr_spmm+pAp"] = (time_kernel(lambda: A.matmat_into(p, ap, dots=pap, dots_row=it_ptr, gate=done_ptr)),
                           A.nnz * (s + 4) + 4 * (n + 1) + 2 * n * k * s)
    out["cg_update_r"] = (time_kernel(lambda: lib.call("cola_cg_update_r_f32", be.ptr(r), be.ptr(ap), n, k, k,
                                                       be.ptr(ctl), be.ptr(gamma), be.ptr(pap), be.ptr(gamma),
                                                       be.stream_ptr())), 3 * n * k * s)
    out["cg_update_xp"] = (time_kernel(lambda: lib.call("cola_cg_update_xp_f32", be.ptr(x), be.ptr(r), be.ptr(p), n, k,
                                                        k, be.ptr(ctl), be.ptr(gamma), be.ptr(pap),
                                                        be.stream_ptr())), 5 * n * k * s)
    return out


def run_ours(args):
    import cola_b200 as cb
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    lib = cb.backend.lib()
    g = 256 if args.small else GRID
    k = K_RHS
    data, rows, cols, shape = laplacian_coo(g, torch.float32, dev)
    A = cb.PSD(cb.ops.Sparse(data, rows, cols, shape))
    n, nnz = shape[0], int(data.numel())
    del rows, cols
    B_host = rhs_block(n, k, seed=rank).pin_memory()
    B = B_host.to(dev)
    alg = cb.linalg.CG(tol=1e-30, max_iters=ITERS)

    def step():
        x, info = alg(A, B)
        return x, info

    for _ in range(args.warmup):
        step()
    # ---- timed region: inputs resident in HBM; vectors (1.07 GB each) >> 126 MB L2, no flush needed
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    launches0 = lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        e0.record()
        iters_done = 0
        for _ in range(args.steps):
            x, info = step()
            iters_done += info["iterations"] - 1
        e1.record()
        torch.cuda.synchronize()
    launches = lib.launch_count() - launches0
    elapsed_ms = e0.elapsed_time(e1)
    # ---- e2e: host buffers, H2D + D2H inside the timed region, through the public API
    x_host = torch.empty((n, k), dtype=torch.float32).pin_memory()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    x_ref = x                                                   # solution of the same block from the timed region
    e2e_mode = "double-buffered"
    t0 = time.perf_counter()
    try:
        e2e_iters = e2e_double_buffered(alg, A, B_host, x_host, dev, args.steps)
        e2e_s = time.perf_counter() - t0
        # the copies ran on a second stream: check that what arrived on the host is the solution
        err = float((x_host.to(dev) - x_ref).norm() / x_ref.norm())
        if not err < 1e-3:
            raise RuntimeError(f"double-buffered e2e returned a different solution (relative error {err:.2e})")
    except Exception as exc:                                   # measure the plain serial form instead
        e2e_mode = f"serial (double-buffered path failed: {type(exc).__name__}: {exc})"[:200]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_iters = 0
        for _ in range(args.steps):
            Bd = B_host.to(dev, non_blocking=True)
            x, info = alg(A, Bd)
            x_host.copy_(x, non_blocking=True)
            torch.cuda.synchronize()
            e2e_iters += info["iterations"] - 1
        e2e_s = time.perf_counter() - t0
    t = torch.tensor([elapsed_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
    assert iters_done == args.steps * ITERS, (iters_done, args.steps, ITERS)
    value = world * iters_done / (elapsed_ms * 1e-3)
    e2e_value = world * e2e_iters / (e2e_ms * 1e-3)
    # ---- roofline of the dominant kernel + per-kernel breakdown (CUDA events, same process)
    peak, peak_src = peaks()
    br = cg_kernel_breakdown(A, n, k, dev)
    kern = {name: {"ms": ms, "algorithmic_bytes": by, "GBps": by / ms * 1e-6} for name, (ms, by) in br.items()}
    dom = max(kern, key=lambda nm: kern[nm]["ms"])
    iter_bytes = nnz * 8 + 4 * (n + 1) + 11 * n * k * 4   # B_param + 11 n k s
    ms_iter = elapsed_ms / iters_done
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["GBps"], "peak": peak, "unit": "GB/s",
                "frac": kern[dom]["GBps"] / peak, "traffic": ncu_traffic(dom, g, k), "peak_source": peak_src,
                "kernels": kern,
                "iteration": {"algorithmic_bytes": iter_bytes, "ms": ms_iter, "GBps": iter_bytes / ms_iter * 1e-6,
                              "frac": iter_bytes / ms_iter * 1e-6 / peak,
                              "sum_kernel_ms": sum(v["ms"] for v in kern.values())}}
    # ---- CPU baseline: the oracle on the box's host cores, bounded sample (same workload, few iterations)
    cpu = None
    if not args.no_cpu_baseline:
        cpu = cpu_reference_rate(g, k, sample_iters=2 if not args.small else 10)
    out = {
        "metric": "CG iters/s (4M-row CSR, 64 RHS)", "value": value, "unit": "iterations/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"cfg2: CG on CSR 5-point Laplacian {g}x{g} grid (n={n}, nnz={nnz}, int32 indices), "
                               f"{k} RHS per GPU, fp32, {ITERS} fixed iterations per solve (tol=1e-30)",
                   "iters_per_step": ITERS, "rhs_per_gpu": k, "sharding": "RHS blocks per rank, operator replicated",
                   "l2": "no flush: every vector block is 1.07 GB >> 126 MB L2"},
        "clocks": clocks.summary(), "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": "iterations/s", "h2d_bytes_per_step": n * k * 4,
                "d2h_bytes_per_step": n * k * 4, "mode": e2e_mode},
        "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def cpu_reference_rate(g, k, sample_iters, threads=None):
    """CG iterations/s of the CPU oracle (torch CPU, all host threads) on the same operator and RHS block."""
    from oracle import krylov_oracle as ko
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    data, rows, cols, shape = laplacian_coo(g, torch.float32, "cpu")
    A = ko.SparseOp(data, rows, cols, shape)
    B = rhs_block(shape[0], k, seed=0)
    t0 = time.perf_counter()
    _, _, its, info = ko.cg(A, B, tol=1e-30, max_iters=sample_iters)
    loop_s = info["iteration_time"] * info["iterations"]
    return {"value": its / loop_s, "unit": "iterations/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{its} CG iterations of the full workload (n={shape[0]}, {k} RHS) with the CPU oracle "
                      f"(oracle/krylov_oracle.py, torch {torch.__version__} CPU), loop time {loop_s:.2f} s, "
                      f"wall incl. setup {time.perf_counter() - t0:.2f} s"}


def run_reference(args):
    """--impl reference: the reference's algorithm on the host cores (CPU oracle port; rank 0 only)."""
    if int(os.environ.get("RANK", 0)) != 0:
        return
    from oracle import krylov_oracle as ko
    g = 256 if args.small else GRID
    k = K_RHS
    torch.set_num_threads(os.cpu_count())
    data, rows, cols, shape = laplacian_coo(g, torch.float32, "cpu")
    A = ko.SparseOp(data, rows, cols, shape)
    B = rhs_block(shape[0], k, seed=0)
    per_step_iters = 1 if not args.small else 10     # bounded sample: 1 full-size iteration per step (~4 s)
    steps = min(args.steps, 8) if not args.small else args.steps
    warm = min(args.warmup, 1)
    for _ in range(warm):
        ko.cg(A, B, tol=1e-30, max_iters=per_step_iters)
    loop_s, its = 0.0, 0
    t0 = time.perf_counter()
    for _ in range(steps):
        _, _, n_it, info = ko.cg(A, B, tol=1e-30, max_iters=per_step_iters)
        # body time only: iteration_time is (loop wall)/(cond evaluations)
        loop_s += info["iteration_time"] * info["iterations"]
        its += n_it
    wall = time.perf_counter() - t0
    value = its / loop_s
    out = {
        "impl": "reference", "metric": "CG iters/s (4M-row CSR, 64 RHS)", "value": value, "unit": "iterations/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", 1)), "steps": steps, "warmup": warm,
        "ms_per_step": loop_s / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"cfg2: CG on CSR 5-point Laplacian {g}x{g} grid (n={shape[0]}, nnz={int(data.numel())}, int32 "
                               f"indices), {k} RHS per GPU, fp32, {ITERS} fixed iterations per solve (tol=1e-30)",
                   "iters_per_step": per_step_iters, "rhs_per_gpu": k,
                   "sample": f"each step = {per_step_iters} full-size CG iteration(s) of that solve on the host cores"},
        "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{its} full-size CG iterations, loop {loop_s:.1f} s, wall {wall:.1f} s"},
        "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def ncu_traffic(kernel, g, k):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full`
    capture (profiles/r1_ncu_full_traffic.json); None when there is no capture for this kernel / workload size."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_ncu_full_traffic.json")
    if not (os.path.exists(path) and g == 2048 and k == 64):
        return None
    rec = json.load(open(path)).get("kernels", {}).get(kernel)
    return None if rec is None else rec["traffic_bytes"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--small", action="store_true", help="256x256 grid (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
