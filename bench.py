"""bench.py -- headline benchmark of the Krylov hot path (BASELINE.json: "CG iters/s (4M-row CSR, 64 RHS); SLQ logdet s;
% HBM roofline; 1/2/4/8 GPU").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload all|cfg2] [--small]
                    [--no-cpu-baseline] [--no-secondary]

Primary line (`metric`, `value`, `e2e`, `roofline`, `cpu_baseline`, `parity_check`): one *step* = one CG solve
(`cola_b200.linalg.CG(tol=1e-30, max_iters=ITERS)`) on BASELINE config 2: 5-point Laplacian on a 2048x2048 grid as CSR
(n = 4,194,304, nnz = 20,963,328, int32 indices), 64 right-hand sides, fp32, ITERS fixed iterations per solve.
`value` = CG iterations per second (whole job, inputs resident in HBM); at N > 1 every rank solves its own 64-RHS block
of the same operator (RHS sharding, no data-path collective) and value counts all ranks' iterations.  `e2e` = same metric
with the right-hand sides in pinned HOST memory and the solution copied back, every step's copies inside the timed
region, double-buffered on two copy streams under the neighbouring steps' iterations (checked against the
device-resident solution afterwards; a failure falls back to the serial copy-solve-copy form, see `e2e.mode`).
`roofline` = the dominant kernel's algorithmic bytes / CUDA-event time vs the measured HBM peak, with every kernel of the
iteration and the whole iteration beside it.  `parity_check` = the first residual norms of the timed solve against the
CPU oracle's trace on the same operator and right-hand sides.  `cpu_baseline` = the unmodified reference
(baseline/_ref, `kind: "reference"`; the oracle port when that install is absent) on the host cores, bounded sample.

`secondary` (same invocation, each with its own clocks and roofline; `--workload cfg2` or `--no-secondary` skips them):
  cfg3           CG on Kronecker(64x64 x3) + 0.1 I, 128 RHS, fp32 (tcgen05 mode contractions): iterations/s
  slq            BASELINE config 4: SLQ log-determinant, 100 Lanczos steps x 1024 probes, Kronecker(128,128,64) + Diagonal,
                 n = 2^20, fp32: seconds.  At N > 1 the FIXED 1024 probes are sharded over the ranks (strong scaling, one
                 all-reduce); time = max over ranks
  cfg5           Lanczos eig top-64 with full reorthogonalisation, graph Laplacian 2^24 nodes, fp64, m = 128: seconds
                 (replicas only: rank 0 runs it)
  gpu_reference  the unmodified reference on torch-CUDA tensors of the same B200 on config 2 (eager ATen + cuSPARSE)

`--impl reference` times the unmodified reference's CPU path alone on the same `config` (rank 0 only).
"""
import argparse
import gc
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
# stdout carries ONE JSON line: keep NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION in some images) out of it
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"

ITERS = 50          # CG iterations per solve (step)
GRID = 2048         # cfg2 grid side
K_RHS = 64
REF_ITERS = 1       # full-size CG iterations per step of the reference arm (bounded sample, ~1.4 s each on 16 host cores)
SLQ_PROBES, SLQ_M, SLQ_DIMS = 1024, 100, (128, 128, 64)     # BASELINE config 4
CFG5_LOG2N, CFG5_M = 24, 128                                # BASELINE config 5


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

    def __init__(self, index, background=True):
        self.index, self.rows, self.proc, self.thread, self.stop = index, [], None, None, threading.Event()
        self.source = None
        # background=False: no sampling thread; the timed code calls .sample() itself at points of its choosing (still
        # inside the timed region, under load).  For launch-latency-sensitive regions (CUDA-graph replays): an NVML query
        # from a second thread stalled those by tens of milliseconds per sample on some boxes (cfg3: 1806-2326 it/s with
        # the thread, 2442-2480 without, same box, back to back).
        self.background = background
        self._nvml = None

    def sample(self):
        if self._nvml is None:
            return
        pynvml, h, mx = self._nvml
        try:
            sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            mask = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            self.rows.append((sm, mx, [n for n, bit in self.REASONS if mask & bit]))
        except Exception:
            pass

    def __enter__(self):
        # every timed region of this file sits inside one of these: no cyclic garbage collection in it (what timeit does too):
        # a collection pauses the enqueueing thread for tens of milliseconds, longer than the 30 ms of work a CG batch queues
        gc.collect()
        self._gc_was_on = gc.isenabled()
        gc.disable()
        if os.environ.get("COLA_BENCH_NO_CLOCKS"):
            return self
        # a sample every 0.6 s (the first 30 ms in): on some boxes an NVML query stalls kernel LAUNCHES for milliseconds (long kernels do not
        # notice, graph replays and copy-stream hand-offs do: cfg3 measured 1650 vs 2350-2560 it/s, e2e 417 vs 463)
        # (six back-to-back cfg2 runs on one box: 3 samples in the 0.93 s region 534 / 540 / 520 / 537 / 490 / 496 it/s with e2e a
        # steady 515-516 (no sampler there); 2 samples: 538 / 526 / 537 / 536 / 539 / 530; none: 534 / 511 / 536)
        period = float(os.environ.get("COLA_BENCH_CLOCK_MS", "600")) * 1e-3
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self._nvml = (pynvml, h, mx)
            self.source = "nvml"
            if not self.background:
                return self

            def loop():
                self.stop.wait(min(0.03, period))     # first sample inside the region (under load), not at its very start
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
        if getattr(self, "_gc_was_on", False):
            gc.enable()
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
        if not self.rows and self._nvml is not None:
            self.sample()                              # a region shorter than the first wait: one sample as it ends

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": self.source}
        reasons = sorted({n for _, _, rs in self.rows for n in rs})
        return {"sm_mhz": statistics.median(r[0] for r in self.rows), "sm_max_mhz": max(r[1] for r in self.rows),
                "reasons": reasons, "samples": len(self.rows), "source": self.source}


def shared_config(g, n, nnz, k):
    """The `config` object BOTH arms print (the driver compares them key by key)."""
    return {"workload": f"cfg2: CG on CSR 5-point Laplacian {g}x{g} grid (n={n}, nnz={nnz}, int32 indices), "
                        f"{k} RHS per GPU, fp32, {ITERS} fixed iterations per solve (tol=1e-30)",
            "iters_per_step": ITERS, "rhs_per_gpu": k, "sharding": "RHS blocks per rank, operator replicated",
            "l2": "no flush: every vector block is 1.07 GB >> 126 MB L2",
            "reference_arm_sample": f"--impl reference: each step = {REF_ITERS} full-size CG iteration(s) of this solve with "
                                    "the unmodified reference on the host cores (rank 0); value = iterations / loop time"}


def e2e_double_buffered(alg, A, B_host, x_host, dev, steps):
    """End-to-end steps through the public API with HOST buffers: every step copies its right-hand-side block from
    pinned host memory to the device and its solution back, all inside the timed region.  The copies run on two side
    streams (H2D and D2H use different copy engines): step s+1's H2D and step s's D2H overlap step s+1's iterations
    (two device input buffers, the usual prefetch of a data loader), instead of idling the GPU ~40 ms per 110 ms solve."""
    main = torch.cuda.current_stream()
    s_in = torch.cuda.Stream(device=dev)
    s_out = torch.cuda.Stream(device=dev)
    n, k = B_host.shape
    Bd = [torch.empty((n, k), dtype=B_host.dtype, device=dev) for _ in range(2)]
    loaded = [torch.cuda.Event() for _ in range(2)]            # H2D into buffer i finished
    released = [torch.cuda.Event() for _ in range(2)]          # the solve that read buffer i finished
    iters = 0
    with torch.cuda.stream(s_in):
        Bd[0].copy_(B_host, non_blocking=True)
        loaded[0].record(s_in)
    for s in range(steps):
        cur, nxt = s % 2, (s + 1) % 2
        if s + 1 < steps:
            with torch.cuda.stream(s_in):
                if s >= 1:
                    s_in.wait_event(released[nxt])             # buffer nxt was the input of solve s-1
                Bd[nxt].copy_(B_host, non_blocking=True)
                loaded[nxt].record(s_in)
        main.wait_event(loaded[cur])
        x, info = alg(A, Bd[cur])
        released[cur].record(main)
        x.record_stream(s_out)                                 # keep x's memory until the copy stream has read it
        with torch.cuda.stream(s_out):
            s_out.wait_event(released[cur])
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


class DistCtx:
    def __init__(self):
        self.rank = int(os.environ.get("RANK", 0))
        self.world = int(os.environ.get("WORLD_SIZE", 1))
        self.local = int(os.environ.get("LOCAL_RANK", 0))
        self.dist = None
        self.group = None
        torch.cuda.set_device(self.local)
        self.dev = torch.device(f"cuda:{self.local}")
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist, self.group = dist, dist.group.WORLD

    def barrier(self):
        torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()

    def max_over_ranks(self, *vals):
        t = torch.tensor(list(vals), dtype=torch.float64, device=self.dev)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def close(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()


def run_ours(args):
    import cola_b200 as cb
    ctx = DistCtx()
    rank, world, dev = ctx.rank, ctx.world, ctx.dev
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
    ctx.barrier()
    launches0 = lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(ctx.local) as clocks:
        e0.record()
        iters_done = 0
        for _ in range(args.steps):
            x, info = step()
            iters_done += info["iterations"] - 1
        e1.record()
        torch.cuda.synchronize()
    launches = lib.launch_count() - launches0
    elapsed_ms = e0.elapsed_time(e1)
    trace_ours = [float(v) for v in info["errors"][:2]]         # ||r_2||, ||r_3|| (column means) of the last timed solve
    # ---- e2e: host buffers, H2D + D2H inside the timed region, through the public API
    x_host = torch.empty((n, k), dtype=torch.float32).pin_memory()
    x_ref = x                                                   # solution of the same block from the timed region
    e2e_mode = "double-buffered (H2D and D2H on separate copy streams)"
    # warm-up of the end-to-end path itself (untimed, like the W warm-up steps of the resident path): the copy streams, the
    # two device input buffers and the extra 1 GiB solution blocks the allocator has to find while a copy still holds the
    # previous one cost 70-150 ms ONCE (scripts/diag_e2e.py: 110.8 ms per step in a first 4-step run, 104.5 ms after it)
    try:
        e2e_double_buffered(alg, A, B_host, x_host, dev, 2)
    except Exception:
        pass
    ctx.barrier()
    gc.collect()
    gc.disable()                                                # as in the resident region (ClockSampler.__enter__)
    t0 = time.perf_counter()
    try:
        e2e_iters = e2e_double_buffered(alg, A, B_host, x_host, dev, args.steps)
        e2e_s = time.perf_counter() - t0
        gc.enable()
        # the copies ran on side streams: check that what arrived on the host is the solution
        err = float((x_host.to(dev) - x_ref).norm() / x_ref.norm())
        if not err < 1e-3:
            raise RuntimeError(f"double-buffered e2e returned a different solution (relative error {err:.2e})")
    except Exception as exc:                                   # measure the plain serial form instead
        gc.enable()
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
    elapsed_ms, e2e_ms = ctx.max_over_ranks(elapsed_ms, e2e_s * 1e3)
    del x_host, x_ref, x
    out = None
    if rank == 0:
        assert iters_done == args.steps * ITERS, (iters_done, args.steps, ITERS)
        value = world * iters_done / (elapsed_ms * 1e-3)
        e2e_value = world * e2e_iters / (e2e_ms * 1e-3)
        # ---- roofline of the dominant kernel + per-kernel breakdown (CUDA events, same process)
        peak, peak_src = peaks()
        br = cg_kernel_breakdown(A, n, k, dev)
        kern = {name: {"ms": ms, "algorithmic_bytes": by, "GBps": by / ms * 1e-6, "frac": by / ms * 1e-6 / peak}
                for name, (ms, by) in br.items()}
        dom = max(kern, key=lambda nm: kern[nm]["ms"])
        iter_bytes = nnz * 8 + 4 * (n + 1) + 11 * n * k * 4   # B_param + 11 n k s
        ms_iter = elapsed_ms / iters_done
        roofline = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["GBps"], "peak": peak, "unit": "GB/s",
                    "frac": kern[dom]["GBps"] / peak, "traffic": ncu_traffic(dom, g, k), "peak_source": peak_src,
                    "kernels": kern,
                    "iteration": {"algorithmic_bytes": iter_bytes, "ms": ms_iter, "GBps": iter_bytes / ms_iter * 1e-6,
                                  "frac": iter_bytes / ms_iter * 1e-6 / peak,
                                  "sum_kernel_ms": sum(v["ms"] for v in kern.values())}}
        out = {
            "metric": "CG iters/s (4M-row CSR, 64 RHS)", "value": value, "unit": "iterations/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": shared_config(g, n, nnz, k),
            "clocks": clocks.summary(), "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "iterations/s", "h2d_bytes_per_step": n * k * 4,
                    "d2h_bytes_per_step": n * k * 4, "mode": e2e_mode, "frac_of_value": e2e_value / value},
            "roofline": roofline,
        }
    del A, B, B_host, data
    torch.cuda.empty_cache()
    # ---- the other BASELINE configs, same invocation (every rank takes part in the sharded one)
    if args.workload == "all" and not args.no_secondary and not args.small:
        secondary = {}
        for name, fn in (("cfg3", secondary_cfg3), ("slq", secondary_slq), ("cfg5", secondary_cfg5),
                         ("gpu_reference", secondary_gpu_reference)):
            try:
                res = fn(ctx, cb)
            except Exception as exc:                            # a secondary workload must not cost the headline line
                res = {"error": f"{type(exc).__name__}: {exc}"[:300]}
            if rank == 0 and res is not None:
                secondary[name] = res
            torch.cuda.empty_cache()
        if rank == 0:
            out["secondary"] = secondary
            flat = {"slq_logdet_s": secondary.get("slq", {}).get("seconds"),
                    "cfg3_iters_per_s": secondary.get("cfg3", {}).get("iters_per_s"),
                    "cfg5_s": secondary.get("cfg5", {}).get("seconds"),
                    "gpu_reference_iters_per_s": secondary.get("gpu_reference", {}).get("iters_per_s")}
            out["secondary"].update(flat)
    if rank == 0:
        # ---- CPU baseline (bounded sample) and the parity check of the timed solve against the oracle's trace
        if not args.no_cpu_baseline:
            cpu, trace32 = cpu_reference_rate(g, k, sample_iters=3 if not args.small else 10)
            out["cpu_baseline"] = cpu
            out["parity_check"] = parity_check(g, k, trace_ours, trace32)
        else:
            out["cpu_baseline"] = None
        print(json.dumps(out))
    ctx.close()


# ----------------------------------------------------------------------------------------------- secondary workloads
def _factor(d, seed, dev):
    gen = torch.Generator().manual_seed(seed)
    G = torch.randn(d, d, generator=gen)
    return (G @ G.T / d + 0.5 * torch.eye(d)).to(dev)


def _timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3, out


TF32_PEAK_TFLOPS = 1166.0   # 128*256*8 MAC / 130.8 cycles * 2 * 148 SMs * 1.965 GHz (profiles/r2_umma_rate.log)


def secondary_cfg3(ctx, cb, iters=100, reps=25):
    """BASELINE config 3 (every rank its own 128-RHS block; rank 0 reports its own rate x world = weak scaling)."""
    dev = ctx.dev
    lib = cb.backend.lib()
    Fs = [_factor(64, i, dev) for i in range(3)]
    K = cb.ops.Kronecker(*[cb.PSD(cb.ops.Dense(F)) for F in Fs])
    A = cb.PSD(K + 0.1 * cb.ops.I_like(K))
    n, k = 64**3, 128
    B = torch.randn(n, k, generator=torch.Generator().manual_seed(ctx.rank)).to(dev)
    alg = cb.linalg.CG(tol=1e-30, max_iters=iters)
    for _ in range(5):                                          # a 100-iteration solve is ~40 ms: let the clocks ramp
        alg(A, B)
    ctx.barrier()
    l0 = lib.launch_count()
    with ClockSampler(ctx.local, background=False) as clocks:    # sampled from this thread, twice (see ClockSampler)
        def solves():
            out = None
            for i in range(reps):
                out = alg(A, B)
                if i in (reps // 3, (2 * reps) // 3):    # two samples under load (an NVML query costs up to ~25 ms on some boxes)
                    clocks.sample()
            return out
        s, (x, info) = _timed(solves)
    launches = lib.launch_count() - l0
    (s,) = ctx.max_over_ranks(s)
    if ctx.rank != 0:
        return None
    peak, _ = peaks()
    by = 3 * 64 * 64 * 4 + 11 * n * k * 4                       # factors + 11 vector passes
    rate = ctx.world * iters * reps / s
    r_true = float((torch.linalg.norm(B - A @ x, dim=0) / torch.linalg.norm(B, dim=0)).mean())
    # the matmat alone: HBM bound (X in, Y out, factors) and the 3xTF32 tensor work it issues
    Y = torch.empty_like(B)
    mm_ms = time_kernel(lambda: A.matmat_into(B, Y), reps=50, warm=5)
    mm_by = 3 * 64 * 64 * 4 + 2 * n * k * 4
    mm_flop = 3 * 2 * 64 * n * k                                # three mode contractions, fp32-equivalent flops
    return {"workload": "cfg3: CG, Kronecker(64x64 x3)+0.1 I, n=262144, 128 RHS per GPU, fp32 (tcgen05 3xTF32 mode contractions), "
                        f"{iters} fixed iterations per solve, {reps} solves timed",
            "iters_per_s": rate, "ms_per_iter": s / (iters * reps) * 1e3, "scaling": "weak", "n_gpus": ctx.world,
            "roofline": {"bound": "hbm", "algorithmic_bytes_per_iter": by, "achieved": by * iters * reps / s * 1e-9, "peak": peak,
                         "unit": "GB/s", "frac": by * iters * reps / s * 1e-9 / peak,
                         "matmat": {"ms": mm_ms, "algorithmic_bytes": mm_by, "GBps": mm_by / mm_ms * 1e-6,
                                    "frac_hbm": mm_by / mm_ms * 1e-6 / peak, "fp32_equiv_TFLOPs": mm_flop / mm_ms * 1e-9,
                                    "tf32_mma_TFLOPs_issued": 3 * mm_flop / mm_ms * 1e-9,
                                    "tf32_peak_TFLOPs": TF32_PEAK_TFLOPS, "frac_tensor": 3 * mm_flop / mm_ms * 1e-9 / TF32_PEAK_TFLOPS,
                                    "tf32_peak_source": "measured on this pool: scripts/umma_rate.cu, tcgen05.mma kind::tf32 "
                                                        "128x256x8 at 130.8 cycles per SM (profiles/r2_umma_rate.log)"}},
            "final_mean_rel_residual": r_true, "recurrence_residual": float(info["errors"][-1]),
            "clocks": clocks.summary(), "gpu_launches": int(launches)}


def secondary_slq(ctx, cb, probes=SLQ_PROBES, m=SLQ_M):
    """BASELINE config 4: the FIXED `probes` Hutchinson probes are sharded over the ranks (strong scaling), one all-reduce."""
    dev = ctx.dev
    lib = cb.backend.lib()
    dims = SLQ_DIMS
    Fs = [_factor(d, i, dev) for i, d in enumerate(dims)]
    n = dims[0] * dims[1] * dims[2]
    dg = (torch.rand(n, generator=torch.Generator().manual_seed(3)) + 0.5).to(dev)
    K = cb.ops.Kronecker(*[cb.PSD(cb.ops.Dense(F)) for F in Fs])
    A = cb.PSD(K + cb.ops.Diagonal(dg))
    chunk = 64
    vtol = 1.0 / (probes ** 0.5)
    slq = cb.linalg.stochastic_lanczos_quad
    # warm-up (one chunk, unsharded): kernels loaded, allocator holds the basis block
    slq(A, torch.log, max_iters=m, tol=1e-7, vtol=1.0 / (chunk ** 0.5) * 0.9999, key=1, probe_chunk_size=chunk)
    ctx.barrier()
    l0 = lib.launch_count()
    with ClockSampler(ctx.local) as clocks:
        s, val = _timed(lambda: slq(A, torch.log, max_iters=m, tol=1e-7, vtol=vtol * 0.9999, key=42,
                                    probe_chunk_size=chunk, group=ctx.group))
    launches = lib.launch_count() - l0
    (s,) = ctx.max_over_ranks(s)
    if ctx.rank != 0:
        return None
    peak, _ = peaks()
    by_model = (2 * m * m + 16 * m) * n * 4                     # SURVEY 8d per probe: 4 basis sweeps per step (two CGS passes)
    by_moved = (1.5 * m * m + 16 * m) * n * 4                   # what this code moves: 3 sweeps per step (update + dots fused)
    per_gpu = probes / ctx.world
    return {"workload": f"cfg4: SLQ logdet, Lanczos {m} steps x {probes} probes (fixed total, sharded over {ctx.world} GPU(s), "
                        f"one all-reduce), Kronecker{dims}+Diagonal, n=2^20, fp32, probe chunk {chunk}",
            "seconds": s, "n_gpus": ctx.world, "scaling": "strong", "probes": probes, "logdet_estimate": float(val),
            "roofline": {"bound": "hbm", "unit": "GB/s", "peak": peak,
                         "algorithmic_bytes_per_probe": by_model, "achieved": by_model * per_gpu / s * 1e-9,
                         "frac": by_model * per_gpu / s * 1e-9 / peak,
                         "moved_bytes_per_probe": by_moved, "achieved_moved": by_moved * per_gpu / s * 1e-9,
                         "frac_moved": by_moved * per_gpu / s * 1e-9 / peak,
                         "note": "frac: SURVEY 8d's 4-sweep byte model per GPU; frac_moved: the 3 sweeps the code performs"},
            "clocks": clocks.summary(), "gpu_launches": int(launches)}


def secondary_cfg5(ctx, cb, log2n=CFG5_LOG2N, m=CFG5_M):
    """BASELINE config 5 (single start vector: replicas only -- rank 0 runs it, the others wait)."""
    if ctx.rank != 0:
        ctx.barrier()
        return None
    dev = ctx.dev
    lib = cb.backend.lib()
    n = 1 << log2n
    g = torch.Generator(device=dev).manual_seed(7)
    a = torch.randint(0, n, (8 * n,), device=dev, generator=g)
    b = torch.randint(0, n, (8 * n,), device=dev, generator=g)
    keep = a != b
    a, b = a[keep], b[keep]
    key = torch.unique(torch.cat([a * n + b, b * n + a]))
    r, c = key // n, key % n
    del key, a, b, keep
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
    l0 = lib.launch_count()
    with ClockSampler(ctx.local) as clocks:
        s, (ev, V) = _timed(lambda: cb.linalg.eig(L, 64, "LM", alg))
    launches = lib.launch_count() - l0
    peak, _ = peaks()
    by = (2 * m * m + 16 * m) * n * 8 + m * (nnz * 12 + 4 * (n + 1))
    by_moved = (1.5 * m * m + 16 * m) * n * 8 + m * (nnz * 12 + 4 * (n + 1))
    vtop = V.to_dense()[:, -1].contiguous()
    res = float(torch.linalg.norm(L @ vtop - ev[-1] * vtop) / ev[-1])
    out = {"workload": f"cfg5: Lanczos eig top-64, full reorthogonalisation, graph Laplacian 2^{log2n} nodes (nnz={nnz}), fp64, "
                       f"m={m}, single start vector (replicas only)",
           "seconds": s, "iters_per_s": m / s, "lambda_max": float(ev[-1]), "top_ritz_rel_residual": res,
           "roofline": {"bound": "hbm", "unit": "GB/s", "peak": peak, "algorithmic_bytes": by, "achieved": by / s * 1e-9,
                        "frac": by / s * 1e-9 / peak, "moved_bytes": by_moved, "frac_moved": by_moved / s * 1e-9 / peak},
           "clocks": clocks.summary(), "gpu_launches": int(launches)}
    del L, V, vtop
    ctx.barrier()
    return out


def load_reference():
    """`cola` = the unmodified reference (baseline/_ref on the GPU box, the live tree in the build container) or None."""
    try:
        from baseline.install_ref import import_reference
        return import_reference()
    except Exception:
        return None


def reference_sparse(cola, data, rows, cols, shape):
    """The reference's Sparse on (row, col)-sorted COO input.  Its constructor sorts with a non-stable argsort and can
    misalign values and indices; the intended CSR values are `data` itself, so the INSTANCE is
    repaired when that happened (the reference source is untouched)."""
    S = cola.ops.Sparse(data, rows, cols, shape)
    if not torch.equal(S.col_indices.to(torch.int32), S.A.col_indices()):
        S.data, S.row_indices, S.col_indices = data, rows, cols
        S.A = torch.sparse_csr_tensor(S.A.crow_indices(), S.A.col_indices(), data, size=shape)
    return S


def secondary_gpu_reference(ctx, cb, iters=10):
    """The unmodified reference's CG on torch-CUDA tensors of the same B200 (config 2): eager ATen kernels + cuSPARSE
    SpMM, two host syncs per iteration -- the 'reference backend on a GPU' bar of BASELINE.md."""
    if ctx.rank != 0:
        ctx.barrier()
        return None
    cola = load_reference()
    if cola is None:
        ctx.barrier()
        return {"error": "reference not installed (baseline/_ref)"}
    from cola.linalg.inverse.cg import CG
    dev = ctx.dev
    data, rows, cols, shape = laplacian_coo(GRID, torch.float32, "cpu")
    A = cola.PSD(reference_sparse(cola, data.to(dev), rows.to(dev), cols.to(dev), shape))
    B = rhs_block(shape[0], K_RHS, 0).to(dev)
    l0 = cb.backend.lib().launch_count()
    CG(tol=1e-30, max_iters=2)(A, B)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    x, info = CG(tol=1e-30, max_iters=iters)(A, B)
    torch.cuda.synchronize()
    s = time.perf_counter() - t0
    loop_s = info["iteration_time"] * info["iterations"]
    assert cb.backend.lib().launch_count() == l0                 # none of this repo's kernels on that path
    out = {"workload": f"cfg2 with the unmodified reference (cola.linalg CG, torch backend) on CUDA tensors, {iters} iterations",
           "iters_per_s": iters / loop_s, "ms_per_iter": loop_s / iters * 1e3, "wall_s_incl_setup": s,
           "impl": os.path.dirname(cola.__file__)}
    del A, B, x
    ctx.barrier()
    return out


# ----------------------------------------------------------------------------------------------- CPU arm
def reference_cpu_cg(g, k, iters, steps=1, warm=0, threads=None):
    """`steps` solves of `iters` CG iterations each with the unmodified reference on the host cores (the oracle port
    when the reference is not installed).  Returns (iterations, loop seconds, wall seconds, last trace, kind, cores)."""
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    data, rows, cols, shape = laplacian_coo(g, torch.float32, "cpu")
    B = rhs_block(shape[0], k, seed=0)
    cola = load_reference()
    if cola is not None:
        from cola.linalg.inverse.cg import CG
        A = cola.PSD(reference_sparse(cola, data, rows, cols, shape))
        solve = lambda: CG(tol=1e-30, max_iters=iters)(A, B)[1]     # noqa: E731
        kind = "reference"
    else:
        from oracle import krylov_oracle as ko
        A = ko.SparseOp(data, rows, cols, shape)
        solve = lambda: ko.cg(A, B, tol=1e-30, max_iters=iters)[3]  # noqa: E731
        kind = "port"
    for _ in range(warm):
        solve()
    its, loop_s = 0, 0.0
    t0 = time.perf_counter()
    for _ in range(steps):
        info = solve()
        # the reference's own clock: iteration_time = (loop wall) / (cond evaluations)  (torch_tqdm.py:38,50)
        loop_s += info["iteration_time"] * info["iterations"]
        its += info["iterations"] - 1
    wall = time.perf_counter() - t0
    return its, loop_s, wall, [float(v) for v in info["errors"]], kind, torch.get_num_threads()


def cpu_reference_rate(g, k, sample_iters):
    its, loop_s, wall, trace, kind, cores = reference_cpu_cg(g, k, sample_iters)
    what = "the unmodified reference (baseline/_ref, cola.linalg CG on torch CPU)" if kind == "reference" else \
        "the CPU oracle (oracle/krylov_oracle.py)"
    return ({"value": its / loop_s, "unit": "iterations/s", "cores": cores, "kind": kind,
             "sample": f"{its} CG iterations of the full workload (n={g * g}, {k} RHS) with {what}, torch {torch.__version__}, "
                       f"loop time {loop_s:.2f} s, wall incl. setup {wall:.2f} s"}, trace)


def parity_check(g, k, trace_ours, trace32, tol=1e-5):
    """First residual norms of the timed GPU solve against the oracle on the same operator and right-hand sides
    (rank 0's block, seed 0).  At this size the fp32 CPU trace is itself ~4e-4 from exact arithmetic (sequential fp32
    column sums over 4M rows; tests/test_gpu_fullscale.py), so the bar (1e-5) is applied against the oracle run in
    fp64 on the same fp32 inputs, and the fp32 CPU trace is reported with its own distance from that run."""
    from oracle import krylov_oracle as ko
    data, rows, cols, shape = laplacian_coo(g, torch.float32, "cpu")
    B = rhs_block(shape[0], k, seed=0)
    m = len(trace_ours)
    _, _, _, info64 = ko.cg(ko.SparseOp(data.double(), rows, cols, shape), B.double(), tol=1e-30, max_iters=m + 1)
    t64 = [float(v) for v in info64["errors"][:m]]
    t32 = [float(v) for v in trace32[:m]]
    rel = lambda a, b: max(abs(x - y) / abs(y) for x, y in zip(a, b))     # noqa: E731
    err64 = rel(trace_ours, t64)
    return {"what": f"info['errors'][:{m}] of the timed solve vs the CPU oracle (fp64 arithmetic, same fp32 inputs)",
            "rel_err": err64, "tol": tol, "ok": bool(err64 < tol),
            "ours": trace_ours, "oracle_fp64": t64, "cpu_fp32": t32,
            "rel_err_vs_cpu_fp32": rel(trace_ours, t32) if len(t32) == m else None,
            "cpu_fp32_own_distance_from_fp64": rel(t32, t64) if len(t32) == m else None}


def run_reference(args):
    """--impl reference: the unmodified reference's CG on the host cores (rank 0 only)."""
    world = int(os.environ.get("WORLD_SIZE", 1))
    if int(os.environ.get("RANK", 0)) != 0:
        return
    g = 256 if args.small else GRID
    k = K_RHS
    its, loop_s, wall, trace, kind, cores = reference_cpu_cg(g, k, REF_ITERS, steps=args.steps, warm=args.warmup)
    n = g * g
    nnz = 5 * n - 4 * g
    value = its / loop_s
    out = {
        "impl": "reference", "metric": "CG iters/s (4M-row CSR, 64 RHS)", "value": value, "unit": "iterations/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": loop_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": shared_config(g, n, nnz, k),
        "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": cores, "kind": kind,
                         "sample": f"{its} full-size CG iterations ({args.steps} steps x {REF_ITERS}), loop {loop_s:.1f} s, "
                                   f"wall {wall:.1f} s"},
        "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def ncu_traffic(kernel, g, k):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full`
    capture (profiles/r2_ncu_full_traffic.json, else round 1's); None when there is no capture for this kernel / size."""
    here = os.path.dirname(os.path.abspath(__file__))
    if not (g == 2048 and k == 64):
        return None
    for name in ("r2_ncu_full_traffic.json", "r1_ncu_full_traffic.json"):
        path = os.path.join(here, "profiles", name)
        if os.path.exists(path):
            rec = json.load(open(path)).get("kernels", {}).get(kernel)
            if rec is not None:
                return rec["traffic_bytes"]
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="all", choices=["all", "cfg2"],
                    help="all: config 2 (the headline line) + the secondary configs; cfg2: the headline line only")
    ap.add_argument("--small", action="store_true", help="256x256 grid (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
