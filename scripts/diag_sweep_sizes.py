"""CG sweeps (update_r: 3 passes, update_xp: 5 passes) against the block size: where does the small-block penalty come from?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cola_b200 as cb
from cola_b200 import backend as be
dev = torch.device("cuda:0")
dt, sx = torch.float32, "f32"
lib, st = be.lib(), be.stream_ptr
k = int(os.environ.get("K", 128))
for logn in [int(v) for v in os.environ.get("LOGN", "16,17,18,19,20,21").split(",")]:
    n = 1 << logn
    r = torch.randn(n, k, device=dev); ap = torch.randn(n, k, device=dev); x = torch.zeros(n, k, device=dev); p = torch.randn(n, k, device=dev)
    gamma = torch.ones((1000, k), dtype=torch.float64, device=dev); pap = torch.ones((1000, k), dtype=torch.float64, device=dev)
    ctl = be.small_ints([0, 0, 900, k], dev)
    def upd_r(): lib.call(f"cola_cg_update_r_{sx}", be.ptr(r), be.ptr(ap), n, k, k, be.ptr(ctl), be.ptr(gamma), be.ptr(pap), be.ptr(gamma), st())
    def upd_xp(): lib.call(f"cola_cg_update_xp_{sx}", be.ptr(x), be.ptr(r), be.ptr(p), n, k, k, be.ptr(ctl), be.ptr(gamma), be.ptr(pap), st())
    def dots(): be.col_dots(r, ap, gamma[5])
    def axpby(): be.axpby(r, ap, 0.5, 0.5)
    out = []
    for name, fn, passes in (("update_r", upd_r, 3), ("update_xp", upd_xp, 5), ("col_dots", dots, 2), ("axpby", axpby, 3)):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 40
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        out.append(f"{name} {us:7.1f} us {passes * n * k * 4 / us * 1e-3:6.0f} GB/s")
    print(f"n=2^{logn} ({n*k*4>>20} MB blocks), CTAs/SM={os.environ.get('COLA_SWEEP_CTAS','8')}: " + " | ".join(out), flush=True)
