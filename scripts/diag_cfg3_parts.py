"""cfg3 CG iteration, piece by piece inside CUDA graphs: where do the microseconds between the kernels go?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cola_b200 as cb
from cola_b200 import backend as be

dev = torch.device("cuda:0")
def factor(d, seed):
    gg = torch.Generator().manual_seed(seed)
    G = torch.randn(d, d, generator=gg)
    return (G @ G.T / d + 0.5 * torch.eye(d)).to(dev)
Fs = [factor(64, i) for i in range(3)]
K = cb.ops.Kronecker(*[cb.PSD(cb.ops.Dense(F)) for F in Fs])
A = cb.PSD(K + 0.1 * cb.ops.I_like(K))
A.plan()
n, k = 64 ** 3, int(os.environ.get("K", 128))
dt, sx = torch.float32, "f32"
g = torch.Generator().manual_seed(0)
max_iters = 100000
b = torch.randn(n, k, generator=g).to(dev)
r = b / b.norm(dim=0); x = torch.zeros_like(b); p = r.clone(); ap = torch.empty_like(b)
gamma = torch.zeros((max_iters + 2, k), dtype=torch.float64, device=dev)
pap = torch.zeros((max_iters + 1, k), dtype=torch.float64, device=dev)
tol_eff = torch.full((k, ), -1.0, dtype=dt, device=dev)
ctl = be.small_ints([0, 0, max_iters, k], dev)
be.col_dots(r, r, gamma[0])
lib, st = be.lib(), be.stream_ptr
it_ptr, done_ptr = ctl[0:1], ctl[1:2]

def mm(): A.matmat_into(p, ap, dots=pap, dots_row=it_ptr, gate=done_ptr)
def mm_plain(): K.matmat_into(p, ap)
def mm_shift(): A.matmat_into(p, ap)
def upd_r(): lib.call(f"cola_cg_update_r_{sx}", be.ptr(r), be.ptr(ap), n, k, k, be.ptr(ctl), be.ptr(gamma), be.ptr(pap), be.ptr(gamma), st())
def upd_xp(): lib.call(f"cola_cg_update_xp_{sx}", be.ptr(x), be.ptr(r), be.ptr(p), n, k, k, be.ptr(ctl), be.ptr(gamma), be.ptr(pap), st())
def adv(): lib.call(f"cola_cg_advance_{sx}", be.ptr(ctl), be.ptr(gamma), be.ptr(tol_eff), 1, st())
def full(): mm(); upd_r(); upd_xp(); adv()
def no_adv(): mm(); upd_r(); upd_xp()
def sweeps(): upd_r(); upd_xp()

NB = 16
def graph_time(fn, name, reps=6):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            for _ in range(NB): fn()
        gr.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps): gr.replay()
        e1.record(s); torch.cuda.synchronize()
        tg = e0.elapsed_time(e1) / (reps * NB) * 1e3
        e0.record(s)
        for _ in range(reps * NB): fn()
        e1.record(s); torch.cuda.synchronize()
        ts = e0.elapsed_time(e1) / (reps * NB) * 1e3
    print(f"{name:28s} graph {tg:7.1f} us   stream {ts:7.1f} us", flush=True)
    return tg

t = {}
for name, fn in (("matmat+shift+dots", mm), ("matmat+shift", mm_shift), ("matmat plain", mm_plain), ("update_r", upd_r), ("update_xp", upd_xp),
                 ("advance", adv), ("r + xp", sweeps), ("matmat + r + xp", no_adv), ("full iteration", full)):
    t[name] = graph_time(fn, name)
print("sum of parts", t["matmat+shift+dots"] + t["update_r"] + t["update_xp"] + t["advance"], "full", t["full iteration"])
