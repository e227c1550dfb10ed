"""Diagnostic: where does the time between CG kernels go? (not part of the product)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cola_b200 as cb
import cola_b200.linalg.cg as cgmod
from bench import laplacian_coo, rhs_block

dev = torch.device("cuda:0")
data, rows, cols, shape = laplacian_coo(2048, torch.float32, dev)
A = cb.PSD(cb.ops.Sparse(data, rows, cols, shape))
B = rhs_block(shape[0], 64, 0).to(dev)
for ce in (16, 50, 1000):
    cgmod.CHECK_EVERY = ce
    alg = cb.linalg.CG(tol=1e-30, max_iters=50)
    for _ in range(2):
        alg(A, B)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(4):
        alg(A, B)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"CHECK_EVERY={ce}: {e0.elapsed_time(e1)/200:.3f} ms/iter (host enqueue+poll wall {(t1-t0)*1e3/200:.3f} ms/iter)")

# raw enqueue cost of one iteration's launches (no sync)
import cola_b200.backend as be
lib = be.lib()
n, k = B.shape
x = torch.zeros_like(B); r = B.clone(); p = B.clone(); ap = torch.empty_like(B)
gamma = torch.ones((1002, k), dtype=torch.float64, device=dev); pap = torch.ones((1001, k), dtype=torch.float64, device=dev)
tol_eff = torch.zeros(k, device=dev)
ctl = torch.tensor([0, 0, 1000, k], dtype=torch.int32, device=dev)
itp, dnp = ctl[0:1], ctl[1:2]
def one():
    A.matmat_into(p, ap, dots=pap, dots_row=itp, gate=dnp)
    lib.call("cola_cg_update_xr_f32", be.ptr(x), be.ptr(r), be.ptr(p), be.ptr(ap), n, k, k, be.ptr(ctl), be.ptr(gamma), be.ptr(pap), be.ptr(gamma), be.stream_ptr())
    lib.call("cola_cg_update_p_f32", be.ptr(r), be.ptr(p), n, k, k, be.ptr(ctl), be.ptr(gamma), be.stream_ptr())
    lib.call("cola_cg_advance_f32", be.ptr(ctl), be.ptr(gamma), be.ptr(tol_eff), 1, be.stream_ptr())
one(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): one()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"enqueue {1e3*(t1-t0)/20:.3f} ms/iter, drain total {1e3*(t2-t0)/20:.3f} ms/iter")
# graph capture of 10 iterations
ctl.copy_(torch.tensor([0, 0, 1000, k], dtype=torch.int32))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    one()
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        for _ in range(10): one()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): g.replay()
e1.record(); torch.cuda.synchronize()
print(f"graph replay: {e0.elapsed_time(e1)/50:.3f} ms/iter; ctl={ctl.tolist()}")
