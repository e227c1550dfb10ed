"""Why does the cfg3 CG iteration time vary between boxes?  Per-solve timing, graph captures, host time per batch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cola_b200 as cb
from cola_b200.linalg import cg as cgmod
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
def factor(d, seed):
    gg = torch.Generator().manual_seed(seed)
    G = torch.randn(d, d, generator=gg)
    return (G @ G.T / d + 0.5 * torch.eye(d)).to(dev)
Fs = [factor(64, i) for i in range(3)]
K = cb.ops.Kronecker(*[cb.PSD(cb.ops.Dense(F)) for F in Fs])
A = cb.PSD(K + 0.1 * cb.ops.I_like(K))
n, k = 64**3, 128
B = torch.randn(n, k, generator=g).to(dev)
captures = [0]
orig = torch.cuda.CUDAGraph
class Counting(orig):
    def __new__(cls, *a, **kw):
        captures[0] += 1
        return orig.__new__(cls, *a, **kw)
torch.cuda.CUDAGraph = Counting
polls = [0, 0.0]
orig_read = cb.backend.read_small
def timed_read(t):
    t0 = time.perf_counter(); r = orig_read(t); polls[0] += 1; polls[1] += time.perf_counter() - t0; return r
cb.backend.read_small = timed_read
alg = cb.linalg.CG(tol=1e-30, max_iters=100)
for i in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); x, info = alg(A, B); e1.record(); torch.cuda.synchronize()
    ws = A.__dict__.get("_cg_workspace")
    print(f"solve {i}: wall {1e3*(time.perf_counter()-t0):.2f} ms, device {e0.elapsed_time(e1):.2f} ms, graph captures so far {captures[0]}, "
          f"polls {polls[0]} ({1e3*polls[1]:.2f} ms in polls), ws graph {'yes' if ws and ws['graph'] is not None else 'no'}", flush=True)
    polls[0], polls[1] = 0, 0.0
