import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cola_b200 as cb
from cola_b200 import backend as be
dev = torch.device("cuda:0")
torch.set_printoptions(linewidth=200, precision=3, sci_mode=False)
D, k = 2, 32
n = 64**D
I = torch.eye(64, device=dev).contiguous()
F2 = (2 * torch.eye(64, device=dev)).contiguous()
X = torch.arange(n * k, device=dev, dtype=torch.float32).reshape(n, k) / 1000.0
Y = torch.full_like(X, 7.0)
ws = torch.full((2 * n * 32,), 5.0, device=dev)
facs = (ctypes.c_void_p * D)(I.data_ptr(), F2.data_ptr())
ldf = (ctypes.c_int64 * D)(64, 64)
be.lib().call("cola_kron_matmat_tc_f32", D, facs, ldf, be.ptr(X), be.ptr(Y), k, be.ptr(ws), ctypes.c_float(1.0),
              ctypes.c_float(0.0), None, 0, None, None, None, be.stream_ptr())
torch.cuda.synchronize()
print("Y sentinel left:", int((Y == 7.0).sum()), "of", Y.numel(), " zeros:", int((Y == 0).sum()))
print("ws0 sentinel left:", int((ws[:n*32] == 5.0).sum()), " zeros:", int((ws[:n*32] == 0).sum()))
w0 = ws[:n * 32].reshape(n, 32)
print("ws0[0:2,:6]", w0[0:2, :6], "X[0:2,:6]", X[0:2, :6])
print("Y[0:2,:6]", Y[0:2, :6])
nz = (w0 != 0) & (w0 != 5.0)
print("ws0 nonzero count", int(nz.sum()))
if int(nz.sum()) > 0:
    idx = nz.nonzero()[:10]
    print(idx, w0[nz][:10])
