import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cola_b200 as cb
dev = torch.device("cuda:0")
D, k = 3, 128
Fs = [torch.randn(64, 64, device=dev) / 8 + 0.5 * torch.eye(64, device=dev) for _ in range(D)]
K = cb.ops.Kronecker(*[cb.ops.Dense(F) for F in Fs])
X = torch.randn(64**3, k, device=dev); Y = torch.empty_like(X)
for _ in range(4):
    K.matmat_into(X, Y)
torch.cuda.synchronize()
