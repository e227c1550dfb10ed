import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cola_b200 as cb
dev = torch.device("cuda:0")
torch.set_printoptions(linewidth=200, precision=3, sci_mode=False)
D, k = 2, 32
n = 64**D
def run(Fs, X):
    K = cb.ops.Kronecker(*[cb.ops.Dense(F.contiguous()) for F in Fs])
    Y = torch.empty_like(X)
    K.matmat_into(X, Y)
    torch.cuda.synchronize()
    return Y
def ref(Fs, X):
    E = X.double().reshape(*([64] * D), X.shape[1])
    for i, F in enumerate(Fs):
        E = torch.moveaxis(torch.tensordot(F.double(), torch.moveaxis(E, i, 0), dims=1), 0, i)
    return E.reshape(n, X.shape[1])
I = torch.eye(64, device=dev)
# X[(i1,i2), r] = 1000*i1 + 10*i2 + 0.01*r  -> identifies any permutation
i1 = torch.arange(64, device=dev).repeat_interleave(64).float(); i2 = torch.arange(64, device=dev).repeat(64).float()
X = (100 * i1 + i2)[:, None] + 0.01 * torch.arange(k, device=dev).float()[None, :]
Y = run([I, I], X)
print("identity: max err", float((Y - X).abs().max()))
print("Y[0:3, :8]", Y[0:3, :8]); print("Y[64:66, :8]", Y[64:66, :8]); print("Y[4095, :8]", Y[4095, :8])
# shift matrix S: (S x)[a] = x[a-1]
S = torch.zeros(64, 64, device=dev); S[torch.arange(1, 64), torch.arange(0, 63)] = 1.0
Y = run([I, S], X); R = ref([I, S], X)
print("I (x) S: max err", float((Y.double() - R).abs().max())); print(Y[0:4, :4], R[0:4, :4])
Y = run([S, I], X); R = ref([S, I], X)
print("S (x) I: max err", float((Y.double() - R).abs().max())); print(Y[62:67, :4], R[62:67, :4])
torch.manual_seed(0)
F = torch.randn(64, 64, device=dev)
Xr = torch.randn(n, k, device=dev)
Y = run([I, F], Xr); R = ref([I, F], Xr)
print("I (x) F rel err", float((Y.double() - R).norm() / R.norm()))
Y = run([F, I], Xr); R = ref([F, I], Xr)
print("F (x) I rel err", float((Y.double() - R).norm() / R.norm()))
Y = run([I, F.T], Xr); R = ref([I, F], Xr)
print("I (x) F^T vs F ref rel err", float((Y.double() - R).norm() / R.norm()))
