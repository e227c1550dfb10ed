"""Locates and classifies wrong atoms of the fused tcgen05 Kronecker matmat (bring-up diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cola_b200 as cb

dev = torch.device("cuda:0")
torch.manual_seed(0)
reps = int(os.environ.get("DIAG_REPS", "12"))
cases = [(3, 128, True), (3, 128, False), (2, 128, True), (3, 64, True)]
for D, k, fused in cases:
    Fs = [torch.randn(64, 64, device=dev) / 8 + 0.5 * torch.eye(64, device=dev) for _ in range(D)]
    K = cb.ops.Kronecker(*[cb.ops.Dense(F) for F in Fs])
    n = 64**D
    X = torch.randn(n, k, device=dev)
    A = (K + 0.1 * cb.ops.I_like(K)) if fused else K
    E = X.double().reshape(*([64] * D), k)
    for i, F in enumerate(Fs):
        E = torch.moveaxis(torch.tensordot(F.double(), torch.moveaxis(E, i, 0), dims=1), 0, i)
    plain = E.reshape(n, k)
    ref = plain + (0.1 * X.double() if fused else 0)
    nbad_runs = 0
    for rep in range(reps):
        Y = torch.full_like(X, float("nan"))
        A.matmat_into(X, Y)
        torch.cuda.synchronize()
        d = (Y.double() - ref).abs()
        bad = (d > 1e-4 * ref.abs().max()) | ~torch.isfinite(Y)
        if not bad.any():
            continue
        nbad_runs += 1
        # atoms of the last mode: 64 consecutive rows x one 32-column block
        ba = bad.reshape(n // 64, 64, k // 32, 32).any(dim=3).any(dim=1).nonzero()
        msg = f"D={D} k={k} fused={fused} rep={rep}: bad elements {int(bad.sum())}, nan {int((~torch.isfinite(Y)).sum())}, atoms {ba.shape[0]}:"
        for p, cb_ in ba[:4].tolist():
            blk = Y[p * 64:(p + 1) * 64, cb_ * 32:(cb_ + 1) * 32].double()
            rblk = ref[p * 64:(p + 1) * 64, cb_ * 32:(cb_ + 1) * 32]
            pl = plain[p * 64:(p + 1) * 64, cb_ * 32:(cb_ + 1) * 32]
            xb = X[p * 64:(p + 1) * 64, cb_ * 32:(cb_ + 1) * 32].double()
            e_rel = float((blk - rblk).norm() / rblk.norm())
            # which hypothesis explains the block: plain (epilogue lost), another atom's result (stale operand / accumulator)
            e_plain = float((blk - pl).norm() / rblk.norm())
            refs = ref[:, cb_ * 32:(cb_ + 1) * 32].reshape(n // 64, 64, 32)
            dist = (refs - blk[None]).flatten(1).norm(dim=1) / rblk.norm()
            pbest = int(dist.argmin())
            # Kx of another atom with this atom's x (accumulator of another tile, own epilogue operand)
            pls = plain[:, cb_ * 32:(cb_ + 1) * 32].reshape(n // 64, 64, 32)
            dist2 = (pls + 0.1 * xb[None] * (1 if fused else 0) - blk[None]).flatten(1).norm(dim=1) / rblk.norm()
            pbest2 = int(dist2.argmin())
            rows_bad = bad[p * 64:(p + 1) * 64, cb_ * 32:(cb_ + 1) * 32].any(dim=1).sum().item()
            cols_bad = bad[p * 64:(p + 1) * 64, cb_ * 32:(cb_ + 1) * 32].any(dim=0).sum().item()
            msg += (f"\n    atom p={p} ({p // 64},{p % 64}) colblock {cb_}: rel err {e_rel:.2e}, rows {rows_bad} cols {cols_bad}, vs plain {e_plain:.2e}, "
                    f"nearest ref atom {pbest} ({float(dist[pbest]):.2e}), nearest Kx-of-other-atom {pbest2} ({float(dist2[pbest2]):.2e})")
        print(msg, flush=True)
    print(f"D={D} k={k} fused={fused}: {nbad_runs}/{reps} runs with wrong atoms", flush=True)
