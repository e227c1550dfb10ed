"""Backward-pass cases (cg_bwd cola/linalg/inverse/cg.py:72-86, slq_bwd cola/linalg/tbd/slq.py:10-31).

A case = named parameter tensors (the leaves that require grad), a `spec(params)` in the tuple language of
tests/problems.py (so the same case builds a reference operator, an oracle operator and a cola_b200 operator), a
right-hand-side block B and loss weights W:  loss = sum(W * solve(A(theta), B))  for the CG cases, the SLQ estimate
of logdet itself for the SLQ cases.  tests/golden/make_golden_bwd.py records the REFERENCE's gradients
(tests/golden/bwd_*.npz); the oracle, the host logic and the GPU kernels are tested against them."""
import numpy as np
import torch

from tests import problems as pb


def _spd(n, dtype, seed):
    return pb.spd_dense(n, dtype, seed, lo=0.5)


def case(name):
    dt = torch.float32 if name.endswith("f32") else torch.float64
    if name.startswith("bwd_cg_diag"):            # the reference's own test (tests/algorithms/test_cg.py:20-62), wider
        n = 32
        d = pb.t(pb.rs(1).uniform(1.0, 4.0, size=n), dt)
        return dict(kind="cg", params={"diag": d}, spec=lambda p: ("diag", p["diag"]), ann="psd",
                    B=pb.randn_np((n, 3), dt, 2), W=pb.randn_np((n, 3), dt, 3), tol=1e-10, max_iters=200)
    if name.startswith("bwd_cg_dense"):
        n = 24
        return dict(kind="cg", params={"M": _spd(n, dt, 4)}, spec=lambda p: ("dense", p["M"]), ann="psd",
                    B=pb.randn_np((n, 4), dt, 5), W=pb.randn_np((n, 4), dt, 6), tol=1e-10, max_iters=300)
    if name.startswith("bwd_cg_csr_shift"):       # CSR values + c*I + Diagonal: the cfg2-shaped composition
        g = 6
        data, rows, cols, shape = pb.laplacian_2d_coo(g, dt)
        n = g * g
        d = pb.t(pb.rs(7).uniform(0.5, 1.5, size=n), dt)
        return dict(kind="cg", params={"vals": data, "diag": d},
                    spec=lambda p: ("sum", [("csr", p["vals"], rows, cols, shape), ("scaled_identity", 0.3, n), ("diag", p["diag"])]),
                    ann="psd", B=pb.randn_np((n, 5), dt, 8), W=pb.randn_np((n, 5), dt, 9), tol=1e-10, max_iters=300)
    if name.startswith("bwd_cg_kron"):            # Kronecker factors + Diagonal: the cfg3 / cfg4-shaped composition
        dims = (4, 3, 5)
        Fs = {f"F{i}": _spd(d, dt, 10 + i) for i, d in enumerate(dims)}
        n = int(np.prod(dims))
        dg = pb.t(pb.rs(14).uniform(0.5, 1.5, size=n), dt)
        return dict(kind="cg", params={**Fs, "diag": dg},
                    spec=lambda p: ("sum", [("kron", [("dense", p[f"F{i}"]) for i in range(len(dims))]), ("diag", p["diag"])]),
                    ann="psd", B=pb.randn_np((n, 3), dt, 15), W=pb.randn_np((n, 3), dt, 16), tol=1e-10, max_iters=400)
    if name.startswith("bwd_cg_product"):         # Product[Dense, Dense] + c*I  (L L^T + c I)
        n = 20
        Lm = pb.t(np.tril(pb.rs(17).normal(size=(n, n))) / np.sqrt(n) + np.eye(n), dt)
        return dict(kind="cg", params={"L": Lm},
                    spec=lambda p: ("sum", [("product", [("dense", p["L"]), ("dense", p["L"].T.contiguous())]), ("scaled_identity", 0.5, n)]),
                    ann="psd", B=pb.randn_np((n, 2), dt, 18), W=pb.randn_np((n, 2), dt, 19), tol=1e-10, max_iters=300)
    if name.startswith("bwd_cg_blockdiag"):
        b1, b2 = _spd(5, dt, 20), _spd(7, dt, 21)
        n = 5 * 2 + 7
        return dict(kind="cg", params={"b1": b1, "b2": b2},
                    spec=lambda p: ("blockdiag", [("dense", p["b1"]), ("dense", p["b2"])], [2, 1]),
                    ann="psd", B=pb.randn_np((n, 3), dt, 22), W=pb.randn_np((n, 3), dt, 23), tol=1e-10, max_iters=300)
    if name.startswith("bwd_slq_dense"):
        n = 30
        return dict(kind="slq", params={"M": _spd(n, dt, 24)}, spec=lambda p: ("dense", p["M"]), ann="psd",
                    max_iters=30, tol=1e-7, vtol=0.25, key=5)
    if name.startswith("bwd_slq_csr_diag"):
        g = 5
        data, rows, cols, shape = pb.laplacian_2d_coo(g, dt)
        n = g * g
        d = pb.t(pb.rs(25).uniform(0.5, 1.5, size=n), dt)
        return dict(kind="slq", params={"vals": data, "diag": d},
                    spec=lambda p: ("sum", [("csr", p["vals"], rows, cols, shape), ("diag", p["diag"])]), ann="psd",
                    max_iters=25, tol=1e-7, vtol=0.25, key=11)
    raise KeyError(name)


BWD_CASES = ["bwd_cg_diag_f64", "bwd_cg_dense_f64", "bwd_cg_dense_f32", "bwd_cg_csr_shift_f64", "bwd_cg_kron_f64",
             "bwd_cg_kron_f32", "bwd_cg_product_f64", "bwd_cg_blockdiag_f64", "bwd_slq_dense_f64", "bwd_slq_csr_diag_f64"]
