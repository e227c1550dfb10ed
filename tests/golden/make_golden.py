"""Generate golden vectors from the REAL reference (wilson-labs/cola at /root/reference).

Runs only in the build container (the reference does not travel to the GPU box).
It imports the unmodified reference through two import shims for packages missing
from this image (tests/golden/refshim/{plum,optree}); nothing is written to
/root/reference.  Output: tests/golden/*.npz + MANIFEST.json, committed.

    python tests/golden/make_golden.py

The fixtures hold OUTPUTS only; inputs are regenerated from seeds by tests/problems.py
(an input checksum is stored so drift is detected).
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(HERE, "refshim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import cola  # noqa: E402  (the reference)
from cola.linalg.decompositions.decompositions import Arnoldi, Lanczos  # noqa: E402
from cola.linalg.inverse.cg import CG  # noqa: E402
from cola.linalg.inverse.gmres import GMRES  # noqa: E402
from cola.linalg.tbd.slq import stochastic_lanczos_quad  # noqa: E402
from cola.linalg.trace.diagonal_estimation import Hutch  # noqa: E402
from cola.linalg.unary.unary import LanczosUnary  # noqa: E402
from cola.ops import operators as rops  # noqa: E402

from tests import problems as pb  # noqa: E402
from tests.golden_cases import (ARNOLDI_CASES, CG_CASES, DIAG_CASES, GMRES_CASES, LANCZOS_CASES,  # noqa: E402
                                MATMAT_PROBLEMS, NEXT_CG_CASES, NEXT_MATMAT_PROBLEMS, PCG_CASES, POWER_CASES,
                                UNARY_CASES)

assert cola.__file__.startswith("/root/reference"), cola.__file__


def aligned_sparse(data, rows, cols, shape):
    """Reference defect workaround: Sparse.__init__ (operators.py:64-75) sorts with a non-stable
    `argsort(row_indices)`; when that permutes entries inside a row, its `data` no longer lines up with
    scipy's canonical `indices`.  Inputs here are pre-sorted by (row, col), so the intended CSR values are
    `data` itself: put them back (instance attributes only; the reference source is untouched)."""
    S = rops.Sparse(data, rows, cols, shape)
    if not torch.equal(S.col_indices.to(torch.int32), S.A.col_indices()):
        REPAIRED.append(tuple(shape))
        S.data, S.row_indices, S.col_indices = data, rows, cols
        S.A = torch.sparse_csr_tensor(S.A.crow_indices(), S.A.col_indices(), data, size=shape)
    return S


REPAIRED = []


def to_reference(spec, ann=None):
    def rec(s, dtype_hint=None):
        k = s[0]
        if k == "dense":
            return rops.Dense(s[1])
        if k == "csr":
            return aligned_sparse(s[1], s[2], s[3], s[4])
        if k == "diag":
            return rops.Diagonal(s[1])
        if k == "scaled_identity":
            return s[1] * rops.Identity((s[2], s[2]), dtype_hint)
        if k == "scale":
            return s[1] * rec(s[2])
        if k == "kron":
            return rops.Kronecker(*[rec(x) for x in s[1]])
        if k == "blockdiag":
            return rops.BlockDiag(*[rec(x) for x in s[1]], multiplicities=s[2])
        if k == "product":
            return rops.Product(*[rec(x) for x in s[1]])
        if k == "psd":
            return cola.PSD(rec(s[1]))
        if k == "kronsum":
            return rops.KronSum(*[rec(x) for x in s[1]])
        if k == "tridiag":
            return rops.Tridiagonal(*s[1:])
        if k == "sum":
            first = rec(s[1][0])
            out = first
            for x in s[1][1:]:
                out = out + rec(x, first.dtype)
            return out
        raise KeyError(k)

    A = rec(spec)
    if ann == "psd":
        A = cola.PSD(A)
    elif ann == "sa":
        A = cola.SelfAdjoint(A)
    return A


def checksum(spec, B):
    h = hashlib.sha256()

    def rec(s):
        for x in s:
            if torch.is_tensor(x):
                h.update(x.numpy().tobytes())
            elif isinstance(x, (list, tuple)):
                rec(x)
            else:
                h.update(repr(x).encode())

    rec(spec)
    h.update(B.numpy().tobytes())
    return h.hexdigest()[:16]


OUT = {}
MANIFEST = {"reference": "wilson-labs/cola @ /root/reference", "torch": torch.__version__, "cases": {}}


def save(case, **arrays):
    arrays = {k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items()}
    np.savez_compressed(os.path.join(HERE, case + ".npz"), **arrays)
    MANIFEST["cases"][case] = {k: list(v.shape) for k, v in arrays.items()}


# --------------------------------------------------------------------------- operator matmats

def gen_matmat():
    for name in MATMAT_PROBLEMS:
        P = pb.problem(name)
        A = to_reference(P["spec"], P["ann"])
        X = pb.randn_np((A.shape[1], 6), P["dtype"], 100)
        save("matmat_" + name, Y=A @ X, y=A @ X[:, 0].contiguous(), sum=checksum(P["spec"], P["B"]))


# --------------------------------------------------------------------------- CG

def gen_cg():
    for case, (name, tol, iters) in CG_CASES.items():
        P = pb.problem(name)
        A = to_reference(P["spec"], P["ann"])
        x, info = CG(tol=tol, max_iters=iters)(A, P["B"])
        save(case, x=x, errors=info["errors"], iterations=info["iterations"], sum=checksum(P["spec"], P["B"]))
    # through the dispatch surface: solve() with an explicit CG (inv.py:23-39,66-69)
    P = pb.problem("dense96_f32")
    A = to_reference(P["spec"], P["ann"])
    x = cola.linalg.solve(A, P["B"], CG(tol=1e-6, max_iters=500))
    save("solve_dense96_f32", x=x)
    # x0 given
    x0 = pb.randn_np(tuple(P["B"].shape), P["dtype"], 77)
    x, info = CG(tol=1e-6, max_iters=500, x0=x0)(A, P["B"])
    save("cg_dense96_f32_x0", x=x, errors=info["errors"], iterations=info["iterations"])


# --------------------------------------------------------------------------- Lanczos

def gen_lanczos():
    for case, (name, m, tol, batched) in LANCZOS_CASES.items():
        P = pb.problem(name)
        A = to_reference(P["spec"], P["ann"])
        start = P["B"] if (batched or P["B"].dim() == 1) else P["B"][:, 0].contiguous()
        Q, T, info = Lanczos(start_vector=start, max_iters=m, tol=tol)(A)
        if start.dim() == 1:
            Qd = Q.to_dense()
            save(case, Q=Qd[::16] if Qd.shape[0] > 1000 else Qd, alpha=T.alpha[:, 0], beta=T.beta[:, 0],
                 errors=info["errors"], iterations=info["iterations"])
        else:
            save(case, Q=Q.A.A if hasattr(Q.A, "A") else Q.A, alpha=T.alpha[..., 0], beta=T.beta[..., 0],
                 errors=info["errors"], iterations=info["iterations"])
    # default start vector: randn(n, key=PRNGKey(42))  (lanczos.py:209-213) + eig() LM slice (eigs.py:106-111)
    P = pb.problem("graph2k_f64")
    A = to_reference(P["spec"], P["ann"])
    vals, vecs = cola.linalg.eig(A, 6, "LM", Lanczos(max_iters=48, tol=1e-12))
    save("eig_graph2k_f64_default_start", eigvals=vals, eigvecs=vecs.to_dense()[::16])
    # early-termination known answers restated from the reference's tests (tests/algorithms/test_lanczos.py:268-300)
    A = rops.Dense(torch.diag(torch.tensor([4., 2., 1.])))
    Q, T, info = Lanczos(start_vector=torch.tensor([[1.0, 0.0, 0.0]]).T, max_iters=3, tol=1e-7)(cola.SelfAdjoint(A))
    save("lanczos_case_early", alpha=T.alpha[..., 0], beta=T.beta[..., 0], iterations=info["iterations"])


# --------------------------------------------------------------------------- Arnoldi

def gen_arnoldi():
    for case, (name, m, tol, batched) in ARNOLDI_CASES.items():
        P = pb.problem(name)
        A = to_reference(P["spec"], P["ann"])
        start = P["B"] if batched else P["B"][:, 0].contiguous()
        Q, H, info = Arnoldi(start_vector=start, max_iters=m, tol=tol)(A)
        if batched:
            save(case, Q=Q.A.A if hasattr(Q.A, "A") else Q.A, H=H.A, errors=info["errors"],
                 iterations=info["iterations"])
        else:
            save(case, Q=Q.to_dense(), H=H.to_dense(), errors=info["errors"], iterations=info["iterations"])
    P = pb.problem("nonsym48_f64")
    A = to_reference(P["spec"], P["ann"])
    vals, vecs = cola.linalg.eig(A, 48, "LM", Arnoldi(start_vector=P["B"][:, 0].contiguous(), max_iters=48, tol=1e-12))
    order = np.argsort(np.abs(vals.numpy()))
    save("eig_arnoldi_nonsym48_f64", eigvals_sorted_abs=np.abs(vals.numpy())[order])


# --------------------------------------------------------------------------- GMRES
def gen_gmres():
    for case, (name, m, tol, vec) in GMRES_CASES.items():
        P = pb.problem(name)
        A = to_reference(P["spec"], P["ann"])
        b = P["B"][:, 0].contiguous() if vec else P["B"]
        x, info = GMRES(tol=tol, max_iters=m)(A, b)
        save(case, x=x, errors=info["errors"], iterations=info["iterations"])
    # with an initial guess, and through solve() (inv.py:23-39)
    P = pb.problem("nonsym48_f64")
    A = to_reference(P["spec"], P["ann"])
    x0 = pb.randn_np(tuple(P["B"].shape), P["dtype"], 78)
    # 20 steps: the full-space run (48 steps on n = 48) squares an ill-conditioned H and is not a usable pin
    x, info = GMRES(tol=1e-12, max_iters=20, x0=x0)(A, P["B"])
    save("gmres_nonsym48_f64_x0", x=x, iterations=info["iterations"])
    save("solve_gmres_nonsym48_f64", x=cola.linalg.solve(A, P["B"], GMRES(tol=1e-12, max_iters=20)))


# --------------------------------------------------------------------------- preconditioned CG
def gen_pcg():
    from cola.linalg.preconditioning.preconditioners import NystromPrecond
    for case, (name, rank, tol, iters) in PCG_CASES.items():
        P_ = pb.problem(name)
        A = to_reference(P_["spec"], P_["ann"])
        Nys = NystromPrecond(A, rank=rank, key=A.xnp.PRNGKey(3))
        x, info = CG(tol=tol, max_iters=iters, P=Nys)(A, P_["B"])
        save(case, x=x, errors=info["errors"], iterations=info["iterations"], PB=Nys @ P_["B"], Lambda=Nys.Lambda)


# --------------------------------------------------------------------------- power iteration
def gen_power():
    from cola.linalg.eig.power_iteration import PowerIteration
    for case, (name, tol, iters) in POWER_CASES.items():
        P_ = pb.problem(name)
        A = to_reference(P_["spec"], P_["ann"])
        v, emax, info = PowerIteration(tol=tol, max_iter=iters, key=A.xnp.PRNGKey(11))(A)
        save(case, v=v, eigmax=emax, errors=info["errors"], iterations=info["iterations"])
    P_ = pb.problem("dense96_f64")
    A = to_reference(P_["spec"], P_["ann"])
    save("eigmax_dense96_f64", eigmax=cola.linalg.eigmax(A, PowerIteration(tol=1e-9, max_iter=400, key=A.xnp.PRNGKey(11))))


# --------------------------------------------------------------------------- SLQ / Hutch / f(A)v
def gen_stochastic():
    for name, m, vtol in [("kron884_diag_f32", 25, 0.25), ("kron465_diag_f64", 30, 0.2), ("lap24_f64", 40, 0.25)]:
        P = pb.problem(name)
        A = to_reference(P["spec"], P["ann"])
        key = A.xnp.PRNGKey(42)
        val = stochastic_lanczos_quad(A, torch.log, max_iters=m, tol=1e-7, vtol=vtol, key=key)
        dense = torch.linalg.slogdet(A.to_dense().double())[1]
        save("slq_" + name, logdet=val, dense_logdet=dense, key=key, num_samples=max(int(1 / vtol**2), 1))
    for name, m in [("kron884_diag_f32", 25), ("kron465_diag_f64", 30)]:
        P = pb.problem(name)
        A = to_reference(P["spec"], P["ann"])
        # f(A) V  (unary.py:46-60)
        F = LanczosUnary(A, torch.log, max_iters=m, tol=1e-7)
        save("logA_matmat_" + name, Y=F @ P["B"])
        # logdet(A, Lanczos, Hutch): hutchinson_diag_estimate over LanczosUnary (logdet.py:111-117)
        key = A.xnp.PRNGKey(42)
        val = cola.linalg.logdet(A, Lanczos(max_iters=m, tol=1e-7), Hutch(tol=2e-2, max_iters=3, key=key))
        F = LanczosUnary(A, torch.log, max_iters=m, tol=1e-7)
        dg = Hutch(tol=2e-2, max_iters=3, key=key)(F, 0)
        save("hutch_logdet_" + name, logdet=val, diag=dg, key=key)
    # plain Hutchinson diagonal of an explicit operator, rademacher probes
    P = pb.problem("dense96_f64")
    A = to_reference(P["spec"], P["ann"])
    dg = Hutch(tol=5e-2, max_iters=4, rand="rademacher", key=A.xnp.PRNGKey(7))(A, 0)
    save("hutch_diag_dense96_f64", diag=dg)


# --------------------------------------------------------------------------- SURVEY 8f items 3-4
def gen_next():
    from cola.linalg.trace.diagonal_estimation import Exact
    for name in NEXT_MATMAT_PROBLEMS:
        P = pb.problem(name)
        A = to_reference(P["spec"], P["ann"])
        X = pb.randn_np((A.shape[1], 6), P["dtype"], 100)
        save("matmat_" + name, Y=A @ X, y=A @ X[:, 0].contiguous(), sum=checksum(P["spec"], P["B"]))
    for case, (name, tol, iters) in NEXT_CG_CASES.items():
        P = pb.problem(name)
        A = to_reference(P["spec"], P["ann"])
        x, info = CG(tol=tol, max_iters=iters)(A, P["B"])
        save(case, x=x, errors=info["errors"], iterations=info["iterations"], sum=checksum(P["spec"], P["B"]))
    for case, (name, fn, alg, m, tol) in UNARY_CASES.items():
        P = pb.problem(name)
        A = to_reference(P["spec"], P["ann"])
        alg = (Arnoldi if alg == "arnoldi" else Lanczos)(max_iters=m, tol=tol)
        F = getattr(cola.linalg, fn)(A, alg)
        Y = F @ P["B"]
        save(case, Y=Y, kind=type(F).__name__.split("[")[0])
    for case, (name, k, alg) in DIAG_CASES.items():
        P = pb.problem(name)
        A = to_reference(P["spec"], P["ann"])
        alg = Exact() if alg == "exact" else Hutch(tol=2e-2, max_iters=4, key=A.xnp.PRNGKey(9))
        save(case, diag=cola.linalg.diag(A, k, alg), dense_diag=torch.diagonal(A.to_dense(), offset=k))
    # the Lanczos | Arnoldi rule of slogdet returns (tr/|tr|, |tr|)  (logdet.py:111-117): an operator with det < 1
    P = pb.problem("dense96_f64")
    A = to_reference(P["spec"], P["ann"])
    sign, mag = cola.linalg.slogdet(A, Lanczos(max_iters=40, tol=1e-12), Hutch(tol=2e-2, max_iters=2, key=A.xnp.PRNGKey(42)))
    save("slogdet_lanczos_dense96_f64", sign=sign, logdet=mag, dense_logdet=torch.linalg.slogdet(A.to_dense())[1])


GENERATORS = dict(matmat=gen_matmat, cg=gen_cg, lanczos=gen_lanczos, arnoldi=gen_arnoldi, gmres=gen_gmres, pcg=gen_pcg,
                  power=gen_power, stochastic=gen_stochastic, next=gen_next)

if __name__ == "__main__":
    torch.set_num_threads(1)
    which = sys.argv[1:] or list(GENERATORS)       # `make_golden.py next` regenerates one family only
    manifest_path = os.path.join(HERE, "MANIFEST.json")
    previous = json.load(open(manifest_path)) if (sys.argv[1:] and os.path.exists(manifest_path)) else {}
    MANIFEST["cases"].update(previous.get("cases", {}))
    repaired = dict(previous.get("sparse_instances_repaired", {}))
    for name in which:
        before = len(REPAIRED)
        GENERATORS[name]()
        repaired[name] = len(REPAIRED) - before
    MANIFEST["sparse_instances_repaired"] = repaired     # per generator family
    with open(manifest_path, "w") as fh:
        json.dump(MANIFEST, fh, indent=1, sort_keys=True)
    print("wrote", len(MANIFEST["cases"]), "cases;", len(REPAIRED), "Sparse instances repaired")
