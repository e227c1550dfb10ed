"""Backward passes: cg_bwd (cola/linalg/inverse/cg.py:72-86) and slq_bwd (cola/linalg/tbd/slq.py:10-31).

  * the oracle's restatement against the gradients the REAL reference's autograd produced
    (tests/golden/bwd_*.npz, tests/golden/make_golden_bwd.py);
  * the native rules (cola_b200/autograd.py) on CPU tensors with the kernels replaced by tests/host_harness.py:
    the chain rule over the operator algebra, the Function plumbing, the leaf rules' index conventions;
  * (-m gpu) the native rules on the B200 through the C ABI: csrc/param_grad.cu.
The reference cannot differentiate a Sparse operator at all (torch CSR SpMM has no autograd: "Sparse CSR tensors do not
have strides"), so the two Sparse cases are pinned to the oracle only (which takes the dense-equivalent vjp restricted
to the pattern); the reference also returns no gradient for the right-hand side, so dB is checked against the oracle."""
import os

import numpy as np
import pytest
import torch

import cola_b200
from tests import problems as pb
from tests.bwd_cases import BWD_CASES, case
from tests.host_harness import emulated_kernels

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a = a.detach().cpu().double() if torch.is_tensor(a) else torch.as_tensor(np.asarray(a), dtype=torch.float64)
    b = b.detach().cpu().double() if torch.is_tensor(b) else torch.as_tensor(np.asarray(b), dtype=torch.float64)
    den = 0.5 * (a.norm() + b.norm())
    return 0.0 if den == 0 else float((a - b).norm() / den)


def load(name):
    path = os.path.join(GOLDEN, name + ".npz")
    return np.load(path) if os.path.exists(path) else None


def oracle_grads(name):
    """(loss, x or None, {d_param}, dB or None) from the oracle's restatement of the two backward rules."""
    from oracle import krylov_oracle as ko
    C = case(name)
    names = list(C["params"])

    def make_op(ps):
        return pb.to_oracle(C["spec"](dict(zip(names, ps))))

    params = [C["params"][k] for k in names]
    A = make_op(params)
    if C["kind"] == "cg":
        x, *_ = ko.cg(A, C["B"], tol=C["tol"], max_iters=C["max_iters"])
        d_params, db = ko.cg_bwd(make_op, params, x, C["W"], tol=C["tol"], max_iters=C["max_iters"])
        return float((C["W"] * x).sum()), x, dict(zip(names, d_params)), db
    loss = ko.slq(A, torch.log, max_iters=C["max_iters"], tol=C["tol"], vtol=C["vtol"], key=C["key"])
    num = max(int(1 / C["vtol"]**2), 1)
    d_params = ko.slq_bwd(make_op, params, torch.tensor(1.0, dtype=A.dtype), num, key=C["key"])
    return float(loss), None, dict(zip(names, d_params)), None


def native_grads(name, dev):
    cb = cola_b200
    C = case(name)
    params = {k: v.clone().to(dev).requires_grad_(True) for k, v in C["params"].items()}
    A = pb.to_b200(C["spec"](params), dev, C["ann"])
    dB = None
    if C["kind"] == "cg":
        B = C["B"].clone().to(dev).requires_grad_(True)
        x, info = cb.linalg.CG(tol=C["tol"], max_iters=C["max_iters"])(A, B)
        assert info["iterations"] > 0 and x.requires_grad
        loss = (C["W"].to(dev) * x).sum()
        loss.backward()
        dB = B.grad
    else:
        x = None
        loss = cb.linalg.stochastic_lanczos_quad(A, torch.log, max_iters=C["max_iters"], tol=C["tol"], vtol=C["vtol"],
                                                 key=C["key"])
        loss.backward()
    return float(loss), x, {k: p.grad for k, p in params.items()}, dB


def tols(name):
    return (2e-4, 2e-3) if name.endswith("f32") else (1e-9, 1e-7)


@pytest.mark.parametrize("name", BWD_CASES)
def test_oracle_backward_matches_reference_autograd(name):
    g = load(name)
    if g is None:
        pytest.skip("the reference's own backward raises on Sparse operators (no fixture); see the module docstring")
    t_fwd, t_bwd = tols(name)
    loss, x, grads, _ = oracle_grads(name)
    assert abs(loss - float(g["loss"])) <= t_fwd * max(1.0, abs(float(g["loss"])))
    if x is not None:
        assert rel(x, g["x"]) < t_fwd
    for k, d in grads.items():
        assert rel(d, g["d_" + k]) < t_bwd, (name, k, rel(d, g["d_" + k]))


def _check_native(name, dev):
    t_fwd, t_bwd = tols(name)
    loss, x, grads, dB = native_grads(name, dev)
    o_loss, o_x, o_grads, o_dB = oracle_grads(name)
    g = load(name)
    ref_loss = float(g["loss"]) if g is not None else o_loss
    assert abs(loss - ref_loss) <= t_fwd * max(1.0, abs(ref_loss)), (loss, ref_loss)
    for k, d in grads.items():
        assert d is not None, (name, k)
        want = g["d_" + k] if g is not None else o_grads[k]
        assert rel(d, want) < t_bwd, (name, k, rel(d, want))
        assert rel(d, o_grads[k]) < t_bwd, (name, k, "oracle", rel(d, o_grads[k]))
    if dB is not None:
        assert rel(dB, o_dB) < t_bwd


@pytest.mark.parametrize("name", BWD_CASES)
def test_native_backward_host_logic(name, monkeypatch):
    with emulated_kernels():
        _check_native(name, "cpu")


def test_parameters_and_refusal_outside_rules():
    """`parameters` lists float leaves only; an operator without a gradient rule raises instead of returning a result
    that silently lacks grad_fn."""
    with emulated_kernels():
        ops = cola_b200.ops
        data, rows, cols, shape = pb.laplacian_2d_coo(4, torch.float64)
        S = ops.Sparse(data, rows, cols, shape)
        assert [tuple(p.shape) for p in cola_b200.autograd.parameters(S)] == [tuple(data.shape)]
        M = torch.eye(16, dtype=torch.float64).requires_grad_(True)
        T = cola_b200.PSD(ops.Triangular(M))
        b = torch.ones(16, 1, dtype=torch.float64)
        x, _ = cola_b200.linalg.CG(tol=1e-8, max_iters=20)(T, b)
        with pytest.raises(NotImplementedError, match="no parameter gradient rule"):
            x.sum().backward()


@pytest.mark.gpu
@pytest.mark.parametrize("name", BWD_CASES)
def test_native_backward_gpu(name, monkeypatch):
    assert torch.cuda.is_available()
    cola_b200.backend.lib()
    monkeypatch.setattr(cola_b200.rng, "PROBE_DEVICE", "cpu")
    _check_native(name, "cuda:0")


@pytest.mark.gpu
def test_param_grad_kernels_at_scale():
    """The three kernels at sizes that exercise split-K, ragged tiles, wide / narrow rows: against fp64 torch."""
    be = cola_b200.backend
    dev = "cuda:0"
    g = torch.Generator().manual_seed(0)
    for dt, tol in [(torch.float32, 2e-6), (torch.float64, 1e-13)]:
        # SDDMM on a 2-D Laplacian pattern, 64 columns
        data, rows, cols, shape = pb.laplacian_2d_coo(96, dt)
        S = cola_b200.ops.Sparse(data.to(dev), rows.to(dev), cols.to(dev), shape)
        n = shape[0]
        G = torch.randn(n, 64, dtype=dt, generator=g).to(dev)
        V = torch.randn(n, 64, dtype=dt, generator=g).to(dev)
        out = torch.empty_like(S.data)
        be.sddmm_csr(S.indptr, S.indices, n, G, V, 0.5, out)
        want = 0.5 * (G.double()[S.row_indices.long()] * V.double()[S.col_indices.long()]).sum(1)
        assert rel(out, want) < tol
        # row dots with offsets, k = 1, 7, 64
        for k in (1, 7, 64):
            Gk, Vk = G[:, :k].contiguous(), V[:, :k].contiguous()
            o = torch.empty(n - 3, dtype=dt, device=dev)
            be.row_dots(Gk, 3, Vk, 0, n - 3, -2.0, o)
            assert rel(o, -2.0 * (Gk.double()[3:] * Vk.double()[:n - 3]).sum(1)) < tol
        # gram: Dense shape (ragged 100 x 70, K = 33) and mode-Gram shape (64 x 64 over pre = 64, post = 64 * 32)
        Gd, Zd = torch.randn(100, 33, dtype=dt, generator=g).to(dev), torch.randn(70, 33, dtype=dt, generator=g).to(dev)
        C = torch.zeros(100, 70, dtype=torch.float64, device=dev)
        be.gram_nt(Gd, 0, Zd, 0, 100, 70, 1, 33, 1.5, C)
        assert rel(C, 1.5 * Gd.double() @ Zd.double().T) < 1e-13
        pre, d, post = 64, 64, 64 * 32
        Gm = torch.randn(pre, d, post, dtype=dt, generator=g).to(dev)
        Zm = torch.randn(pre, d, post, dtype=dt, generator=g).to(dev)
        C = torch.zeros(d, d, dtype=torch.float64, device=dev)
        be.gram_nt(Gm, 0, Zm, 0, d, d, pre, post, 1.0, C)
        assert rel(C, torch.einsum("pat,pjt->aj", Gm.double(), Zm.double())) < 1e-12
