"""Backward-pass golden vectors from the REAL reference's autograd (wilson-labs/cola at /root/reference):
cg_bwd (cola/linalg/inverse/cg.py:72-86) and slq_bwd (cola/linalg/tbd/slq.py:10-31) through
`iterative_autograd` (cola/utils/custom_autodiff.py).  Same conventions as make_golden.py (runs only in the build
container; outputs only).

    python tests/golden/make_golden_bwd.py
"""
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
from cola.linalg.inverse.cg import CG  # noqa: E402
from cola.linalg.tbd.slq import stochastic_lanczos_quad  # noqa: E402

from tests.bwd_cases import BWD_CASES, case  # noqa: E402
from tests.golden.make_golden import to_reference  # noqa: E402

assert cola.__file__.startswith("/root/reference"), cola.__file__


def run(name):
    C = case(name)
    params = {k: v.clone().requires_grad_(True) for k, v in C["params"].items()}
    A = to_reference(C["spec"](params), C["ann"])
    out = {}
    if C["kind"] == "cg":
        # the reference propagates no gradient to the right-hand side (custom_autodiff.py:40-43 keeps dA only)
        x, info = CG(tol=C["tol"], max_iters=C["max_iters"])(A, C["B"])
        loss = (C["W"] * x).sum()
        loss.backward()
        out["x"] = x.detach()
    else:
        loss = stochastic_lanczos_quad(A, torch.log, max_iters=C["max_iters"], tol=C["tol"], vtol=C["vtol"], key=C["key"])
        loss.backward()
    out["loss"] = loss.detach()
    for k, p in params.items():
        assert p.grad is not None, (name, k)
        out["d_" + k] = p.grad
    return out


if __name__ == "__main__":
    only = sys.argv[1:]
    for name in BWD_CASES:
        if only and name not in only:
            continue
        try:
            out = run(name)
        except Exception as e:  # noqa: BLE001
            print(f"{name}: reference FAILED: {type(e).__name__}: {str(e)[:300]}")
            continue
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **{k: v.numpy() for k, v in out.items()})
        print(name, {k: tuple(v.shape) for k, v in out.items()}, "loss", float(out["loss"]))
