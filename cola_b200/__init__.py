"""cola_b200: a B200-native (sm_100a) Krylov engine behind CoLA's operator / algorithm API.

    import cola_b200 as cola
    A = cola.PSD(cola.ops.Sparse(data, rows, cols, shape))        # CUDA tensors
    x = cola.linalg.solve(A, b, cola.linalg.CG(tol=1e-6))

Same names, arguments, return values and error behaviour as wilson-labs/cola for the Krylov hot path
(CG / Lanczos / Arnoldi / SLQ-Hutchinson and the Dense, Sparse, Kronecker, BlockDiag, Diagonal, Sum, Product
matmats).  All arithmetic runs in hand-written CUDA kernels loaded from cola_b200/csrc/libcola_b200.so through
the C ABI in include/cola_b200.h; there is no CPU or eager-torch fallback.

`cola_b200.install()` rebinds the reference's own loop functions and structured matmats to this engine for CUDA
float32/float64 operators (cola_b200/plugin.py), which makes it a drop-in inside an existing wilson-labs/cola
program.
"""
from . import autograd, backend, linalg, ops, plugin, rng, sharding
from .ops import (PSD, Hermitian, LinearOperator, SelfAdjoint, Stiefel, Unitary, block_diag, densify, kron, kronsum,
                  lazify)

from .plugin import from_cola, install, uninstall

__all__ = ["autograd", "backend", "linalg", "ops", "plugin", "rng", "sharding", "install", "uninstall", "from_cola", "PSD", "SelfAdjoint", "Hermitian", "Stiefel", "Unitary",
           "LinearOperator", "lazify", "kron", "kronsum", "densify", "block_diag"]
