"""Case tables shared by tests/golden/make_golden.py (which holds the authoritative copy used with
the reference) and the parity tests.  Kept in a reference-free module so the GPU box can import it."""
MATMAT_PROBLEMS = ["dense96_f32", "dense96_f64", "lap24_f32", "lap24_f64", "lap16_shift_f32", "kron888_f32",
                   "kron465_diag_f64", "kron884_diag_f32", "blockdiag_f32", "product_f64", "nonsym48_f32",
                   "graph2k_f64"]

CG_CASES = {  # case -> (problem, tol, max_iters)
    "cg_cfg1_dense1024": ("cfg1_dense1024", 1e-6, 1000),
    "cg_dense96_f32": ("dense96_f32", 1e-6, 500),
    "cg_dense96_f64": ("dense96_f64", 1e-11, 500),
    "cg_lap24_f32": ("lap24_f32", 1e-30, 60),
    "cg_lap24_f64": ("lap24_f64", 1e-30, 60),
    "cg_lap24_f64_conv": ("lap24_f64", 1e-9, 1000),
    "cg_lap16_shift_f32": ("lap16_shift_f32", 1e-6, 200),
    "cg_kron888_f32": ("kron888_f32", 1e-6, 200),
    "cg_kron465_diag_f64": ("kron465_diag_f64", 1e-10, 200),
    "cg_kron884_diag_f32": ("kron884_diag_f32", 1e-30, 40),
    "cg_blockdiag_f32": ("blockdiag_f32", 1e-6, 100),
    "cg_product_f64": ("product_f64", 1e-10, 200),
}

LANCZOS_CASES = {  # case -> (problem, max_iters, tol, batched?)
    "lanczos_dense96_f32": ("dense96_f32", 30, 1e-7, True),
    "lanczos_dense96_f64": ("dense96_f64", 30, 1e-12, True),
    "lanczos_lap24_f64_vec": ("lap24_f64", 40, 1e-12, False),
    "lanczos_graph2k_f64_vec": ("graph2k_f64", 48, 1e-12, False),
    "lanczos_kron884_diag_f32": ("kron884_diag_f32", 25, 1e-7, True),
    "lanczos_kron465_diag_f64": ("kron465_diag_f64", 20, 1e-12, True),
}

ARNOLDI_CASES = {
    "arnoldi_nonsym48_f32": ("nonsym48_f32", 20, 1e-7, True),
    "arnoldi_nonsym48_f64": ("nonsym48_f64", 20, 1e-12, True),
    "arnoldi_nonsym48_f64_vec": ("nonsym48_f64", 48, 1e-12, False),
}

GMRES_CASES = {  # case -> (problem, max_iters, tol, vector?)
    "gmres_nonsym48_f64": ("nonsym48_f64", 48, 1e-12, False),
    "gmres_nonsym48_f32": ("nonsym48_f32", 30, 1e-7, False),
    "gmres_nonsym48_f64_vec": ("nonsym48_f64", 20, 1e-12, True),
    "gmres_lap24_f64": ("lap24_f64", 40, 1e-12, False),
    "gmres_kron465_diag_f64": ("kron465_diag_f64", 25, 1e-12, False),
}

PCG_CASES = {  # case -> (problem, rank, tol, max_iters)   CG with P = NystromPrecond(A, rank, key=PRNGKey(3))
    "pcg_dense96_f64": ("dense96_f64", 12, 1e-11, 500),
    "pcg_dense96_f32": ("dense96_f32", 12, 1e-6, 500),
    "pcg_lap24_f64": ("lap24_f64", 24, 1e-9, 1000),
    "pcg_kron465_diag_f64": ("kron465_diag_f64", 10, 1e-10, 200),
}

POWER_CASES = {  # case -> (problem, tol, max_iter)   power_iteration(A, key=PRNGKey(11))
    "power_dense96_f64": ("dense96_f64", 1e-9, 400),
    "power_dense96_f32": ("dense96_f32", 1e-5, 400),
    "power_kron465_diag_f64": ("kron465_diag_f64", 1e-8, 300),
    "power_lap24_f64_capped": ("lap24_f64", 1e-12, 25),
}

# ---- f(A)v operators, exact / off-diagonal estimators, KronSum and Tridiagonal matmats
NEXT_MATMAT_PROBLEMS = ["kronsum465_f64", "kronsum884_f32", "tridiag200_f64", "tridiag200_shift_f32"]

NEXT_CG_CASES = {  # case -> (problem, tol, max_iters)
    "cg_kronsum465_f64": ("kronsum465_f64", 1e-10, 200),
    "cg_kronsum884_f32": ("kronsum884_f32", 1e-6, 200),
    "cg_tridiag200_shift_f32": ("tridiag200_shift_f32", 1e-6, 200),
}

UNARY_CASES = {  # case -> (problem, function, algorithm, max_iters, tol)   F = cola.linalg.<function>(A, alg); Y = F @ B
    "expA_arnoldi_nonsym48_f64": ("nonsym48_f64", "exp", "arnoldi", 20, 1e-12),
    "logA_arnoldi_nonsym48_f64": ("nonsym48_f64", "log", "arnoldi", 48, 1e-12),
    "sqrtA_arnoldi_nonsym48_f32": ("nonsym48_f32", "sqrt", "arnoldi", 20, 1e-7),
    "expA_arnoldi_tridiag200_f64": ("tridiag200_f64", "exp", "arnoldi", 30, 1e-12),
    "sqrtA_lanczos_kron465_diag_f64": ("kron465_diag_f64", "sqrt", "lanczos", 30, 1e-7),
    "isqrtA_lanczos_kron884_diag_f32": ("kron884_diag_f32", "isqrt", "lanczos", 25, 1e-7),
    "expA_lanczos_kronsum465_f64": ("kronsum465_f64", "exp", "lanczos", 8, 1e-12),   # exp(KronSum) -> Kronecker of exp(factor); full Krylov spaces
    "sqrtA_lanczos_kron888_pure_f32": ("kron888_pure_f32", "sqrt", "lanczos", 8, 1e-7),   # full Krylov spaces (d = 8)   # pow(Kronecker) -> Kronecker of pow(factor)
}

DIAG_CASES = {  # case -> (problem, k, algorithm)   cola.linalg.diag(A, k, alg)
    "diag_exact_tridiag200_f64_k0": ("tridiag200_f64", 0, "exact"),
    "diag_exact_tridiag200_f64_k1": ("tridiag200_f64", 1, "exact"),
    "diag_exact_tridiag200_f64_km1": ("tridiag200_f64", -1, "exact"),
    "diag_exact_lap16_shift_f32_k0": ("lap16_shift_f32", 0, "exact"),     # ragged last block (n = 256)
    "diag_exact_kron465_diag_f64_k0": ("kron465_diag_f64", 0, "exact"),
    # structure rules (diag_trace.py:56-119): with Hutch these are exact or partly exact, term by term
    "diag_hutch_kron465_diag_f64_k0": ("kron465_diag_f64", 0, "hutch"),      # Sum[Kronecker, Diagonal]: exact
    "diag_hutch_blockdiag_f32_k0": ("blockdiag_f32", 0, "hutch"),            # BlockDiag of Dense: exact
    "diag_hutch_kronsum465_f64_k0": ("kronsum465_f64", 0, "hutch"),          # KronSum of Dense: exact
    "diag_hutch_lap16_shift_f32_k0": ("lap16_shift_f32", 0, "hutch"),        # Sum[CSR (Hutch), c*I (Hutch), Diagonal (exact)]
    "diag_hutch_kron888_f32_k0": ("kron888_f32", 0, "hutch"),                # Sum[Kronecker (exact), c*I (Hutch)]
    "diag_hutch_product_f64_k2": ("product_f64", 2, "hutch"),
    "diag_hutch_product_f64_km3": ("product_f64", -3, "hutch"),
}
