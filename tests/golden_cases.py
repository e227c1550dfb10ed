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
