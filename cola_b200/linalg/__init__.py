from .algorithm_base import Algorithm, Auto, IterativeOperatorWInfo
from .arnoldi import arnoldi, arnoldi_eigs, arnoldi_fact
from .cg import CG, cg, release_cg_workspace, run_batched_cg
from .dispatch import (LSTSQ, LU, Arnoldi, Cholesky, Eigh, Exact, Lanczos, TriangularInv, apply_unary, diag, eig, eigmax,
                       eigmin, pinv, exact_diag, exp, get_slice,
                       inv, isqrt, log, logdet, pow, slogdet, solve, sqrt, trace)
from .gmres import GMRES, gmres, gmres_fwd
from .lanczos import lanczos, lanczos_eigs, lanczos_fact
from .power_iteration import PowerIteration, power_iteration
from .preconditioners import NystromPrecond, get_nys_approx
from .unary import ArnoldiUnary
from .stochastic import (Hutch, LanczosUnary, hutchinson_diag_estimate, slq_fwd, slq_per_probe,
                         stochastic_lanczos_quad)
