/*
 * cola_b200.h -- C ABI of the B200-native Krylov engine (libcola_b200.so).
 *
 * This is the drop-in boundary for CoLA's Krylov hot path.  The reference
 * (wilson-labs/cola) is pure Python over eager torch ops and has no native
 * interface of its own; every entry point below therefore names the reference
 * *Python* function / torch call sites it replaces (paths relative to the
 * reference checkout).  INTEGRATION.md shows the ctypes binding a cola
 * maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative COLA_E_* for bad
 *     arguments, or a positive cudaError_t; cola_last_error() gives the text.
 *     No exceptions cross this boundary.
 *   - all data pointers are DEVICE pointers unless the comment says "host".
 *     The caller owns every buffer; the library never allocates, frees or
 *     retains device memory.
 *   - dense blocks of vectors are row-major (n, k) with the RHS / probe
 *     index fastest and an explicit leading dimension `ld` (elements).
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it,
 *     nothing synchronises the host.
 *   - reductions (column dots) accumulate into DOUBLE accumulators that the
 *     caller zeroes; kernels add to them with fp64 atomics, so results do not
 *     depend on launch geometry beyond 1e-16 relative.
 *   - suffix _f32 / _f64 = the arithmetic type of the path.
 */
#ifndef COLA_B200_H
#define COLA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COLA_OK 0
#define COLA_E_BADARG (-1)
#define COLA_E_UNSUPPORTED (-2)
#define COLA_E_NOGPU (-3)

/* ---- library ---------------------------------------------------------- */
int cola_version(void);               /* ABI version, bump on any signature change          */
const char* cola_last_error(void);    /* host string describing the last non-zero status    */
int cola_device_info(int* sm_count, int* cc_major, int* cc_minor); /* COLA_E_NOGPU without a device */
/* Kernels launched by this library since load (the `gpu_launches` evidence). */
int64_t cola_launch_count(void);

/* ---- device-side gating ------------------------------------------------------
 * Krylov loops run without host synchronisation: a batch of iterations is enqueued (or captured in a
 * CUDA graph) before the host knows where the stopping rule fires.  Every kernel therefore takes an
 * optional `gate` (device int32*, may be NULL): the launch is a no-op when *gate != 0.  For CG the gate
 * is &ctl->done.  Matmats that feed a per-iteration accumulator also take `dots_row` (device int32*,
 * may be NULL): the column dots go to dots[(*dots_row) * k + c] (for CG, &ctl->it).
 */

/* ---- fused operator epilogue ------------------------------------------
 * Every matmat computes, for a square "core" operator K (Dense / CSR / Kronecker / BlockDiag):
 *     Y[i,:] = alpha * (K X)[i,:] + (shift + diag[i]) * X[i,:]   (+ Y[i,:] if accumulate)
 * which is how Sum[K, c*I, Diagonal(d)] / ScalarMul compositions
 * (cola/ops/operators.py:84-127,138-191,323-348) collapse into one HBM pass.
 * If `dots` != NULL the kernel also adds  sum_i X[i,c] * Y[i,c]  into dots[c]
 * (k doubles): the p^T A p of CG (cola/linalg/inverse/cg.py:157-158) or the Lanczos
 * <w, v_i> (cola/linalg/decompositions/lanczos.py:245) in the same pass.
 * shift/diag/dots require a square operator.  diag may be NULL.
 */

/* Sparse CSR  Y = A X.   Replaces Sparse._matmat -> torch.sparse_csr @ dense
 * (cola/ops/operators.py:77-78, indices int32 per :73-74).  nnz = rowptr[n_rows]; max_row_nnz = longest row
 * (0 if unknown): both only steer tiling / prefetch depth, never results. */
int cola_csr_spmm_f32(const int32_t* rowptr, const int32_t* colidx, const float* vals, int64_t n_rows,
                      int64_t n_cols, int64_t nnz, int64_t max_row_nnz, const float* X, int64_t ldx, int64_t k, float* Y, int64_t ldy, float alpha,
                      float shift, const float* diag, int accumulate, double* dots, const int32_t* dots_row,
                      const int32_t* gate, void* stream);
int cola_csr_spmm_f64(const int32_t* rowptr, const int32_t* colidx, const double* vals, int64_t n_rows,
                      int64_t n_cols, int64_t nnz, int64_t max_row_nnz, const double* X, int64_t ldx, int64_t k, double* Y, int64_t ldy, double alpha,
                      double shift, const double* diag, int accumulate, double* dots, const int32_t* dots_row,
                      const int32_t* gate, void* stream);

/* Staged CSR SpMM for wide X blocks on patterns with long column runs (stencil / banded matrices; same replacement as
 * cola_csr_spmm_*: Sparse._matmat, cola/ops/operators.py:77-78, + cg.py:157-158 through `dots`).  The pattern comes in
 * the tile-local form cola_b200/csr_tiles.py builds once: a tile = `strips` strips of `strip_rows` consecutive rows,
 * `stride` rows apart for the first n_tiles2d tiles (rows [0, rows2d)), consecutive rows after; rec = 32 int32 per tile
 * [nz_begin, nz_padded (multiple of 4), n_runs (-1: gather from global), n_distinct, 0 x4, (col0, slot0<<16|len) x12];
 * rp = 2*RT+4 int32 per tile (RT = strip_rows*strips): local row pointers [0,RT], n_runs again at [RT+3], the slot
 * of each row's own X row at [RT+4, 2RT+4); idx = per non-zero the slot of its X row among the tile's staged
 * rows (the column for irregular tiles); vals in the same padded order.  cap_rows / cap_nz = largest n_distinct /
 * nz_padded of any tile (ring-stage size).  X and Y contiguous (ldx == k), k*sizeof(T) a multiple of 16.
 * COLA_E_UNSUPPORTED when two ring stages do not fit shared memory. */
int cola_csr_spmm_tiled_f32(const int32_t* rec, const int32_t* rp, const int32_t* idx, const float* vals, int64_t n_rows,
                            int64_t n_tiles, int64_t n_tiles2d, int64_t rows2d, int64_t stride, int64_t strip_rows,
                            int64_t strips, int64_t cap_rows, int64_t cap_nz, const float* X, int64_t k, float* Y, int64_t ldy,
                            float alpha, float shift, const float* diag, int accumulate, double* dots,
                            const int32_t* dots_row, const int32_t* gate, void* stream);
int cola_csr_spmm_tiled_f64(const int32_t* rec, const int32_t* rp, const int32_t* idx, const double* vals, int64_t n_rows,
                            int64_t n_tiles, int64_t n_tiles2d, int64_t rows2d, int64_t stride, int64_t strip_rows,
                            int64_t strips, int64_t cap_rows, int64_t cap_nz, const double* X, int64_t k, double* Y,
                            int64_t ldy, double alpha, double shift, const double* diag, int accumulate, double* dots,
                            const int32_t* dots_row, const int32_t* gate, void* stream);

/* Batched mode contraction  out[p,a,q] = alpha * sum_j M[a,j] in[p,j,q].
 * M is (d_out, d_in) row-major with leading dimension ldm.  `in` is (pre, d_in, post) contiguous,
 * `out` (pre, d_out, post) contiguous; they must not alias.
 *   Dense._matmat      (operators.py:26-28)    : pre=1, post=k
 *   Kronecker._matmat  (operators.py:216-223)  : one call per factor, no moveaxis/reshape copies
 *   BlockDiag._matmat  (operators.py:299-310)  : pre=multiplicity, post=k, one call per block
 * The epilogue (shift/diag/accumulate/dots) acts on the flattened (pre*d_out, post) matrix, row = p*d_out+a,
 * and needs d_out == d_in for shift/diag/dots; `epi_x` is the tensor the epilogue's X refers to (the
 * matmat input, laid out like `out`); pass NULL when shift/diag/dots are unused.  dots has `post` entries. */
int cola_mode_contract_f32(const float* M, int64_t ldm, int64_t d_out, int64_t d_in, int64_t pre, int64_t post,
                           const float* in, float* out, float alpha, float shift, const float* diag,
                           const float* epi_x, int accumulate, double* dots, const int32_t* dots_row,
                           const int32_t* gate, void* stream);
int cola_mode_contract_f64(const double* M, int64_t ldm, int64_t d_out, int64_t d_in, int64_t pre, int64_t post,
                           const double* in, double* out, double alpha, double shift, const double* diag,
                           const double* epi_x, int accumulate, double* dots, const int32_t* dots_row,
                           const int32_t* gate, void* stream);

/* Kronecker matmat on the tensor cores: Y = alpha * (F_1 (x) ... (x) F_D) X (+ epilogue), every factor 64x64
 * fp32 row-major with leading dimension ldf[i], X / Y (64^D, k) row-major with k a multiple of 32.
 * tcgen05.mma kind::tf32 with 3xTF32 error compensation (fp32-grade accuracy), TMEM accumulators, TMA-staged
 * 128B-swizzled tiles; all D mode contractions of all column chunks in ONE cooperative launch (2 <= D <= 3).  Replaces
 * Kronecker._matmat (operators.py:216-223) for the GP-kernel shapes of BASELINE config 3.  `factors` / `ldf` are HOST
 * arrays (of device pointers / of leading dimensions); `workspace` is a device buffer of
 * cola_kron_tc_workspace_bytes(n, D) bytes, 128-byte aligned.  cola_kron_tc_supported returns 1 when the shapes
 * qualify (otherwise: cola_mode_contract_tc_f32 / cola_mode_contract_* per mode). */
int cola_kron_tc_supported(int64_t n_factors, const int64_t* dims, int64_t k);
int64_t cola_kron_tc_workspace_bytes(int64_t n, int64_t n_factors);
int cola_kron_matmat_tc_f32(int64_t n_factors, const float* const* factors, const int64_t* ldf, const float* X,
                            float* Y, int64_t k, float* workspace, float alpha, float shift, const float* diag,
                            int accumulate, double* dots, const int32_t* dots_row, const int32_t* gate,
                            void* stream);

/* One mode contraction on the tensor cores (tcgen05 3xTF32, same tiles as cola_kron_matmat_tc_f32) for a SQUARE factor of
 * size d = 64 or 128:  out[p, a, l, r] = alpha * sum_j M[a, j] in[p, j, l, r]  with in / out (pre, d, L, k) row-major,
 * k a multiple of 32 and pre * L a multiple of 4.  It takes the modes of a Kronecker chain the fused kernel does not
 * (BASELINE config 4: Kronecker(128, 128, 64)); the epilogue (shift / diag / dots / accumulate, as cola_mode_contract_*)
 * needs L == 1 (the last factor).  in and out must not alias.  cola_mode_contract_tc_supported says whether a shape is
 * taken. */
int cola_mode_contract_tc_supported(int64_t d, int64_t pre, int64_t L, int64_t k);
int cola_mode_contract_tc_f32(const float* M, int64_t ldm, int64_t d, int64_t pre, int64_t L, int64_t k, const float* in,
                              float* out, float alpha, float shift, const float* diag, const float* epi_x, int accumulate,
                              double* dots, const int32_t* dots_row, const int32_t* gate, void* stream);

/* Operators with no core (Diagonal, ScalarMul*Identity, sums of those):
 *   Y = (shift + diag[i]) * X  (+Y).   Diagonal._matmat / ScalarMul._matmat (operators.py:97-98,338-339). */
int cola_diag_matmat_f32(const float* X, int64_t ldx, float* Y, int64_t ldy, int64_t n, int64_t k, float shift,
                         const float* diag, int accumulate, double* dots, const int32_t* dots_row,
                         const int32_t* gate, void* stream);
int cola_diag_matmat_f64(const double* X, int64_t ldx, double* Y, int64_t ldy, int64_t n, int64_t k, double shift,
                         const double* diag, int accumulate, double* dots, const int32_t* dots_row,
                         const int32_t* gate, void* stream);

/* ---- column reductions / scalings on (n,k) blocks, leading dimension ld ---------------- */
/* dots[c] += sum_i X[i,c]*Y[i,c]   (xnp.sum(conj(a)*b, axis=-2), xnp.norm(.,axis=-2)**2) */
int cola_col_dots_f32(const float* X, const float* Y, int64_t n, int64_t k, int64_t ld, double* dots,
                      const int32_t* gate, void* stream);
int cola_col_dots_f64(const double* X, const double* Y, int64_t n, int64_t k, int64_t ld, double* dots,
                      const int32_t* gate, void* stream);
/* mode 0: Y[i,c] = a*s[c]*X[i,c]
 * mode 1: Y[i,c] = X[i,c] / safe(s[c])     (|s|<1e-40 -> 1e-40: do_safe_div, cg.py:173-178)
 * mode 2: Y[i,c] = X[i,c] / s[c]           (unguarded, lanczos.py:240-241)
 * mode 3: Y[i,c] = X[i,c] / max(s[c], a)   (clip(norm, min=tol/2), arnoldi.py:315)
 * s[c] = sqrt(sq[c]) if take_sqrt else sq[c]; sq is k doubles.  Y may alias X. */
int cola_col_scale_f32(const float* X, float* Y, int64_t n, int64_t k, int64_t ld, const double* sq, int take_sqrt,
                       int mode, float a, const int32_t* gate, void* stream);
int cola_col_scale_f64(const double* X, double* Y, int64_t n, int64_t k, int64_t ld, const double* sq, int take_sqrt,
                       int mode, double a, const int32_t* gate, void* stream);
/* Y = a*X + b*Y elementwise (residual r0 = b - A x0, cg.py:123). */
int cola_axpby_f32(const float* X, float* Y, int64_t n, int64_t k, int64_t ld, float a, float b, const int32_t* gate,
                   void* stream);
int cola_axpby_f64(const double* X, double* Y, int64_t n, int64_t k, int64_t ld, double a, double b,
                   const int32_t* gate, void* stream);

/* ---- CG iteration (cola/linalg/inverse/cg.py:94-178, preconditioner = Identity) ------------------
 * Device-resident control block: the whole loop, including the stopping rule, runs without host syncs.
 * Iteration `it` uses accumulator rows gamma[it*k..], pAp[it*k..], gamma[(it+1)*k..] (all zeroed by the
 * caller before the solve; gamma row 0 holds <r0,r0>).  After the solve gamma[j*k+c] = ||r_j[:,c]||^2 is
 * the full residual trace from which info['errors'] (cola/utils/torch_tqdm.py:35-62) is rebuilt on the host. */
typedef struct {
  int32_t it;        /* iterations completed                                                        */
  int32_t done;      /* set when the reference's cond_fun (cg.py:133-138) turns false               */
  int32_t max_iters;
  int32_t k;
} cola_cg_ctl_t;

/* One CG iteration minus the matmat, as two sweeps + one tiny scalar kernel (10 vector passes per iteration
 * including the matmat's 2, instead of the 11 of the textbook x/r-update + p-update split):
 *   r :      alpha = safe(gamma[it]/pAp[it]) (0 where ||r||<1e-40);  R -= alpha AP;  gamma[it+1] += <R,R>
 *   xp:      beta  = safe(gamma[it+1]/gamma[it]) (0 where converged);  X += alpha P;  P = R + beta P
 *   advance: it += increment;  done = !(any(sqrt(gamma[it]) > tol_eff) && it < max_iters)
 * Each kernel reads ctl->it / ctl->done on the device and is a no-op once done. */
int cola_cg_update_r_f32(float* R, const float* AP, int64_t n, int64_t k, int64_t ld, const cola_cg_ctl_t* ctl,
                         const double* gamma, const double* pAp, double* gamma_w, void* stream);
int cola_cg_update_r_f64(double* R, const double* AP, int64_t n, int64_t k, int64_t ld, const cola_cg_ctl_t* ctl,
                         const double* gamma, const double* pAp, double* gamma_w, void* stream);
int cola_cg_update_xp_f32(float* X, const float* R, float* P, int64_t n, int64_t k, int64_t ld,
                          const cola_cg_ctl_t* ctl, const double* gamma, const double* pAp, void* stream);
int cola_cg_update_xp_f64(double* X, const double* R, double* P, int64_t n, int64_t k, int64_t ld,
                          const cola_cg_ctl_t* ctl, const double* gamma, const double* pAp, void* stream);
/* tol_eff[c] = tol*||r0[:,c]|| + tol in the path's dtype (cg.py:101). */
int cola_cg_tol_f32(const double* gamma0, float tol, float* tol_eff, int64_t k, void* stream);
int cola_cg_tol_f64(const double* gamma0, double tol, double* tol_eff, int64_t k, void* stream);
int cola_cg_advance_f32(cola_cg_ctl_t* ctl, const double* gamma, const float* tol_eff, int increment, void* stream);
int cola_cg_advance_f64(cola_cg_ctl_t* ctl, const double* gamma, const double* tol_eff, int increment, void* stream);

/* Publish a few bytes of device state (the CG control block, a row of squared norms) to MAPPED PINNED HOST memory
 * from a kernel, in stream order: `host_mapped` is a cudaHostAlloc'ed buffer (directly addressable from the device
 * under unified addressing), nbytes a multiple of 4.  This is how the loops poll their stopping rules (the two
 * device->host syncs per iteration of cola/utils/torch_tqdm.py:42,88): the host waits on an event and reads its own
 * memory, so the poll never queues behind a large DMA transfer on the copy engines (a 16-byte cudaMemcpy does: it
 * cost 16 % of the end-to-end solve when 1 GiB result copies were in flight). */
int cola_publish_bytes(const void* src, void* host_mapped, int64_t nbytes, void* stream);

/* The opposite direction: four 32-bit words passed as kernel arguments are stored to dst[0..3] in stream order.  The
 * loops build their control blocks (cola_cg_ctl_t: iteration, done flag, cap, columns) with it: a 16-byte pageable
 * cudaMemcpy would queue on the H2D copy engine behind the next solve's 1 GiB right-hand-side block. */
int cola_store_i32x4(int32_t* dst, int32_t a, int32_t b, int32_t c, int32_t d, void* stream);

/* ---- Full reorthogonalisation (Lanczos CGS2: lanczos.py:287-296; Arnoldi MGS: arnoldi.py:304-311) ------
 * Krylov basis layout (B200-native, differs from the reference's (b, n, m+2) with the Krylov index
 * fastest): V is (n_vec, n, b) contiguous, i.e. vector j is an (n, b) row-major block at V + j*vstride,
 * which is exactly the matmat operand layout, so no transposes are ever made.
 *   dots:   C[j*b + c] += sum_i V[j][i,c] * W[i,c]   for j in [j0, j1)          (C doubles, zeroed by caller)
 *   update: W[i,c] = W[i,c] + sign * sum_{j in [j0,j1)} coef(j,c) * V[j][i,c]
 *           coef = C[j*b+c] (doubles, rounded to the path dtype first)           (sign = -1 for Gram-Schmidt)
 *           if wnorm2 != NULL also wnorm2[c] += sum_i W_new[i,c]^2
 */
int cola_reorth_dots_f32(const float* V, int64_t vstride, int64_t j0, int64_t j1, const float* W, int64_t n,
                         int64_t b, double* C, const int32_t* gate, void* stream);
int cola_reorth_dots_f64(const double* V, int64_t vstride, int64_t j0, int64_t j1, const double* W, int64_t n,
                         int64_t b, double* C, const int32_t* gate, void* stream);
int cola_reorth_update_f32(const float* V, int64_t vstride, int64_t j0, int64_t j1, float* W, int64_t n, int64_t b,
                           const double* C, float sign, double* wnorm2, const int32_t* gate, void* stream);
int cola_reorth_update_f64(const double* V, int64_t vstride, int64_t j0, int64_t j1, double* W, int64_t n, int64_t b,
                           const double* C, double sign, double* wnorm2, const int32_t* gate, void* stream);

/* Fused middle step of CGS2 (lanczos.py:287-296): W += sign * V C1 and C2[j*b+c] += sum_i V[j][i,c] * W_new[i,c] with
 * the basis read ONCE (a row-chunk of all vectors is kept in shared memory): 3 sweeps over V per Lanczos step
 * instead of 4.  Returns COLA_E_UNSUPPORTED when the shape does not fit (b*sizeof(T) not a multiple of 16 bytes
 * -- a single column is folded when n allows --, too many vectors for the per-thread register slices, or a chunk
 * too large for shared memory); callers then use cola_reorth_update_* + cola_reorth_dots_*. */
int cola_reorth_update_dots_f32(const float* V, int64_t vstride, int64_t j0, int64_t j1, float* W, int64_t n,
                                int64_t b, const double* C1, float sign, double* C2, const int32_t* gate,
                                void* stream);
int cola_reorth_update_dots_f64(const double* V, int64_t vstride, int64_t j0, int64_t j1, double* W, int64_t n,
                                int64_t b, const double* C1, double sign, double* C2, const int32_t* gate,
                                void* stream);

/* Eigenvalues and FIRST eigenvector components of b symmetric tridiagonal m x m matrices (the Lanczos T of
 * slq.py:42-51, which the reference hands to a dense batched eigh): implicit-shift QL in fp64, one thread per
 * matrix, O(m^2).  Arrays are [i][matrix] with row stride ld >= b.  In: d = diagonal (m rows), e = off-diagonal
 * (e[i] couples i and i+1; m rows of storage, the last is scratch).  Out: d = eigenvalues (unordered),
 * z = first components (signs arbitrary), e destroyed, status[matrix] = 1 if an eigenvalue did not converge. */
int cola_tridiag_eig_first_row_f64(double* d, double* e, double* z, int64_t m, int64_t b, int64_t ld, int32_t* status,
                                   const int32_t* gate, void* stream);

/* Lanczos three-term step after the matmat (lanczos.py:245-248), fused:
 *   W -= alpha[c] * Vi + beta_prev[c] * Vim1   with alpha[c] = (T)alpha_acc[c] (the <w,v_i> dots of the matmat),
 *   beta_prev[c] = (T)sqrt(beta_prev_sq[c]).  Vim1 / beta_prev_sq may be NULL (first step). */
int cola_lanczos_three_term_f32(float* W, const float* Vi, const float* Vim1, int64_t n, int64_t b,
                                const double* alpha_acc, const double* beta_prev_sq, const int32_t* gate,
                                void* stream);
int cola_lanczos_three_term_f64(double* W, const double* Vi, const double* Vim1, int64_t n, int64_t b,
                                const double* alpha_acc, const double* beta_prev_sq, const int32_t* gate,
                                void* stream);

/* Arnoldi modified Gram-Schmidt link (arnoldi.py:304-311), one launch per basis vector:
 *   if Qprev: W -= (T)hprev[c] * Qprev;   if Qcur: hcur[c] += sum_i Qcur[i,c] * W[i,c];
 *   if wnorm2: wnorm2[c] += sum_i W[i,c]^2 (last link). */
int cola_mgs_link_f32(float* W, const float* Qprev, const double* hprev, const float* Qcur, double* hcur,
                      double* wnorm2, int64_t n, int64_t b, const int32_t* gate, void* stream);
int cola_mgs_link_f64(double* W, const double* Qprev, const double* hprev, const double* Qcur, double* hcur,
                      double* wnorm2, int64_t n, int64_t b, const int32_t* gate, void* stream);

/* The whole MGS chain of one Arnoldi step (arnoldi.py:304-316) in one cooperative launch:
 *   for j in [0, n_links):  H[j*ldh + c] += sum_i Q_j[i,c] * W[i,c];  W -= (T)H[j*ldh + c] * Q_j     (Q_j = Q + j*q_stride)
 *   then wnorm2[c] += sum_i W[i,c]^2 (may be NULL)
 * in exact modified-Gram-Schmidt order (same arithmetic as n_links + 1 cola_mgs_link_* launches), grid-wide syncs
 * instead of launch boundaries, W kept in L2, each Q_j read from DRAM once.  H rows must be zero on entry.
 * COLA_E_UNSUPPORTED for blocks wider than 256 16-byte vectors or when a cooperative launch is refused: use the links. */
int cola_mgs_chain_f32(float* W, const float* Q, int64_t q_stride, int64_t n_links, double* H, int64_t ldh,
                       double* wnorm2, int64_t n, int64_t b, const int32_t* gate, void* stream);
int cola_mgs_chain_f64(double* W, const double* Q, int64_t q_stride, int64_t n_links, double* H, int64_t ldh,
                       double* wnorm2, int64_t n, int64_t b, const int32_t* gate, void* stream);

/* ---- parameter gradients of the backward passes ---------------------------------------------
 * cg_bwd (cola/linalg/inverse/cg.py:72-86) and slq_bwd (cola/linalg/tbd/slq.py:10-31) end in
 * xnp.vjp_derivs(fun = theta -> A(theta) @ V, primals = theta, duals = G) (cola/backends/torch_fns.py:244-260), which
 * the reference leaves to torch autograd over its eager matmat.  Per leaf of the hot-path operators that vjp is one of
 * the three contractions below; G and V are (n, k) row-major blocks with leading dimensions ldg / ldv. */

/* Sparse values (operators.py:48-81):  out_vals[e] (+)= alpha * sum_c G[row(e), c] * V[colidx[e], c]
 * over the CSR pattern (rowptr, colidx) -- a sampled dense-dense product; out_vals is aligned with `vals`. */
int cola_sddmm_csr_f32(const int32_t* rowptr, const int32_t* colidx, int64_t n_rows, const float* G, int64_t ldg,
                       const float* V, int64_t ldv, int64_t k, float alpha, float* out_vals, int accumulate, void* stream);
int cola_sddmm_csr_f64(const int32_t* rowptr, const int32_t* colidx, int64_t n_rows, const double* G, int64_t ldg,
                       const double* V, int64_t ldv, int64_t k, double alpha, double* out_vals, int accumulate,
                       void* stream);

/* Diagonal (operators.py:323-348) and the bands of a Tridiagonal (:351-372, with row-shifted G / V pointers):
 *   out[i] (+)= alpha * sum_c G[i, c] * V[i, c],  i < n.
 * With out_sq != NULL also  out_sq[i] (+)= sum_c (G[i, c] * V[i, c])^2 : the two running sums of the Hutchinson diagonal
 * estimator, sum_probes z*(Az) and its square (cola/linalg/trace/diagonal_estimation.py:190-199), in one pass. */
int cola_row_dots_f32(const float* G, int64_t ldg, const float* V, int64_t ldv, int64_t n, int64_t k, float alpha, float* out,
                      float* out_sq, int accumulate, void* stream);
int cola_row_dots_f64(const double* G, int64_t ldg, const double* V, int64_t ldv, int64_t n, int64_t k, double alpha,
                      double* out, double* out_sq, int accumulate, void* stream);

/* Dense (operators.py:12-38; pre = 1, post = k: C = G V^T) and one factor of a Kronecker / KronSum (:198-275; G and Z
 * viewed as (pre, d, post) like cola_mode_contract's operands: the "mode Gram" C = sum_p G_p Z_p^T):
 *   C[a, j] += alpha * sum_{p < pre} sum_{t < post} G[(p*d_g + a)*post + t] * Z[(p*d_z + j)*post + t]
 * C is a (d_g, d_z) DOUBLE accumulator with leading dimension ldc that the caller zeroes (split-K, fp64 atomics). */
int cola_gram_nt_f32(const float* G, const float* Z, int64_t d_g, int64_t d_z, int64_t pre, int64_t post, double alpha,
                     double* C, int64_t ldc, void* stream);
int cola_gram_nt_f64(const double* G, const double* Z, int64_t d_g, int64_t d_z, int64_t pre, int64_t post, double alpha,
                     double* C, int64_t ldc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COLA_B200_H */
