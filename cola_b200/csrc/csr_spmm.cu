// CSR SpMM over a block of right-hand sides with the fused operator epilogue and the p^T A p / <w,v>
// column dots in the same pass.  Replaces Sparse._matmat (cola/ops/operators.py:77-78) and, inside CG,
// the separate `sum(conj(p) * Ap)` reduction (cola/linalg/inverse/cg.py:157-158).
//
// Mapping (HBM/L2-bound integer+float gather work; no tensor cores on purpose):
//   * a CTA owns a tile of consecutive rows; it first stages the tile's rowptr slice and its whole
//     colidx/vals range into shared memory with coalesced loads, which removes the
//     rowptr -> colidx -> X dependent-load chain from the inner loop;
//   * a group of `lanes` threads owns one row at a time; each thread owns VEC consecutive RHS columns, so a
//     row of X is fetched as one contiguous k*sizeof(T) segment (256 B for 64 fp32 RHS) per non-zero and
//     neighbouring rows' segments hit in L1/L2;
//   * Y is written once, streaming; per-column dot partials stay in registers for the whole kernel.
// Algorithmic HBM bytes: nnz*(sizeof(T)+4) + 4(n+1) + 2*n*k*sizeof(T)   (SURVEY.md section 8d).
#include "sweep.cuh"

namespace cola {

constexpr int kCsrThreads = 256;

template <typename T>
struct CsrArgs {
  const int32_t* rowptr; const int32_t* colidx; const T* vals;
  int64_t n_rows;
  const T* X; int64_t ldx; int64_t k;
  T* Y; int64_t ldy;
  T alpha, shift; const T* diag; int accumulate;
  double* dots; const int32_t* dots_row; int64_t k_full; const int32_t* gate;
  int lanes, groups, rows_per_tile, cap;
  int need_x;
};

template <typename T, int VEC>
__global__ void __launch_bounds__(kCsrThreads) csr_spmm_kernel(CsrArgs<T> a) {
  if (a.gate != nullptr && *a.gate != 0) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: vals[cap] | cols[cap] | rp[rows_per_tile+1] | red[256*VEC doubles]
  T* s_vals = reinterpret_cast<T*>(smem_raw);
  int32_t* s_cols = reinterpret_cast<int32_t*>(s_vals + a.cap);
  int32_t* s_rp = s_cols + a.cap;
  const int tid = threadIdx.x;
  const int g = tid / a.lanes, l = tid - g * a.lanes;
  const int64_t c0 = (int64_t)l * VEC;
  const bool col_ok = (g < a.groups) && (c0 < a.k);

  double dacc[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) dacc[v] = 0.0;

  const int64_t n_tiles = (a.n_rows + a.rows_per_tile - 1) / a.rows_per_tile;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * a.rows_per_tile;
    const int rows = (int)min((int64_t)a.rows_per_tile, a.n_rows - row0);
    __syncthreads();  // previous tile's consumers are done with smem
    for (int i = tid; i <= rows; i += kCsrThreads) s_rp[i] = a.rowptr[row0 + i];
    __syncthreads();
    const int32_t base = s_rp[0];
    const int32_t tile_nnz = s_rp[rows] - base;
    const bool staged = tile_nnz <= a.cap;
    if (staged) {
      for (int i = tid; i < tile_nnz; i += kCsrThreads) {
        s_cols[i] = __ldcs(a.colidx + base + i);
        s_vals[i] = __ldcs(a.vals + base + i);
      }
    }
    __syncthreads();
    if (!col_ok) continue;
    for (int r = g; r < rows; r += a.groups) {
      const int64_t row = row0 + r;
      const int32_t s = s_rp[r] - base, e = s_rp[r + 1] - base;
      T acc[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[v] = (T)0;
      if (staged) {
#pragma unroll 4
        for (int32_t j = s; j < e; ++j) {
          const int64_t c = s_cols[j];
          const T w = s_vals[j];
          Vec<T, VEC> x = ldg<T, VEC>(a.X + c * a.ldx + c0);
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[v] += w * x.v[v];
        }
      } else {  // tile too heavy for the staging buffer: read the CSR arrays directly
#pragma unroll 4
        for (int32_t j = s; j < e; ++j) {
          const int64_t c = a.colidx[base + j];
          const T w = a.vals[base + j];
          Vec<T, VEC> x = ldg<T, VEC>(a.X + c * a.ldx + c0);
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[v] += w * x.v[v];
        }
      }
      Vec<T, VEC> y, xo;
      if (a.need_x) xo = ldg<T, VEC>(a.X + row * a.ldx + c0);
      const T d = a.diag ? a.diag[row] : (T)0;
      Vec<T, VEC> yo;
      if (a.accumulate) yo = ldg<T, VEC>(a.Y + row * a.ldy + c0);
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        T t = a.alpha * acc[v];
        if (a.shift != (T)0) t += a.shift * xo.v[v];
        if (a.diag) t += d * xo.v[v];
        if (a.accumulate) t += yo.v[v];
        y.v[v] = t;
        if (a.dots) dacc[v] += (double)xo.v[v] * (double)t;
      }
      stg_stream<T, VEC>(a.Y + row * a.ldy + c0, y);
    }
  }

  if (a.dots) {
    double* red = reinterpret_cast<double*>(s_rp + a.rows_per_tile + 2);
    red = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(red) + 7) & ~(uintptr_t)7);
    double* out = a.dots + (a.dots_row ? (int64_t)(*a.dots_row) * a.k_full : 0);
    block_col_reduce<VEC>(red, dacc, col_ok, tid, g, l, a.lanes, a.groups, c0, a.k, -1, out);
  }
}

template <typename T>
int csr_spmm(const int32_t* rowptr, const int32_t* colidx, const T* vals, int64_t n_rows, int64_t n_cols,
             int64_t nnz, const T* X, int64_t ldx, int64_t k, T* Y, int64_t ldy, T alpha, T shift, const T* diag,
             int accumulate, double* dots, const int32_t* dots_row, const int32_t* gate, cudaStream_t st) {
  COLA_REQUIRE(rowptr && colidx && vals && X && Y, "csr_spmm: null pointer");
  COLA_REQUIRE(ldx >= k && ldy >= k, "csr_spmm: leading dimension < k");
  COLA_REQUIRE(X != Y, "csr_spmm: X and Y must not alias");
  const bool epi = (shift != (T)0) || diag || dots;
  COLA_REQUIRE(!epi || n_rows == n_cols, "csr_spmm: shift/diag/dots need a square operator");
  if (n_rows <= 0 || k <= 0) return COLA_OK;
  int vec = pick_vec<T>(k, ldx, X, Y);
  if (ldy % vec) vec = 1;
  const int64_t slab_max = (int64_t)kCsrThreads;  // at most 256 lanes -> 256*VEC columns per launch
  int rc = COLA_OK;
  for (int64_t c = 0; c < k && rc == COLA_OK; c += slab_max * vec) {
    CsrArgs<T> a;
    a.rowptr = rowptr; a.colidx = colidx; a.vals = vals; a.n_rows = n_rows;
    a.X = X + c; a.ldx = ldx; a.k = (k - c < slab_max * vec) ? (k - c) : slab_max * vec;
    a.Y = Y + c; a.ldy = ldy; a.alpha = alpha; a.shift = shift; a.diag = diag; a.accumulate = accumulate;
    a.dots = dots ? dots + c : nullptr; a.dots_row = dots_row; a.k_full = k; a.gate = gate;
    a.need_x = epi ? 1 : 0;
    int64_t need = (a.k + vec - 1) / vec;
    a.lanes = (int)need;
    a.groups = kCsrThreads / a.lanes;
    // tile size: ~6 rows per group in flight, but keep the staged nnz range near half the buffer
    double avg = n_rows > 0 ? (double)nnz / (double)n_rows : 1.0;
    if (avg < 1.0) avg = 1.0;
    const int cap_max = sizeof(T) == 4 ? 4096 : 2048;   // staging buffer stays under the 48 KB default
    int rpg = (int)((cap_max / 2) / (avg * a.groups));
    if (rpg < 1) rpg = 1;
    if (rpg > 8) rpg = 8;
    a.rows_per_tile = a.groups * rpg;
    int64_t want = (int64_t)(2.0 * avg * a.rows_per_tile) + 64;   // 2x the mean tile, heavier tiles take the direct path
    a.cap = (int)(want < 512 ? 512 : (want > cap_max ? cap_max : want));
    size_t smem = (size_t)a.cap * (sizeof(T) + 4) + (size_t)(a.rows_per_tile + 4) * 4 + 8 +
                  (dots ? (size_t)kCsrThreads * vec * sizeof(double) : 0);
    int64_t n_tiles = (n_rows + a.rows_per_tile - 1) / a.rows_per_tile;
    COLA_DISPATCH_VEC(T, vec, ({
      int per_sm = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, csr_spmm_kernel<T, VEC>, kCsrThreads, smem);
      if (per_sm < 1) per_sm = 1;
      int64_t grid = (int64_t)sm_count() * per_sm;   // persistent: whole waves of resident CTAs
      if (grid > n_tiles) grid = n_tiles;
      csr_spmm_kernel<T, VEC><<<(unsigned)grid, kCsrThreads, smem, st>>>(a);
    }));
    rc = cuda_status("csr_spmm");
  }
  return rc;
}

}  // namespace cola

using namespace cola;
extern "C" {
int cola_csr_spmm_f32(const int32_t* rowptr, const int32_t* colidx, const float* vals, int64_t n_rows,
                      int64_t n_cols, int64_t nnz, const float* X, int64_t ldx, int64_t k, float* Y, int64_t ldy,
                      float alpha, float shift, const float* diag, int accumulate, double* dots,
                      const int32_t* dots_row, const int32_t* gate, void* stream) {
  return csr_spmm<float>(rowptr, colidx, vals, n_rows, n_cols, nnz, X, ldx, k, Y, ldy, alpha, shift, diag, accumulate,
                         dots, dots_row, gate, reinterpret_cast<cudaStream_t>(stream));
}
int cola_csr_spmm_f64(const int32_t* rowptr, const int32_t* colidx, const double* vals, int64_t n_rows,
                      int64_t n_cols, int64_t nnz, const double* X, int64_t ldx, int64_t k, double* Y, int64_t ldy,
                      double alpha, double shift, const double* diag, int accumulate, double* dots,
                      const int32_t* dots_row, const int32_t* gate, void* stream) {
  return csr_spmm<double>(rowptr, colidx, vals, n_rows, n_cols, nnz, X, ldx, k, Y, ldy, alpha, shift, diag,
                          accumulate, dots, dots_row, gate, reinterpret_cast<cudaStream_t>(stream));
}
}
