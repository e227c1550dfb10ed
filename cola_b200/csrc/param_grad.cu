// Parameter-gradient kernels of the backward passes: the vjp of  theta -> A(theta) @ V  with cotangent G
// that cg_bwd (cola/linalg/inverse/cg.py:72-86) and slq_bwd (cola/linalg/tbd/slq.py:10-31) hand to torch autograd
// (xnp.vjp_derivs, cola/backends/torch_fns.py:244-260).  For the operators of the hot path that vjp is, per leaf:
//   Sparse values   d vals[e] = sum_c G[row(e), c] V[col(e), c]          (sampled dense-dense product, SDDMM)
//   Diagonal        d diag[i] = sum_c G[i, c] V[i, c]                    (row dots)
//   Dense           d M[a, j] = sum_c G[a, c] V[j, c]                    (G V^T)
//   Kronecker / KronSum factor i
//                   d F[a, j] = sum_{p, q} G[p, a, q] Z[p, j, q]         (mode Gram: a sum of pre products G_p Z_p^T)
// The last two are one kernel (gram_nt) with a batch-sum dimension.  Everything accumulates in fp64.
#include "common.cuh"

namespace cola {

// ---- SDDMM on a CSR pattern: one warp per row, the row of G read once per non-zero from L1 ----------------
template <typename T>
__global__ void __launch_bounds__(256)
    sddmm_csr_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, int64_t n_rows,
                     const T* __restrict__ G, int64_t ldg_, const T* __restrict__ V, int64_t ldv, int64_t k, T alpha,
                     T* __restrict__ out, int accumulate) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = warp0; row < n_rows; row += nwarps) {
    const int32_t e0 = rowptr[row], e1 = rowptr[row + 1];
    const T* __restrict__ g = G + row * ldg_;
    for (int32_t e = e0; e < e1; ++e) {
      const T* __restrict__ v = V + (int64_t)colidx[e] * ldv;
      double acc = 0.0;
      for (int64_t c = lane; c < k; c += 32) acc += (double)g[c] * (double)v[c];
      acc = warp_sum(acc);
      if (lane == 0) {
        const T r = alpha * (T)acc;
        out[e] = accumulate ? out[e] + r : r;
      }
    }
  }
}

// ---- row dots: LPR lanes per row (power of two <= 32) --------------------------------------------------------
template <typename T, int LPR>
__global__ void __launch_bounds__(256)
    row_dots_kernel(const T* __restrict__ G, int64_t ldg_, const T* __restrict__ V, int64_t ldv, int64_t n, int64_t k,
                    T alpha, T* __restrict__ out, T* __restrict__ out_sq, int accumulate) {
  const int sub = threadIdx.x % LPR;
  const int64_t r0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const int64_t stride = ((int64_t)gridDim.x * blockDim.x) / LPR;
  for (int64_t row = r0; row < n; row += stride) {
    const T* __restrict__ g = G + row * ldg_;
    const T* __restrict__ v = V + row * ldv;
    double acc = 0.0, acc2 = 0.0;
    for (int64_t c = sub; c < k; c += LPR) {
      const double pr = (double)g[c] * (double)v[c];
      acc += pr;
      acc2 += pr * pr;
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) {
      acc += __shfl_xor_sync(0xffffffffu, acc, o);
      acc2 += __shfl_xor_sync(0xffffffffu, acc2, o);
    }
    if (sub == 0) {
      const T r = alpha * (T)acc;
      out[row] = accumulate ? out[row] + r : r;
      if (out_sq != nullptr) {
        const T r2 = (T)acc2;
        out_sq[row] = accumulate ? out_sq[row] + r2 : r2;
      }
    }
  }
}

// ---- C[a, j] += alpha * sum_p sum_t Gm[(p*da + a)*post + t] * Zm[(p*dz + j)*post + t] -------------------------
// 64x64 output tile per CTA, 4x4 per thread, the reduction range [k_begin, k_end) of the flattened (p, t) index per
// grid.z slice (split-K); fp64 atomics into the caller-zeroed accumulator.
constexpr int kGT = 64, kGK = 32;
template <typename T>
__global__ void __launch_bounds__(256)
    gram_nt_kernel(const T* __restrict__ Gm, const T* __restrict__ Zm, int64_t da, int64_t dz, int64_t pre, int64_t post,
                   double alpha, double* __restrict__ C, int64_t ldc, int64_t k_per_split) {
  __shared__ T Gs[kGK][kGT + 4];
  __shared__ T Zs[kGK][kGT + 4];
  const int64_t a0 = (int64_t)blockIdx.x * kGT, j0 = (int64_t)blockIdx.y * kGT;
  const int64_t K = pre * post;
  const int64_t kb = (int64_t)blockIdx.z * k_per_split;
  const int64_t ke = (kb + k_per_split < K) ? kb + k_per_split : K;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  double acc[4][4] = {};
  for (int64_t kc = kb; kc < ke; kc += kGK) {
    // cooperative load: 64 rows x 32 reduction indices per operand, lanes along the (contiguous) reduction index
#pragma unroll
    for (int i = 0; i < (kGT * kGK) / 256; ++i) {
      const int idx = i * 256 + threadIdx.x;
      const int r = idx / kGK, t = idx % kGK;
      const int64_t kk = kc + t;
      T g = (T)0, z = (T)0;
      if (kk < ke) {
        const int64_t p = kk / post, tt = kk - p * post;
        if (a0 + r < da) g = Gm[(p * da + a0 + r) * post + tt];
        if (j0 + r < dz) z = Zm[(p * dz + j0 + r) * post + tt];
      }
      Gs[t][r] = g;
      Zs[t][r] = z;
    }
    __syncthreads();
#pragma unroll 8
    for (int t = 0; t < kGK; ++t) {
      T ga[4], zb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { ga[i] = Gs[t][ty * 4 + i]; zb[i] = Zs[t][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += (double)ga[i] * (double)zb[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t a = a0 + ty * 4 + i, jj = j0 + tx * 4 + j;
      if (a < da && jj < dz) atomicAdd(C + a * ldc + jj, alpha * acc[i][j]);
    }
}

template <typename T>
static int sddmm_launch(const int32_t* rowptr, const int32_t* colidx, int64_t n_rows, const T* G, int64_t ldg_, const T* V,
                        int64_t ldv, int64_t k, T alpha, T* out, int accumulate, void* stream) {
  COLA_REQUIRE(rowptr && colidx && G && V && out, "sddmm_csr: null pointer");
  COLA_REQUIRE(n_rows >= 0 && k >= 1 && ldg_ >= k && ldv >= k, "sddmm_csr: bad shape");
  if (n_rows == 0) return COLA_OK;
  const int64_t want = (n_rows + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * 16;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  sddmm_csr_kernel<T><<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(rowptr, colidx, n_rows, G, ldg_, V, ldv, k,
                                                                               alpha, out, accumulate);
  return cuda_status("sddmm_csr");
}

template <typename T>
static int row_dots_launch(const T* G, int64_t ldg_, const T* V, int64_t ldv, int64_t n, int64_t k, T alpha, T* out,
                           T* out_sq, int accumulate, void* stream) {
  COLA_REQUIRE(G && V && out, "row_dots: null pointer");
  COLA_REQUIRE(n >= 0 && k >= 1 && ldg_ >= k && ldv >= k, "row_dots: bad shape");
  if (n == 0) return COLA_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int lpr = 1;
  while (lpr < 32 && lpr < k) lpr <<= 1;
  const int64_t want = (n * lpr + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  switch (lpr) {
    case 1: row_dots_kernel<T, 1><<<grid, 256, 0, st>>>(G, ldg_, V, ldv, n, k, alpha, out, out_sq, accumulate); break;
    case 2: row_dots_kernel<T, 2><<<grid, 256, 0, st>>>(G, ldg_, V, ldv, n, k, alpha, out, out_sq, accumulate); break;
    case 4: row_dots_kernel<T, 4><<<grid, 256, 0, st>>>(G, ldg_, V, ldv, n, k, alpha, out, out_sq, accumulate); break;
    case 8: row_dots_kernel<T, 8><<<grid, 256, 0, st>>>(G, ldg_, V, ldv, n, k, alpha, out, out_sq, accumulate); break;
    case 16: row_dots_kernel<T, 16><<<grid, 256, 0, st>>>(G, ldg_, V, ldv, n, k, alpha, out, out_sq, accumulate); break;
    default: row_dots_kernel<T, 32><<<grid, 256, 0, st>>>(G, ldg_, V, ldv, n, k, alpha, out, out_sq, accumulate); break;
  }
  return cuda_status("row_dots");
}

template <typename T>
static int gram_nt_launch(const T* Gm, const T* Zm, int64_t da, int64_t dz, int64_t pre, int64_t post, double alpha, double* C,
                          int64_t ldc, void* stream) {
  COLA_REQUIRE(Gm && Zm && C, "gram_nt: null pointer");
  COLA_REQUIRE(da >= 1 && dz >= 1 && pre >= 1 && post >= 1 && ldc >= dz, "gram_nt: bad shape");
  const int64_t ta = (da + kGT - 1) / kGT, tj = (dz + kGT - 1) / kGT;
  COLA_REQUIRE(tj <= 65535, "gram_nt: output too wide");
  const int64_t K = pre * post;
  // split the reduction until the grid fills the machine twice over; whole kGK chunks per slice
  int64_t splits = ((int64_t)sm_count() * 2 + ta * tj - 1) / (ta * tj);
  const int64_t max_splits = (K + 8 * kGK - 1) / (8 * kGK);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  int64_t per = (K + splits - 1) / splits;
  per = (per + kGK - 1) / kGK * kGK;
  splits = (K + per - 1) / per;
  dim3 grid((unsigned)ta, (unsigned)tj, (unsigned)splits);
  gram_nt_kernel<T><<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(Gm, Zm, da, dz, pre, post, alpha, C, ldc, per);
  return cuda_status("gram_nt");
}

}  // namespace cola

using namespace cola;

extern "C" {

int cola_sddmm_csr_f32(const int32_t* rowptr, const int32_t* colidx, int64_t n_rows, const float* G, int64_t ldg,
                       const float* V, int64_t ldv, int64_t k, float alpha, float* out_vals, int accumulate, void* stream) {
  return sddmm_launch<float>(rowptr, colidx, n_rows, G, ldg, V, ldv, k, alpha, out_vals, accumulate, stream);
}
int cola_sddmm_csr_f64(const int32_t* rowptr, const int32_t* colidx, int64_t n_rows, const double* G, int64_t ldg,
                       const double* V, int64_t ldv, int64_t k, double alpha, double* out_vals, int accumulate,
                       void* stream) {
  return sddmm_launch<double>(rowptr, colidx, n_rows, G, ldg, V, ldv, k, alpha, out_vals, accumulate, stream);
}
int cola_row_dots_f32(const float* G, int64_t ldg, const float* V, int64_t ldv, int64_t n, int64_t k, float alpha, float* out,
                      float* out_sq, int accumulate, void* stream) {
  return row_dots_launch<float>(G, ldg, V, ldv, n, k, alpha, out, out_sq, accumulate, stream);
}
int cola_row_dots_f64(const double* G, int64_t ldg, const double* V, int64_t ldv, int64_t n, int64_t k, double alpha,
                      double* out, double* out_sq, int accumulate, void* stream) {
  return row_dots_launch<double>(G, ldg, V, ldv, n, k, alpha, out, out_sq, accumulate, stream);
}
int cola_gram_nt_f32(const float* G, const float* Z, int64_t d_g, int64_t d_z, int64_t pre, int64_t post, double alpha,
                     double* C, int64_t ldc, void* stream) {
  return gram_nt_launch<float>(G, Z, d_g, d_z, pre, post, alpha, C, ldc, stream);
}
int cola_gram_nt_f64(const double* G, const double* Z, int64_t d_g, int64_t d_z, int64_t pre, int64_t post, double alpha,
                     double* C, int64_t ldc, void* stream) {
  return gram_nt_launch<double>(G, Z, d_g, d_z, pre, post, alpha, C, ldc, stream);
}

}  // extern "C"
