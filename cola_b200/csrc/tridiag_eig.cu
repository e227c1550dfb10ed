// Eigenvalues and first eigenvector components of many small symmetric tridiagonal matrices.
//
// stochastic Lanczos quadrature (cola/linalg/tbd/slq.py:42-51) needs, per probe, the eigenvalues of the
// Lanczos tridiagonal T and tau = first row of its eigenvector matrix: sum_j tau_j^2 f(lambda_j).  The reference
// gets them from a dense batched `eigh` (O(m^3) per probe, 9 % of BASELINE config 4 when done with the library
// routine on the device).  For a tridiagonal matrix the implicit-shift QL iteration delivers exactly these
// quantities in O(m^2): one thread per probe, fp64 throughout, the rotations applied to the first row only
// (Golub-Welsch).  Arrays are laid out [i][probe] (probe fastest) so the threads of a warp touch consecutive
// addresses.
#include "common.cuh"

namespace cola {

// d: in diagonal / out eigenvalues (m x ld); e: in off-diagonal e[i] couples i and i+1 (rows 0..m-2), destroyed;
// z: out first components (m x ld); status[probe] = 0 ok, 1 = an eigenvalue did not converge in 60 sweeps.
__global__ void tridiag_ql_first_row_kernel(double* __restrict__ d, double* __restrict__ e, double* __restrict__ z,
                                            int m, int64_t b, int64_t ld, int32_t* __restrict__ status,
                                            const int32_t* __restrict__ gate) {
  if (gate != nullptr && *gate != 0) return;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= b) return;
#define D(i) d[(int64_t)(i) * ld + t]
#define E(i) e[(int64_t)(i) * ld + t]
#define Z(i) z[(int64_t)(i) * ld + t]
  for (int i = 0; i < m; ++i) Z(i) = (i == 0) ? 1.0 : 0.0;
  E(m - 1) = 0.0;
  int bad = 0;
  const double eps = 2.220446049250313e-16;
  for (int l = 0; l < m; ++l) {
    int iter = 0;
    int mm;
    do {
      for (mm = l; mm < m - 1; ++mm) {
        const double dd = fabs(D(mm)) + fabs(D(mm + 1));
        if (fabs(E(mm)) <= eps * dd) break;
      }
      if (mm != l) {
        if (iter++ == 60) { bad = 1; break; }
        double g = (D(l + 1) - D(l)) / (2.0 * E(l));
        double r = hypot(g, 1.0);
        g = D(mm) - D(l) + E(l) / (g + copysign(r, g));
        double s = 1.0, c = 1.0, p = 0.0;
        int i;
        for (i = mm - 1; i >= l; --i) {
          double f = s * E(i);
          const double bb = c * E(i);
          r = hypot(f, g);
          E(i + 1) = r;
          if (r == 0.0) {
            D(i + 1) -= p;
            E(mm) = 0.0;
            break;
          }
          s = f / r;
          c = g / r;
          g = D(i + 1) - p;
          r = (D(i) - g) * s + 2.0 * c * bb;
          p = s * r;
          D(i + 1) = g + p;
          g = c * r - bb;
          f = Z(i + 1);
          Z(i + 1) = s * Z(i) + c * f;
          Z(i) = c * Z(i) - s * f;
        }
        if (r == 0.0 && i >= l) continue;
        D(l) -= p;
        E(l) = g;
        E(mm) = 0.0;
      }
    } while (mm != l);
    if (bad) break;
  }
  if (status != nullptr) status[t] = bad;
#undef D
#undef E
#undef Z
}

}  // namespace cola

using namespace cola;
extern "C" {
int cola_tridiag_eig_first_row_f64(double* d, double* e, double* z, int64_t m, int64_t b, int64_t ld, int32_t* status,
                                   const int32_t* gate, void* stream) {
  COLA_REQUIRE(d && e && z, "tridiag_eig_first_row: null pointer");
  COLA_REQUIRE(m >= 1 && m < (1 << 20) && ld >= b, "tridiag_eig_first_row: bad sizes");
  if (b <= 0) return COLA_OK;
  const int threads = 32;   // one warp per CTA: the probes of a chunk spread over as many SMs as possible
  const int64_t blocks = (b + threads - 1) / threads;
  tridiag_ql_first_row_kernel<<<(unsigned)blocks, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      d, e, z, (int)m, b, ld, status, gate);
  return cuda_status("tridiag_eig_first_row");
}
}
