// Vector sweeps of the Krylov loops: column dots / scalings, CG x-r-p updates and control block,
// Lanczos three-term step, Arnoldi MGS links, core-less (diagonal) matmat.  All are instances of
// sweep_kernel (sweep.cuh): one HBM pass over each operand, fp64 per-column reductions.
#include "sweep.cuh"

namespace cola {

thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

int sm_count() {
  static int cached = 0;
  if (cached) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) == cudaSuccess &&
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) {
    cached = n;
    return n;
  }
  cudaGetLastError();
  return kSMsFallback;
}

// Shape of a sweep.  An (n,1) block with ld==1 is a flat vector: view it as (n/VEC, VEC) so accesses stay
// 16 bytes wide, and fold every lane onto column 0 for scalars / reductions (colmask = 0).
struct Shape {
  int64_t n, k, ld;
  int vec;
  int64_t colmask;
};
template <typename T>
Shape shape_of(int64_t n, int64_t k, int64_t ld, const void* a, const void* b = nullptr, const void* c = nullptr,
               const void* d = nullptr) {
  Shape s{n, k, ld, 1, (int64_t)-1};
  if (k == 1 && ld == 1) {
    for (int v = 16 / (int)sizeof(T); v > 1; v >>= 1) {
      if (n % v == 0 && pick_vec<T>(v, v, a, b, c, d) == v) return Shape{n / v, v, v, v, 0};
    }
    return s;
  }
  s.vec = pick_vec<T>(k, ld, a, b, c, d);
  return s;
}

__device__ __forceinline__ bool gate_open(const int32_t* gate) { return gate == nullptr || *gate == 0; }

// ---- dots[c] += sum_i X[i,c] * Y[i,c] -------------------------------------------------------------
template <typename T, int VEC>
struct DotsOp {
  static constexpr int NACC = 1;
  struct Regs { Vec<T, VEC> x, y; };
  const T* X; const T* Y; int64_t ld; double* dots; const int32_t* gate; bool same;
  __device__ bool enabled() const { return gate_open(gate); }
  __device__ void setup(int64_t) {}
  __device__ void load(int64_t row, int64_t c0, Regs& r) const {
    r.x = ldg_stream<T, VEC>(X + row * ld + c0);
    if (!same) r.y = ldg_stream<T, VEC>(Y + row * ld + c0);
  }
  __device__ void finish(int64_t, int64_t, Regs& r, double (&acc)[1][VEC]) const {
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[0][v] += (double)r.x.v[v] * (double)(same ? r.x.v[v] : r.y.v[v]);
  }
  __device__ double* out(int) const { return dots; }
};

template <typename T>
int col_dots(const T* X, const T* Y, int64_t n, int64_t k, int64_t ld, double* dots, const int32_t* gate,
             cudaStream_t st) {
  COLA_REQUIRE(X && Y && dots, "col_dots: null pointer");
  COLA_REQUIRE(ld >= k, "col_dots: ld < k");
  Shape s = shape_of<T>(n, k, ld, X, Y);
  int rc = COLA_OK;
  COLA_DISPATCH_VEC(T, s.vec, (rc = launch_sweep<T, VEC>(s.n, s.k, s.colmask, st, [&](int64_t c) {
    return DotsOp<T, VEC>{X + c, Y + c, s.ld, dots + (c & s.colmask), gate, X == Y};
  }, "col_dots")));
  return rc;
}

// ---- Y = a * X * s[c]   or   a * X / safe(s[c]) ------------------------------------------------------
template <typename T, int VEC>
struct ScaleOp {
  static constexpr int NACC = 0;
  struct Regs { Vec<T, VEC> x; };
  const T* X; T* Y; int64_t ld; const double* sq; int take_sqrt, mode; T a; int64_t colmask; const int32_t* gate;
  T f[VEC];
  __device__ bool enabled() const { return gate_open(gate); }
  __device__ void setup(int64_t c0) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      double s = sq[(c0 + v) & colmask];
      T sc = (T)(take_sqrt ? sqrt(s) : s);
      f[v] = (mode == 0) ? a * sc : safe_div<T>(a, sc);
    }
  }
  __device__ void load(int64_t row, int64_t c0, Regs& r) const { r.x = ldg_stream<T, VEC>(X + row * ld + c0); }
  __device__ void finish(int64_t row, int64_t c0, Regs& r, double (&)[1][VEC]) const {
    Vec<T, VEC> y;
    // mode 1 mirrors `num / denom` elementwise (cg.py:177): divide, do not multiply by a reciprocal
#pragma unroll
    for (int v = 0; v < VEC; ++v) y.v[v] = r.x.v[v] * f[v];
    stg<T, VEC>(Y + row * ld + c0, y);
  }
  __device__ double* out(int) const { return nullptr; }
};

// exact-division variant used for the RHS normalisation b / ||b|| (cg.py:96-97)
template <typename T, int VEC>
struct DivOp {
  static constexpr int NACC = 0;
  struct Regs { Vec<T, VEC> x; };
  const T* X; T* Y; int64_t ld; const double* sq; int take_sqrt; int mode; T floor_; int64_t colmask;
  const int32_t* gate;
  T d[VEC];
  __device__ bool enabled() const { return gate_open(gate); }
  __device__ void setup(int64_t c0) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      double s = sq[(c0 + v) & colmask];
      T sc = (T)(take_sqrt ? sqrt(s) : s);
      if (mode == 1) d[v] = (fabs((double)sc) < 1e-40) ? (T)1e-40 : sc;  // do_safe_div, cg.py:173-178
      else if (mode == 3) d[v] = sc < floor_ ? floor_ : sc;              // clip(norm, min), arnoldi.py:315
      else d[v] = sc;                                                    // plain division, lanczos.py:240-241
    }
  }
  __device__ void load(int64_t row, int64_t c0, Regs& r) const { r.x = ldg_stream<T, VEC>(X + row * ld + c0); }
  __device__ void finish(int64_t row, int64_t c0, Regs& r, double (&)[1][VEC]) const {
    Vec<T, VEC> y;
#pragma unroll
    for (int v = 0; v < VEC; ++v) y.v[v] = r.x.v[v] / d[v];
    stg<T, VEC>(Y + row * ld + c0, y);
  }
  __device__ double* out(int) const { return nullptr; }
};

template <typename T>
int col_scale(const T* X, T* Y, int64_t n, int64_t k, int64_t ld, const double* sq, int take_sqrt, int mode, T a,
              const int32_t* gate, cudaStream_t st) {
  COLA_REQUIRE(X && Y && sq, "col_scale: null pointer");
  COLA_REQUIRE(mode >= 0 && mode <= 3, "col_scale: mode must be 0..3");
  Shape s = shape_of<T>(n, k, ld, X, Y);
  int rc = COLA_OK;
  if (mode == 0) {
    COLA_DISPATCH_VEC(T, s.vec, (rc = launch_sweep<T, VEC>(s.n, s.k, s.colmask, st, [&](int64_t c) {
      return ScaleOp<T, VEC>{X + c, Y + c, s.ld, sq + (c & s.colmask), take_sqrt, 0, a, s.colmask, gate, {}};
    }, "col_scale")));
  } else {
    COLA_DISPATCH_VEC(T, s.vec, (rc = launch_sweep<T, VEC>(s.n, s.k, s.colmask, st, [&](int64_t c) {
      return DivOp<T, VEC>{X + c, Y + c, s.ld, sq + (c & s.colmask), take_sqrt, mode, a, s.colmask, gate, {}};
    }, "col_div")));
  }
  return rc;
}

// ---- Y = a X + b Y -----------------------------------------------------------------------------------
template <typename T, int VEC>
struct AxpbyOp {
  static constexpr int NACC = 0;
  struct Regs { Vec<T, VEC> x, y; };
  const T* X; T* Y; int64_t ld; T a, b; const int32_t* gate;
  __device__ bool enabled() const { return gate_open(gate); }
  __device__ void setup(int64_t) {}
  __device__ void load(int64_t row, int64_t c0, Regs& r) const {
    r.x = ldg_stream<T, VEC>(X + row * ld + c0);
    if (b != (T)0) r.y = ldg_stream<T, VEC>(Y + row * ld + c0);
  }
  __device__ void finish(int64_t row, int64_t c0, Regs& r, double (&)[1][VEC]) const {
    Vec<T, VEC> y;
#pragma unroll
    for (int v = 0; v < VEC; ++v) y.v[v] = (b != (T)0) ? a * r.x.v[v] + b * r.y.v[v] : a * r.x.v[v];
    stg<T, VEC>(Y + row * ld + c0, y);
  }
  __device__ double* out(int) const { return nullptr; }
};

template <typename T>
int axpby(const T* X, T* Y, int64_t n, int64_t k, int64_t ld, T a, T b, const int32_t* gate, cudaStream_t st) {
  COLA_REQUIRE(X && Y, "axpby: null pointer");
  Shape s = shape_of<T>(n, k, ld, X, Y);
  int rc = COLA_OK;
  COLA_DISPATCH_VEC(T, s.vec, (rc = launch_sweep<T, VEC>(s.n, s.k, s.colmask, st, [&](int64_t c) {
    return AxpbyOp<T, VEC>{X + c, Y + c, s.ld, a, b, gate};
  }, "axpby")));
  return rc;
}

// ---- core-less matmat: Y = (shift + diag[i]) X (+Y), dots += <X,Y> ----------------------------------
template <typename T, int VEC>
struct DiagOp {
  static constexpr int NACC = 1;
  struct Regs { Vec<T, VEC> x, y; T d; };
  const T* X; T* Y; int64_t ldx, ldy; T shift; const T* diag; int accumulate; double* dots; const int32_t* gate;
  const int32_t* dots_row; int64_t k_full;
  __device__ bool enabled() const { return gate_open(gate); }
  __device__ void setup(int64_t) {}
  __device__ void load(int64_t row, int64_t c0, Regs& r) const {
    r.x = ldg_stream<T, VEC>(X + row * ldx + c0);
    if (accumulate) r.y = ldg_stream<T, VEC>(Y + row * ldy + c0);
    r.d = diag ? diag[row] : (T)0;
  }
  __device__ void finish(int64_t row, int64_t c0, Regs& r, double (&acc)[1][VEC]) const {
    Vec<T, VEC> y;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      T t = (diag ? r.d * r.x.v[v] : (T)0);
      if (shift != (T)0) t = diag ? t + shift * r.x.v[v] : shift * r.x.v[v];
      y.v[v] = accumulate ? r.y.v[v] + t : t;
      acc[0][v] += (double)r.x.v[v] * (double)y.v[v];
    }
    stg<T, VEC>(Y + row * ldy + c0, y);
  }
  __device__ double* out(int) const {
    if (!dots) return nullptr;
    return dots + (dots_row ? (int64_t)(*dots_row) * k_full : 0);
  }
};

template <typename T>
int diag_matmat(const T* X, int64_t ldx, T* Y, int64_t ldy, int64_t n, int64_t k, T shift, const T* diag,
                int accumulate, double* dots, const int32_t* dots_row, const int32_t* gate, cudaStream_t st) {
  COLA_REQUIRE(X && Y, "diag_matmat: null pointer");
  int vec = pick_vec<T>(k, ldx, X, Y);
  if (ldy % vec) vec = 1;
  int rc = COLA_OK;
  COLA_DISPATCH_VEC(T, vec, (rc = launch_sweep<T, VEC>(n, k, -1, st, [&](int64_t c) {
    return DiagOp<T, VEC>{X + c, Y + c, ldx, ldy, shift, diag, accumulate, dots ? dots + c : nullptr, gate,
                          dots_row, k};
  }, "diag_matmat")));
  return rc;
}

// ---- CG iteration minus the matmat, as two sweeps (10 vector passes per iteration with the matmat's 2,
// instead of the 11 of the textbook split x/r-update + p-update: P is read once, by the sweep that needs the
// old direction for BOTH x += alpha p and p = r + beta p).
//   r-sweep : alpha = safe(gamma/pAp) (0 where ||r|| < 1e-40);  R -= alpha AP;  gamma[it+1] += <R,R>
//   xp-sweep: beta = safe(gamma[it+1]/gamma[it]) (0 where converged);  X += alpha P;  P = R + beta P
// (cg.py:141-170; alpha/beta/has_converged are recomputed per thread from the device-resident accumulators)
template <typename T>
__device__ __forceinline__ T cg_alpha(const double* gamma, const double* pAp, int64_t it, int64_t k_full, int64_t c) {
  const T g = (T)gamma[it * k_full + c];
  const T q = (T)pAp[it * k_full + c];
  const bool conv = (T)sqrt(gamma[it * k_full + c]) < (T)1e-40;  // has_converged, cg.py:144
  return conv ? (T)0 : safe_div<T>(g, q);
}

template <typename T, int VEC>
struct CgROp {
  static constexpr int NACC = 1;
  struct Regs { Vec<T, VEC> r, ap; };
  T* R; const T* AP; int64_t ld; const cola_cg_ctl_t* ctl; const double* gamma; const double* pAp;
  double* gamma_w; int64_t k_full; int64_t colmask;
  T alpha[VEC];
  __device__ bool enabled() const { return ctl->done == 0; }
  __device__ void setup(int64_t c0) {
    const int64_t it = ctl->it;
#pragma unroll
    for (int v = 0; v < VEC; ++v) alpha[v] = cg_alpha<T>(gamma, pAp, it, k_full, (c0 + v) & colmask);
  }
  __device__ void load(int64_t row, int64_t c0, Regs& r) const {
    const int64_t o = row * ld + c0;
    r.r = ldg_stream<T, VEC>(R + o);
    r.ap = ldg_stream<T, VEC>(AP + o);
  }
  __device__ void finish(int64_t row, int64_t c0, Regs& r, double (&acc)[1][VEC]) const {
    Vec<T, VEC> rn;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      rn.v[v] = r.r.v[v] - alpha[v] * r.ap.v[v];
      acc[0][v] += (double)rn.v[v] * (double)rn.v[v];
    }
    stg<T, VEC>(R + row * ld + c0, rn);
  }
  __device__ double* out(int) const { return gamma_w + ((int64_t)ctl->it + 1) * k_full; }
};

template <typename T>
int cg_update_r(T* R, const T* AP, int64_t n, int64_t k, int64_t ld, const cola_cg_ctl_t* ctl, const double* gamma,
                const double* pAp, double* gamma_w, cudaStream_t st) {
  COLA_REQUIRE(R && AP && ctl && gamma && pAp && gamma_w, "cg_update_r: null pointer");
  Shape s = shape_of<T>(n, k, ld, R, AP);
  int rc = COLA_OK;
  COLA_DISPATCH_VEC(T, s.vec, (rc = launch_sweep<T, VEC>(s.n, s.k, s.colmask, st, [&](int64_t c) {
    int64_t cc = c & s.colmask;
    return CgROp<T, VEC>{R + c, AP + c, s.ld, ctl, gamma + cc, pAp + cc, gamma_w + cc, k, s.colmask, {}};
  }, "cg_update_r")));
  return rc;
}

template <typename T, int VEC>
struct CgXpOp {
  static constexpr int NACC = 0;
  struct Regs { Vec<T, VEC> x, r, p; };
  T* X; const T* R; T* P; int64_t ld; const cola_cg_ctl_t* ctl; const double* gamma; const double* pAp;
  int64_t k_full; int64_t colmask;
  T alpha[VEC], beta[VEC];
  __device__ bool enabled() const { return ctl->done == 0; }
  __device__ void setup(int64_t c0) {
    const int64_t it = ctl->it;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int64_t c = (c0 + v) & colmask;
      alpha[v] = cg_alpha<T>(gamma, pAp, it, k_full, c);
      const T g0 = (T)gamma[it * k_full + c];
      const T g1 = (T)gamma[(it + 1) * k_full + c];
      const bool conv = (T)sqrt(gamma[it * k_full + c]) < (T)1e-40;
      beta[v] = conv ? (T)0 : safe_div<T>(g1, g0);
    }
  }
  __device__ void load(int64_t row, int64_t c0, Regs& r) const {
    const int64_t o = row * ld + c0;
    r.x = ldg_stream<T, VEC>(X + o);
    r.r = ldg_stream<T, VEC>(R + o);
    r.p = ldg_stream<T, VEC>(P + o);
  }
  __device__ void finish(int64_t row, int64_t c0, Regs& r, double (&)[1][VEC]) const {
    const int64_t o = row * ld + c0;
    Vec<T, VEC> xn, pn;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      xn.v[v] = r.x.v[v] + alpha[v] * r.p.v[v];
      pn.v[v] = r.r.v[v] + beta[v] * r.p.v[v];
    }
    stg_stream<T, VEC>(X + o, xn);
    stg<T, VEC>(P + o, pn);   // P is gathered by the matmat right after: keep it cacheable
  }
  __device__ double* out(int) const { return nullptr; }
};

template <typename T>
int cg_update_xp(T* X, const T* R, T* P, int64_t n, int64_t k, int64_t ld, const cola_cg_ctl_t* ctl,
                 const double* gamma, const double* pAp, cudaStream_t st) {
  COLA_REQUIRE(X && R && P && ctl && gamma && pAp, "cg_update_xp: null pointer");
  Shape s = shape_of<T>(n, k, ld, X, R, P);
  int rc = COLA_OK;
  COLA_DISPATCH_VEC(T, s.vec, (rc = launch_sweep<T, VEC>(s.n, s.k, s.colmask, st, [&](int64_t c) {
    int64_t cc = c & s.colmask;
    return CgXpOp<T, VEC>{X + c, R + c, P + c, s.ld, ctl, gamma + cc, pAp + cc, k, s.colmask, {}, {}};
  }, "cg_update_xp")));
  return rc;
}

// ---- CG control block ---------------------------------------------------------------------------------
template <typename T>
__global__ void cg_tol_kernel(const double* gamma0, T tol, T* tol_eff, int64_t k) {
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c < k) tol_eff[c] = tol * (T)sqrt(gamma0[c]) + tol;  // cg.py:101
}

template <typename T>
__global__ void cg_advance_kernel(cola_cg_ctl_t* ctl, const double* gamma, const T* tol_eff, int increment) {
  __shared__ int any_large;
  if (ctl->done) return;
  if (threadIdx.x == 0) any_large = 0;
  __syncthreads();
  const int it = ctl->it + (increment ? 1 : 0);
  const int k = ctl->k;
  int mine = 0;
  for (int c = threadIdx.x; c < k; c += blockDim.x) {
    T rs = (T)sqrt(gamma[(int64_t)it * k + c]);
    // `rs > tol` is false for NaN, exactly like torch.any(rs > tol)  (cg.py:133-138)
    if (rs > tol_eff[c]) mine = 1;
  }
  if (mine) atomicOr(&any_large, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    ctl->it = it;
    ctl->done = (any_large && it < ctl->max_iters) ? 0 : 1;
  }
}

// ---- small device -> mapped-pinned-host publication (poll of the stopping rules without a DMA transfer) --------
__global__ void publish_kernel(const uint32_t* __restrict__ src, volatile uint32_t* dst, int64_t n_words) {
  for (int64_t i = threadIdx.x; i < n_words; i += blockDim.x) dst[i] = src[i];
  __threadfence_system();
}

// ---- four 32-bit words from kernel arguments into device memory (control blocks built without a DMA transfer) ----
__global__ void store_i32x4_kernel(int32_t* dst, int32_t a, int32_t b, int32_t c, int32_t d) {
  dst[0] = a; dst[1] = b; dst[2] = c; dst[3] = d;
}

// ---- Lanczos three-term step: W -= alpha Vi + beta_prev Vim1   (lanczos.py:245-248) ---------------
template <typename T, int VEC>
struct ThreeTermOp {
  static constexpr int NACC = 0;
  struct Regs { Vec<T, VEC> w, a, b; };
  T* W; const T* Vi; const T* Vim1; int64_t ld; const double* alpha_acc; const double* beta_prev_sq;
  int64_t colmask; const int32_t* gate;
  T al[VEC], be[VEC];
  __device__ bool enabled() const { return gate_open(gate); }
  __device__ void setup(int64_t c0) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      int64_t c = (c0 + v) & colmask;
      al[v] = (T)alpha_acc[c];
      be[v] = beta_prev_sq ? (T)sqrt(beta_prev_sq[c]) : (T)0;
    }
  }
  __device__ void load(int64_t row, int64_t c0, Regs& r) const {
    const int64_t o = row * ld + c0;
    r.w = ldg_stream<T, VEC>(W + o);
    r.a = ldg_stream<T, VEC>(Vi + o);
    if (Vim1) r.b = ldg_stream<T, VEC>(Vim1 + o);
  }
  __device__ void finish(int64_t row, int64_t c0, Regs& r, double (&)[1][VEC]) const {
    Vec<T, VEC> w;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      // aux = diag*V_i + subdiag*V_{i-1}; new_vec -= aux  (same association as the reference)
      T aux = al[v] * r.a.v[v];
      if (Vim1) aux = aux + be[v] * r.b.v[v];
      w.v[v] = r.w.v[v] - aux;
    }
    stg<T, VEC>(W + row * ld + c0, w);
  }
  __device__ double* out(int) const { return nullptr; }
};

template <typename T>
int lanczos_three_term(T* W, const T* Vi, const T* Vim1, int64_t n, int64_t b, const double* alpha_acc,
                       const double* beta_prev_sq, const int32_t* gate, cudaStream_t st) {
  COLA_REQUIRE(W && Vi && alpha_acc, "lanczos_three_term: null pointer");
  Shape s = shape_of<T>(n, b, b, W, Vi, Vim1);
  int rc = COLA_OK;
  COLA_DISPATCH_VEC(T, s.vec, (rc = launch_sweep<T, VEC>(s.n, s.k, s.colmask, st, [&](int64_t c) {
    int64_t cc = c & s.colmask;
    return ThreeTermOp<T, VEC>{W + c, Vi + c, Vim1 ? Vim1 + c : nullptr, s.ld, alpha_acc + cc,
                               beta_prev_sq ? beta_prev_sq + cc : nullptr, s.colmask, gate, {}, {}};
  }, "lanczos_three_term")));
  return rc;
}

// ---- Arnoldi MGS link: W -= hprev Qprev;  hcur += <Qcur, W>;  wnorm2 += <W,W>   (arnoldi.py:304-316) --
template <typename T, int VEC>
struct MgsOp {
  static constexpr int NACC = 2;
  struct Regs { Vec<T, VEC> w, qp, qc; };
  T* W; const T* Qprev; const double* hprev; const T* Qcur; double* hcur; double* wnorm2; int64_t ld;
  int64_t colmask; const int32_t* gate;
  T hp[VEC];
  __device__ bool enabled() const { return gate_open(gate); }
  __device__ void setup(int64_t c0) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) hp[v] = Qprev ? (T)hprev[(c0 + v) & colmask] : (T)0;
  }
  __device__ void load(int64_t row, int64_t c0, Regs& r) const {
    const int64_t o = row * ld + c0;
    r.w = ldg_stream<T, VEC>(W + o);
    if (Qprev) r.qp = ldg_stream<T, VEC>(Qprev + o);
    if (Qcur) r.qc = ldg_stream<T, VEC>(Qcur + o);
  }
  __device__ void finish(int64_t row, int64_t c0, Regs& r, double (&acc)[2][VEC]) const {
    Vec<T, VEC> w = r.w;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      if (Qprev) w.v[v] = r.w.v[v] - hp[v] * r.qp.v[v];
      if (Qcur) acc[0][v] += (double)r.qc.v[v] * (double)w.v[v];
      acc[1][v] += (double)w.v[v] * (double)w.v[v];
    }
    if (Qprev) stg<T, VEC>(W + row * ld + c0, w);
  }
  __device__ double* out(int a) const { return a == 0 ? (Qcur ? hcur : nullptr) : wnorm2; }
};

template <typename T>
int mgs_link(T* W, const T* Qprev, const double* hprev, const T* Qcur, double* hcur, double* wnorm2, int64_t n,
             int64_t b, const int32_t* gate, cudaStream_t st) {
  COLA_REQUIRE(W, "mgs_link: null W");
  COLA_REQUIRE(!Qprev || hprev, "mgs_link: Qprev needs hprev");
  COLA_REQUIRE(!Qcur || hcur, "mgs_link: Qcur needs hcur");
  Shape s = shape_of<T>(n, b, b, W, Qprev, Qcur);
  int rc = COLA_OK;
  COLA_DISPATCH_VEC(T, s.vec, (rc = launch_sweep<T, VEC>(s.n, s.k, s.colmask, st, [&](int64_t c) {
    int64_t cc = c & s.colmask;
    return MgsOp<T, VEC>{W + c, Qprev ? Qprev + c : nullptr, hprev ? hprev + cc : nullptr, Qcur ? Qcur + c : nullptr,
                         hcur ? hcur + cc : nullptr, wnorm2 ? wnorm2 + cc : nullptr, s.ld, s.colmask, gate, {}};
  }, "mgs_link")));
  return rc;
}

// ---- Arnoldi MGS chain: all links of one step in ONE cooperative launch (arnoldi.py:304-316) ----------------
// for j = 0 .. n_links-1:  h_j = <q_j, w>;  w -= h_j q_j     and finally  wnorm2 += <w, w>
// in exact modified-Gram-Schmidt order, like n_links + 1 mgs_link launches (pass j subtracts h_{j-1} q_{j-1} and
// accumulates <q_j, w> in the same sweep), with a grid-wide sync where a launch boundary was.  A thread owns the same
// elements of w in every pass, so w needs no sync of its own and stays in L2 (ld.global.cg / plain stores; the streaming
// hint is kept for the last use of q_{j-1}); q_j is read from DRAM once (its second use, one pass later, is an L2 hit
// when w and one basis block fit).  The link chain re-read q_{j-1} from DRAM and paid a launch + drain per link.
template <typename T, int VEC>
__global__ void __launch_bounds__(kSweepThreads)
    mgs_chain_kernel(T* W, const T* Q, int64_t q_stride, int n_links, double* H, int64_t ldh, double* wnorm2, int64_t n,
                     int64_t k, int64_t ld, int lanes, int rows_per_pass, int64_t colmask, const int32_t* gate, int cluster) {
  if (!gate_open(gate)) return;                     // uniform over the grid: nobody reaches a grid sync
  cg::grid_group grid = cg::this_grid();
  __shared__ double red[kSweepThreads * VEC];
  const int tid = threadIdx.x;
  const int r = tid / lanes, l = tid - r * lanes;
  const int64_t c0 = (int64_t)l * VEC;
  const bool active = (r < rows_per_pass) && (c0 < k);
  const int64_t stride = (int64_t)gridDim.x * rows_per_pass;
  constexpr bool HINT = sizeof(T) * VEC == 16;       // w is the operand every pass re-reads: evict_last in L2
  uint64_t pol = 0;
  if constexpr (HINT) pol = l2_policy_evict_last();
  auto ldw = [&](const T* p) {
    if constexpr (HINT) return ldg_hint<T, VEC>(p, pol);
    else return ldg_cg<T, VEC>(p);
  };
  for (int j = 0; j <= n_links; ++j) {
    const T* qp = j > 0 ? Q + (int64_t)(j - 1) * q_stride : nullptr;
    const T* qc = j < n_links ? Q + (int64_t)j * q_stride : nullptr;
    T hp[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) hp[v] = qp ? (T)__ldcg(H + (int64_t)(j - 1) * ldh + ((c0 + v) & colmask)) : (T)0;
    double acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.0;
    if (active) {
      auto one = [&](int64_t o, Vec<T, VEC> w, const Vec<T, VEC>& xp, const Vec<T, VEC>& xc) {
        if (qp) {
#pragma unroll
          for (int v = 0; v < VEC; ++v) w.v[v] = w.v[v] - hp[v] * xp.v[v];
          if constexpr (HINT) stg_hint<T, VEC>(W + o, w, pol);
          else stg<T, VEC>(W + o, w);
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] += (double)(qc ? xc.v[v] : w.v[v]) * (double)w.v[v];
      };
      int64_t row = (int64_t)blockIdx.x * rows_per_pass + r;
      for (; row + stride < n; row += 2 * stride) {          // two rows in flight per thread
        const int64_t oa = row * ld + c0, ob = (row + stride) * ld + c0;
        Vec<T, VEC> wa = ldw(W + oa), wb = ldw(W + ob), pa, pb, ca, cb;
        if (qp) { pa = ldg_stream<T, VEC>(qp + oa); pb = ldg_stream<T, VEC>(qp + ob); }
        if (qc) { ca = ldg_stream<T, VEC>(qc + oa); cb = ldg_stream<T, VEC>(qc + ob); }
        one(oa, wa, pa, ca);
        one(ob, wb, pb, cb);
      }
      if (row < n) {
        const int64_t oa = row * ld + c0;
        Vec<T, VEC> wa = ldw(W + oa), pa, ca;
        if (qp) pa = ldg_stream<T, VEC>(qp + oa);
        if (qc) ca = ldg_stream<T, VEC>(qc + oa);
        one(oa, wa, pa, ca);
      }
    }
    double* out = qc ? H + (int64_t)j * ldh : wnorm2;
    if (out != nullptr) block_col_reduce<VEC>(red, acc, active, tid, r, l, lanes, rows_per_pass, c0, k, colmask, out, cluster);
    if (j < n_links) grid.sync();                   // h_j complete (fp64 atomics at L2) before anyone subtracts it
  }
}

template <typename T>
int mgs_chain(T* W, const T* Q, int64_t q_stride, int64_t n_links, double* H, int64_t ldh, double* wnorm2, int64_t n,
              int64_t b, const int32_t* gate, cudaStream_t st) {
  COLA_REQUIRE(W && Q && H, "mgs_chain: null pointer");
  COLA_REQUIRE(n_links >= 1 && n_links < (1 << 20) && ldh >= b, "mgs_chain: bad link count / ldh");
  if (n <= 0 || b <= 0) return COLA_OK;
  Shape s = shape_of<T>(n, b, b, W, Q);
  if (s.vec > 1 && q_stride % s.vec != 0) s = Shape{n, b, b, 1, (int64_t)-1};
  if (s.k > (int64_t)kSweepThreads * s.vec) return fail(COLA_E_UNSUPPORTED, "mgs_chain: block wider than one sweep slab");
  int rc = COLA_OK;
  COLA_DISPATCH_VEC(T, s.vec, ({
    RowMap m = row_map(s.k, VEC, kSweepThreads);
    auto kern = mgs_chain_kernel<T, VEC>;
    static int per_sm = 0;
    if (per_sm == 0) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSweepThreads, 0);
    if (per_sm < 1) return fail(COLA_E_UNSUPPORTED, "mgs_chain: kernel does not fit an SM");
    int64_t tiles = (s.n + m.rows_per_pass - 1) / m.rows_per_pass;
    int64_t grid = (int64_t)sm_count() * (per_sm > 4 ? 4 : per_sm);   // co-resident by construction (grid-wide sync)
    if (grid > tiles) grid = tiles;
    int nl = (int)n_links;
    int64_t qs = q_stride, ldh_ = ldh, n_ = s.n, k_ = s.k, ld_ = s.ld, cm = s.colmask;
    // clusters of 8 CTAs fold their column sums before the atomics (see block_col_reduce): every pass ends in
    // grid x b same-address atomics, which is what a short pass waits for
    int cluster = (grid >= 16 && grid * s.k >= 48 * 1024) ? 8 : 1;   // same threshold as launch_sweep
    cudaError_t e = cudaErrorUnknown;
    if (cluster > 1) {
      const int64_t gridc = grid / cluster * cluster;           // (rounding down keeps the grid co-resident)
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)gridc);
      cfg.blockDim = dim3(kSweepThreads);
      cfg.stream = st;
      cudaLaunchAttribute attr[2];
      attr[0].id = cudaLaunchAttributeCooperative;
      attr[0].val.cooperative = 1;
      attr[1].id = cudaLaunchAttributeClusterDimension;
      attr[1].val.clusterDim.x = (unsigned)cluster;
      attr[1].val.clusterDim.y = 1;
      attr[1].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 2;
      e = cudaLaunchKernelEx(&cfg, kern, W, Q, qs, nl, H, ldh_, wnorm2, n_, k_, ld_, m.lanes, m.rows_per_pass, cm, gate, cluster);
      if (e != cudaSuccess) { cudaGetLastError(); cluster = 1; }
    }
    if (cluster == 1) {
      void* args[] = {&W, &Q, &qs, &nl, &H, &ldh_, &wnorm2, &n_, &k_, &ld_, &m.lanes, &m.rows_per_pass, &cm, &gate, &cluster};
      e = cudaLaunchCooperativeKernel((void*)kern, dim3((unsigned)grid), dim3(kSweepThreads), args, 0, st);
    }
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(COLA_E_UNSUPPORTED, "mgs_chain: cooperative launch refused");
    }
    rc = cuda_status("mgs_chain");
  }));
  return rc;
}

}  // namespace cola

// ======================================================================================================
// C ABI
// ======================================================================================================
using namespace cola;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int cola_version(void) { return 3; }
int cola_publish_bytes(const void* src, void* host_mapped, int64_t nbytes, void* stream) {
  if (!src || !host_mapped) return fail(COLA_E_BADARG, "publish: null pointer");
  if (nbytes <= 0) return COLA_OK;
  if (nbytes % 4 != 0) return fail(COLA_E_BADARG, "publish: nbytes must be a multiple of 4");
  publish_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(static_cast<const uint32_t*>(src),
                                                                          static_cast<volatile uint32_t*>(host_mapped), nbytes / 4);
  return cuda_status("publish");
}
int cola_store_i32x4(int32_t* dst, int32_t a, int32_t b, int32_t c, int32_t d, void* stream) {
  if (!dst) return fail(COLA_E_BADARG, "store_i32x4: null pointer");
  store_i32x4_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dst, a, b, c, d);
  return cuda_status("store_i32x4");
}
const char* cola_last_error(void) { return g_err; }
int64_t cola_launch_count(void) { return g_launches.load(); }

int cola_device_info(int* sms, int* major, int* minor) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return fail(COLA_E_NOGPU, "no CUDA device");
  }
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
    cudaGetLastError();
    return fail(COLA_E_NOGPU, "no CUDA device");
  }
  if (sms) *sms = p.multiProcessorCount;
  if (major) *major = p.major;
  if (minor) *minor = p.minor;
  return COLA_OK;
}

#define COLA_VEC_API(SFX, T)                                                                                          \
  int cola_col_dots_##SFX(const T* X, const T* Y, int64_t n, int64_t k, int64_t ld, double* dots,                    \
                          const int32_t* gate, void* s) {                                                             \
    return col_dots<T>(X, Y, n, k, ld, dots, gate, ST(s));                                                            \
  }                                                                                                                   \
  int cola_col_scale_##SFX(const T* X, T* Y, int64_t n, int64_t k, int64_t ld, const double* sq, int take_sqrt,      \
                           int mode, T a, const int32_t* gate, void* s) {                                             \
    return col_scale<T>(X, Y, n, k, ld, sq, take_sqrt, mode, a, gate, ST(s));                                         \
  }                                                                                                                   \
  int cola_axpby_##SFX(const T* X, T* Y, int64_t n, int64_t k, int64_t ld, T a, T b, const int32_t* gate, void* s) { \
    return axpby<T>(X, Y, n, k, ld, a, b, gate, ST(s));                                                               \
  }                                                                                                                   \
  int cola_diag_matmat_##SFX(const T* X, int64_t ldx, T* Y, int64_t ldy, int64_t n, int64_t k, T shift,              \
                             const T* diag, int accumulate, double* dots, const int32_t* dots_row,                    \
                             const int32_t* gate, void* s) {                                                          \
    return diag_matmat<T>(X, ldx, Y, ldy, n, k, shift, diag, accumulate, dots, dots_row, gate, ST(s));                \
  }                                                                                                                   \
  int cola_cg_update_r_##SFX(T* R, const T* AP, int64_t n, int64_t k, int64_t ld, const cola_cg_ctl_t* ctl,          \
                             const double* gamma, const double* pAp, double* gamma_w, void* s) {                      \
    return cg_update_r<T>(R, AP, n, k, ld, ctl, gamma, pAp, gamma_w, ST(s));                                          \
  }                                                                                                                   \
  int cola_cg_update_xp_##SFX(T* X, const T* R, T* P, int64_t n, int64_t k, int64_t ld, const cola_cg_ctl_t* ctl,    \
                              const double* gamma, const double* pAp, void* s) {                                      \
    return cg_update_xp<T>(X, R, P, n, k, ld, ctl, gamma, pAp, ST(s));                                                \
  }                                                                                                                   \
  int cola_cg_tol_##SFX(const double* gamma0, T tol, T* tol_eff, int64_t k, void* s) {                               \
    if (!gamma0 || !tol_eff) return fail(COLA_E_BADARG, "cg_tol: null pointer");                                      \
    if (k <= 0) return COLA_OK;                                                                                       \
    cg_tol_kernel<T><<<(unsigned)((k + 127) / 128), 128, 0, ST(s)>>>(gamma0, tol, tol_eff, k);                        \
    return cuda_status("cg_tol");                                                                                     \
  }                                                                                                                   \
  int cola_cg_advance_##SFX(cola_cg_ctl_t* ctl, const double* gamma, const T* tol_eff, int increment, void* s) {     \
    if (!ctl || !gamma || !tol_eff) return fail(COLA_E_BADARG, "cg_advance: null pointer");                           \
    cg_advance_kernel<T><<<1, 256, 0, ST(s)>>>(ctl, gamma, tol_eff, increment);                                       \
    return cuda_status("cg_advance");                                                                                 \
  }                                                                                                                   \
  int cola_lanczos_three_term_##SFX(T* W, const T* Vi, const T* Vim1, int64_t n, int64_t b, const double* alpha_acc, \
                                    const double* beta_prev_sq, const int32_t* gate, void* s) {                       \
    return lanczos_three_term<T>(W, Vi, Vim1, n, b, alpha_acc, beta_prev_sq, gate, ST(s));                            \
  }                                                                                                                   \
  int cola_mgs_link_##SFX(T* W, const T* Qprev, const double* hprev, const T* Qcur, double* hcur, double* wnorm2,    \
                          int64_t n, int64_t b, const int32_t* gate, void* s) {                                       \
    return mgs_link<T>(W, Qprev, hprev, Qcur, hcur, wnorm2, n, b, gate, ST(s));                                       \
  }                                                                                                        \
  int cola_mgs_chain_##SFX(T* W, const T* Q, int64_t q_stride, int64_t n_links, double* H, int64_t ldh, double* wnorm2, \
                           int64_t n, int64_t b, const int32_t* gate, void* s) {                                        \
    return mgs_chain<T>(W, Q, q_stride, n_links, H, ldh, wnorm2, n, b, gate, ST(s));                                    \
  }

COLA_VEC_API(f32, float)
COLA_VEC_API(f64, double)

}  // extern "C"
