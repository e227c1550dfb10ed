// Full reorthogonalisation as tall-skinny kernels (north_star item 4): C = V^T W and W -= V C over the
// Krylov vectors stored so far.  Replaces do_gram (cola/linalg/decompositions/lanczos.py:293-296), which
// materialises two (b, n, m+2) temporaries per pass and sweeps all m+2 columns whether filled or not, and
// the Python MGS loop of Arnoldi when used as projection (arnoldi.py:304-311 uses mgs_link instead).
//
// Layout: V is (n_vec, n, b): vector j is an (n, b) row-major block at V + j*vstride (the matmat operand
// layout).  Both kernels are pure HBM streams over V[j0:j1]:
//   reorth_dots    reads W once per CTA row-chunk into shared memory, then each warp owns whole Krylov
//                  vectors j (no cross-warp reduction, no atomics in the inner loop) and streams V[j]'s chunk
//                  with 16-byte loads; partial C lives in shared memory as fp64 and is flushed once per CTA.
//   reorth_update  one thread owns (row, VEC columns), loops over j with coefficients in shared memory,
//                  fuses ||w||^2 (the Lanczos beta, lanczos.py:252) into the same pass.
// Algorithmic bytes per call: (j1-j0) * n*b*s  (+ n*b*s for W, twice for update).
#include "sweep.cuh"

namespace cola {

constexpr int kDotsThreads = 512;  // 16 warps: each warp owns Krylov vectors j = j0 + warp (mod 16)
constexpr int kDotsWarps = kDotsThreads / 32;

template <typename T>
struct RoArgs {
  const T* V; int64_t vstride, j0, j1; const T* Wc; T* W; int64_t n, b;
  double* C; const double* Cc; T sign; double* wnorm2; const int32_t* gate;
  int lanes_per_row, rows_per_chunk, vec_path;
};

// accumulate type: fp32 partials over one chunk (<= rows_per_chunk terms per lane) are promoted to fp64
// when they leave the registers; fp64 paths stay fp64 throughout.
template <typename T, int VEC, int NC>
__global__ void __launch_bounds__(kDotsThreads) reorth_dots_kernel(RoArgs<T> a) {
  if (a.gate != nullptr && *a.gate != 0) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int64_t nj = a.j1 - a.j0, b = a.b;
  double* csm = reinterpret_cast<double*>(smem_raw);            // nj * b doubles
  T* wsm = reinterpret_cast<T*>(csm + nj * b);                  // rows_per_chunk * b
  const int tid = threadIdx.x, lane = tid % 32, warp = tid / 32;
  for (int64_t i = tid; i < nj * b; i += kDotsThreads) csm[i] = 0.0;

  const int Lr = a.lanes_per_row;          // power of two <= 32
  const int rows_per_step = 32 / Lr;
  const int rsub = lane / Lr, cl = lane % Lr;
  const int64_t n_chunks = (a.n + a.rows_per_chunk - 1) / a.rows_per_chunk;

  for (int64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
    const int64_t row0 = ch * a.rows_per_chunk;
    const int rows = (int)min((int64_t)a.rows_per_chunk, a.n - row0);
    __syncthreads();
    {  // stage W chunk (contiguous rows*b elements)
      const T* src = a.Wc + row0 * b;
      const int64_t cnt = (int64_t)rows * b;
      if (VEC > 1) {
        for (int64_t i = (int64_t)tid * VEC; i < cnt; i += (int64_t)kDotsThreads * VEC)
          stg<T, VEC>(wsm + i, ldg_stream<T, VEC>(src + i));
      } else {
        for (int64_t i = tid; i < cnt; i += kDotsThreads) wsm[i] = src[i];
      }
    }
    __syncthreads();
    for (int64_t j = a.j0 + warp; j < a.j1; j += kDotsWarps) {
      const T* vj = a.V + j * a.vstride + row0 * b;
      T acc[NC][VEC];
#pragma unroll
      for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[c][v] = (T)0;
#pragma unroll 4
      for (int r = rsub; r < rows; r += rows_per_step) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const int64_t col = ((int64_t)c * Lr + cl) * VEC;
          if (col < b) {
            Vec<T, VEC> x = ldg_stream<T, VEC>(vj + (int64_t)r * b + col);
            Vec<T, VEC> w = *reinterpret_cast<const Vec<T, VEC>*>(wsm + (int64_t)r * b + col);
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[c][v] += x.v[v] * w.v[v];
          }
        }
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          double s = (double)acc[c][v];
          for (int o = Lr; o < 32; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          const int64_t col = ((int64_t)c * Lr + cl) * VEC + v;
          if (rsub == 0 && col < b) csm[(j - a.j0) * b + col] += s;  // this warp is the only writer of row j
        }
      }
    }
  }
  __syncthreads();
  for (int64_t i = tid; i < nj * b; i += kDotsThreads) {
    double s = csm[i];
    if (s != 0.0) atomicAdd(a.C + a.j0 * b + i, s);
  }
}

constexpr int kUpdThreads = 256;

template <typename T, int VEC>
__global__ void __launch_bounds__(kUpdThreads) reorth_update_kernel(RoArgs<T> a, int lanes, int rows_per_pass) {
  if (a.gate != nullptr && *a.gate != 0) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int64_t nj = a.j1 - a.j0, b = a.b;
  T* coef = reinterpret_cast<T*>(smem_raw);  // nj * b, already multiplied by sign
  double* red = reinterpret_cast<double*>(smem_raw + ((nj * b * sizeof(T) + 15) / 16) * 16);
  const int tid = threadIdx.x;
  for (int64_t i = tid; i < nj * b; i += kUpdThreads) coef[i] = a.sign * (T)a.Cc[a.j0 * b + i];
  __syncthreads();
  const int r = tid / lanes, l = tid - r * lanes;
  const int64_t c0 = (int64_t)l * VEC;
  const bool active = (r < rows_per_pass) && (c0 < b);
  double nacc[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) nacc[v] = 0.0;
  if (active) {
    for (int64_t row = (int64_t)blockIdx.x * rows_per_pass + r; row < a.n; row += (int64_t)gridDim.x * rows_per_pass) {
      const int64_t o = row * b + c0;
      Vec<T, VEC> w = ldg_stream<T, VEC>(a.W + o);
      const T* vp = a.V + a.j0 * a.vstride + o;
      // sequential over j like `sum(V * c, axis=-1)`; 8 independent loads in flight per thread
      int64_t j = 0;
      for (; j + 8 <= nj; j += 8) {
        Vec<T, VEC> x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) x[u] = ldg_stream<T, VEC>(vp + (j + u) * a.vstride);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          Vec<T, VEC> cf = *reinterpret_cast<const Vec<T, VEC>*>(coef + (j + u) * b + c0);
#pragma unroll
          for (int v = 0; v < VEC; ++v) w.v[v] += cf.v[v] * x[u].v[v];
        }
      }
      for (; j < nj; ++j) {
        Vec<T, VEC> x = ldg_stream<T, VEC>(vp + j * a.vstride);
        Vec<T, VEC> cf = *reinterpret_cast<const Vec<T, VEC>*>(coef + j * b + c0);
#pragma unroll
        for (int v = 0; v < VEC; ++v) w.v[v] += cf.v[v] * x.v[v];
      }
      stg<T, VEC>(a.W + o, w);
#pragma unroll
      for (int v = 0; v < VEC; ++v) nacc[v] += (double)w.v[v] * (double)w.v[v];
    }
  }
  if (a.wnorm2) block_col_reduce<VEC>(red, nacc, active, tid, r, l, lanes, rows_per_pass, c0, b, -1, a.wnorm2);
}

static inline int next_pow2(int64_t x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

template <typename T>
int reorth_dots(const T* V, int64_t vstride, int64_t j0, int64_t j1, const T* W, int64_t n, int64_t b, double* C,
                const int32_t* gate, cudaStream_t st) {
  COLA_REQUIRE(V && W && C, "reorth_dots: null pointer");
  COLA_REQUIRE(j1 >= j0 && j0 >= 0, "reorth_dots: bad vector range");
  if (j1 == j0 || n <= 0 || b <= 0) return COLA_OK;
  const int maxv = 16 / (int)sizeof(T);
  // vector path: b*sizeof(T) multiple of 16 B, rows 16 B aligned, and b/VEC a power of two or a multiple of 32
  int vec = 1;
  {
    int v = pick_vec<T>(b, b, V, W);
    if (vstride % v) v = 1;
    int64_t per_row = b / v;
    if (v == maxv && ((per_row <= 32 && (per_row & (per_row - 1)) == 0) || per_row % 32 == 0)) vec = v;
  }
  int64_t per_row = (b + vec - 1) / vec;
  int Lr = per_row >= 32 ? 32 : next_pow2(per_row);
  int nc = (int)((per_row + Lr - 1) / Lr);
  COLA_REQUIRE(nc <= 8, "reorth_dots: probe block too wide (b > 256*VEC); split the block");
  // shared-memory budget: nj*b doubles for C + a W chunk; split the j range if C does not fit
  int dev = 0, smem_max = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (smem_max <= 0) smem_max = 48 * 1024;
  const int64_t budget = (int64_t)smem_max - 2048;
  int64_t w_rows = (32 * 1024) / (b * (int64_t)sizeof(T));
  if (w_rows < 32 / Lr) w_rows = 32 / Lr;
  if (w_rows > 1024) w_rows = 1024;
  const int64_t w_bytes = w_rows * b * (int64_t)sizeof(T);
  int64_t max_nj = (budget - w_bytes) / (b * 8);
  COLA_REQUIRE(max_nj >= 1, "reorth_dots: probe block too wide for shared memory");
  int rc = COLA_OK;
  for (int64_t ja = j0; ja < j1 && rc == COLA_OK; ja += max_nj) {
    int64_t jb = ja + max_nj < j1 ? ja + max_nj : j1;
    RoArgs<T> a{};
    a.V = V; a.vstride = vstride; a.j0 = ja; a.j1 = jb; a.Wc = W; a.n = n; a.b = b; a.C = C; a.gate = gate;
    a.lanes_per_row = Lr; a.rows_per_chunk = (int)w_rows;
    size_t smem = (size_t)((jb - ja) * b * 8 + w_bytes);
    int64_t n_chunks = (n + w_rows - 1) / w_rows;
    int per_sm = (int)(budget / (int64_t)smem);
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    int64_t grid = (int64_t)sm_count() * per_sm;
    if (grid > n_chunks) grid = n_chunks;
#define COLA_LAUNCH_DOTS(VECV, NCV)                                                                              \
  do {                                                                                                           \
    auto kern = reorth_dots_kernel<T, VECV, NCV>;                                                                \
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
    kern<<<(unsigned)grid, kDotsThreads, smem, st>>>(a);                                                         \
  } while (0)
#define COLA_DOTS_NC(VECV)                                   \
  do {                                                       \
    if (nc == 1) COLA_LAUNCH_DOTS(VECV, 1);                  \
    else if (nc == 2) COLA_LAUNCH_DOTS(VECV, 2);             \
    else if (nc <= 4) COLA_LAUNCH_DOTS(VECV, 4);             \
    else COLA_LAUNCH_DOTS(VECV, 8);                          \
  } while (0)
    if (vec == maxv && vec > 1) {
      if constexpr (sizeof(T) == 4) COLA_DOTS_NC(4); else COLA_DOTS_NC(2);
    } else {
      COLA_DOTS_NC(1);
    }
    rc = cuda_status("reorth_dots");
  }
  return rc;
}

template <typename T>
int reorth_update(const T* V, int64_t vstride, int64_t j0, int64_t j1, T* W, int64_t n, int64_t b, const double* C,
                  T sign, double* wnorm2, const int32_t* gate, cudaStream_t st) {
  COLA_REQUIRE(V && W && C, "reorth_update: null pointer");
  COLA_REQUIRE(j1 >= j0 && j0 >= 0, "reorth_update: bad vector range");
  if (n <= 0 || b <= 0) return COLA_OK;
  int vec = pick_vec<T>(b, b, V, W);
  if (vstride % vec) vec = 1;
  COLA_REQUIRE(b <= (int64_t)kUpdThreads * vec, "reorth_update: probe block too wide (b > 256*VEC); split it");
  int dev = 0, smem_max = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (smem_max <= 0) smem_max = 48 * 1024;
  const int64_t red_bytes = (int64_t)kUpdThreads * 4 * 8;
  int64_t max_nj = ((int64_t)smem_max - 2048 - red_bytes) / (b * (int64_t)sizeof(T));
  if (max_nj > 256) max_nj = 256;  // keeps >= 2 CTAs per SM for typical b
  COLA_REQUIRE(max_nj >= 1, "reorth_update: probe block too wide for shared memory");
  int rc = COLA_OK;
  int64_t ja = j0;
  do {
    int64_t jb = ja + max_nj < j1 ? ja + max_nj : j1;
    RoArgs<T> a{};
    a.V = V; a.vstride = vstride; a.j0 = ja; a.j1 = jb; a.W = W; a.n = n; a.b = b; a.Cc = C; a.sign = sign;
    a.wnorm2 = (jb == j1) ? wnorm2 : nullptr; a.gate = gate;
    RowMap m = row_map(b, vec, kUpdThreads);
    size_t smem = (size_t)(((jb - ja) * b * sizeof(T) + 15) / 16 * 16 + red_bytes);
    int64_t tiles = (n + m.rows_per_pass - 1) / m.rows_per_pass;
    int64_t grid = (int64_t)sm_count() * 4;
    if (grid > tiles) grid = tiles;
#define COLA_LAUNCH_UPD(VECV)                                                                                    \
  do {                                                                                                           \
    auto kern = reorth_update_kernel<T, VECV>;                                                                   \
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
    kern<<<(unsigned)grid, kUpdThreads, smem, st>>>(a, m.lanes, m.rows_per_pass);                                \
  } while (0)
    if constexpr (sizeof(T) == 4) {
      if (vec == 4) COLA_LAUNCH_UPD(4); else if (vec == 2) COLA_LAUNCH_UPD(2); else COLA_LAUNCH_UPD(1);
    } else {
      if (vec == 2) COLA_LAUNCH_UPD(2); else COLA_LAUNCH_UPD(1);
    }
    rc = cuda_status("reorth_update");
    ja = jb;
  } while (ja < j1 && rc == COLA_OK);
  return rc;
}

}  // namespace cola

using namespace cola;
extern "C" {
#define COLA_RO_API(SFX, T)                                                                                          \
  int cola_reorth_dots_##SFX(const T* V, int64_t vstride, int64_t j0, int64_t j1, const T* W, int64_t n, int64_t b, \
                             double* C, const int32_t* gate, void* stream) {                                         \
    return reorth_dots<T>(V, vstride, j0, j1, W, n, b, C, gate, reinterpret_cast<cudaStream_t>(stream));             \
  }                                                                                                                  \
  int cola_reorth_update_##SFX(const T* V, int64_t vstride, int64_t j0, int64_t j1, T* W, int64_t n, int64_t b,     \
                               const double* C, T sign, double* wnorm2, const int32_t* gate, void* stream) {         \
    return reorth_update<T>(V, vstride, j0, j1, W, n, b, C, sign, wnorm2, gate,                                      \
                            reinterpret_cast<cudaStream_t>(stream));                                                 \
  }
COLA_RO_API(f32, float)
COLA_RO_API(f64, double)
}
