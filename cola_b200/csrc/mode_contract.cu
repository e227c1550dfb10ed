// Batched mode contraction  out[p,a,q] = alpha * sum_j M[a,j] in[p,j,q]  (+ fused operator epilogue).
//
// One primitive serves three reference operators without any of their moveaxis/reshape/concat copies:
//   Dense._matmat      cola/ops/operators.py:26-28     pre = 1,            post = k
//   Kronecker._matmat  cola/ops/operators.py:216-223   one call per factor (pre = prod earlier dims,
//                                                      post = prod later dims * k)
//   BlockDiag._matmat  cola/ops/operators.py:299-310   pre = multiplicity, post = k, one call per block
//
// This file is the exact-arithmetic SIMT path (fp32 and fp64, any shape).  Two kernels:
//   mc_tile_kernel   64x64x16 shared-memory tiled GEMM per (p, a-tile, q-tile), 4x4 register micro-tiles;
//   mc_thin_kernel   post <= 8 (GEMV-like: single RHS solves, Lanczos with one start vector): one warp per
//                    output row, lanes stride over j so M is read coalesced exactly once.
// The fp32 tensor-core path for large Kronecker factors lives in kron_tc.cu.
#include <cstdlib>
#include "sweep.cuh"

namespace cola {

template <typename T>
struct McArgs {
  const T* M; int64_t ldm, d_out, d_in, pre, post;
  const T* in; T* out;
  T alpha, shift; const T* diag; const T* epi_x; int accumulate;
  double* dots; const int32_t* dots_row; const int32_t* gate;
};

constexpr int BM = 64, BN = 64, BK = 16, MC_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(MC_THREADS) mc_tile_kernel(McArgs<T> a) {
  if (a.gate != nullptr && *a.gate != 0) return;
  __shared__ T As[BK][BM + 4];
  __shared__ T Bs[BK][BN + 4];
  __shared__ double red[MC_THREADS * 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;  // thread owns rows ty*4..+3, cols tx*4..+3 of the tile
  const int64_t na = (a.d_out + BM - 1) / BM, nq = (a.post + BN - 1) / BN;
  double dacc[4] = {0.0, 0.0, 0.0, 0.0};
  const bool epi = (a.shift != (T)0) || a.diag || a.dots;

  for (int64_t qt = blockIdx.y; qt < nq; qt += gridDim.y) {
    const int64_t q0 = qt * BN;
    for (int64_t t = blockIdx.x; t < a.pre * na; t += gridDim.x) {
      const int64_t p = t / na, a0 = (t - p * na) * BM;
      const T* inp = a.in + p * a.d_in * a.post;
      T acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = (T)0;

      for (int64_t k0 = 0; k0 < a.d_in; k0 += BK) {
        // M tile (BM x BK), stored transposed; 1024 elements / 256 threads
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          int idx = tid + e * MC_THREADS;
          int kk = idx % BK, mm = idx / BK;
          int64_t ga = a0 + mm, gk = k0 + kk;
          As[kk][mm] = (ga < a.d_out && gk < a.d_in) ? a.M[ga * a.ldm + gk] : (T)0;
        }
        // in tile (BK x BN): q contiguous -> coalesced
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          int idx = tid + e * MC_THREADS;
          int nn = idx % BN, kk = idx / BN;
          int64_t gk = k0 + kk, gq = q0 + nn;
          Bs[kk][nn] = (gk < a.d_in && gq < a.post) ? inp[gk * a.post + gq] : (T)0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
          T av[4], bv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
          for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] += av[i] * bv[j];
        }
        __syncthreads();
      }
      // epilogue on the flattened (pre*d_out, post) matrix
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t ga = a0 + ty * 4 + i;
        if (ga >= a.d_out) continue;
        const int64_t row = p * a.d_out + ga;
        const T d = a.diag ? a.diag[row] : (T)0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int64_t gq = q0 + tx * 4 + j;
          if (gq >= a.post) continue;
          const int64_t o = row * a.post + gq;
          T tval = a.alpha * acc[i][j];
          T xo = (T)0;
          if (epi) {
            xo = a.epi_x[o];
            if (a.shift != (T)0) tval += a.shift * xo;
            if (a.diag) tval += d * xo;
          }
          if (a.accumulate) tval += a.out[o];
          a.out[o] = tval;
          if (a.dots) dacc[j] += (double)xo * (double)tval;
        }
      }
    }
  }
  if (a.dots) {  // host guarantees gridDim.y == nq here, so this thread's 4 columns are fixed
    const int64_t q0 = (int64_t)blockIdx.y * BN;
    double* out = a.dots + (a.dots_row ? (int64_t)(*a.dots_row) * a.post : 0) + q0;
    block_col_reduce<4>(red, dacc, true, tid, ty, tx, 16, 16, (int64_t)tx * 4, a.post - q0, -1, out);
  }
}

constexpr int THIN_MAX = 8;

template <typename T>
__global__ void __launch_bounds__(MC_THREADS) mc_thin_kernel(McArgs<T> a) {
  if (a.gate != nullptr && *a.gate != 0) return;
  __shared__ double red[(MC_THREADS / 32) * THIN_MAX];
  const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
  const int nwarps = MC_THREADS / 32;
  const int post = (int)a.post;
  const bool epi = (a.shift != (T)0) || a.diag || a.dots;
  double dacc[THIN_MAX];
#pragma unroll
  for (int q = 0; q < THIN_MAX; ++q) dacc[q] = 0.0;
  const int64_t total = a.pre * a.d_out;
  for (int64_t row = (int64_t)blockIdx.x * nwarps + warp; row < total; row += (int64_t)gridDim.x * nwarps) {
    const int64_t p = row / a.d_out, ga = row - p * a.d_out;
    const T* inp = a.in + p * a.d_in * a.post;
    const T* mrow = a.M + ga * a.ldm;
    T acc[THIN_MAX];
#pragma unroll
    for (int q = 0; q < THIN_MAX; ++q) acc[q] = (T)0;
    for (int64_t j = lane; j < a.d_in; j += 32) {
      const T m = mrow[j];
#pragma unroll
      for (int q = 0; q < THIN_MAX; ++q)
        if (q < post) acc[q] += m * inp[j * post + q];
    }
#pragma unroll
    for (int q = 0; q < THIN_MAX; ++q) {
      if (q < post) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
      }
    }
    if (lane == 0) {
      const T d = a.diag ? a.diag[row] : (T)0;
      for (int q = 0; q < post; ++q) {
        const int64_t o = row * post + q;
        T tval = a.alpha * acc[q];
        T xo = (T)0;
        if (epi) {
          xo = a.epi_x[o];
          if (a.shift != (T)0) tval += a.shift * xo;
          if (a.diag) tval += d * xo;
        }
        if (a.accumulate) tval += a.out[o];
        a.out[o] = tval;
        dacc[q] += (double)xo * (double)tval;
      }
    }
  }
  if (a.dots) {
    if (lane == 0)
      for (int q = 0; q < post; ++q) red[warp * THIN_MAX + q] = dacc[q];
    __syncthreads();
    if (threadIdx.x < post) {
      double s = 0.0;
      for (int w = 0; w < nwarps; ++w) s += red[w * THIN_MAX + threadIdx.x];
      double* out = a.dots + (a.dots_row ? (int64_t)(*a.dots_row) * a.post : 0);
      atomicAdd(out + threadIdx.x, s);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// fp32 factors with d_out % 128 == 0, d_in % 8 == 0, post % 4 == 0 (the 128-wide Kronecker factors of BASELINE
// config 4): 128x128x8 tiles, 8x8 register micro-tiles (two 4-wide halves per dimension so every shared-memory
// read is a conflict-free / broadcast LDS.128), global->register prefetch of the next k-slab while the current
// one is multiplied (double-buffered shared memory, one barrier per slab).  Per slab step a thread issues
// 4 LDS.128 for 64 FFMA, against 8 scalar LDS for 16 FFMA in mc_tile_kernel.
// ---------------------------------------------------------------------------------------------------
constexpr int GB_M = 128, GB_N = 128, GB_K = 8;

__global__ void __launch_bounds__(MC_THREADS, 2) mc_big_kernel(McArgs<float> a) {
  if (a.gate != nullptr && *a.gate != 0) return;
  __shared__ __align__(16) float As[2][GB_K][GB_M + 4];
  __shared__ __align__(16) float Bs[2][GB_K][GB_N];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int64_t na = a.d_out / GB_M, nq = (a.post + GB_N - 1) / GB_N;
  const int64_t n_tiles = a.pre * na * nq;
  const bool epi = (a.shift != 0.f) || a.diag;
  // loader roles
  const int a_row = tid / 2, a_kq = (tid % 2) * 4;            // M tile 128 x 8: one float4 along k per thread
  const int b_kk = tid / 32, b_nq = (tid % 32) * 4;           // in tile 8 x 128: one float4 along q per thread
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int64_t qt = t % nq, rest = t / nq;
    const int64_t p = rest / na, a0 = (rest - p * na) * GB_M;
    const int64_t q0 = qt * GB_N;
    const float* inp = a.in + p * a.d_in * a.post + q0;
    const float* mp = a.M + (a0 + a_row) * a.ldm + a_kq;
    const bool b_ok = q0 + b_nq < a.post;                     // post % 4 == 0: a float4 is all in or all out
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    float4 ra = *reinterpret_cast<const float4*>(mp);
    float4 rb = b_ok ? *reinterpret_cast<const float4*>(inp + (int64_t)b_kk * a.post + b_nq) : make_float4(0.f, 0.f, 0.f, 0.f);
    int cur = 0;
    As[0][a_kq + 0][a_row] = ra.x; As[0][a_kq + 1][a_row] = ra.y; As[0][a_kq + 2][a_row] = ra.z; As[0][a_kq + 3][a_row] = ra.w;
    *reinterpret_cast<float4*>(&Bs[0][b_kk][b_nq]) = rb;
    __syncthreads();
    for (int64_t k0 = 0; k0 < a.d_in; k0 += GB_K) {
      const bool more = k0 + GB_K < a.d_in;
      if (more) {
        ra = *reinterpret_cast<const float4*>(mp + k0 + GB_K);
        rb = b_ok ? *reinterpret_cast<const float4*>(inp + (k0 + GB_K + b_kk) * a.post + b_nq) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int kk = 0; kk < GB_K; ++kk) {
        const float4 a_lo = *reinterpret_cast<const float4*>(&As[cur][kk][ty * 4]);
        const float4 a_hi = *reinterpret_cast<const float4*>(&As[cur][kk][64 + ty * 4]);
        const float4 b_lo = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * 4]);
        const float4 b_hi = *reinterpret_cast<const float4*>(&Bs[cur][kk][64 + tx * 4]);
        const float av[8] = {a_lo.x, a_lo.y, a_lo.z, a_lo.w, a_hi.x, a_hi.y, a_hi.z, a_hi.w};
        const float bv[8] = {b_lo.x, b_lo.y, b_lo.z, b_lo.w, b_hi.x, b_hi.y, b_hi.z, b_hi.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] += av[i] * bv[j];
      }
      if (more) {
        const int nxt = cur ^ 1;
        As[nxt][a_kq + 0][a_row] = ra.x; As[nxt][a_kq + 1][a_row] = ra.y; As[nxt][a_kq + 2][a_row] = ra.z; As[nxt][a_kq + 3][a_row] = ra.w;
        *reinterpret_cast<float4*>(&Bs[nxt][b_kk][b_nq]) = rb;
        __syncthreads();
        cur = nxt;
      }
    }
    __syncthreads();   // the next tile's prologue overwrites buffer 0
    // epilogue: rows ty*4+i (+64), columns tx*4.. (+64), float4 stores
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t ga = a0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
      const int64_t row = p * a.d_out + ga;
      const float d = a.diag ? a.diag[row] : 0.f;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int64_t gq = q0 + h * 64 + tx * 4;
        if (gq >= a.post) continue;
        const int64_t o = row * a.post + gq;
        float4 v = make_float4(a.alpha * acc[i][h * 4 + 0], a.alpha * acc[i][h * 4 + 1], a.alpha * acc[i][h * 4 + 2],
                               a.alpha * acc[i][h * 4 + 3]);
        if (epi) {
          const float4 xo = *reinterpret_cast<const float4*>(a.epi_x + o);
          const float f = a.shift + d;
          v.x += f * xo.x; v.y += f * xo.y; v.z += f * xo.z; v.w += f * xo.w;
        }
        if (a.accumulate) {
          const float4 y = *reinterpret_cast<const float4*>(a.out + o);
          v.x += y.x; v.y += y.y; v.z += y.z; v.w += y.w;
        }
        *reinterpret_cast<float4*>(a.out + o) = v;
      }
    }
  }
}

static bool big_ok(const McArgs<float>& a) {
  return !a.dots && a.d_out % GB_M == 0 && a.d_in % GB_K == 0 && a.post % 4 == 0 && a.post >= 64 && a.ldm % 4 == 0 &&
         ((uintptr_t)a.M % 16 == 0) && ((uintptr_t)a.in % 16 == 0) && ((uintptr_t)a.out % 16 == 0) &&
         (!a.epi_x || (uintptr_t)a.epi_x % 16 == 0) && getenv("COLA_MC_NO_BIG") == nullptr;
}
static bool big_ok(const McArgs<double>&) { return false; }
static void launch_big(const McArgs<float>& a, cudaStream_t st) {
  const int64_t n_tiles = a.pre * (a.d_out / GB_M) * ((a.post + GB_N - 1) / GB_N);
  int64_t grid = (int64_t)sm_count() * 2 * 8;                 // a few waves of 2 CTAs/SM; tiles are uniform
  if (grid > n_tiles) grid = n_tiles;
  mc_big_kernel<<<(unsigned)grid, MC_THREADS, 0, st>>>(a);
}
static void launch_big(const McArgs<double>&, cudaStream_t) {}

// ---------------------------------------------------------------------------------------------------
// Short-and-wide factors (d_out <= 64, d_in huge, pre = 1): the U^T r product of a Nystrom preconditioner
// (preconditioners.py:128-130), r x n times n x k.  Tiling the output gives one tile; instead the K range is split
// over the whole grid, every CTA accumulates a full d_out x post partial in registers (4 x 16 per thread) and adds
// it to the zeroed output with one atomic per element.  Summation order across CTAs is not fixed, so results are
// reproducible to rounding, not bitwise.
// ---------------------------------------------------------------------------------------------------
constexpr int SK_M = 64, SK_N = 256;

template <typename T>
__global__ void __launch_bounds__(MC_THREADS) mc_zero_kernel(T* out, int64_t count, const int32_t* gate) {
  if (gate != nullptr && *gate != 0) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (T)0;
}

// Thread tiles: 4 output rows x VEC columns (one 16-byte vector), tile t = (row group t / QG, column group t % QG);
// a thread owns tiles t = slot, slot + MC_THREADS, ... and -- when there are fewer tiles than threads -- the thread
// groups split the k rows of a slab between them (partials meet in the atomics anyway).  Per k a thread issues one
// or two 16-byte shared loads per operand for 4 x VEC FMAs.
constexpr int SK_BK = 32;     // k rows staged per step
constexpr int SK_NT = 4;      // tiles per thread (64 x 256 fp32 outputs = 1024 tiles)

template <typename T, int VEC>
__global__ void __launch_bounds__(MC_THREADS, 2) mc_splitk_kernel(McArgs<T> a, int64_t k_per_cta, int AG, int QG) {
  if (a.gate != nullptr && *a.gate != 0) return;
  extern __shared__ __align__(16) unsigned char sk_smem[];
  const int AP = AG * 4, QP = QG * VEC;
  T* As = reinterpret_cast<T*>(sk_smem);                 // [SK_BK][AP]   As[kk][m] = M[m][k0 + kk]
  T* Bs = As + SK_BK * AP;                               // [SK_BK][QP]
  const int tid = threadIdx.x;
  const int d_out = (int)a.d_out, post = (int)a.post;
  const int tiles = AG * QG;
  const int groups = tiles >= MC_THREADS ? 1 : MC_THREADS / tiles;     // k-splitting thread groups
  const int grp = tiles >= MC_THREADS ? 0 : tid / tiles;
  const int slot = tiles >= MC_THREADS ? tid : tid - grp * tiles;
  const bool live = grp < groups;
  const int64_t k_begin = (int64_t)blockIdx.x * k_per_cta;
  const int64_t k_end = min(a.d_in, k_begin + k_per_cta);
  if (k_begin >= k_end) return;
  T acc[SK_NT][4][VEC];
#pragma unroll
  for (int t = 0; t < SK_NT; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[t][i][v] = (T)0;
  const bool vec_in = (post % VEC == 0) && ((reinterpret_cast<uintptr_t>(a.in) % 16) == 0);
  for (int64_t k0 = k_begin; k0 < k_end; k0 += SK_BK) {
    // factor slab: AP x SK_BK, read along k (contiguous in M), stored transposed
    for (int idx = tid; idx < AP * SK_BK; idx += MC_THREADS) {
      const int kk = idx % SK_BK, mm = idx / SK_BK;
      const int64_t gk = k0 + kk;
      As[kk * AP + mm] = (mm < d_out && gk < k_end) ? a.M[(int64_t)mm * a.ldm + gk] : (T)0;
    }
    if (vec_in) {
      for (int idx = tid; idx < SK_BK * QG; idx += MC_THREADS) {
        const int kk = idx / QG, qg = idx - kk * QG;
        const int64_t gk = k0 + kk;
        Vec<T, VEC> x;
        if (gk < k_end) {
          x = ldg_stream<T, VEC>(a.in + gk * a.post + (int64_t)qg * VEC);
        } else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) x.v[v] = (T)0;
        }
        *reinterpret_cast<Vec<T, VEC>*>(Bs + kk * QP + qg * VEC) = x;
      }
    } else {
      for (int idx = tid; idx < SK_BK * QP; idx += MC_THREADS) {
        const int kk = idx / QP, nn = idx - kk * QP;
        const int64_t gk = k0 + kk;
        Bs[kk * QP + nn] = (gk < k_end && nn < post) ? a.in[gk * a.post + nn] : (T)0;
      }
    }
    __syncthreads();
    if (live) {
#pragma unroll
      for (int t = 0; t < SK_NT; ++t) {
        const int tile = slot + t * MC_THREADS;
        if (tile < tiles) {
          const int ag = tile / QG, qg = tile - ag * QG;
          const T* ap = As + ag * 4;
          const T* bp = Bs + qg * VEC;
#pragma unroll 4
          for (int kk = grp; kk < SK_BK; kk += groups) {
            const Vec<T, VEC> bv = *reinterpret_cast<const Vec<T, VEC>*>(bp + kk * QP);
            T av[4];
            if constexpr (sizeof(T) == 4) {
              const Vec<T, 4> a4 = *reinterpret_cast<const Vec<T, 4>*>(ap + kk * AP);
#pragma unroll
              for (int i = 0; i < 4; ++i) av[i] = a4.v[i];
            } else {
              const Vec<T, 2> a0 = *reinterpret_cast<const Vec<T, 2>*>(ap + kk * AP);
              const Vec<T, 2> a1 = *reinterpret_cast<const Vec<T, 2>*>(ap + kk * AP + 2);
              av[0] = a0.v[0]; av[1] = a0.v[1]; av[2] = a1.v[0]; av[3] = a1.v[1];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int v = 0; v < VEC; ++v) acc[t][i][v] += av[i] * bv.v[v];
          }
        }
      }
    }
    __syncthreads();
  }
  if (!live) return;
#pragma unroll
  for (int t = 0; t < SK_NT; ++t) {
    const int tile = slot + t * MC_THREADS;
    if (tile >= tiles) continue;
    const int ag = tile / QG, qg = tile - ag * QG;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int ga = ag * 4 + i;
      if (ga >= d_out) continue;
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const int q = qg * VEC + v;
        if (q < post) atomicAdd(a.out + (int64_t)ga * a.post + q, a.alpha * acc[t][i][v]);
      }
    }
  }
}

template <typename T>
static bool splitk_ok(const McArgs<T>& a) {
  const bool epi = (a.shift != (T)0) || a.diag || a.dots || a.accumulate;
  const int64_t tiles = ((a.d_out + 3) / 4) * ((a.post + 16 / (int64_t)sizeof(T) - 1) / (16 / (int64_t)sizeof(T)));
  return !epi && a.pre == 1 && a.d_out <= SK_M && a.post <= SK_N && tiles <= (int64_t)SK_NT * MC_THREADS &&
         a.d_in >= 8192 && a.d_in >= 64 * a.d_out &&
         getenv("COLA_MC_NO_SPLITK") == nullptr;
}

template <typename T>
int mode_contract(const T* M, int64_t ldm, int64_t d_out, int64_t d_in, int64_t pre, int64_t post, const T* in,
                  T* out, T alpha, T shift, const T* diag, const T* epi_x, int accumulate, double* dots,
                  const int32_t* dots_row, const int32_t* gate, cudaStream_t st) {
  COLA_REQUIRE(M && in && out, "mode_contract: null pointer");
  COLA_REQUIRE(in != out, "mode_contract: in and out must not alias");
  COLA_REQUIRE(ldm >= d_in, "mode_contract: ldm < d_in");
  const bool epi = (shift != (T)0) || diag || dots;
  COLA_REQUIRE(!epi || (epi_x && d_out == d_in), "mode_contract: shift/diag/dots need epi_x and a square factor");
  if (pre <= 0 || post <= 0 || d_out <= 0) return COLA_OK;
  McArgs<T> a{M, ldm, d_out, d_in, pre, post, in, out, alpha, shift, diag, epi_x, accumulate, dots, dots_row, gate};
  if (post <= THIN_MAX) {
    int64_t rows = pre * d_out;
    int64_t grid = (rows + (MC_THREADS / 32) - 1) / (MC_THREADS / 32);
    int64_t cap = (int64_t)sm_count() * 8;
    if (grid > cap) grid = cap;
    mc_thin_kernel<T><<<(unsigned)grid, MC_THREADS, 0, st>>>(a);
    return cuda_status("mode_contract(thin)");
  }
  if (big_ok(a)) {
    launch_big(a, st);
    return cuda_status("mode_contract(big)");
  }
  if (splitk_ok(a)) {
    const int64_t count = d_out * post;
    mc_zero_kernel<T><<<(unsigned)((count + MC_THREADS - 1) / MC_THREADS), MC_THREADS, 0, st>>>(out, count, gate);
    int64_t ctas = (int64_t)sm_count() * 4;
    int64_t k_per_cta = (d_in + ctas - 1) / ctas;
    k_per_cta = (k_per_cta + SK_BK - 1) / SK_BK * SK_BK;
    ctas = (d_in + k_per_cta - 1) / k_per_cta;
    constexpr int VEC = 16 / (int)sizeof(T);
    const int AG = (int)((d_out + 3) / 4), QG = (int)((post + VEC - 1) / VEC);
    const size_t smem = (size_t)SK_BK * (AG * 4 + QG * VEC) * sizeof(T);
    auto kern = mc_splitk_kernel<T, VEC>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<(unsigned)ctas, MC_THREADS, smem, st>>>(a, k_per_cta, AG, QG);
    return cuda_status("mode_contract(split-k)");
  }
  const int64_t na = (d_out + BM - 1) / BM, nq = (post + BN - 1) / BN;
  COLA_REQUIRE(!dots || nq <= 65535, "mode_contract: dots with post > 4M columns unsupported");
  int64_t gy = nq < 65535 ? nq : 65535;
  int64_t gx = pre * na;
  int64_t want = ((int64_t)sm_count() * 4 + gy - 1) / gy;  // ~4 CTAs per SM in total
  if (dots && gx > want) gx = want;                        // persistent over (p, a-tile): fixed columns per thread
  if (gx > 2147483647LL) gx = 2147483647LL;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)gy);
  mc_tile_kernel<T><<<grid, MC_THREADS, 0, st>>>(a);
  return cuda_status("mode_contract(tile)");
}

}  // namespace cola

using namespace cola;
extern "C" {
int cola_mode_contract_f32(const float* M, int64_t ldm, int64_t d_out, int64_t d_in, int64_t pre, int64_t post,
                           const float* in, float* out, float alpha, float shift, const float* diag,
                           const float* epi_x, int accumulate, double* dots, const int32_t* dots_row,
                           const int32_t* gate, void* stream) {
  return mode_contract<float>(M, ldm, d_out, d_in, pre, post, in, out, alpha, shift, diag, epi_x, accumulate, dots,
                              dots_row, gate, reinterpret_cast<cudaStream_t>(stream));
}
int cola_mode_contract_f64(const double* M, int64_t ldm, int64_t d_out, int64_t d_in, int64_t pre, int64_t post,
                           const double* in, double* out, double alpha, double shift, const double* diag,
                           const double* epi_x, int accumulate, double* dots, const int32_t* dots_row,
                           const int32_t* gate, void* stream) {
  return mode_contract<double>(M, ldm, d_out, d_in, pre, post, in, out, alpha, shift, diag, epi_x, accumulate, dots,
                               dots_row, gate, reinterpret_cast<cudaStream_t>(stream));
}
}
