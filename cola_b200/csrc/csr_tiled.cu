// Staged CSR SpMM for wide right-hand-side blocks on patterns with long column runs (stencil / banded matrices:
// BASELINE config 2).  Replaces Sparse._matmat (cola/ops/operators.py:77-78) + the p^T A p reduction (cg.py:157-158) like
// csr_spmm.cu, from the tile-local form of the pattern built by cola_b200/csr_tiles.py:
//   * a tile is S strips of R consecutive rows, `stride` rows apart (the pattern's dominant far diagonal), so the rows
//     of X gathered across that diagonal belong to the neighbouring strips of the SAME tile;
//   * the distinct rows of X a tile touches arrive in shared memory ONCE, as a few contiguous runs, by bulk copies
//     (cp.async.bulk, one per run, issued by a producer warp two tiles ahead into a ring); the non-zeros address them by
//     slot, so the row loop reads shared memory only: no register gather, no L1 lottery;
//   * L2 -> SM traffic for a 5-point stencil drops from 3 rows of X per output row (register-gather kernel: 3.2 GB per
//     cfg2 SpMM at the ~8 TB/s the L2 fabric delivers) to 1.3.
// Tiles the record marks irregular (too many runs / too many distinct rows) gather from global memory in the same loop.
// X must be contiguous (ldx == k) and k * sizeof(T) a multiple of 16.
#include <cstdlib>

#include "sweep.cuh"

namespace cola {

constexpr int kTlMaxConsumers = 768;             // consumer threads (a launch parameter) + one producer warp <= 800
constexpr int kTlRec = 32;                       // record words per tile (cola_b200/csr_tiles.py)

template <typename T>
struct TiledArgs {
  const int32_t* rec; const int32_t* rp; const int32_t* idx; const T* vals;
  int64_t n_rows, n_tiles, n_tiles2d, rows2d, stride, tiles_per_blk;
  int strip_rows, strips, rp_stride, cap_rows, cap_nz, n_stages;
  const T* X; int64_t k; T* Y; int64_t ldy;
  T alpha, shift; const T* diag; int accumulate;
  double* dots; const int32_t* dots_row; int64_t k_full; const int32_t* gate;
  int lanes;                                     // threads per row: k * sizeof(T) / 16
  int consumers;                                 // consumer threads (whole warps); the producer is the warp after them
  int stage_bytes, off_idx, off_val, off_rp;     // ring stage layout (bytes)
};

__device__ __forceinline__ uint32_t tl_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tl_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void tl_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tl_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tl_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ bool tl_mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tl_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// per-thread view of the tile being consumed
template <typename T>
struct TlRows {
  const unsigned char* sx;       // staged X rows, already offset to this thread's 16-byte chunk
  const int32_t* rp; const int32_t* sidx; const T* sval;
  const T* Xc; T* Yc;
  int64_t row0, stride;
  int RT, R, g, groups, s0, r0, ds, dr;
};

// Rows g, g + groups, ... of one tile.  REGULAR: sidx / the own-row table hold BYTE offsets of staged rows (slot * row
// bytes, from csr_tiles.py); else sidx holds columns and the rows are gathered from global memory.
template <typename T, bool EPI, bool DOTS, bool REGULAR>
__device__ __forceinline__ void tl_tile_rows(const TiledArgs<T>& a, const TlRows<T>& w, T* facc) {
  constexpr int VEC = 16 / (int)sizeof(T);
  const int32_t* self_off = w.rp + w.RT + 4;
  int s = w.s0, r = w.r0;
  for (int lr = w.g; lr < w.RT; lr += w.groups) {
    const int64_t row = w.row0 + s * w.stride + r;
    r += w.dr; s += w.ds;
    if (r >= w.R) { r -= w.R; ++s; }
    if (row >= a.n_rows) continue;
    const int32_t e0 = w.rp[lr], cnt = w.rp[lr + 1] - e0;
    Vec<T, VEC> xo, yo;
    T sd = a.shift;
    if constexpr (EPI) {                                                  // own row: the epilogue's operand
      if constexpr (REGULAR) xo = *reinterpret_cast<const Vec<T, VEC>*>(w.sx + self_off[lr]);
      else xo = ldg<T, VEC>(w.Xc + row * a.k);
      if (a.diag) sd += a.diag[row];
    }
    T* yp = w.Yc + row * a.ldy;
    if (a.accumulate) yo = ldg<T, VEC>(yp);
    const int32_t* pi = w.sidx + e0;
    const T* pv = w.sval + e0;
    T acc[VEC];
#pragma unroll
    for (int q = 0; q < VEC; ++q) acc[q] = (T)0;
#pragma unroll 4
    for (int32_t j = 0; j < cnt; ++j) {
      const T wv = pv[j];
      Vec<T, VEC> x;
      if constexpr (REGULAR) x = *reinterpret_cast<const Vec<T, VEC>*>(w.sx + pi[j]);
      else x = ldg<T, VEC>(w.Xc + (int64_t)pi[j] * a.k);
#pragma unroll
      for (int q = 0; q < VEC; ++q) acc[q] += wv * x.v[q];
    }
    Vec<T, VEC> y;
    if constexpr (EPI) {
#pragma unroll
      for (int q = 0; q < VEC; ++q) y.v[q] = a.alpha * acc[q] + sd * xo.v[q];
    } else {
#pragma unroll
      for (int q = 0; q < VEC; ++q) y.v[q] = a.alpha * acc[q];
    }
    if (a.accumulate) {
#pragma unroll
      for (int q = 0; q < VEC; ++q) y.v[q] += yo.v[q];
    }
    if constexpr (EPI && DOTS) {
#pragma unroll
      for (int q = 0; q < VEC; ++q) facc[q] += xo.v[q] * y.v[q];
    }
    stg_stream<T, VEC>(yp, y);
  }
}

template <typename T, bool EPI, bool DOTS>
__global__ void __launch_bounds__(800, 1) csr_spmm_tiled_kernel(TiledArgs<T> a) {
  if (a.gate != nullptr && *a.gate != 0) return;
  constexpr int VEC = 16 / (int)sizeof(T);
  extern __shared__ __align__(128) unsigned char tl_smem[];
  const uint32_t sbase = tl_smem_u32(tl_smem);
  // layout: ring stages | barriers (full[n_stages], empty[n_stages]) | dot scratch
  const uint32_t bar0 = sbase + (uint32_t)a.n_stages * (uint32_t)a.stage_bytes;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.n_stages; ++s) {
      tl_mbar_init(bar0 + 8 * s, 1);
      tl_mbar_init(bar0 + 8 * (a.n_stages + s), a.consumers / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  double* s_dots = reinterpret_cast<double*>(tl_smem + (size_t)a.n_stages * a.stage_bytes + 16 * a.n_stages + 64);
  if (DOTS) {
    for (int i = threadIdx.x; i < a.lanes * VEC; i += blockDim.x) s_dots[i] = 0.0;
  }
  __syncthreads();

  // this CTA's tiles: one contiguous chunk (neighbouring tiles share halo rows: L2 locality, contiguous records)
  const int64_t per = (a.n_tiles + gridDim.x - 1) / gridDim.x;
  const int64_t t_begin = (int64_t)blockIdx.x * per;
  const int64_t t_end = t_begin + per < a.n_tiles ? t_begin + per : a.n_tiles;
  const uint32_t row_bytes = (uint32_t)(a.k * (int64_t)sizeof(T));

  if (warp == a.consumers / 32) {
    // ===== producer warp: lane j copies run j; lanes 12 / 13 / 14 the slots, values and row pointers =====
    int32_t v = (t_begin < t_end) ? a.rec[t_begin * kTlRec + lane] : 0;
    int it = 0;
    for (int64_t t = t_begin; t < t_end; ++t, ++it) {
      const int32_t cur = v;
      if (t + 1 < t_end) v = a.rec[(t + 1) * kTlRec + lane];        // next record while this tile's copies are issued
      const int s = it % a.n_stages;
      const uint32_t full = bar0 + 8 * s, empty = bar0 + 8 * (a.n_stages + s);
      const int32_t nz_begin = __shfl_sync(0xffffffffu, cur, 0), nz_pad = __shfl_sync(0xffffffffu, cur, 1);
      const int32_t n_runs = __shfl_sync(0xffffffffu, cur, 2), n_dist = __shfl_sync(0xffffffffu, cur, 3);
      const int32_t col0 = __shfl_sync(0xffffffffu, cur, (8 + 2 * lane) & 31);
      const int32_t sl = __shfl_sync(0xffffffffu, cur, (9 + 2 * lane) & 31);
      if (lane == 0) {
        while (!tl_mbar_try(empty, ((it / a.n_stages) & 1) ^ 1)) __nanosleep(100);   // (a spinning warp costs its scheduler issue slots)
      }
      __syncwarp();
      const uint32_t stage = sbase + (uint32_t)s * (uint32_t)a.stage_bytes;
      if (lane == 0) {
        const uint32_t x_bytes = n_runs > 0 ? (uint32_t)n_dist * row_bytes : 0u;
        tl_mbar_expect_tx(full, x_bytes + (uint32_t)nz_pad * (4u + (uint32_t)sizeof(T)) + (uint32_t)a.rp_stride * 4u);
      }
      __syncwarp();
      if (lane < n_runs) {
        const uint32_t slot0 = (uint32_t)sl >> 16, len = (uint32_t)sl & 0xFFFFu;
        tl_bulk_load(stage + slot0 * row_bytes, reinterpret_cast<const char*>(a.X) + (size_t)col0 * row_bytes, len * row_bytes, full);
      } else if (lane == 12) {
        if (nz_pad > 0) tl_bulk_load(stage + a.off_idx, a.idx + nz_begin, (uint32_t)nz_pad * 4u, full);
      } else if (lane == 13) {
        if (nz_pad > 0) tl_bulk_load(stage + a.off_val, a.vals + nz_begin, (uint32_t)nz_pad * (uint32_t)sizeof(T), full);
      } else if (lane == 14) {
        tl_bulk_load(stage + a.off_rp, a.rp + t * a.rp_stride, (uint32_t)a.rp_stride * 4u, full);
      }
    }
  } else {
    // ===== consumers: `lanes` threads per row, one 16-byte column chunk each =====
    const int tid = threadIdx.x;
    const int g = tid / a.lanes, l = tid - g * a.lanes;
    const int groups = a.consumers / a.lanes;
    const bool col_ok = g < groups;
    TlRows<T> w;
    w.RT = a.strip_rows * a.strips; w.R = a.strip_rows; w.g = g; w.groups = groups;
    w.s0 = g / a.strip_rows; w.r0 = g - w.s0 * a.strip_rows;          // local row g + i * groups as (strip, row in strip),
    w.ds = groups / a.strip_rows; w.dr = groups - w.ds * a.strip_rows;  // advanced without a division per row
    w.Xc = a.X + (size_t)l * VEC; w.Yc = a.Y + (size_t)l * VEC;
    double dacc[VEC];
#pragma unroll
    for (int q = 0; q < VEC; ++q) dacc[q] = 0.0;
    int it = 0, s = 0;
    uint32_t parity = 0;
    for (int64_t t = t_begin; t < t_end; ++t, ++it) {
      tl_mbar_wait(bar0 + 8 * s, parity);
      const unsigned char* stage = tl_smem + (size_t)s * a.stage_bytes;
      w.sx = stage + l * 16;
      w.rp = reinterpret_cast<const int32_t*>(stage + a.off_rp);
      w.sidx = reinterpret_cast<const int32_t*>(stage + a.off_idx);
      w.sval = reinterpret_cast<const T*>(stage + a.off_val);
      if (t < a.n_tiles2d) {
        const int64_t blk = t / a.tiles_per_blk, c = t - blk * a.tiles_per_blk;
        w.row0 = blk * a.strips * a.stride + c * a.strip_rows;
        w.stride = a.stride;
      } else {
        w.row0 = a.rows2d + (t - a.n_tiles2d) * w.RT;
        w.stride = a.strip_rows;
      }
      T facc[VEC];
#pragma unroll
      for (int q = 0; q < VEC; ++q) facc[q] = (T)0;
      if (col_ok) {
        if (w.rp[w.RT + 3] >= 0) tl_tile_rows<T, EPI, DOTS, true>(a, w, facc);
        else tl_tile_rows<T, EPI, DOTS, false>(a, w, facc);
      }
      if constexpr (DOTS) {
#pragma unroll
        for (int q = 0; q < VEC; ++q) dacc[q] += (double)facc[q];
      }
      __syncwarp();
      if (lane == 0) tl_mbar_arrive(bar0 + 8 * (a.n_stages + s));   // this warp no longer reads the stage
      if (++s == a.n_stages) { s = 0; parity ^= 1; }
    }
    if constexpr (DOTS) {
      if (col_ok) {
#pragma unroll
        for (int q = 0; q < VEC; ++q) atomicAdd(s_dots + l * VEC + q, dacc[q]);
      }
    }
  }
  if constexpr (DOTS) {
    __syncthreads();
    double* out = a.dots + (a.dots_row ? (int64_t)(*a.dots_row) * a.k_full : 0);
    for (int i = threadIdx.x; i < a.lanes * VEC; i += blockDim.x)
      if (i < a.k) atomicAdd(out + i, s_dots[i]);
  }
}

template <typename T>
static int csr_spmm_tiled(const int32_t* rec, const int32_t* rp, const int32_t* idx, const T* vals, int64_t n_rows,
                          int64_t n_tiles, int64_t n_tiles2d, int64_t rows2d, int64_t stride, int64_t strip_rows,
                          int64_t strips, int64_t cap_rows, int64_t cap_nz, const T* X, int64_t k, T* Y, int64_t ldy, T alpha,
                          T shift, const T* diag, int accumulate, double* dots, const int32_t* dots_row, const int32_t* gate,
                          cudaStream_t st) {
  COLA_REQUIRE(rec && rp && idx && vals && X && Y, "csr_spmm_tiled: null pointer");
  COLA_REQUIRE(X != Y, "csr_spmm_tiled: X and Y must not alias");
  constexpr int VEC = 16 / (int)sizeof(T);
  COLA_REQUIRE(k >= VEC && k % VEC == 0 && k / VEC <= 256 && ldy % VEC == 0, "csr_spmm_tiled: k must be whole 16-byte chunks");
  COLA_REQUIRE(((uintptr_t)X % 16 == 0) && ((uintptr_t)Y % 16 == 0) && ((uintptr_t)idx % 16 == 0) &&
               ((uintptr_t)vals % 16 == 0) && ((uintptr_t)rp % 16 == 0), "csr_spmm_tiled: 16-byte alignment");
  COLA_REQUIRE(strip_rows >= 1 && strips >= 1 && n_tiles >= 0, "csr_spmm_tiled: bad tile geometry");
  if (n_rows <= 0 || n_tiles == 0) return COLA_OK;
  TiledArgs<T> a;
  a.rec = rec; a.rp = rp; a.idx = idx; a.vals = vals; a.n_rows = n_rows; a.n_tiles = n_tiles; a.n_tiles2d = n_tiles2d;
  a.rows2d = rows2d; a.stride = stride; a.tiles_per_blk = n_tiles2d > 0 ? stride / strip_rows : 1;
  a.strip_rows = (int)strip_rows; a.strips = (int)strips; a.rp_stride = (int)(2 * strip_rows * strips + 4);
  a.cap_rows = (int)cap_rows; a.cap_nz = (int)cap_nz;
  a.X = X; a.k = k; a.Y = Y; a.ldy = ldy; a.alpha = alpha; a.shift = shift; a.diag = diag; a.accumulate = accumulate;
  a.dots = dots; a.dots_row = dots_row; a.k_full = k; a.gate = gate;
  a.lanes = (int)(k / VEC);
  a.consumers = 768;
  if (const char* e = getenv("COLA_SPMM_TILE_WARPS")) { const int v = atoi(e); if (v >= 1 && v <= kTlMaxConsumers / 32) a.consumers = 32 * v; }
  while (a.consumers < a.lanes) a.consumers += 32;
  const int64_t row_bytes = k * (int64_t)sizeof(T);
  auto up = [](int64_t x, int64_t m) { return (x + m - 1) / m * m; };
  a.off_idx = (int)up(cap_rows * row_bytes, 128);
  a.off_val = a.off_idx + (int)up(cap_nz * 4, 128);
  a.off_rp = a.off_val + (int)up(cap_nz * (int64_t)sizeof(T), 128);
  a.stage_bytes = a.off_rp + (int)up((int64_t)a.rp_stride * 4, 128);
  int dev = 0, smem_max = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const int64_t extra = 16 * 8 + 64 + (int64_t)a.lanes * VEC * 8 + 128;
  int n_stages = (int)(((int64_t)smem_max - extra) / a.stage_bytes);
  if (n_stages > 6) n_stages = 6;
  if (n_stages < 2) return fail(COLA_E_UNSUPPORTED, "csr_spmm_tiled: a ring of two stages does not fit shared memory");
  a.n_stages = n_stages;
  const size_t smem = (size_t)n_stages * a.stage_bytes + extra;
  const bool epi = (shift != (T)0) || diag || dots;
  int64_t grid = sm_count();
  if (grid > n_tiles) grid = n_tiles;
#define COLA_TILED_LAUNCH(EPIV, DOTSV)                                                              \
  do {                                                                                              \
    auto kern = csr_spmm_tiled_kernel<T, EPIV, DOTSV>;                                              \
    static int attr_smem = 0;                                                                       \
    if ((int)smem > attr_smem) {                                                                    \
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);           \
      attr_smem = (int)smem;                                                                        \
    }                                                                                               \
    kern<<<(unsigned)grid, a.consumers + 32, smem, st>>>(a);                                              \
  } while (0)
  if (dots) COLA_TILED_LAUNCH(true, true);
  else if (epi) COLA_TILED_LAUNCH(true, false);
  else COLA_TILED_LAUNCH(false, false);
  return cuda_status("csr_spmm_tiled");
}

}  // namespace cola

using namespace cola;
extern "C" {
int cola_csr_spmm_tiled_f32(const int32_t* rec, const int32_t* rp, const int32_t* idx, const float* vals, int64_t n_rows,
                            int64_t n_tiles, int64_t n_tiles2d, int64_t rows2d, int64_t stride, int64_t strip_rows,
                            int64_t strips, int64_t cap_rows, int64_t cap_nz, const float* X, int64_t k, float* Y, int64_t ldy,
                            float alpha, float shift, const float* diag, int accumulate, double* dots,
                            const int32_t* dots_row, const int32_t* gate, void* stream) {
  return csr_spmm_tiled<float>(rec, rp, idx, vals, n_rows, n_tiles, n_tiles2d, rows2d, stride, strip_rows, strips, cap_rows,
                               cap_nz, X, k, Y, ldy, alpha, shift, diag, accumulate, dots, dots_row, gate,
                               reinterpret_cast<cudaStream_t>(stream));
}
int cola_csr_spmm_tiled_f64(const int32_t* rec, const int32_t* rp, const int32_t* idx, const double* vals, int64_t n_rows,
                            int64_t n_tiles, int64_t n_tiles2d, int64_t rows2d, int64_t stride, int64_t strip_rows,
                            int64_t strips, int64_t cap_rows, int64_t cap_nz, const double* X, int64_t k, double* Y,
                            int64_t ldy, double alpha, double shift, const double* diag, int accumulate, double* dots,
                            const int32_t* dots_row, const int32_t* gate, void* stream) {
  return csr_spmm_tiled<double>(rec, rp, idx, vals, n_rows, n_tiles, n_tiles2d, rows2d, stride, strip_rows, strips, cap_rows,
                                cap_nz, X, k, Y, ldy, alpha, shift, diag, accumulate, dots, dots_row, gate,
                                reinterpret_cast<cudaStream_t>(stream));
}
}
