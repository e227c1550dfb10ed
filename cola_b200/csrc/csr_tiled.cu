// Staged CSR SpMM for wide right-hand-side blocks on patterns with long column runs (stencil / banded matrices:
// BASELINE config 2).  Replaces Sparse._matmat (cola/ops/operators.py:77-78) + the p^T A p reduction (cg.py:157-158) like
// csr_spmm.cu, from the tile-local form of the pattern built by cola_b200/csr_tiles.py:
//   * a tile is S strips of R consecutive rows, `stride` rows apart (the pattern's dominant far diagonal), so the rows
//     of X gathered across that diagonal belong to the neighbouring strips of the SAME tile;
//   * the distinct rows of X a tile touches (its own rows included: the fused epilogue's operand) arrive in shared memory
//     ONCE, as a few contiguous runs, by bulk copies (cp.async.bulk, one per run, issued by a producer warp into a ring of
//     stages with full / empty mbarriers); the non-zeros address them by byte offset, so the row loop reads shared memory
//     only: no register gather, no L1 lottery;
//   * L2 -> SM traffic for a 5-point stencil drops from 3 rows of X per output row (register-gather kernel: 3.2 GB per
//     cfg2 SpMM at the ~8 TB/s the L2 fabric delivered it) to 1.3, and the tiles of a CTA run block-fastest (vertical
//     neighbours back to back) so the halo strips are L2 hits: X crosses DRAM 1.05 times;
//   * <x, y> partials stay in registers (fp32 over <= 8 tiles, then fp64) and meet in shared memory once per CTA.
// Tiles the record marks irregular (too many runs / too many distinct rows) gather from global memory in the same loop.
// X must be contiguous (ldx == k) and k * sizeof(T) a multiple of 16.  cfg2 (fp32, k = 64): 0.41 ms = 5.7 TB/s of
// algorithmic bytes (register-gather kernel: 0.58 ms); bring-up table in profiles/r2_cg_cfg2_summary.md.
#include <cstdlib>

#include "sweep.cuh"

namespace cola {

#ifndef TL_UNROLL
#define TL_UNROLL 2      // non-zeros of a row in flight (measured on cfg2: see profiles/r2_cg_cfg2_summary.md)
#endif
constexpr int kTlRec = 32;                       // record words per tile (cola_b200/csr_tiles.py)

template <typename T>
struct TiledArgs {
  const int32_t* rec; const int32_t* rp; const int32_t* idx; const T* vals;
  int64_t n_rows, n_tiles, n_tiles2d, rows2d, stride, tiles_per_blk, n_blk;
  int strip_rows, strips, rp_stride, cap_rows, cap_nz, n_stages;
  const T* X; int64_t k; T* Y; int64_t ldy;
  T alpha, shift; const T* diag; int accumulate;
  double* dots; const int32_t* dots_row; int64_t k_full; const int32_t* gate;
  int lanes;                                     // threads per row: k * sizeof(T) / 16 / (chunks per thread)
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
  const unsigned char* sx;       // staged X rows, already offset to this thread's first 16-byte chunk
  const int32_t* rp; const int32_t* sidx; const T* sval;
  const T* Xc; T* Yc;
  int64_t row0, stride;
  int RT, R, g, groups, s0, r0, ds, dr;
};

// Rows g, g + groups, ... of one tile.  A thread owns CH 16-byte chunks of its row, `lanes` chunks apart (neighbouring
// lanes read neighbouring chunks: no bank conflicts, whole sectors per store), so the per-non-zero overhead (value,
// offset, address) is paid once per CH * 16 bytes.  REGULAR: sidx / the own-row table hold BYTE offsets of staged rows
// (slot * row bytes, from csr_tiles.py); else sidx holds columns and the rows are gathered from global memory.
template <typename T, bool EPI, bool DOTS, bool REGULAR, int CH>
__device__ __forceinline__ void tl_tile_rows(const TiledArgs<T>& a, const TlRows<T>& w, T* facc) {
  constexpr int VEC = 16 / (int)sizeof(T);
  const int32_t* self_off = w.rp + w.RT + 4;
  const int cstep = a.lanes * 16;                 // bytes between a thread's chunks
  const int cstep_e = a.lanes * VEC;              // ... in elements
  int s = w.s0, r = w.r0;
  for (int lr = w.g; lr < w.RT; lr += w.groups) {
    const int64_t row = w.row0 + s * w.stride + r;
    r += w.dr; s += w.ds;
    if (r >= w.R) { r -= w.R; ++s; }
    if (row >= a.n_rows) continue;
    const int32_t e0 = w.rp[lr], cnt = w.rp[lr + 1] - e0;
    Vec<T, VEC> xo[CH];
    T sd = a.shift;
    if constexpr (EPI) {                                                  // own row: the epilogue's operand
      if constexpr (REGULAR) {
        const unsigned char* px = w.sx + self_off[lr];
#pragma unroll
        for (int c = 0; c < CH; ++c) xo[c] = *reinterpret_cast<const Vec<T, VEC>*>(px + c * cstep);
      } else {
#pragma unroll
        for (int c = 0; c < CH; ++c) xo[c] = ldg<T, VEC>(w.Xc + row * a.k + c * cstep_e);
      }
      if (a.diag) sd += a.diag[row];
    }
    T* yp = w.Yc + row * a.ldy;
    const int32_t* pi = w.sidx + e0;
    const T* pv = w.sval + e0;
    T acc[CH][VEC];
#pragma unroll
    for (int c = 0; c < CH; ++c)
#pragma unroll
      for (int q = 0; q < VEC; ++q) acc[c][q] = (T)0;
#pragma unroll 2
      for (int32_t j = 0; j < cnt; ++j) {
        const T wv = pv[j];
        Vec<T, VEC> x[CH];
        if constexpr (REGULAR) {
          const unsigned char* px = w.sx + pi[j];
#pragma unroll
          for (int c = 0; c < CH; ++c) x[c] = *reinterpret_cast<const Vec<T, VEC>*>(px + c * cstep);
        } else {
          const T* px = w.Xc + (int64_t)pi[j] * a.k;
#pragma unroll
          for (int c = 0; c < CH; ++c) x[c] = ldg<T, VEC>(px + c * cstep_e);
        }
#pragma unroll
        for (int c = 0; c < CH; ++c)
#pragma unroll
          for (int q = 0; q < VEC; ++q) acc[c][q] += wv * x[c].v[q];
      }
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      Vec<T, VEC> y;
      if constexpr (EPI) {
#pragma unroll
        for (int q = 0; q < VEC; ++q) y.v[q] = a.alpha * acc[c][q] + sd * xo[c].v[q];
      } else {
#pragma unroll
        for (int q = 0; q < VEC; ++q) y.v[q] = a.alpha * acc[c][q];
      }
      if (a.accumulate) {
        const Vec<T, VEC> yo = ldg<T, VEC>(yp + c * cstep_e);
#pragma unroll
        for (int q = 0; q < VEC; ++q) y.v[q] += yo.v[q];
      }
      if constexpr (EPI && DOTS) {
#pragma unroll
        for (int q = 0; q < VEC; ++q) facc[c * VEC + q] += xo[c].v[q] * y.v[q];
      }
      stg_stream<T, VEC>(yp + c * cstep_e, y);
    }
  }
}

// position u of a CTA's tile sequence -> tile.  The 2-D tiles run block-fastest: consecutive tiles of a CTA are vertical
// neighbours (same strip columns, next block of strips), so the halo strip a tile reads above / below itself is the
// previous / next tile's own strip: still in L2 (measured: X read 1.25x -> ~1.05x from DRAM).
template <typename T>
__device__ __forceinline__ int64_t tl_tile_of(const TiledArgs<T>& a, int64_t u) {
  if (u >= a.n_tiles2d || a.n_blk == 0) return u;
  const int64_t c = u / a.n_blk, b = u - c * a.n_blk;
  return b * a.tiles_per_blk + c;
}

template <typename T, bool EPI, bool DOTS, int CH>
__global__ void __launch_bounds__(CH == 2 ? 544 : 768) __maxnreg__(CH == 2 ? 96 : 80) csr_spmm_tiled_kernel(TiledArgs<T> a) {
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
    for (int i = threadIdx.x; i < a.k; i += blockDim.x) s_dots[i] = 0.0;
  }
  __syncthreads();

  // this CTA's tiles: one contiguous chunk of the tile sequence
  const int64_t per = (a.n_tiles + gridDim.x - 1) / gridDim.x;
  const int64_t u_begin = (int64_t)blockIdx.x * per;
  const int64_t u_end = u_begin + per < a.n_tiles ? u_begin + per : a.n_tiles;
  const uint32_t row_bytes = (uint32_t)(a.k * (int64_t)sizeof(T));

  if (warp == a.consumers / 32) {
    // ===== producer warp: lane j copies run j; lanes 12 / 13 / 14 the offsets, values and row pointers =====
    int32_t v = (u_begin < u_end) ? a.rec[tl_tile_of(a, u_begin) * kTlRec + lane] : 0;
    int s = 0;
    uint32_t parity = 1;
    for (int64_t u = u_begin; u < u_end; ++u) {
      const int64_t t = tl_tile_of(a, u);
      const int32_t cur = v;
      if (u + 1 < u_end) v = a.rec[tl_tile_of(a, u + 1) * kTlRec + lane];   // next record while this tile's copies are issued
      const uint32_t full = bar0 + 8 * s, empty = bar0 + 8 * (a.n_stages + s);
      const int32_t nz_begin = __shfl_sync(0xffffffffu, cur, 0), nz_pad = __shfl_sync(0xffffffffu, cur, 1);
      const int32_t n_runs = __shfl_sync(0xffffffffu, cur, 2), n_dist = __shfl_sync(0xffffffffu, cur, 3);
      const int32_t col0 = __shfl_sync(0xffffffffu, cur, (8 + 2 * lane) & 31);
      const int32_t sl = __shfl_sync(0xffffffffu, cur, (9 + 2 * lane) & 31);
      if (lane == 0) {
        while (!tl_mbar_try(empty, parity)) __nanosleep(100);   // (a spinning warp costs its scheduler issue slots)
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
      if (++s == a.n_stages) { s = 0; parity ^= 1; }
    }
  } else {
    // ===== consumers: `lanes` threads per row, CH 16-byte column chunks each =====
    const int tid = threadIdx.x;
    const int g = tid / a.lanes, l = tid - g * a.lanes;
    const int groups = a.consumers / a.lanes;
    const bool col_ok = g < groups;
    TlRows<T> w;
    w.RT = a.strip_rows * a.strips; w.R = a.strip_rows; w.g = g; w.groups = groups;
    w.s0 = g / a.strip_rows; w.r0 = g - w.s0 * a.strip_rows;          // local row g + i * groups as (strip, row in strip),
    w.ds = groups / a.strip_rows; w.dr = groups - w.ds * a.strip_rows;  // advanced without a division per row
    w.Xc = a.X + (size_t)l * VEC; w.Yc = a.Y + (size_t)l * VEC;
    T facc[CH * VEC];                                                  // <x, y> partials: fp32 over a few tiles, then fp64
#pragma unroll
    for (int q = 0; q < CH * VEC; ++q) facc[q] = (T)0;
    double dacc[CH * VEC];
#pragma unroll
    for (int q = 0; q < CH * VEC; ++q) dacc[q] = 0.0;
    int s = 0, since_fold = 0;
    uint32_t parity = 0;
    for (int64_t u = u_begin; u < u_end; ++u) {
      const int64_t t = tl_tile_of(a, u);
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
      if (col_ok) {
        if (w.rp[w.RT + 3] >= 0) tl_tile_rows<T, EPI, DOTS, true, CH>(a, w, facc);
        else tl_tile_rows<T, EPI, DOTS, false, CH>(a, w, facc);
      }
      __syncwarp();
      if (lane == 0) tl_mbar_arrive(bar0 + 8 * (a.n_stages + s));   // this warp no longer reads the stage
      if (++s == a.n_stages) { s = 0; parity ^= 1; }
      if constexpr (DOTS) {
        if (++since_fold == 8 || u + 1 == u_end) {                  // fp32 partials of <= 8 tiles, folded in fp64 registers
          since_fold = 0;
#pragma unroll
          for (int q = 0; q < CH * VEC; ++q) { dacc[q] += (double)facc[q]; facc[q] = (T)0; }
        }
      }
    }
    if constexpr (DOTS) {                                           // once per CTA (shared fp64 atomics are CAS loops)
      if (col_ok) {
#pragma unroll
        for (int c = 0; c < CH; ++c)
#pragma unroll
          for (int q = 0; q < VEC; ++q) atomicAdd(s_dots + (c * a.lanes + l) * VEC + q, dacc[c * VEC + q]);
      }
    }
  }
  if constexpr (DOTS) {
    __syncthreads();
    double* out = a.dots + (a.dots_row ? (int64_t)(*a.dots_row) * a.k_full : 0);
    for (int i = threadIdx.x; i < a.k; i += blockDim.x) atomicAdd(out + i, s_dots[i]);
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
  a.n_blk = n_tiles2d > 0 ? n_tiles2d / a.tiles_per_blk : 1;
  if (const char* e = getenv("COLA_SPMM_TILE_ORDER")) { if (atoi(e) == 0) a.n_blk = 0; }
  a.strip_rows = (int)strip_rows; a.strips = (int)strips; a.rp_stride = (int)(2 * strip_rows * strips + 4);
  a.cap_rows = (int)cap_rows; a.cap_nz = (int)cap_nz;
  a.X = X; a.k = k; a.Y = Y; a.ldy = ldy; a.alpha = alpha; a.shift = shift; a.diag = diag; a.accumulate = accumulate;
  a.dots = dots; a.dots_row = dots_row; a.k_full = k; a.gate = gate;
  const int chunks = (int)(k / VEC);
  // chunks per thread: 2 halves the per-non-zero overhead (value, offset, address) per byte; measured on cfg2 (fp32, k = 64):
  // 1 chunk x 23 warps 0.482 ms, 2 x 16 0.411 ms, 4 x 8..15 >= 0.60 ms
  int ch = chunks % 2 == 0 ? 2 : 1;
  if (const char* e = getenv("COLA_SPMM_TILE_CH")) { const int v = atoi(e); if ((v == 1 || v == 2) && chunks % v == 0) ch = v; }
  a.lanes = chunks / ch;
  const int max_consumers = ch == 2 ? 512 : 736;   // + the producer warp (registers are per SM sub-partition: 5 warps x 96, 6 x 80)
  a.consumers = max_consumers;
  if (const char* e = getenv("COLA_SPMM_TILE_WARPS")) { const int v = atoi(e); if (v >= 1 && 32 * v <= max_consumers) a.consumers = 32 * v; }
  COLA_REQUIRE(a.lanes <= max_consumers, "csr_spmm_tiled: row too wide");
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
  const int64_t extra = 16 * 8 + 64 + k * 8 + 128;
  int n_stages = (int)(((int64_t)smem_max - extra) / a.stage_bytes);
  if (n_stages > 6) n_stages = 6;
  if (n_stages < 2) return fail(COLA_E_UNSUPPORTED, "csr_spmm_tiled: a ring of two stages does not fit shared memory");
  a.n_stages = n_stages;
  const size_t smem = (size_t)n_stages * a.stage_bytes + extra;
  const bool epi = (shift != (T)0) || diag || dots;
  int64_t grid = sm_count();
  if (grid > n_tiles) grid = n_tiles;
#define COLA_TILED_LAUNCH(EPIV, DOTSV, CHV)                                                      \
  do {                                                                                              \
    auto kern = csr_spmm_tiled_kernel<T, EPIV, DOTSV, CHV>;                                    \
    static int attr_smem = 0;                                                                       \
    if ((int)smem > attr_smem) {                                                                    \
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);           \
      attr_smem = (int)smem;                                                                        \
    }                                                                                               \
    kern<<<(unsigned)grid, a.consumers + 32, smem, st>>>(a);                                        \
  } while (0)
#define COLA_TILED_EPI(CHV)                                  \
  do {                                                       \
    if (dots) COLA_TILED_LAUNCH(true, true, CHV);            \
    else if (epi) COLA_TILED_LAUNCH(true, false, CHV);       \
    else COLA_TILED_LAUNCH(false, false, CHV);               \
  } while (0)
  if (ch == 2) COLA_TILED_EPI(2);
  else COLA_TILED_EPI(1);
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
