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
#include <cstdlib>

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
  int l2_prefetch, row_bytes;
  int l2_hints;   // SpMV: evict_last on the X gathers, evict-first on row pointers and Y (COLA_SPMV_L2_HINTS=0: off)
};

// One staged non-zero: element offset of its X row (col * ldx, computed ONCE per non-zero while staging
// instead of once per lane in the inner loop) and its value.
template <typename T>
struct alignas(sizeof(T) == 4 ? 8 : 16) Nz {
  uint32_t off;
  T val;
};

// OFF32: every X row offset fits 32 bits (n_cols * ldx < 2^32), the common case; otherwise offsets are
// recomputed in 64 bits from the column index.
// NZL: threads that share one row AND one column chunk and split the row's non-zeros between them (their
// partial sums are combined with warp shuffles).  NZL = 1 for wide RHS blocks (the row's k*sizeof(T) bytes already
// fill a half warp); NZL = 8 for SpMV-like shapes (k <= 4), where a thread-per-row walk would expose one
// memory round trip per non-zero.
template <typename T, int VEC, bool EPI, bool DOTS, bool OFF32, int NZL>
__global__ void __launch_bounds__(kCsrThreads, 4) csr_spmm_kernel(CsrArgs<T> a) {
  if (a.gate != nullptr && *a.gate != 0) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: nz[cap] | rp[rows_per_tile+1] | red[256*VEC doubles]
  Nz<T>* s_nz = reinterpret_cast<Nz<T>*>(smem_raw);
  int32_t* s_rp = reinterpret_cast<int32_t*>(s_nz + a.cap);
  const int tid = threadIdx.x;
  const int gw = a.lanes * NZL;                       // threads per row
  const int g = tid / gw, rem = tid - g * gw;
  const int z = rem / a.lanes, l = rem - z * a.lanes;  // non-zero lane, column lane
  const int c0 = l * VEC;
  const bool col_ok = (g < a.groups) && (c0 < a.k);
  const T* __restrict__ Xc = a.X + c0;
  T* __restrict__ Yc = a.Y + c0;
  const uint32_t ldx32 = (uint32_t)a.ldx;
  const uint32_t gmask = NZL > 1 ? (gw >= 32 ? 0xffffffffu : (((1u << gw) - 1u) << ((tid & 31) & ~(gw - 1)))) : 0u;

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
        Nz<T> e;
        const uint32_t c = (uint32_t)__ldcs(a.colidx + base + i);
        e.off = OFF32 ? c * ldx32 : c;
        e.val = __ldcs(a.vals + base + i);
        s_nz[i] = e;
        if (a.l2_prefetch) {
          // The gather below is latency-bound (ncu: long-scoreboard stalls, DRAM at 36 %).  Requesting every X row
          // of the tile into L2 now -- one thread per non-zero, no registers held, all requests in flight at
          // once -- turns the dependent loads of the row loop into L2 hits.
          const char* xr = reinterpret_cast<const char*>(a.X + (size_t)c * (size_t)a.ldx);
          for (int b = 0; b < a.row_bytes; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(xr + b));
        }
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
        for (int32_t j = s + z; j < e; j += NZL) {
          const Nz<T> nz = s_nz[j];
          const Vec<T, VEC> x = ldg<T, VEC>(OFF32 ? (Xc + nz.off) : (Xc + (uint64_t)nz.off * (uint64_t)a.ldx));
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[v] += nz.val * x.v[v];
        }
      } else {  // tile too heavy for the staging buffer: read the CSR arrays directly (rare, kept small)
#pragma unroll 1
        for (int32_t j = s + z; j < e; j += NZL) {
          const int64_t c = a.colidx[base + j];
          const T w = a.vals[base + j];
          Vec<T, VEC> x = ldg<T, VEC>(Xc + c * a.ldx);
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[v] += w * x.v[v];
        }
      }
      if constexpr (NZL > 1) {   // combine the non-zero lanes' partial sums (group-local shuffle mask)
#pragma unroll
        for (int v = 0; v < VEC; ++v)
          for (int o = a.lanes; o < gw; o <<= 1) acc[v] += __shfl_xor_sync(gmask, acc[v], o);
        if (z != 0) continue;
      }
      Vec<T, VEC> y;
      if constexpr (EPI) {
        const Vec<T, VEC> xo = ldg<T, VEC>(Xc + row * a.ldx);   // own row: an L1 hit (it is one of the gathered rows)
        const T d = a.diag ? a.diag[row] : (T)0;
        Vec<T, VEC> yo;
        if (a.accumulate) yo = ldg<T, VEC>(Yc + row * a.ldy);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          T t = a.alpha * acc[v];
          if (a.shift != (T)0) t += a.shift * xo.v[v];
          if (a.diag) t += d * xo.v[v];
          if (a.accumulate) t += yo.v[v];
          y.v[v] = t;
          if constexpr (DOTS) dacc[v] += (double)xo.v[v] * (double)t;
        }
      } else {
        if (a.accumulate) {
          Vec<T, VEC> yo = ldg<T, VEC>(Yc + row * a.ldy);
#pragma unroll
          for (int v = 0; v < VEC; ++v) y.v[v] = a.alpha * acc[v] + yo.v[v];
        } else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) y.v[v] = a.alpha * acc[v];
        }
      }
      stg_stream<T, VEC>(Yc + row * a.ldy, y);
    }
  }

  if constexpr (DOTS) {
    double* red = reinterpret_cast<double*>(s_rp + a.rows_per_tile + 2);
    red = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(red) + 7) & ~(uintptr_t)7);
    double* out = a.dots + (a.dots_row ? (int64_t)(*a.dots_row) * a.k_full : 0);
    // partials live in the z == 0 threads; reduce over (group, nz-lane) pairs as "rows" of width lanes
    block_col_reduce<VEC>(red, dacc, col_ok && z == 0, tid, tid / a.lanes, l, a.lanes, kCsrThreads / a.lanes,
                          (int64_t)c0, a.k, -1, out);
  }
}

// ---------------------------------------------------------------------------------------------------
// Software-pipelined variant for wide right-hand-side blocks (NZL = 1): the CSR staging and the L2 prefetch of the
// gathered X rows run PF tiles AHEAD of the row loop, so the row loop's dependent loads are L2 hits instead of
// DRAM round trips and no staging phase (rowptr -> colidx -> X, two exposed DRAM latencies plus two barriers per
// tile in the kernel above) sits on the critical path:
//     iteration i:  wait(cp.async) + one barrier | prefetch.L2 X rows of tile i+PF | cp.async colidx/vals of tile
//                   i+PF+1 -> smem ring | cp.async rowptr slice of tile i+PF+2 | row loop of tile i
// The ring holds PF+2 tiles of (colidx, vals) and PF+3 rowptr slices; cp.async (LDGSTS) lands in shared memory
// without holding registers or stalling the issuing warp.  ncu of the non-pipelined kernel on cfg2: long_scoreboard
// 9.6 of 16.5 warps/issue, DRAM 36 %, L2 25 % -- a latency-bound gather; this removes the DRAM part of that latency.
// The row loop gathers up to BATCH non-zeros of a row in one predicated batch (one round trip for rows of <= BATCH
// non-zeros instead of ceil(nnz/4) + remainder).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int BYTES>
__device__ __forceinline__ void cp_async_small(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// CH: 16-byte column chunks per thread.  Thread l of a row group owns columns [(h*lanes + l)*VEC, +VEC) for h < CH, so
// every load instruction of the group covers lanes*16 contiguous bytes.  CH = 2 halves the per-non-zero overhead
// (shared-memory read of (col, val), address arithmetic, loop control) per FMA: the first kernel executed ~110 warp
// instructions per row against ~13 of loads + FMAs + stores and ran at 54 % issue utilisation -- issue-bound as much as
// latency-bound.
// 16 bytes of X at base + col * pitch: one mad.wide.u32 (IMAD.WIDE.U32) for the address + one 128-bit read-only load
template <typename T>
__device__ __forceinline__ Vec<T, 16 / (int)sizeof(T)> gather16(uint64_t base, uint32_t col, uint32_t pitch) {
  Vec<T, 16 / (int)sizeof(T)> r;
  if constexpr (sizeof(T) == 4) {
    asm volatile("{ .reg .u64 a; mad.wide.u32 a, %4, %5, %6; ld.global.nc.v4.f32 {%0,%1,%2,%3}, [a]; }"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "r"(col), "r"(pitch), "l"(base));
  } else {
    asm volatile("{ .reg .u64 a; mad.wide.u32 a, %2, %3, %4; ld.global.nc.v2.f64 {%0,%1}, [a]; }"
                 : "=d"(r.v[0]), "=d"(r.v[1]) : "r"(col), "r"(pitch), "l"(base));
  }
  return r;
}

template <typename T, int CH, bool EPI, bool DOTS, int PF, int BATCH, int MINB>
__global__ void __launch_bounds__(kCsrThreads, MINB) csr_spmm_pipe_kernel(CsrArgs<T> a) {
  if (a.gate != nullptr && *a.gate != 0) return;
  constexpr int VEC = 16 / (int)sizeof(T);
  constexpr int NZS = PF + 2, RPS = PF + 3;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: nz[NZS][cap] (col, val) | rp[RPS][rows_per_tile + 1] | red[256*VEC doubles]
  Nz<T>* s_nz = reinterpret_cast<Nz<T>*>(smem_raw);
  int32_t* s_rp = reinterpret_cast<int32_t*>(s_nz + (size_t)NZS * a.cap);
  const int tid = threadIdx.x;
  const int g = tid / a.lanes, l = tid - g * a.lanes;
  const bool col_ok = g < a.groups;
  const uint32_t ldxb = (uint32_t)(a.ldx * (int64_t)sizeof(T));      // row pitch of X in bytes (< 2^32, checked by the launcher)
  const uint32_t chunk_b = (uint32_t)a.lanes * 16u;                   // byte distance between a thread's column chunks
  // The thread's two column-chunk base addresses, made opaque to the compiler: otherwise it keeps "parameter + lane
  // offset" symbolic and re-derives every gather address with LDC.64 + IMAD.WIDE + IADD3 + IADD3.X (+ shifts); as an
  // opaque 64-bit register a gather address is ONE IMAD.WIDE.U32 (col * pitch + base).
  uint64_t xb0 = reinterpret_cast<uint64_t>(a.X) + (uint64_t)l * 16, xb1 = xb0 + chunk_b;
  asm volatile("" : "+l"(xb0), "+l"(xb1));
  const char* __restrict__ Xb = reinterpret_cast<const char*>(a.X) + (size_t)l * 16;
  char* __restrict__ Yb = reinterpret_cast<char*>(a.Y) + (size_t)l * 16;
  const size_t ldyb = (size_t)a.ldy * sizeof(T);
  const int rpt = a.rows_per_tile, rps = rpt + 1, cap = a.cap;
  const int64_t n_tiles = (a.n_rows + rpt - 1) / rpt;

  // fp64 column accumulators: thread-private slots in shared memory (slot q of thread t at [q*256 + t]: conflict-free),
  // touched once per tile -- 16 registers less in the row loop
  double* s_dacc = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(s_rp + (size_t)RPS * rps + 1) + 7) & ~(uintptr_t)7);
  if constexpr (DOTS) {
#pragma unroll
    for (int q = 0; q < CH * VEC; ++q) s_dacc[q * kCsrThreads + tid] = 0.0;
  }

  auto tile_of = [&](int i) -> int64_t { return (int64_t)blockIdx.x + (int64_t)i * gridDim.x; };
  auto rows_of = [&](int64_t t) -> int { return (int)min((int64_t)rpt, a.n_rows - t * rpt); };
  auto load_rp = [&](int i) {
    const int64_t t = tile_of(i);
    if (t >= n_tiles) return;
    const int rows = rows_of(t);
    int32_t* dst = s_rp + (i % RPS) * rps;
    const int32_t* src = a.rowptr + t * rpt;
    for (int j = tid; j <= rows; j += kCsrThreads) cp_async_small<4>(dst + j, src + j);
  };
  auto stage = [&](int i) {   // rowptr slice of tile i has landed and is visible
    const int64_t t = tile_of(i);
    if (t >= n_tiles) return;
    const int32_t* rp = s_rp + (i % RPS) * rps;
    const int32_t base = rp[0], nnz = rp[rows_of(t)] - base;
    if (nnz > cap) return;    // heavy tile: its row loop reads the CSR arrays directly
    Nz<T>* dn = s_nz + (size_t)(i % NZS) * cap;
    for (int j = tid; j < nnz; j += kCsrThreads) {
      cp_async_small<4>(&dn[j].off, a.colidx + base + j);
      cp_async_small<(int)sizeof(T)>(&dn[j].val, a.vals + base + j);
    }
  };
  auto prefetch = [&](int i) {   // staged colidx of tile i has landed and is visible
    const int64_t t = tile_of(i);
    if (t >= n_tiles) return;
    const int32_t* rp = s_rp + (i % RPS) * rps;
    const int32_t nnz = rp[rows_of(t)] - rp[0];
    if (nnz > cap) return;
    const Nz<T>* dn = s_nz + (size_t)(i % NZS) * cap;
    for (int j = tid; j < nnz; j += kCsrThreads) {
      const char* xr = reinterpret_cast<const char*>(a.X) + (uint64_t)dn[j].off * (uint64_t)ldxb;
      for (int b = 0; b < a.row_bytes; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(xr + b));
    }
  };

  // prologue: rowptr slices of tiles 0..PF+1, then the CSR ranges of tiles 0..PF
  for (int i = 0; i <= PF + 1; ++i) load_rp(i);
  cp_async_commit();
  cp_async_wait_all();
  __syncthreads();
  for (int i = 0; i <= PF; ++i) stage(i);
  cp_async_commit();

  for (int i = 0; tile_of(i) < n_tiles; ++i) {
    cp_async_wait_all();
    __syncthreads();   // staged data of tiles <= i+PF and rowptr of tile i+PF+1 visible; tile i-1's row loop finished
    if (a.l2_prefetch) {
      if (i == 0)
        for (int q = 0; q < PF; ++q) prefetch(q);
      prefetch(i + PF);
    }
    stage(i + PF + 1);
    load_rp(i + PF + 2);
    cp_async_commit();
    if (!col_ok) continue;

    const int64_t tile = tile_of(i);
    const int64_t row0 = tile * rpt;
    const int rows = rows_of(tile);
    const int32_t* rp = s_rp + (i % RPS) * rps;
    const int32_t base = rp[0];
    const bool staged = rp[rows] - base <= cap;
    const Nz<T>* tn = s_nz + (size_t)(i % NZS) * cap;
    // per-tile dot partials in T (fp32: a handful of rows per thread), folded into the fp64 accumulators once per
    // tile: the two F2F.F64.F32 conversions per element of a per-element fp64 product cost 0.13 ms of the 0.78 ms kernel
    T facc[CH][VEC];
#pragma unroll
    for (int h = 0; h < CH; ++h)
#pragma unroll
      for (int v = 0; v < VEC; ++v) facc[h][v] = (T)0;

    for (int r = g; r < rows; r += a.groups) {
      const int64_t row = row0 + r;
      const int32_t s = rp[r] - base, e = rp[r + 1] - base;
      T acc[CH][VEC];
#pragma unroll
      for (int h = 0; h < CH; ++h)
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[h][v] = (T)0;
      if (staged) {
        for (int32_t j = s; j < e; j += BATCH) {
          // one predicated batch: all loads of up to BATCH non-zeros in flight, then predicated FMAs (no zero fill)
          Vec<T, VEC> x[BATCH][CH];
          T w[BATCH];
#pragma unroll
          for (int u = 0; u < BATCH; ++u) {
            if (j + u < e) {
              const Nz<T> nz = tn[j + u];
              w[u] = nz.val;
              x[u][0] = gather16<T>(xb0, nz.off, ldxb);
              if constexpr (CH > 1) x[u][1] = gather16<T>(xb1, nz.off, ldxb);
            }
          }
#pragma unroll
          for (int u = 0; u < BATCH; ++u) {
            if (j + u < e) {
#pragma unroll
              for (int h = 0; h < CH; ++h)
#pragma unroll
                for (int v = 0; v < VEC; ++v) acc[h][v] += w[u] * x[u][h].v[v];
            }
          }
        }
      } else {
#pragma unroll 1
        for (int32_t j = s; j < e; ++j) {
          const uint32_t c = (uint32_t)a.colidx[base + j];
          const T w = a.vals[base + j];
          const char* xp = Xb + (uint64_t)c * (uint64_t)ldxb;
#pragma unroll
          for (int h = 0; h < CH; ++h) {
            const Vec<T, VEC> x = ldg<T, VEC>(reinterpret_cast<const T*>(xp + h * chunk_b));
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[h][v] += w * x.v[v];
          }
        }
      }
      const char* xrow = Xb + (uint64_t)row * (uint64_t)ldxb;
      char* yrow = Yb + (size_t)row * ldyb;
      T sd = (T)0;
      if constexpr (EPI) sd = a.shift + (a.diag ? a.diag[row] : (T)0);      // (shift + diag_i) x_i as one term
#pragma unroll
      for (int h = 0; h < CH; ++h) {
        Vec<T, VEC> y;
        if constexpr (EPI) {
          const Vec<T, VEC> xo = ldg<T, VEC>(reinterpret_cast<const T*>(xrow + h * chunk_b));   // own row: an L1 hit
#pragma unroll
          for (int v = 0; v < VEC; ++v) y.v[v] = a.alpha * acc[h][v] + sd * xo.v[v];
          if (a.accumulate) {
            const Vec<T, VEC> yo = ldg<T, VEC>(reinterpret_cast<const T*>(yrow + h * chunk_b));
#pragma unroll
            for (int v = 0; v < VEC; ++v) y.v[v] += yo.v[v];
          }
          if constexpr (DOTS) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) facc[h][v] += xo.v[v] * y.v[v];
          }
        } else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) y.v[v] = a.alpha * acc[h][v];
          if (a.accumulate) {
            const Vec<T, VEC> yo = ldg<T, VEC>(reinterpret_cast<const T*>(yrow + h * chunk_b));
#pragma unroll
            for (int v = 0; v < VEC; ++v) y.v[v] += yo.v[v];
          }
        }
        stg_stream<T, VEC>(reinterpret_cast<T*>(yrow + h * chunk_b), y);
      }
    }
    if constexpr (DOTS) {
#pragma unroll
      for (int h = 0; h < CH; ++h)
#pragma unroll
        for (int v = 0; v < VEC; ++v) s_dacc[(h * VEC + v) * kCsrThreads + tid] += (double)facc[h][v];
    }
  }
  cp_async_wait_all();

  if constexpr (DOTS) {
    double* red = s_dacc + (size_t)CH * VEC * kCsrThreads;
    double* out = a.dots + (a.dots_row ? (int64_t)(*a.dots_row) * a.k_full : 0);
#pragma unroll
    for (int h = 0; h < CH; ++h) {
      double dacc[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) dacc[v] = s_dacc[(h * VEC + v) * kCsrThreads + tid];
      block_col_reduce<VEC>(red, dacc, col_ok, tid, g, l, a.lanes, kCsrThreads / a.lanes,
                            (int64_t)(h * a.lanes + l) * VEC, a.k, -1, out);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// SpMV-like shapes (k <= 4: single-RHS CG, Lanczos with one start vector, cfg5): no shared-memory staging, no
// block barriers.  SUB threads share a row: each reads a strided subset of the row's (colidx, val) pairs straight
// from global memory (consecutive lanes -> consecutive non-zeros: coalesced), gathers X, and the partial sums are
// combined with sub-warp shuffles.  With ~32 registers the SM holds 2048 threads = 256 independent rows in
// flight, which is what hides the rowptr -> colidx -> X dependent-load chain.
// ---------------------------------------------------------------------------------------------------
// gather load that asks L2 to keep the line (evict_last): the slice of X a column strip gathers from is the only reusable
// data of the sweep; everything else (CSR arrays, row pointers, Y) streams through with evict-first hints
template <typename T>
__device__ __forceinline__ T ldg_keep(const T* p, uint64_t pol) {
  if constexpr (sizeof(T) == 8) {
    double v;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return (T)v;
  } else {
    float v;
    asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return (T)v;
  }
}

template <typename T, int KMAX, int SUB, bool EPI, bool DOTS>
__global__ void __launch_bounds__(256) csr_spmv_kernel(CsrArgs<T> a) {
  if (a.gate != nullptr && *a.gate != 0) return;
  uint64_t keep = 0;
  if (a.l2_hints) keep = l2_policy_evict_last();
  const int tid = threadIdx.x;
  const int sub = tid % SUB;
  const int64_t grp = ((int64_t)blockIdx.x * blockDim.x + tid) / SUB;
  const int64_t n_grp = (int64_t)gridDim.x * blockDim.x / SUB;
  const uint32_t gmask = SUB >= 32 ? 0xffffffffu : (((1u << SUB) - 1u) << ((tid & 31) & ~(SUB - 1)));
  const int k = (int)a.k;
  double dacc[KMAX];
#pragma unroll
  for (int c = 0; c < KMAX; ++c) dacc[c] = 0.0;
  for (int64_t row = grp; row < a.n_rows; row += n_grp) {
    const int32_t s = a.l2_hints ? __ldcs(a.rowptr + row) : a.rowptr[row], e = a.l2_hints ? __ldcs(a.rowptr + row + 1) : a.rowptr[row + 1];
    if (!EPI && a.accumulate && s == e) continue;   // a vertical strip without entries in this row: Y stays as it is
    T acc[KMAX];
#pragma unroll
    for (int c = 0; c < KMAX; ++c) acc[c] = (T)0;
#pragma unroll 4
    for (int32_t j = s + sub; j < e; j += SUB) {
      const int64_t col = __ldcs(a.colidx + j);
      const T w = __ldcs(a.vals + j);
      const T* xr = a.X + col * a.ldx;
#pragma unroll
      for (int c = 0; c < KMAX; ++c)
        if (c < k) acc[c] += w * (a.l2_hints ? ldg_keep<T>(xr + c, keep) : __ldg(xr + c));   // read-only path: 32-byte sector fills
    }
#pragma unroll
    for (int c = 0; c < KMAX; ++c)
      for (int o = 1; o < SUB; o <<= 1) acc[c] += __shfl_xor_sync(gmask, acc[c], o);
    if (sub == 0) {
      const T d = (EPI && a.diag) ? a.diag[row] : (T)0;
#pragma unroll
      for (int c = 0; c < KMAX; ++c) {
        if (c < k) {
          T t = a.alpha * acc[c];
          if constexpr (EPI) {
            const T xo = a.X[row * a.ldx + c];
            if (a.shift != (T)0) t += a.shift * xo;
            if (a.diag) t += d * xo;
            if (a.accumulate) t += a.l2_hints ? __ldcs(a.Y + row * a.ldy + c) : a.Y[row * a.ldy + c];
            if constexpr (DOTS) dacc[c] += (double)xo * (double)t;
          } else if (a.accumulate) {
            t += a.l2_hints ? __ldcs(a.Y + row * a.ldy + c) : a.Y[row * a.ldy + c];
          }
          if (a.l2_hints) __stcs(a.Y + row * a.ldy + c, t);
          else a.Y[row * a.ldy + c] = t;
        }
      }
    }
  }
  if constexpr (DOTS) {
    __shared__ double red[8][KMAX];
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int c = 0; c < KMAX; ++c) {
      double v = dacc[c];   // non-zero only in sub == 0 threads
      v = warp_sum(v);
      if (lane == 0) red[warp][c] = v;
    }
    __syncthreads();
    if (tid < k) {
      double v = 0.0;
      for (int w = 0; w < 8; ++w) v += red[w][tid];
      double* out = a.dots + (a.dots_row ? (int64_t)(*a.dots_row) * a.k_full : 0);
      atomicAdd(out + tid, v);
    }
  }
}

template <typename T>
int csr_spmm(const int32_t* rowptr, const int32_t* colidx, const T* vals, int64_t n_rows, int64_t n_cols,
             int64_t nnz, int64_t max_row_nnz, const T* X, int64_t ldx, int64_t k, T* Y, int64_t ldy, T alpha, T shift, const T* diag,
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
    static const bool no_pf = getenv("COLA_CSR_NO_PREFETCH") != nullptr;   // A/B knob
    a.row_bytes = (int)(a.k * (int64_t)sizeof(T));
    a.l2_prefetch = (!no_pf && a.row_bytes >= 32) ? 1 : 0;
    int64_t need = (a.k + vec - 1) / vec;
    a.lanes = (int)need;
    // SpMV-like shapes: share a row between 8 (or 4 / 2) threads when the column lanes are a small power of two
    int nzl = 1;
    if (a.lanes <= 4 && (a.lanes & (a.lanes - 1)) == 0 && a.k == (int64_t)a.lanes * vec && nnz >= 2 * n_rows)
      nzl = 8 / a.lanes >= 2 ? 8 / a.lanes : 1;
    a.groups = kCsrThreads / (a.lanes * nzl);
    // tile size: ~6 rows per group in flight, but keep the staged nnz range near half the buffer
    double avg = n_rows > 0 ? (double)nnz / (double)n_rows : 1.0;
    if (avg < 1.0) avg = 1.0;
    const int cap_max = sizeof(T) == 4 ? 4096 : 3072;   // staging buffer: 32 KB (fp32) / 48 KB (fp64) of (offset, value)
    int rpg = (int)((cap_max / 2) / (avg * a.groups));
    if (rpg < 1) rpg = 1;
    if (rpg > 8) rpg = 8;
    a.rows_per_tile = a.groups * rpg;
    int64_t want = (int64_t)(2.0 * avg * a.rows_per_tile) + 64;   // 2x the mean tile, heavier tiles take the direct path
    a.cap = (int)(want < 512 ? 512 : (want > cap_max ? cap_max : want));
    size_t smem = (size_t)a.cap * sizeof(Nz<T>) + (size_t)(a.rows_per_tile + 4) * 4 + 8 +
                  (dots ? (size_t)kCsrThreads * vec * sizeof(double) : 0);
    int64_t n_tiles = (n_rows + a.rows_per_tile - 1) / a.rows_per_tile;
    if (k <= 4) {   // SpMV-like: sub-warp-per-row kernel, no staging
      a.k = k;
      static const int l2_hints = [] { const char* e = getenv("COLA_SPMV_L2_HINTS"); return (e && atoi(e) == 0) ? 0 : 1; }();
      a.l2_hints = (k <= 2) ? l2_hints : 0;   // cfg5 graph, 45 MB column strips: k = 1 2.26 -> 2.13 ms, k = 2 4.00 -> 3.86, k = 4 7.09 -> 7.44 (off there)
      const double avg_nnz = n_rows > 0 ? (double)nnz / (double)n_rows : 1.0;
      int subw = avg_nnz >= 24 ? 8 : (avg_nnz >= 6 ? 4 : 2);   // measured on the cfg5 graph (17 nnz/row): 4 beats 8 by 4 %
      if (const char* e = getenv("COLA_SPMV_SUB")) { const int v = atoi(e); if (v == 2 || v == 4 || v == 8 || v == 16) subw = v; }
      int64_t groups_needed = n_rows;
      int64_t blocks = (groups_needed * subw + 255) / 256;
      int64_t cap_blocks = (int64_t)sm_count() * 64;   // many short CTAs: the row chains of different CTAs overlap (-9 %)
      if (const char* e = getenv("COLA_SPMV_CTAS")) { const int v = atoi(e); if (v >= 1 && v <= 100000) cap_blocks = (int64_t)sm_count() * v; }
      if (blocks > cap_blocks) blocks = cap_blocks;
#define COLA_SPMV_LAUNCH(SUBV)                                                                            \
  do {                                                                                                    \
    if (dots) csr_spmv_kernel<T, 4, SUBV, true, true><<<(unsigned)blocks, 256, 0, st>>>(a);               \
    else if (epi) csr_spmv_kernel<T, 4, SUBV, true, false><<<(unsigned)blocks, 256, 0, st>>>(a);          \
    else csr_spmv_kernel<T, 4, SUBV, false, false><<<(unsigned)blocks, 256, 0, st>>>(a);                  \
  } while (0)
      if (subw == 16) COLA_SPMV_LAUNCH(16); else if (subw == 8) COLA_SPMV_LAUNCH(8); else if (subw == 4) COLA_SPMV_LAUNCH(4); else COLA_SPMV_LAUNCH(2);
      rc = cuda_status("csr_spmv");
      break;
    }
    const bool off32 = (double)n_cols * (double)ldx < 4294967296.0;
    // wide blocks whose rows split into whole 16-byte chunks, two per thread: the software-pipelined kernel
    static const int env_pipe = getenv("COLA_CSR_PIPE") ? atoi(getenv("COLA_CSR_PIPE")) : 1;      // 0: first kernel (A/B)
    static const int env_rpg = getenv("COLA_CSR_RPG") ? atoi(getenv("COLA_CSR_RPG")) : 8;
    constexpr int kFullVec = 16 / (int)sizeof(T);
    const bool pipe = env_pipe && nzl == 1 && vec == kFullVec && a.k % (2 * kFullVec) == 0 &&
                      (double)n_cols * (double)ldx * sizeof(T) < 4294967296.0 && a.k / (2 * kFullVec) <= kCsrThreads;
    if (pipe) {
      a.lanes = (int)(a.k / (2 * kFullVec));
      a.groups = kCsrThreads / a.lanes;
      a.rows_per_tile = a.groups * (env_rpg > 0 && env_rpg <= 16 ? env_rpg : 8);
      constexpr int pf = 1;
      // measured on cfg2 (B200): prefetch.global.L2 of the gathered rows one tile ahead changes nothing (0.607 vs
      // 0.626 ms): the row loop's ~1 us average miss latency is queueing in the memory system, not a cold L2
      a.l2_prefetch = getenv("COLA_CSR_PIPE_PREFETCH") ? 1 : 0;
      int64_t want2 = (int64_t)(1.5 * avg * a.rows_per_tile) + 64;
      const int cap_max2 = sizeof(T) == 4 ? 3072 : 2048;
      a.cap = (int)(want2 < 256 ? 256 : (want2 > cap_max2 ? cap_max2 : want2));
      n_tiles = (n_rows + a.rows_per_tile - 1) / a.rows_per_tile;
      smem = (size_t)(pf + 2) * a.cap * sizeof(Nz<T>) + (size_t)(pf + 3) * (a.rows_per_tile + 1) * 4 + 16 +
             (dots ? (size_t)kCsrThreads * (2 + 1) * kFullVec * sizeof(double) : 0);   // fp64 accumulator slots + reduction scratch
    }
#define COLA_CSR_PIPE_LAUNCH2(EPIV, DOTSV, PFV, BV, MB)                                                       \
  do {                                                                                                        \
    auto kern = csr_spmm_pipe_kernel<T, 2, EPIV, DOTSV, PFV, BV, MB>;                                         \
    static int attr_smem = 0;                                                                                 \
    if (smem > 48 * 1024 && (int)smem > attr_smem) {                                                          \
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                     \
      attr_smem = (int)smem;                                                                                  \
    }                                                                                                         \
    int per_sm = 0;                                                                                           \
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kCsrThreads, smem);                          \
    if (per_sm < 1) per_sm = 1;                                                                               \
    int64_t grid = (int64_t)sm_count() * per_sm;                                                              \
    if (grid > n_tiles) grid = n_tiles;                                                                       \
    kern<<<(unsigned)grid, kCsrThreads, smem, st>>>(a);                                                       \
  } while (0)
  /* measured on cfg2: batch 6 / 2 CTAs per SM 0.59 ms; batch 4: 0.68; batch 8 (spills): 0.72; 3 CTAs per SM at 80
     registers (spills): 0.77-0.96; prefetch distance 2: 0.58 */
#define COLA_CSR_PIPE_LAUNCH(EPIV, DOTSV) COLA_CSR_PIPE_LAUNCH2(EPIV, DOTSV, 1, 6, 2)
#define COLA_CSR_LAUNCH(EPIV, DOTSV, OFFV)                                                                    \
  do {                                                                                                        \
    if (pipe) { COLA_CSR_PIPE_LAUNCH(EPIV, DOTSV); break; }                                                   \
    auto kern = nzl == 8 ? csr_spmm_kernel<T, VEC, EPIV, DOTSV, OFFV, 8>                                      \
              : nzl == 4 ? csr_spmm_kernel<T, VEC, EPIV, DOTSV, OFFV, 4>                                      \
              : nzl == 2 ? csr_spmm_kernel<T, VEC, EPIV, DOTSV, OFFV, 2>                                      \
                         : csr_spmm_kernel<T, VEC, EPIV, DOTSV, OFFV, 1>;                                     \
    int per_sm = 0;                                                                                           \
    /* raise the dynamic shared-memory limit once per kernel variant: the call is not allowed while a stream   \
       is being captured (the CG loop captures its iteration batches), and the first eager batch sets it */     \
    static int attr_smem[4] = {0, 0, 0, 0};                                                                   \
    const int vi = nzl == 8 ? 3 : nzl == 4 ? 2 : nzl == 2 ? 1 : 0;                                            \
    if (smem > 48 * 1024 && (int)smem > attr_smem[vi]) {                                                      \
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                     \
      attr_smem[vi] = (int)smem;                                                                              \
    }                                                                                                         \
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kCsrThreads, smem);                          \
    if (per_sm < 1) per_sm = 1;                                                                               \
    int64_t grid = (int64_t)sm_count() * per_sm; /* persistent: whole waves of resident CTAs */               \
    if (grid > n_tiles) grid = n_tiles;                                                                       \
    kern<<<(unsigned)grid, kCsrThreads, smem, st>>>(a);                                                       \
  } while (0)
    COLA_DISPATCH_VEC(T, vec, ({
      if (off32) {
        if (dots) COLA_CSR_LAUNCH(true, true, true);
        else if (epi) COLA_CSR_LAUNCH(true, false, true);
        else COLA_CSR_LAUNCH(false, false, true);
      } else {
        if (dots) COLA_CSR_LAUNCH(true, true, false);
        else if (epi) COLA_CSR_LAUNCH(true, false, false);
        else COLA_CSR_LAUNCH(false, false, false);
      }
    }));
    rc = cuda_status("csr_spmm");
  }
  return rc;
}

}  // namespace cola

using namespace cola;
extern "C" {
int cola_csr_spmm_f32(const int32_t* rowptr, const int32_t* colidx, const float* vals, int64_t n_rows,
                      int64_t n_cols, int64_t nnz, int64_t max_row_nnz, const float* X, int64_t ldx, int64_t k, float* Y,
                      int64_t ldy,
                      float alpha, float shift, const float* diag, int accumulate, double* dots,
                      const int32_t* dots_row, const int32_t* gate, void* stream) {
  return csr_spmm<float>(rowptr, colidx, vals, n_rows, n_cols, nnz, max_row_nnz, X, ldx, k, Y, ldy, alpha, shift, diag, accumulate,
                         dots, dots_row, gate, reinterpret_cast<cudaStream_t>(stream));
}
int cola_csr_spmm_f64(const int32_t* rowptr, const int32_t* colidx, const double* vals, int64_t n_rows,
                      int64_t n_cols, int64_t nnz, int64_t max_row_nnz, const double* X, int64_t ldx, int64_t k, double* Y,
                      int64_t ldy,
                      double alpha, double shift, const double* diag, int accumulate, double* dots,
                      const int32_t* dots_row, const int32_t* gate, void* stream) {
  return csr_spmm<double>(rowptr, colidx, vals, n_rows, n_cols, nnz, max_row_nnz, X, ldx, k, Y, ldy, alpha, shift, diag,
                          accumulate, dots, dots_row, gate, reinterpret_cast<cudaStream_t>(stream));
}
}
