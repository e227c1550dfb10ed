// Generic streaming sweep over an (n, k) row-major block with fused per-column reductions.
//
// All vector-update kernels of the Krylov loops (CG x/r/p updates, Lanczos three-term step, MGS links,
// column dots / scalings) are instances of this one template: every element of every operand is read
// exactly once with the widest aligned vector access, per-column partial sums live in registers (a thread
// always owns the same columns), are tree-reduced across the block's rows in shared memory and leave the
// SM as one fp64 atomicAdd per column per block.  HBM-bound by construction: bytes = sum of operand sizes.
#pragma once
#include <cooperative_groups.h>

#include <cstdlib>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace cola {

constexpr int kSweepThreads = 256;

// Block-level reduction of per-thread column partials: thread (r,l) holds VEC partial sums for columns
// c0..c0+VEC-1; partials of equal l are tree-summed over r in shared memory (r need not span a power of two)
// and row 0 issues one fp64 atomicAdd per column.  `red` must hold blockDim.x*VEC doubles.
template <int VEC>
__device__ __forceinline__ void block_col_reduce(double* red, const double (&acc)[VEC], bool active, int tid, int r,
                                                 int l, int lanes, int rows, int64_t c0, int64_t k, int64_t colmask,
                                                 double* out, int cluster = 1) {
  __syncthreads();
#pragma unroll
  for (int v = 0; v < VEC; ++v) red[tid * VEC + v] = active ? acc[v] : 0.0;
  __syncthreads();
  int span = 1;
  while (span < rows) span <<= 1;
  for (int s = span >> 1; s > 0; s >>= 1) {
    if (r < s && r + s < rows) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) red[(r * lanes + l) * VEC + v] += red[((r + s) * lanes + l) * VEC + v];
    }
    __syncthreads();
  }
  if (cluster > 1) {
    // the CTAs of a thread-block cluster fold their column sums through distributed shared memory and leave as ONE
    // atomic per column per cluster: same-address fp64 atomics serialise at L2 (~40 ns each: 1184 CTAs x 128 columns cost
    // a 45 us tail on a 60 us sweep), so an 8-CTA cluster cuts that tail 8x
    cg::cluster_group cl = cg::this_cluster();
    cl.sync();                                            // every CTA's row 0 of `red` is final
    if (cl.block_rank() == 0 && r == 0 && c0 < k) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        double sum = red[l * VEC + v];
        for (unsigned j = 1; j < cl.num_blocks(); ++j) sum += cl.map_shared_rank(red, j)[l * VEC + v];
        atomicAdd(out + ((c0 + v) & colmask), sum);
      }
    }
    cl.sync();                                            // nobody's shared memory goes away (or is reused) under the reader
    return;
  }
  if (r == 0 && c0 < k) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) atomicAdd(out + ((c0 + v) & colmask), red[l * VEC + v]);
  }
}

// Op concept:
//   static constexpr int NACC;                       number of per-column reductions
//   struct Regs;                                     registers holding one row-chunk of every operand
//   __device__ bool enabled()                        false => whole kernel is a no-op (device-side gate)
//   __device__ void setup(int64_t c0)                per-thread column scalars (alpha, beta, ...)
//   __device__ void load(int64_t row, int64_t c0, Regs&)            issue the loads
//   __device__ void finish(int64_t row, int64_t c0, Regs&, acc)     arithmetic, stores, partial sums
//   __device__ double* out(int a)                    accumulator row for reduction a (k doubles) or nullptr
// colmask: -1 normally; 0 when an (n,1) vector is viewed as (n/VEC, VEC) ("folded"): every lane then maps to
// column 0 for per-column scalars and reductions.
template <typename T, int VEC, typename Op>
__global__ void __launch_bounds__(kSweepThreads)
    sweep_kernel(int64_t n, int64_t k, int lanes, int rows_per_pass, int64_t colmask, Op op, int cluster) {
  constexpr int NACC = Op::NACC;
  if (!op.enabled()) return;
  const int tid = threadIdx.x;
  const int r = tid / lanes, l = tid - r * lanes;
  const int64_t c0 = (int64_t)l * VEC;
  const bool active = (r < rows_per_pass) && (c0 < k);
  double acc[NACC > 0 ? NACC : 1][VEC];
#pragma unroll
  for (int a = 0; a < (NACC > 0 ? NACC : 1); ++a)
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[a][v] = 0.0;

  if (active) {
    op.setup(c0);
    const int64_t stride = (int64_t)gridDim.x * rows_per_pass;
    int64_t row = (int64_t)blockIdx.x * rows_per_pass + r;
    // two rows in flight per thread: doubles the bytes outstanding per warp
    for (; row + stride < n; row += 2 * stride) {
      typename Op::Regs ra, rb;
      op.load(row, c0, ra);
      op.load(row + stride, c0, rb);
      op.finish(row, c0, ra, acc);
      op.finish(row + stride, c0, rb, acc);
    }
    if (row < n) {
      typename Op::Regs ra;
      op.load(row, c0, ra);
      op.finish(row, c0, ra, acc);
    }
  }

  if constexpr (NACC > 0) {
    __shared__ double red[kSweepThreads * VEC];
#pragma unroll
    for (int a = 0; a < NACC; ++a) {
      double* out = op.out(a);
      if (out == nullptr) continue;  // uniform across the grid
      block_col_reduce<VEC>(red, acc[a], active, tid, r, l, lanes, rows_per_pass, c0, k, colmask, out, cluster);
    }
  }
}

// Host launcher: splits very wide blocks into column slabs of at most kSweepThreads*VEC columns so that
// the kernel's "one column chunk per thread" invariant holds; `make(c_off)` builds the Op for a slab.
template <typename T, int VEC, typename MakeOp>
int launch_sweep(int64_t n, int64_t k, int64_t colmask, cudaStream_t st, MakeOp make, const char* name) {
  if (n <= 0 || k <= 0) return COLA_OK;
  const int64_t slab = (int64_t)kSweepThreads * VEC;
  for (int64_t c = 0; c < k; c += slab) {
    int64_t kk = (k - c < slab) ? (k - c) : slab;
    RowMap m = row_map(kk, VEC, kSweepThreads);
    int64_t tiles = (n + m.rows_per_pass - 1) / m.rows_per_pass;
    int64_t tiles2 = (tiles + 1) / 2;
    // resident CTAs per SM: 8 x 256 threads = full occupancy for the pure streams; 4 for sweeps that end in column reductions
    // (half the cluster syncs and atomics at the end; measured with the cluster reduction on, profiles/r2_sweep_cluster_reduce.log:
    // update_r 64 MB blocks 37.8 -> 29.5 us, 128 MB 76.8 -> 72.6, 1 GB 517 -> 513; col_dots 53.6 -> 47.9 on 128 MB; axpby loses 2 %)
    static const int per_sm_env = [] { const char* e = getenv("COLA_SWEEP_CTAS"); const int v = e ? atoi(e) : 0; return (v >= 1 && v <= 8) ? v : 0; }();
    const int per_sm = per_sm_env ? per_sm_env : (decltype(make(c))::NACC > 0 ? 4 : 8);
    int64_t grid = (int64_t)sm_count() * per_sm;  // whole waves of resident CTAs
    if (grid > tiles2) grid = tiles2 > 0 ? tiles2 : 1;
    auto op = make(c);
    using OpT = decltype(op);
    // kernels with column reductions run as clusters of 8 CTAs (one fp64 atomic per column per cluster, see
    // block_col_reduce); `out(a)` is uniform over the grid, so every CTA of a cluster takes the same path
    static const int want_cluster = [] { const char* e = getenv("COLA_SWEEP_CLUSTER"); const int v = e ? atoi(e) : 8; return (v == 1 || v == 2 || v == 4 || v == 8) ? v : 8; }();
    // ... when the atomics would be many: measured on B200, grid x columns fp64 atomics cost ~0.3 ns each at the end of the
    // kernel (1184 CTAs x 128 columns: 45 us), a cluster launch ~3 us; below ~48K atomics the plain launch is faster
    int cluster = (OpT::NACC > 0 && grid >= 2 * want_cluster && grid * kk >= 48 * 1024) ? want_cluster : 1;
    if (cluster > 1) {
      grid = (grid + cluster - 1) / cluster * cluster;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)grid);
      cfg.blockDim = dim3(kSweepThreads);
      cfg.dynamicSmemBytes = 0;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = (unsigned)cluster;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      cudaError_t e = cudaLaunchKernelEx(&cfg, sweep_kernel<T, VEC, OpT>, n, kk, m.lanes, m.rows_per_pass, colmask, op, cluster);
      if (e != cudaSuccess) {          // (not expected on sm_100: fall back to the plain launch)
        cudaGetLastError();
        cluster = 1;
      }
    }
    if (cluster == 1)
      sweep_kernel<T, VEC, OpT><<<(unsigned)grid, kSweepThreads, 0, st>>>(n, kk, m.lanes, m.rows_per_pass, colmask, op, 1);
    int rc = cuda_status(name);
    if (rc) return rc;
  }
  return COLA_OK;
}

// dispatch on the vector width
#define COLA_DISPATCH_VEC(T, vec, CALL)          \
  do {                                           \
    if constexpr (sizeof(T) == 4) {              \
      if ((vec) == 4) { constexpr int VEC = 4; CALL; } \
      else if ((vec) == 2) { constexpr int VEC = 2; CALL; } \
      else { constexpr int VEC = 1; CALL; }      \
    } else {                                     \
      if ((vec) == 2) { constexpr int VEC = 2; CALL; } \
      else { constexpr int VEC = 1; CALL; }      \
    }                                            \
  } while (0)

}  // namespace cola
