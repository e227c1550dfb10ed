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
#include <cuda.h>
#include <cstdlib>
#include "sweep.cuh"

namespace cola {

constexpr int kDotsThreads = 512;  // 16 warps: each warp owns Krylov vectors j = j0 + warp (mod 16)
constexpr int kDotsWarps = kDotsThreads / 32;

template <typename T>
struct RoArgs {
  const T* V; int64_t vstride, j0, j1; const T* Wc; T* W; int64_t n, b;
  double* C; const double* Cc; T sign; double* wnorm2; const int32_t* gate;
  int lanes_per_row, rows_per_chunk, vec_path;
  int fold;   // reorth_dots on a single column viewed as (n / fold, fold): the fold partial columns add into C[j]
  int qsplit; // reorth_dots: row parts per (vector, chunk) work item (see the kernel)
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
    // Work items = (vector j, row part q of the chunk), dealt round-robin to the 16 warps.  With one part per vector a
    // basis of 65 vectors costs 5 rounds for 4.06 rounds of work and a basis of 4 keeps 12 warps idle; Q parts per vector
    // even that out (the parts of a vector then meet in shared memory with an atomic add).
    const int Q = a.qsplit;
    const int part_rows = ((rows + Q - 1) / Q + rows_per_step - 1) / rows_per_step * rows_per_step;
    for (int64_t item = warp; item < nj * Q; item += kDotsWarps) {
      const int64_t j = a.j0 + item / Q;
      const int r_begin = (int)(item % Q) * part_rows;
      const int r_end = min(rows, r_begin + part_rows);
      const T* vj = a.V + j * a.vstride + row0 * b;
      T acc[NC][VEC];
#pragma unroll
      for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[c][v] = (T)0;
#pragma unroll 8
      for (int r = r_begin + rsub; r < r_end; r += rows_per_step) {
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
          if (rsub == 0 && col < b) {
            if (Q == 1) csm[(j - a.j0) * b + col] += s;            // this warp is the only writer of row j
            else atomicAdd(csm + (j - a.j0) * b + col, s);
          }
        }
      }
    }
  }
  __syncthreads();
  for (int64_t i = tid; i < nj * b; i += kDotsThreads) {
    double s = csm[i];
    if (s != 0.0) atomicAdd(a.fold ? (a.C + a.j0 + i / b) : (a.C + a.j0 * b + i), s);
  }
}

// ---- reorth_dots for ONE column (a single Lanczos start vector: BASELINE config 5) -------------------------------
// The column is swept as 16-byte vectors.  A thread keeps its kD1Rows vectors of W in registers for a whole chunk and
// walks the basis: per basis vector kD1Rows independent 16-byte loads (two basis vectors in flight), one partial sum, one
// warp reduction, one add into the warp's own row of shared accumulators (no shared atomics: fp64 ones are CAS loops).
// The general kernel above gives a warp whole basis vectors in 4 KB pieces and reduces after every piece: 5.0-5.6 TB/s
// on this shape where the update sweep over the same bytes reaches 7.0.
constexpr int kD1Threads = 256;
constexpr int kD1Rows = 8;            // 16-byte vectors of W per thread

template <typename T>
__global__ void __launch_bounds__(kD1Threads, 2) reorth_dots1_kernel(RoArgs<T> a) {
  if (a.gate != nullptr && *a.gate != 0) return;
  constexpr int VEC = 16 / (int)sizeof(T);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int64_t nj = a.j1 - a.j0;
  double* csm = reinterpret_cast<double*>(smem_raw);                 // [warp][nj]
  const int tid = threadIdx.x, lane = tid % 32, warp = tid / 32;
  constexpr int kWarps = kD1Threads / 32;
  for (int64_t i = tid; i < nj * kWarps; i += kD1Threads) csm[i] = 0.0;
  __syncthreads();
  double* mine = csm + (int64_t)warp * nj;
  const int64_t nv = a.n;                                             // 16-byte vectors per column
  const int64_t chunk = (int64_t)kD1Threads * kD1Rows;
  const int64_t n_chunks = (nv + chunk - 1) / chunk;
  for (int64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
    const int64_t base = ch * chunk + tid;
    Vec<T, VEC> w[kD1Rows];
    bool ok[kD1Rows];
#pragma unroll
    for (int i = 0; i < kD1Rows; ++i) {
      const int64_t r = base + (int64_t)i * kD1Threads;
      ok[i] = r < nv;
      if (ok[i]) w[i] = ldg_stream<T, VEC>(a.Wc + r * VEC);
      else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) w[i].v[v] = (T)0;
      }
    }
    auto dot_one = [&](const T* vj) {
      Vec<T, VEC> x[kD1Rows];
#pragma unroll
      for (int i = 0; i < kD1Rows; ++i) {
        if (ok[i]) x[i] = ldg_stream<T, VEC>(vj + (base + (int64_t)i * kD1Threads) * VEC);
        else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) x[i].v[v] = (T)0;
        }
      }
      T acc = (T)0;
#pragma unroll
      for (int i = 0; i < kD1Rows; ++i)
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc += x[i].v[v] * w[i].v[v];
      return (double)acc;
    };
    int64_t jj = 0;
    for (; jj + 1 < nj; jj += 2) {                                     // two basis vectors in flight
      const T* v0 = a.V + (a.j0 + jj) * a.vstride;
      double s0 = dot_one(v0), s1 = dot_one(v0 + a.vstride);
      s0 = warp_sum(s0);
      s1 = warp_sum(s1);
      if (lane == 0) { mine[jj] += s0; mine[jj + 1] += s1; }
    }
    if (jj < nj) {
      double s0 = warp_sum(dot_one(a.V + (a.j0 + jj) * a.vstride));
      if (lane == 0) mine[jj] += s0;
    }
  }
  __syncthreads();
  for (int64_t j = tid; j < nj; j += kD1Threads) {
    double s = 0.0;
#pragma unroll
    for (int wp = 0; wp < kWarps; ++wp) s += csm[(int64_t)wp * nj + j];
    if (s != 0.0) atomicAdd(a.C + a.j0 + j, s);
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

// ---------------------------------------------------------------------------------------------------
// Fused CGS2 middle step:  W -= V C1   and   C2 = V^T W_new   with V read ONCE.
// do_double_gram (lanczos.py:287-296) is dots -> update -> dots -> update: four sweeps over the basis.  The
// update of pass 1 and the dots of pass 2 touch the same rows, so a CTA keeps a row-chunk of ALL basis vectors in
// shared memory (R rows x nj vectors), finishes W for those rows, and immediately takes the pass-2 partial dots
// from the resident chunk: 3 sweeps instead of 4 (the algorithmic-byte model of SURVEY 8d counts 4).
//   thread (r, l) owns VEC columns of row r for the update; for the dots each warp owns the vectors
//   j = warp (mod 8) and keeps their partials in registers for the whole kernel (fp32 per lane over <= a few
//   thousand rows, combined in fp64: they are the O(eps) second-pass corrections).
// colmask = 0 folds an (n,1) vector viewed as (n/VEC, VEC) onto column 0.
// ---------------------------------------------------------------------------------------------------
constexpr int kFuMaxThreads = 480;   // consumer threads (15 warps) + one producer warp = 512
constexpr int kFuBoxH = 8;          // basis vectors per TMA box

__device__ __forceinline__ uint32_t fu_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fu_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

// One elected thread fills a ring slot: the chunk [row0, row0+R) of every basis vector comes in as 2-D TMA boxes
// (R*b contiguous elements x 16 vectors; 16-byte cp.async tops out near 3.5 TB/s on this part, ~40 clk per warp
// instruction), the chunk of W as one bulk copy.  Rows past the end of a vector and vectors past nj are
// zero-filled by the TMA unit and still count towards the expected bytes.
template <typename T>
__device__ __forceinline__ void fused_fill(const CUtensorMap* vmap, const RoArgs<T>& a, uint32_t stage, uint32_t bar, int R,
                                           int nj_pad, int64_t row0, int lane, int inner) {
  // called by the whole producer warp: lane 0 arms the barrier, lane i issues box i, lane 31 the W copy.
  // inner == 0: 2-D map, a box is (R*b contiguous elements) x kFuBoxH vectors.  inner > 0: every vector is viewed
  // as rows of `inner` elements and a box is inner x (R*b/inner) x kFuBoxH (chunks wider than the 256-element box limit).
  const int64_t b = a.b;
  const int rows = (int)min((int64_t)R, a.n - row0);
  const uint32_t box_bytes = (uint32_t)(kFuBoxH * R * b * sizeof(T));
  const uint32_t w_bytes = (uint32_t)(rows * b * sizeof(T));
  if (lane == 0)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
                 "r"((uint32_t)(nj_pad / kFuBoxH) * box_bytes + w_bytes) : "memory");
  __syncwarp();
  const int jb = lane * kFuBoxH;
  if (jb < nj_pad) {
    const uint32_t dst = stage + (uint32_t)((int64_t)jb * R * b * sizeof(T));
    if (inner == 0) {
      asm volatile(
          "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
          ::"r"(dst), "l"(vmap), "r"(bar), "r"((int)(row0 * b)), "r"(jb) : "memory");
    } else {
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
          ::"r"(dst), "l"(vmap), "r"(bar), "r"(0), "r"((int)(row0 * b / inner)), "r"(jb) : "memory");
    }
  }
  if (lane == 31)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(stage + (uint32_t)((int64_t)nj_pad * R * b * sizeof(T))), "l"(a.W + row0 * b), "r"(w_bytes), "r"(bar) : "memory");
}

#define FU_CONSUMER_SYNC() asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory")

template <typename T, int VEC, int RPT, int KJ>
__global__ void __launch_bounds__(kFuMaxThreads + 32, 1)
    reorth_fused_kernel(const __grid_constant__ CUtensorMap vmap, RoArgs<T> a, int R, int S, int nj_pad,
                        int64_t stage_elems, int lanes, int64_t colmask, int64_t b_out, int inner, int dbg) {
  if (a.gate != nullptr && *a.gate != 0) return;
  // no static shared memory in this kernel, so the dynamic window starts 1024-byte aligned (TMA needs 128);
  // deriving every pointer directly from the array keeps the accesses in the shared address space (LDS/STS)
  extern __shared__ __align__(1024) unsigned char fu_smem[];
  const int64_t nj = a.j1 - a.j0, b = a.b;
  const int tid = threadIdx.x, lane = tid & 31;
  const int nthr = (int)blockDim.x - 32;                       // consumer threads; the last warp is the producer
  T* bufs = reinterpret_cast<T*>(fu_smem);                       // S x ([nj_pad][R][b] | W [R][b])  ring of chunks
  T* coef = bufs + (int64_t)S * stage_elems;                   // [nj][b]  (sign * C1)
  T* part = coef + nj * b;                                     // [G][R][b] partial sums of the update
  T* wnew = part + (int64_t)kFuMaxThreads * (RPT + 1) * VEC;  // [R][b] finished rows of W (R*b <= kFuMaxThreads*RPT*VEC)
  uint64_t* bars = reinterpret_cast<uint64_t*>(wnew + (int64_t)kFuMaxThreads * RPT * VEC);   // full[S] | empty[S]
  const int64_t n_chunks = (a.n + R - 1) / R;
  const uint32_t bars_u32 = fu_smem_u32(bars), bufs_u32 = fu_smem_u32(bufs);
  if (tid == 0) {
    for (int s2 = 0; s2 < S; ++s2) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fu_smem_u32(bars + s2)), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fu_smem_u32(bars + S + s2)), "r"(nthr / 32));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (tid < nthr) {
    for (int64_t i = tid; i < nj * b; i += nthr) {
      const int64_t j = i / b, c = i - j * b;
      coef[i] = a.sign * (T)a.Cc[(a.j0 + j) * b_out + (c & colmask)];
    }
  }
  __syncthreads();                                             // barriers initialised, coef staged

  if (tid >= nthr) {
    // ---- producer warp: keeps the ring full, waits only on "slot consumed" ----
    int slot = 0;
    uint32_t phase = 0;
    int64_t it = 0;
    for (int64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x, ++it) {
      if (it >= S) fu_mbar_wait(bars_u32 + 8 * (S + slot), phase ^ 1);
      fused_fill<T>(&vmap, a, bufs_u32 + (uint32_t)(slot * stage_elems * sizeof(T)), bars_u32 + 8 * slot, R, nj_pad, ch * R,
                    lane, inner);
      if (++slot == S) { slot = 0; phase ^= 1; }
    }
    return;
  }

  // ---- consumers ----
  // thread (g, rs, l): VEC columns (l) of RPT adjacent rows of the chunk, basis vectors j = g, g + G, ...  Its slice
  // of those vectors is read from shared memory ONCE, into registers: first for the update partials, then -- once
  // W_new is complete -- for the pass-2 dot partials of the same vectors.  The coefficients and the dot partials
  // of a thread's vectors live in registers for the whole kernel.
  const int RL = R * lanes;                                    // Vec slots of one vector's chunk
  const int GS = RL / RPT, G = nthr / GS;
  int Q = 1;
  while (Q * Q < G) ++Q;                                       // ~sqrt(G) adders per output in the first level
  if (Q * RL > nthr) Q = nthr / RL;
  T* part2 = part + (int64_t)nthr * RPT * VEC;                 // [Q][R][b] first-level sums
  const int g = tid / GS, t_in = tid - g * GS;
  const int rs = t_in / lanes, l_up = t_in - rs * lanes;
  const int r0 = rs * RPT;
  const int c_up = l_up * VEC;
  const int Rb = R * (int)b;
  const int elem0 = r0 * (int)b + c_up;                        // offset of this thread's first row inside a chunk
  const bool in_group = g < G;
  const int out_r = tid / lanes, out_e = out_r * (int)b + (tid - out_r * lanes) * VEC;   // level-2 output slot
  const int l1_q = tid / RL, l1_o = tid - l1_q * RL;                                     // level-1 adder slot
  T acc[KJ][VEC];
  Vec<T, VEC> cf[KJ];
  int xoff[KJ];                                                // element offset of vector jj's slice; -1 = no vector
#pragma unroll
  for (int jj = 0; jj < KJ; ++jj) {
    const int j = g + jj * G;
    xoff[jj] = (in_group && j < nj && !(dbg & 8)) ? j * Rb + elem0 : -1;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      acc[jj][v] = (T)0;
      cf[jj].v[v] = (in_group && j < nj) ? coef[(int64_t)j * b + c_up + v] : (T)0;
    }
  }
  const int ws_off = nj_pad * Rb;

  int slot = 0;
  uint32_t phase = 0;
  for (int64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
    const T* vs = bufs + slot * stage_elems;
    const T* ws = bufs + slot * stage_elems + ws_off;
    const int64_t row0 = ch * R;
    const int rows = (int)min((int64_t)R, a.n - row0);
    fu_mbar_wait(bars_u32 + 8 * slot, phase);
    // 2a. partial sums of  sum_j coef[j] * V[j]  (rows past the end of the vectors arrive zero-filled)
    Vec<T, VEC> x[KJ][RPT];
    Vec<T, VEC> p[RPT];
#pragma unroll
    for (int rr = 0; rr < RPT; ++rr)
#pragma unroll
      for (int v = 0; v < VEC; ++v) p[rr].v[v] = (T)0;
#pragma unroll
    for (int jj = 0; jj < KJ; ++jj) {
#pragma unroll
      for (int rr = 0; rr < RPT; ++rr) {
        if (xoff[jj] >= 0) {
          x[jj][rr] = *reinterpret_cast<const Vec<T, VEC>*>(vs + xoff[jj] + rr * (int)b);
        } else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) x[jj][rr].v[v] = (T)0;
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) p[rr].v[v] += cf[jj].v[v] * x[jj][rr].v[v];
      }
    }
    if (in_group) {
#pragma unroll
      for (int rr = 0; rr < RPT; ++rr)
        *reinterpret_cast<Vec<T, VEC>*>(part + ((int64_t)g * RL + (r0 + rr) * lanes + l_up) * VEC) = p[rr];
    }
    Vec<T, VEC> w_old;                                          // the output thread's element of W, read while the
    if (tid < RL && out_r < rows) w_old = *reinterpret_cast<const Vec<T, VEC>*>(ws + out_e);   // slot is still held
    if (!(dbg & 32)) FU_CONSUMER_SYNC();
    // Everything this chunk needs from the ring slot is in registers now: hand the slot back to the producer before
    // the reduction and the dots, so that two slots of R rows do the work of three (R is what bounds the kernel).
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bars_u32 + 8 * (S + slot)) : "memory");
    // 2b. W_new = W + sum of the G partials, in two levels so that no thread waits on a long chain of
    //     shared-memory loads: Q threads per output each add ~G/Q partials, then one thread adds those Q.
    if (tid < Q * RL && !(dbg & 1)) {
      const int q = l1_q, o = l1_o;
      Vec<T, VEC> s4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < VEC; ++v) s4[u].v[v] = (T)0;
      for (int i = q; i < G; i += 4 * Q) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int gg = i + u * Q;
          if (gg < G) {
            const Vec<T, VEC> pq = *reinterpret_cast<const Vec<T, VEC>*>(part + ((int64_t)gg * RL + o) * VEC);
#pragma unroll
            for (int v = 0; v < VEC; ++v) s4[u].v[v] += pq.v[v];
          }
        }
      }
#pragma unroll
      for (int v = 0; v < VEC; ++v) s4[0].v[v] = (s4[0].v[v] + s4[1].v[v]) + (s4[2].v[v] + s4[3].v[v]);
      *reinterpret_cast<Vec<T, VEC>*>(part2 + (int64_t)tid * VEC) = s4[0];
    }
    if (!(dbg & 32)) FU_CONSUMER_SYNC();
    if (tid < RL) {
      if (out_r < rows) {
        const int e = out_e;
        Vec<T, VEC> s4[4];
        s4[0] = w_old;
#pragma unroll
        for (int u = 1; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < VEC; ++v) s4[u].v[v] = (T)0;
        for (int i = 0; i < Q; i += 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (i + u < Q) {
              const Vec<T, VEC> pq = *reinterpret_cast<const Vec<T, VEC>*>(part2 + ((int64_t)(i + u) * RL + tid) * VEC);
#pragma unroll
              for (int v = 0; v < VEC; ++v) s4[u].v[v] += pq.v[v];
            }
          }
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) s4[0].v[v] = (s4[0].v[v] + s4[1].v[v]) + (s4[2].v[v] + s4[3].v[v]);
        *reinterpret_cast<Vec<T, VEC>*>(wnew + e) = s4[0];
        if (!(dbg & 2)) stg<T, VEC>(a.W + row0 * b + e, s4[0]);
      }
    }
    if (!(dbg & 32)) FU_CONSUMER_SYNC();
    // 3. pass-2 partial dots of this thread's vectors against its slice of W_new
    if (in_group && !(dbg & 4)) {
#pragma unroll
      for (int rr = 0; rr < RPT; ++rr) {
        if (r0 + rr < rows) {
          const Vec<T, VEC> w = *reinterpret_cast<const Vec<T, VEC>*>(wnew + elem0 + rr * (int)b);
#pragma unroll
          for (int jj = 0; jj < KJ; ++jj)
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[jj][v] += x[jj][rr].v[v] * w.v[v];
        }
      }
    }
    if (++slot == S) { slot = 0; phase ^= 1; }
  }
  // flush: combine the row-slices (and every CTA) per (vector, column) in fp64; the ring is free by now
  FU_CONSUMER_SYNC();
  double* red = reinterpret_cast<double*>(bufs);               // [nj][b]
  for (int64_t i = tid; i < nj * b; i += nthr) red[i] = 0.0;
  FU_CONSUMER_SYNC();
  if (in_group) {
#pragma unroll
    for (int jj = 0; jj < KJ; ++jj) {
      const int j = g + jj * G;
      if (j < nj) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) atomicAdd(red + (int64_t)j * b + c_up + v, (double)acc[jj][v]);
      }
    }
  }
  FU_CONSUMER_SYNC();
  for (int64_t i = tid; i < nj * b; i += nthr) {
    const int64_t j = i / b, c = i - j * b;
    const double sres = red[i];
    if (sres != 0.0) atomicAdd(a.C + (a.j0 + j) * b_out + (c & colmask), sres);
  }
}

typedef CUresult (*TmaEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TmaEncodeFn tma_encode_fn() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && p) return (TmaEncodeFn)p;
  return nullptr;
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
  // A single column (one Lanczos start vector, BASELINE config 5) is contiguous: it is swept as an (n / maxv, maxv)
  // block with 16-byte loads, the maxv partial columns folding into the one coefficient (4.3 -> see DESIGN.md TB/s).
  int fold = 0;
  if (b == 1 && n % maxv == 0 && vstride % maxv == 0 && ((uintptr_t)V % 16 == 0) && ((uintptr_t)W % 16 == 0) &&
      getenv("COLA_REORTH_NO_FOLD") == nullptr) {
    fold = maxv;
    n /= maxv;
    b = maxv;
  }
  static const bool dots1 = [] { const char* e = getenv("COLA_REORTH_DOTS1"); return !(e && atoi(e) == 0); }();
  if (fold && dots1 && (j1 - j0) * (kD1Threads / 32) * 8 <= 96 * 1024) {
    RoArgs<T> a{};
    a.V = V; a.vstride = vstride; a.j0 = j0; a.j1 = j1; a.Wc = W; a.n = n; a.b = b; a.C = C; a.gate = gate; a.fold = fold;
    const size_t smem = (size_t)((j1 - j0) * (kD1Threads / 32) * 8);
    const int64_t chunk = (int64_t)kD1Threads * kD1Rows;
    int64_t grid = (int64_t)sm_count() * 2;
    if (grid > (n + chunk - 1) / chunk) grid = (n + chunk - 1) / chunk;
    auto kern = reorth_dots1_kernel<T>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<(unsigned)grid, kD1Threads, smem, st>>>(a);
    return cuda_status("reorth_dots");
  }
  // vector path: b*sizeof(T) a multiple of 16 B and rows 16 B aligned
  int vec = 1;
  {
    int v = pick_vec<T>(b, b, V, W);
    if (vstride % v) v = 1;
    int64_t per_row = b / v;
    (void)per_row;
    if (v == maxv) vec = v;   // any width that is a whole number of 16-byte vectors: idle lanes are guarded by col < b
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
  if (const char* e = getenv("COLA_REORTH_WROWS")) { const int64_t v = atoll(e); if (v >= 32 && v <= 16384) w_rows = v; }   // A/B knob
  const int64_t w_bytes = w_rows * b * (int64_t)sizeof(T);
  int64_t max_nj = (budget - w_bytes) / (b * 8);
  COLA_REQUIRE(max_nj >= 1, "reorth_dots: probe block too wide for shared memory");
  int rc = COLA_OK;
  for (int64_t ja = j0; ja < j1 && rc == COLA_OK; ja += max_nj) {
    int64_t jb = ja + max_nj < j1 ? ja + max_nj : j1;
    RoArgs<T> a{};
    a.V = V; a.vstride = vstride; a.j0 = ja; a.j1 = jb; a.Wc = W; a.n = n; a.b = b; a.C = C; a.gate = gate;
    a.lanes_per_row = Lr; a.rows_per_chunk = (int)w_rows; a.fold = fold;
    {
      const int64_t njb = jb - ja;
      const int steps = (int)(w_rows / (32 / Lr));              // row steps of a full chunk per warp
      int q = (njb % kDotsWarps == 0) ? 1 : 4;                  // whole rounds already: keep the exclusive accumulation
      while (q > 1 && steps / q < 8) q >>= 1;                   // a part is at least one unrolled burst of 8 steps
      if (getenv("COLA_REORTH_QSPLIT")) q = atoi(getenv("COLA_REORTH_QSPLIT"));
      a.qsplit = q < 1 ? 1 : q;
    }
    size_t smem = (size_t)((jb - ja) * b * 8 + w_bytes);
    int64_t n_chunks = (n + w_rows - 1) / w_rows;
    int per_sm = (int)(budget / (int64_t)smem);
    if (per_sm < 1) per_sm = 1;
    // measured (B200, profiles/r2_reorth_dots_ctas.log): more resident CTAs is not better -- 51 vectors x 64 fp32 columns run
    // at 6.2 TB/s with 2 CTAs per SM, 5.0 with 3, 4.8 with 4; the folded single fp64 column at 4.1 / 5.5 / 4.0 TB/s
    const int per_sm_cap = fold ? 3 : 2;
    if (per_sm > per_sm_cap) per_sm = per_sm_cap;
    if (const char* e = getenv("COLA_REORTH_DOTS_CTAS")) { const int v = atoi(e); if (v >= 1 && v < per_sm) per_sm = v; }   // A/B knob
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

template <typename T>
int reorth_update_dots(const T* V, int64_t vstride, int64_t j0, int64_t j1, T* W, int64_t n, int64_t b,
                       const double* C1, T sign, double* C2, const int32_t* gate, cudaStream_t st) {
  COLA_REQUIRE(V && W && C1 && C2, "reorth_update_dots: null pointer");
  COLA_REQUIRE(j1 > j0 && j0 >= 0, "reorth_update_dots: bad vector range");
  const int64_t nj = j1 - j0;
  constexpr int VEC = 16 / (int)sizeof(T);
  // folded view for a single vector
  int64_t nn = n, bb = b, colmask = -1, b_out = b;
  if (b == 1 && n % VEC == 0) { nn = n / VEC; bb = VEC; colmask = 0; }
  if (bb % VEC != 0 || ((uintptr_t)V % 16) || ((uintptr_t)W % 16) || (vstride % VEC))
    return fail(COLA_E_UNSUPPORTED, "reorth_update_dots: shape not supported by the fused kernel");
  const int64_t vpr = bb / VEC;
  if (vpr > kFuMaxThreads) return fail(COLA_E_UNSUPPORTED, "reorth_update_dots: row wider than one CTA");
  int dev = 0, smem_max = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (smem_max <= 0) smem_max = 48 * 1024;
  const int lanes = (int)vpr;
  const int nthr = kFuMaxThreads;
  const int64_t nj_pad = (nj + kFuBoxH - 1) / kFuBoxH * kFuBoxH;
  const int64_t row_bytes = (nj_pad + 1) * bb * (int64_t)sizeof(T);      // one row of every vector + W
  const int64_t coef_bytes = nj * bb * (int64_t)sizeof(T);
  auto stage_bytes = [&](int64_t r) { return (r * row_bytes + 127) / 128 * 128; };
  // Configuration: RPT rows per thread (1 or 2), KJ vectors per thread (2, 4 or 8), R rows per chunk.  Every
  // vector needs an owner (ceil(nj / G) <= KJ), the ring holds >= 3 chunks, and a chunk of one vector is either
  // one TMA box row (R*b <= 256 elements) or a whole number of 256-element rows of the vector (3-D view).
  // The per-chunk instruction cost is nearly fixed, so the tallest chunk wins; ties go to the smaller KJ.
  int best_rpt = 0, best_kj = 0;
  int64_t best_R = 0;
  int min_stages = 2;   // a slot is handed back as soon as its chunk is in registers, so two slots already overlap
  if (const char* e = getenv("COLA_FU_MIN_STAGES")) { const int v = atoi(e); if (v >= 1 && v <= 8) min_stages = v; }
  const bool can_3d = (nn * bb) % 256 == 0 && 256 % bb == 0 && (vstride % 256) == 0;
  for (int rpt = 2; rpt >= 1; --rpt) {
    for (int kj = 2; kj <= 8; kj *= 2) {
      if (rpt == 2 && kj == 8) continue;                       // register budget
      const int64_t part_bytes = (int64_t)kFuMaxThreads * (2 * rpt + 1) * 16;
      const int64_t budget = (int64_t)smem_max - 1024 - coef_bytes - part_bytes;
      int64_t R = (int64_t)nthr / lanes;                       // one output thread per Vec of the chunk (level 2)
      if (R > 256) R = 256;
      R -= R % rpt;
      for (; R >= rpt; R -= rpt) {
        const bool fits_2d = R * bb <= 256;
        const bool fits_3d = can_3d && (R * bb) % 256 == 0 && R * bb / 256 <= 256;
        if (!fits_2d && !fits_3d) continue;
        const int64_t GS = R * lanes / rpt, G = nthr / GS;
        if (G >= 1 && (nj + G - 1) / G <= kj && budget / stage_bytes(R) >= min_stages) break;
      }
      if (R >= rpt && (R > best_R || (R == best_R && kj < best_kj))) { best_R = R; best_rpt = rpt; best_kj = kj; }
    }
  }
  if (best_R == 0) return fail(COLA_E_UNSUPPORTED, "reorth_update_dots: basis chunk does not fit shared memory / registers");
  const int64_t R = best_R;
  const int64_t part_bytes = (int64_t)kFuMaxThreads * (2 * best_rpt + 1) * 16;
  int64_t S = ((int64_t)smem_max - 1024 - coef_bytes - part_bytes) / stage_bytes(R);
  if (S > 8) S = 8;
  if (const char* e = getenv("COLA_FU_S")) { const int64_t ss = atoi(e); if (ss >= 1 && ss < S) S = ss; }
  static TmaEncodeFn enc = tma_encode_fn();
  if (!enc) return fail(COLA_E_UNSUPPORTED, "reorth_update_dots: cuTensorMapEncodeTiled unavailable");
  CUtensorMap vmap;
  const int inner = R * bb <= 256 ? 0 : 256;
  {
    const CUtensorMapDataType dt = sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
    CUresult r;
    if (inner == 0) {
      if (nn * bb >= ((int64_t)1 << 31)) return fail(COLA_E_UNSUPPORTED, "reorth_update_dots: vector too long for one TMA coordinate");
      cuuint64_t dims[2] = {(cuuint64_t)(nn * bb), (cuuint64_t)nj};
      cuuint64_t strides[1] = {(cuuint64_t)(vstride * (int64_t)sizeof(T))};
      cuuint32_t box[2] = {(cuuint32_t)(R * bb), (cuuint32_t)kFuBoxH};
      cuuint32_t es[2] = {1, 1};
      r = enc(&vmap, dt, 2, (void*)(V + j0 * vstride), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)(nn * bb / inner), (cuuint64_t)nj};
      cuuint64_t strides[2] = {(cuuint64_t)(inner * (int64_t)sizeof(T)), (cuuint64_t)(vstride * (int64_t)sizeof(T))};
      cuuint32_t box[3] = {(cuuint32_t)inner, (cuuint32_t)(R * bb / inner), (cuuint32_t)kFuBoxH};
      cuuint32_t es[3] = {1, 1, 1};
      r = enc(&vmap, dt, 3, (void*)(V + j0 * vstride), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) return fail(COLA_E_UNSUPPORTED, "reorth_update_dots: cuTensorMapEncodeTiled failed");
  }
  RoArgs<T> a{};
  a.V = V; a.vstride = vstride; a.j0 = j0; a.j1 = j1; a.W = W; a.n = nn; a.b = bb; a.Cc = C1; a.C = C2; a.sign = sign;
  a.gate = gate;
  const int64_t stage_elems = stage_bytes(R) / (int64_t)sizeof(T);
  const size_t smem = (size_t)(S * stage_bytes(R) + coef_bytes + part_bytes + 16 * 8);
  const int64_t n_chunks = (nn + R - 1) / R;
  int64_t grid = (int64_t)sm_count();
  if (grid > n_chunks) grid = n_chunks;
  int dbg = 0;
  if (const char* e = getenv("COLA_FU_DBG")) dbg = atoi(e);
#define COLA_FU_LAUNCH(RPT_, KJ_)                                                                              \
  do {                                                                                                         \
    auto kern = reorth_fused_kernel<T, VEC, RPT_, KJ_>;                                                        \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                        \
    kern<<<(unsigned)grid, nthr + 32, smem, st>>>(vmap, a, (int)R, (int)S, (int)nj_pad, stage_elems, lanes,    \
                                                  colmask, b_out, inner, dbg);                                          \
  } while (0)
  if (best_rpt == 2 && best_kj == 2) COLA_FU_LAUNCH(2, 2);
  else if (best_rpt == 2) COLA_FU_LAUNCH(2, 4);
  else if (best_kj == 2) COLA_FU_LAUNCH(1, 2);
  else if (best_kj == 4) COLA_FU_LAUNCH(1, 4);
  else COLA_FU_LAUNCH(1, 8);
#undef COLA_FU_LAUNCH
  return cuda_status("reorth_update_dots");
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
int cola_reorth_update_dots_f32(const float* V, int64_t vstride, int64_t j0, int64_t j1, float* W, int64_t n, int64_t b,
                                const double* C1, float sign, double* C2, const int32_t* gate, void* stream) {
  return reorth_update_dots<float>(V, vstride, j0, j1, W, n, b, C1, sign, C2, gate, reinterpret_cast<cudaStream_t>(stream));
}
int cola_reorth_update_dots_f64(const double* V, int64_t vstride, int64_t j0, int64_t j1, double* W, int64_t n, int64_t b,
                                const double* C1, double sign, double* C2, const int32_t* gate, void* stream) {
  return reorth_update_dots<double>(V, vstride, j0, j1, W, n, b, C1, sign, C2, gate, reinterpret_cast<cudaStream_t>(stream));
}
COLA_RO_API(f32, float)
COLA_RO_API(f64, double)
}
