// Kronecker matmat on the 5th-generation tensor cores (north_star item 3): a chain of batched mode-k
// contractions  out[p, a, l, r] = sum_j F[a, j] * in[p, j, l, r]  with 64x64 fp32 factors, executed as
// tcgen05.mma kind::tf32 with TMEM accumulators and TMA-staged tiles.  Replaces the moveaxis/reshape/GEMM/
// moveaxis chain of Kronecker._matmat (cola/ops/operators.py:216-223).
//
// Formulation (per tile): D'[m, a] = sum_j A'[m, j] * B'[j, a]
//   A' = the input tile: 128 "positions" m x K = 64 contraction indices j.  A position is (atom, r): an atom is
//        a (p, l) pair, r one of 32 consecutive right-hand sides; in memory the 32 r of one (p, j, l) are
//        contiguous (128 B), so one TMA box {32 r, 1 l, 64 j} (SWIZZLE_128B_ATOM_32B) IS one MN-major UMMA atom and a tile
//        is just 4 such boxes (no transposes, no reshapes, whatever mode is being contracted).
//   B' = F^T: K-major (F is row-major [a][j]), loaded once per CTA by TMA with the same 128B swizzle.
//   D' = 128 TMEM lanes (positions) x 64 columns (a), fp32.
// fp32-grade accuracy from tf32 tensor cores: 3xTF32 error compensation.  Every operand x is split on the CUDA
// cores into hi = tf32(x) and lo = tf32(x - hi) (elementwise, in place in the swizzled tile, so the layout
// never matters) and D' = A_lo B_hi + A_hi B_lo + A_hi B_hi  (24 MMAs of 128x64x8 per tile).
// Roofline: the contraction itself is HBM-bound at d = 64 (48 flop/B < tf32 ridge): 64 KB of traffic per tile vs
// 768 tensor-pipe cycles.  The host driver therefore runs all modes on one 32-column chunk of right-hand sides at
// a time so the two intermediates (n*32 floats each) stay in the 126 MB L2, and only X in / Y out cross HBM.
//
// Warp roles (384 threads, 1 CTA/SM, persistent over tiles): warp 0 TMA producer, warp 1 MMA issuer,
// warp 2 TMEM allocator, warps 4-7 operand split, warps 8-11 epilogue (TMEM -> registers -> global, fused
// alpha / shift / diag / pAp-dots on the last mode).
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace cola {

constexpr int kTcThreads = 384;
constexpr int kD = 64;                 // factor size handled by this path
constexpr int kAtomBytes = 64 * 128;   // one TMA box: 64 rows (j) x 32 floats
constexpr int kTileBytes = 4 * kAtomBytes;   // 32 KB: 128 positions x 64 j
constexpr int kStages = 2;
constexpr int kFacBytes = kD * kD * 4;       // 16 KB
// shared memory map (all 1024-byte aligned for the 128B swizzle)
constexpr int kOffFacHi = 0;
constexpr int kOffFacLo = kOffFacHi + kFacBytes;
constexpr int kOffStage = kOffFacLo + kFacBytes;                 // per stage: hi tile | lo tile
constexpr int kOffBars = kOffStage + kStages * 2 * kTileBytes;
constexpr int kSmemBytes = kOffBars + 256 + 1024;                // + alignment slack

struct TcArgs {
  float* out; const float* epi_x; const float* diag;
  int64_t L, pre;           // in viewed as (pre, 64, L, row) with row = in_k floats
  int64_t in_r0;            // first column of the 32-wide RHS chunk inside the source rows (TMA coordinate 0)
  int64_t out_k, out_r0;    // row length of the destination and column offset of the chunk inside it
  int64_t n_tiles;          // pre * L / 4
  float alpha, shift; int accumulate; int last_mode;
  double* dots; const int32_t* dots_row; const int32_t* gate;
  int dbg;   // bring-up knob (COLA_KRON_DBG): 1 = skip split math, 2 = skip MMAs, 4 = skip epilogue stores
};

// ---- PTX wrappers ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
template <int C>
__device__ __forceinline__ void tmem_ld16_at(uint32_t taddr, uint32_t (&r)[32]) {   // columns C*16 .. C*16+15 of a 32-wide half row
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[C * 16 + 0]), "=r"(r[C * 16 + 1]), "=r"(r[C * 16 + 2]), "=r"(r[C * 16 + 3]), "=r"(r[C * 16 + 4]),
        "=r"(r[C * 16 + 5]), "=r"(r[C * 16 + 6]), "=r"(r[C * 16 + 7]), "=r"(r[C * 16 + 8]), "=r"(r[C * 16 + 9]),
        "=r"(r[C * 16 + 10]), "=r"(r[C * 16 + 11]), "=r"(r[C * 16 + 12]), "=r"(r[C * 16 + 13]), "=r"(r[C * 16 + 14]),
        "=r"(r[C * 16 + 15])
      : "r"(taddr + C * 16) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor), version 1.
// layout: 2 = SWIZZLE_128B (16-byte chunks ^ row%8), 1 = SWIZZLE_128B_BASE32B (32-byte chunks ^ row%4), the only
// layout the tensor core accepts for MN-major 32-bit (tf32) operands (measured: every other MN-major tf32
// layout silently yields zeros; scripts/umma_probe.cu).
constexpr uint32_t kLayoutSw128 = 2, kLayoutSw128Base32 = 1;
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)layout << 61;
  return d;
}
// instruction descriptor (InstrDescriptor): D=f32, A=B=tf32, A MN-major, B K-major, M=128, N=64
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (0u << 16) | ((64u >> 3) << 17) |
                            ((128u >> 4) << 24);

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
// elementwise 3xTF32 split of `bytes` of swizzled operand data: hi in place, lo into the twin buffer
__device__ __forceinline__ void split_hi_lo(unsigned char* hi, unsigned char* lo, int bytes, int tid, int nthreads) {
  for (int o = tid * 16; o < bytes; o += nthreads * 16) {
    float4 v = *reinterpret_cast<float4*>(hi + o);
    float4 h, l;
    h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
    l.x = tf32_rna(v.x - h.x); l.y = tf32_rna(v.y - h.y); l.z = tf32_rna(v.z - h.z); l.w = tf32_rna(v.w - h.w);
    *reinterpret_cast<float4*>(hi + o) = h;
    *reinterpret_cast<float4*>(lo + o) = l;
  }
}

__global__ void __launch_bounds__(kTcThreads, 1)
    kron_mode_tc_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_fac,
                        TcArgs a) {
  if (a.gate != nullptr && *a.gate != 0) return;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  // barriers
  const uint32_t bar_full0 = sbase + kOffBars;            // [stage] TMA landed
  const uint32_t bar_ready0 = bar_full0 + 8 * kStages;     // [stage] split done (count 4: one per split warp)
  const uint32_t bar_empty0 = bar_ready0 + 8 * kStages;    // [stage] MMAs that read the stage retired
  const uint32_t bar_tfull0 = bar_empty0 + 8 * kStages;    // [acc]   accumulator complete
  const uint32_t bar_tempty0 = bar_tfull0 + 8 * 2;         // [acc]   accumulator drained (count 4)
  const uint32_t bar_fac = bar_tempty0 + 8 * 2;            // factor landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBars + 200);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full0 + 8 * s, 1);
      mbar_init(bar_ready0 + 8 * s, 4);
      mbar_init(bar_empty0 + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_tfull0 + 8 * s, 1);
      mbar_init(bar_tempty0 + 8 * s, 4);
    }
    mbar_init(bar_fac, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // factor: TMA (2 boxes of 32 k x 64 rows), then everybody splits it into hi/lo
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar_fac, kFacBytes);
    tma_load_2d(sbase + kOffFacHi, &map_fac, bar_fac, 0, 0);
    tma_load_2d(sbase + kOffFacHi + kFacBytes / 2, &map_fac, bar_fac, 32, 0);
  }
  mbar_wait(bar_fac, 0);
  split_hi_lo(smem + kOffFacHi, smem + kOffFacLo, kFacBytes, threadIdx.x, kTcThreads);
  fence_async_smem();
  __syncthreads();

  const int64_t first = blockIdx.x, step = gridDim.x;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int it = 0;
      for (int64_t t = first; t < a.n_tiles; t += step, ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        mbar_wait(bar_empty0 + 8 * s, ph ^ 1);
        mbar_expect_tx(bar_full0 + 8 * s, kTileBytes);
        const uint32_t dst = sbase + kOffStage + s * 2 * kTileBytes;
        for (int at = 0; at < 4; ++at) {
          const int64_t flat = t * 4 + at;
          const int64_t p = flat / a.L, l = flat - p * a.L;
          tma_load_3d(dst + at * kAtomBytes, &map_in, bar_full0 + 8 * s, (int)a.in_r0, (int)l, (int)(p * kD));
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      int it = 0;
      for (int64_t t = first; t < a.n_tiles; t += step, ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        const int acc = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(bar_tempty0 + 8 * acc, aph ^ 1);
        mbar_wait(bar_ready0 + 8 * s, ph);
        tc_fence_after();
        const uint32_t a_hi = sbase + kOffStage + s * 2 * kTileBytes, a_lo = a_hi + kTileBytes;
        const uint32_t b_hi = sbase + kOffFacHi, b_lo = sbase + kOffFacLo;
        const uint32_t d = tmem_base + acc * kD;
        uint32_t accum = 0;
        if (!(a.dbg & 2))
        // small terms first: A_lo*B_hi, A_hi*B_lo, then A_hi*B_hi
#pragma unroll
        for (int term = 0; term < 3; ++term) {
          const uint32_t A0 = (term == 0) ? a_lo : a_hi;
          const uint32_t B0 = (term == 1) ? b_lo : b_hi;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            // A': MN-major SW128_BASE32B: rows (j) of 128 B, K groups of 4 rows 512 B apart (SBO), 8 rows per
            // MMA step (1024 B); the four 32-position atoms along MN are 8 KB apart (LBO)
            const uint64_t ad = make_desc(A0 + kk * 1024, kAtomBytes, 512, kLayoutSw128Base32);
            // B': K-major, two 32-float k-chunks of 8 KB; 32 B per K step inside a chunk; 8-row groups 1 KB apart
            const uint64_t bd = make_desc(B0 + (kk / 4) * (kFacBytes / 2) + (kk % 4) * 32, 16, 1024, kLayoutSw128);
            umma_tf32(d, ad, bd, kIdesc, accum);
            accum = 1;
          }
        }
        umma_commit(bar_empty0 + 8 * s);     // stage may be refilled once these MMAs retire
        umma_commit(bar_tfull0 + 8 * acc);   // accumulator ready for the epilogue
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===== operand split: raw fp32 tile -> hi (in place) + lo =====
    const int tid = threadIdx.x - 128;
    int it = 0;
    for (int64_t t = first; t < a.n_tiles; t += step, ++it) {
      const int s = it % kStages;
      const uint32_t ph = (it / kStages) & 1;
      mbar_wait(bar_full0 + 8 * s, ph);
      unsigned char* hi = smem + kOffStage + s * 2 * kTileBytes;
      if (!(a.dbg & 1)) split_hi_lo(hi, hi + kTileBytes, kTileBytes, tid, 128);
      fence_async_smem();     // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_ready0 + 8 * s);
    }
  } else if (warp >= 8) {
    // ===== epilogue: TMEM -> registers -> global =====
    const int q = warp - 8;                 // TMEM lane quadrant == atom index within the tile
    double dacc = 0.0;
    int it = 0;
    for (int64_t t = first; t < a.n_tiles; t += step, ++it) {
      const int acc = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      const int64_t flat = t * 4 + q;
      const int64_t p = flat / a.L, l = flat - p * a.L;
      // element (a, r) of this atom lives at base + a * row_stride + r
      const int64_t row_stride = a.L * a.out_k;
      const int64_t base = (p * kD * a.L + l) * a.out_k + a.out_r0 + lane;
      mbar_wait(bar_tfull0 + 8 * acc, aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * kD;
      const float* __restrict__ xin = a.epi_x;
      const float* __restrict__ dg = a.diag;
      float* __restrict__ outp = a.out;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[16];
        tmem_ld16(taddr + c * 16, v);
        // operands of the fused epilogue are requested while the TMEM load is in flight (independent loads,
        // all issued before the first use)
        float xv[16], dv[16], ov[16];
        if (a.last_mode) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int64_t o = base + (int64_t)(c * 16 + i) * row_stride;
            xv[i] = xin ? xin[o] : 0.f;
            dv[i] = dg ? dg[(p * kD + c * 16 + i) * a.L + l] : 0.f;
            ov[i] = a.accumulate ? outp[o] : 0.f;
          }
        }
        tmem_ld_wait();
        if (c == 3) {   // accumulator fully read: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty0 + 8 * acc);
        }
        if (a.dbg & 4) continue;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int64_t o = base + (int64_t)(c * 16 + i) * row_stride;
          float y = a.alpha * __uint_as_float(v[i]);
          if (a.last_mode) {
            if (xin != nullptr) {
              if (a.shift != 0.f) y += a.shift * xv[i];
              if (dg != nullptr) y += dv[i] * xv[i];
            }
            if (a.accumulate) y += ov[i];
            // <x, y> of the value that is stored, earlier terms of a Sum included (same contract as the SIMT and CSR
            // epilogues: the dots are those of the whole operator)
            if (xin != nullptr && a.dots != nullptr) dacc += (double)xv[i] * (double)y;
          }
          outp[o] = y;
        }
      }
    }
    if (a.dots != nullptr) {
      double* outp = a.dots + (a.dots_row ? (int64_t)(*a.dots_row) * a.out_k : 0);
      atomicAdd(outp + a.out_r0 + lane, dacc);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(128));
  }
}

// =======================================================================================================
// Fused persistent variant (2 <= D <= 3 factors): ONE launch per matmat.  One CTA per SM walks the phases
// (chunk 0: mode 0, 1, .., D-1; chunk 1: ...) and meets the other CTAs at a device-wide barrier only where a
// phase consumes what the previous one produced (mode i -> mode i+1 of the same chunk).  All D factors stay
// resident in shared memory as hi/lo pairs; the TMA producer runs R tiles ahead in a ring of raw tiles, the split
// warps turn one raw tile into the (single) hi/lo operand pair while the previous tile's accumulator drains
// through 8 epilogue warps.  Compared with one launch per (chunk, mode) this removes 4*D-1 prologues (TMEM
// allocation, barrier init, factor load + split) and keeps the TMA queue full across chunk boundaries.
// =======================================================================================================
constexpr int kFusedThreads = 512;          // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 spare, 4-11 split, 12-15 epilogue
constexpr int kMaxFused = 3;
constexpr int kRing = 2;                    // raw-tile ring depth (TMA runs two tiles ahead of the split warps)
constexpr int kSplitWarps = 8;
struct FusedMaps {
  CUtensorMap in[kMaxFused];    // source of mode i: X (3D over row length k) for i = 0, the dense chunk workspaces after
  CUtensorMap fac[kMaxFused];
};
struct FusedArgs {
  int D;
  int cpc;                      // 32-column blocks per chunk: the intermediates hold 32*cpc columns per row
  int64_t k, n_chunks;
  float* ws0; float* ws1; float* Y; const float* X; const float* diag;
  float alpha, shift; int accumulate;
  double* dots; const int32_t* dots_row; const int32_t* gate;
  unsigned int* sync_counter;   // zeroed by the host before the launch
  int dbg;                      // bring-up knob (COLA_KRON_DBG): 1 skip split math, 2 skip MMAs, 4 skip stores, 8 skip grid barrier
};
// shared-memory map of the fused kernel (1024-byte aligned pieces)
constexpr int kFOffFacHi = 0;                          // current mode's factor, hi | lo
constexpr int kFOffFacLo = kFacBytes;
constexpr int kFOffOp = 2 * kFacBytes;                 // two operand pairs [hi | lo]: tile t+1 is split while tile t is multiplied
constexpr int kFOffRaw = kFOffOp + 4 * kTileBytes;     // ring of raw tiles (a factor travels through it as a half-filled slot)
constexpr int kFOffBars = kFOffRaw + kRing * kTileBytes;
constexpr int kFusedSmem = kFOffBars + 256 + 1024;

__device__ __forceinline__ void grid_arrive_and_wait(unsigned int* counter, unsigned int target) {
  __threadfence();
  atomicAdd(counter, 1u);
  unsigned int v;
  do {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
  } while (v < target);
}

__device__ __forceinline__ void split_to(const unsigned char* raw, unsigned char* hi, unsigned char* lo, int bytes, int tid,
                                         int nthreads) {
  for (int o = tid * 16; o < bytes; o += nthreads * 16) {
    const float4 v = *reinterpret_cast<const float4*>(raw + o);
    float4 h, l;
    h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
    l.x = tf32_rna(v.x - h.x); l.y = tf32_rna(v.y - h.y); l.z = tf32_rna(v.z - h.z); l.w = tf32_rna(v.w - h.w);
    *reinterpret_cast<float4*>(hi + o) = h;
    *reinterpret_cast<float4*>(lo + o) = l;
  }
}

__global__ void __launch_bounds__(kFusedThreads, 1)
    kron_fused_tc_kernel(const __grid_constant__ FusedMaps maps, FusedArgs a) {
  if (a.gate != nullptr && *a.gate != 0) return;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const int D = a.D;
  const uint32_t bar_full0 = sbase + kFOffBars;            // [kRing] ring slot landed (tile or factor)
  const uint32_t bar_rfree0 = bar_full0 + 8 * kRing;       // [kRing] ring slot consumed by the split warps (count 8)
  const uint32_t bar_ready0 = bar_rfree0 + 8 * kRing;      // [2] operand pair written (count 8)
  const uint32_t bar_opfree0 = bar_ready0 + 16;            // [2] MMAs reading the operand pair retired
  const uint32_t bar_tfull0 = bar_opfree0 + 16;            // [2] accumulator complete
  const uint32_t bar_tempty0 = bar_tfull0 + 16;            // [2] accumulator drained (count 4)
  const uint32_t bar_phase = bar_tempty0 + 16;             // epilogue warps finished a phase (count 4)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kFOffBars + 200);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kRing; ++s) { mbar_init(bar_full0 + 8 * s, 1); mbar_init(bar_rfree0 + 8 * s, kSplitWarps); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_ready0 + 8 * s, kSplitWarps);
      mbar_init(bar_opfree0 + 8 * s, 1);
      mbar_init(bar_tfull0 + 8 * s, 1);
      mbar_init(bar_tempty0 + 8 * s, 4);
    }
    mbar_init(bar_phase, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  int64_t n = 1;
  for (int i = 0; i < D; ++i) n *= kD;
  const int64_t n_pos_tiles = n / kD / 4;      // pre * L / 4 position tiles, the same for every mode (a power of two)
  const int pos_sh = 6 * (D - 1) - 2;
  const int64_t pos_mask = n_pos_tiles - 1;
  const int64_t n_tiles = n_pos_tiles * a.cpc; // x column blocks of the chunk
  const int n_phases = (int)a.n_chunks * D;
  const int64_t first = blockIdx.x, step = gridDim.x;

  if (warp == 0) {
    // ===== TMA producer: per phase one factor slot, then the tiles =====
    if (lane == 0) {
      int rit = 0;   // ring items issued
      unsigned int barriers_passed = 0;
      for (int ph = 0; ph < n_phases; ++ph) {
        const int chunk = ph / D, mode = ph - chunk * D;
        const int lsh = 6 * (D - 1 - mode);           // L = 64^(D-1-mode): divisions become shifts
        const int64_t lmask = ((int64_t)1 << lsh) - 1;
        {  // the factor does not depend on other CTAs: request it before the phase barrier
          const int s = rit % kRing;
          mbar_wait(bar_rfree0 + 8 * s, ((rit / kRing) & 1) ^ 1);
          mbar_expect_tx(bar_full0 + 8 * s, kFacBytes);
          const uint32_t dst = sbase + kFOffRaw + s * kTileBytes;
          tma_load_2d(dst, &maps.fac[mode], bar_full0 + 8 * s, 0, 0);
          tma_load_2d(dst + kFacBytes / 2, &maps.fac[mode], bar_full0 + 8 * s, 32, 0);
          ++rit;
        }
        if (ph > 0) mbar_wait(bar_phase, (ph - 1) & 1);   // every phase completion is consumed in order (parity tracking)
        if (mode > 0) {
          // this phase reads what every CTA wrote in the previous one: local epilogue done, then device-wide
          ++barriers_passed;
          if (!(a.dbg & 8)) grid_arrive_and_wait(a.sync_counter, barriers_passed * gridDim.x);
          asm volatile("fence.proxy.async;" ::: "memory");
        }
        const int in_base = (mode == 0) ? chunk * 32 * a.cpc : 0;
        for (int64_t t = first; t < n_tiles; t += step, ++rit) {
          const int64_t cc = t >> pos_sh, tp = t & pos_mask;   // column block inside the chunk, position tile
          const int in_r0 = in_base + (int)cc * 32;
          const int s = rit % kRing;
          mbar_wait(bar_rfree0 + 8 * s, ((rit / kRing) & 1) ^ 1);
          const int nat = (a.dbg & 16) ? 1 : 4;
          mbar_expect_tx(bar_full0 + 8 * s, nat * kAtomBytes);
          const uint32_t dst = sbase + kFOffRaw + s * kTileBytes;
          for (int at = 0; at < nat; ++at) {
            const int64_t flat = tp * 4 + at;
            const int64_t p = flat >> lsh, l = flat & lmask;
            tma_load_3d(dst + at * kAtomBytes, &maps.in[mode], bar_full0 + 8 * s, in_r0, (int)l, (int)(p * kD));
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      int it = 0;
      // A': MN-major SW128_BASE32B: rows (j) of 128 B, K groups of 4 rows 512 B apart (SBO), 8 rows (1024 B) per
      // MMA step, the four 32-position atoms 8 KB apart (LBO).  B': K-major SW128, two 32-float k-chunks of 8 KB,
      // 32 B per K step inside a chunk, 8-row groups 1 KB apart.
      const uint64_t ad_hi0 = make_desc(sbase + kFOffOp, kAtomBytes, 512, kLayoutSw128Base32);
      const uint64_t ad_lo0 = make_desc(sbase + kFOffOp + kTileBytes, kAtomBytes, 512, kLayoutSw128Base32);
      const uint64_t bd_hi = make_desc(sbase + kFOffFacHi, 16, 1024, kLayoutSw128);
      const uint64_t bd_lo = make_desc(sbase + kFOffFacLo, 16, 1024, kLayoutSw128);
      for (int ph = 0; ph < n_phases; ++ph) {
        for (int64_t t = first; t < n_tiles; t += step, ++it) {
          const int acc = it & 1;
          const uint32_t aph = (it >> 1) & 1;
          mbar_wait(bar_tempty0 + 8 * acc, aph ^ 1);
          mbar_wait(bar_ready0 + 8 * acc, aph);
          tc_fence_after();
          const uint32_t d = tmem_base + acc * kD;
          const uint64_t opoff = (uint64_t)((acc * 2 * kTileBytes) >> 4);   // operand pair of this tile
          const uint64_t ad_hi = ad_hi0 + opoff, ad_lo = ad_lo0 + opoff;
          uint32_t accum = 0;
          if (!(a.dbg & 2))
#pragma unroll
          for (int term = 0; term < 3; ++term) {
            // descriptors differ only in the 14-bit start-address field: one 64-bit add per MMA on the issuing thread
            const uint64_t ad0 = (term == 0) ? ad_lo : ad_hi;
            const uint64_t bd0 = (term == 1) ? bd_lo : bd_hi;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              const uint64_t ad = ad0 + (uint64_t)((kk * 1024) >> 4);
              const uint64_t bd = bd0 + (uint64_t)(((kk / 4) * (kFacBytes / 2) + (kk % 4) * 32) >> 4);
              umma_tf32(d, ad, bd, kIdesc, accum);
              accum = 1;
            }
          }
          umma_commit(bar_opfree0 + 8 * acc);
          umma_commit(bar_tfull0 + 8 * acc);
        }
      }
    }
  } else if (warp >= 4 && warp < 4 + kSplitWarps) {
    // ===== operand split: ring slot -> hi / lo operand pair (or -> the factor pair at a phase start) =====
    const int tid = threadIdx.x - 128;
    constexpr int kSplitThreads = kSplitWarps * 32;
    int it = 0, rit = 0;
    for (int ph = 0; ph < n_phases; ++ph) {
      {  // factor of this phase
        const int s = rit % kRing;
        mbar_wait(bar_full0 + 8 * s, (rit / kRing) & 1);
        // every MMA of the previous phase must have retired before the factor changes: the last two tiles
        if (it >= 1) mbar_wait(bar_opfree0 + 8 * ((it - 1) & 1), ((it - 1) >> 1) & 1);
        if (it >= 2) mbar_wait(bar_opfree0 + 8 * ((it - 2) & 1), ((it - 2) >> 1) & 1);
        split_to(smem + kFOffRaw + s * kTileBytes, smem + kFOffFacHi, smem + kFOffFacLo, kFacBytes, tid, kSplitThreads);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_rfree0 + 8 * s);
        ++rit;
      }
      for (int64_t t = first; t < n_tiles; t += step, ++it, ++rit) {
        const int s = rit % kRing;
        const int ob = it & 1;
        mbar_wait(bar_full0 + 8 * s, (rit / kRing) & 1);
        mbar_wait(bar_opfree0 + 8 * ob, ((it >> 1) & 1) ^ 1);   // MMAs of tile it-2 no longer read this operand pair
        unsigned char* op = smem + kFOffOp + ob * 2 * kTileBytes;
        if (!(a.dbg & 1)) split_to(smem + kFOffRaw + s * kTileBytes, op, op + kTileBytes, kTileBytes, tid, kSplitThreads);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar_rfree0 + 8 * s);        // ring slot may be refilled
          mbar_arrive(bar_ready0 + 8 * ob);       // operand pair (and, at a phase start, the factor pair) complete
        }
      }
    }
  } else if (warp >= 4 + kSplitWarps) {
    // ===== epilogue: one warp per TMEM lane quadrant (= atom of the tile) =====
    const int q = warp & 3;
    int it = 0;
    for (int ph = 0; ph < n_phases; ++ph) {
      const int chunk = ph / D, mode = ph - chunk * D;
      const bool last = (mode == D - 1);
      const int64_t L = (int64_t)1 << (6 * (D - 1 - mode));
      float* __restrict__ outp = last ? a.Y : ((mode & 1) ? a.ws1 : a.ws0);
      const int64_t out_k = last ? a.k : 32 * a.cpc, out_base = last ? (int64_t)chunk * 32 * a.cpc : 0;
      const float* __restrict__ xin = (last && (a.shift != 0.f || a.diag != nullptr || a.dots != nullptr)) ? a.X : nullptr;
      const float* __restrict__ dg = last ? a.diag : nullptr;
      const float alpha = last ? a.alpha : 1.f;
      const int accumulate = last ? a.accumulate : 0;
      const int64_t row_stride = L * out_k;
      const int lsh = 6 * (D - 1 - mode);
      const int64_t lmask = ((int64_t)1 << lsh) - 1;
      double dacc = 0.0;
      int64_t dacc_r0 = -1;
      double* const dp = (last && a.dots != nullptr) ? a.dots + (a.dots_row ? (int64_t)(*a.dots_row) * a.k : 0) : nullptr;
      for (int64_t t = first; t < n_tiles; t += step, ++it) {
        const int acc = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        const int64_t cc = t >> pos_sh, tp = t & pos_mask;
        const int64_t out_r0 = out_base + cc * 32;
        const int64_t flat = tp * 4 + q;
        const int64_t p = flat >> lsh, l = flat & lmask;
        const int64_t base = (((p * kD) << lsh) + l) * out_k + out_r0 + lane;
        if (dp != nullptr && out_r0 != dacc_r0) {   // column block changed: flush the per-thread partial (rare)
          if (dacc_r0 >= 0) atomicAdd(dp + dacc_r0 + lane, dacc);
          dacc = 0.0;
          dacc_r0 = out_r0;
        }
        mbar_wait(bar_tfull0 + 8 * acc, aph);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * kD;
        // two halves of 32 columns: the TMEM buffer goes back to the MMA warp as soon as the second half is in
        // registers, before that half's global stores
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          tmem_ld16_at<0>(taddr + h * 32, v);
          tmem_ld16_at<1>(taddr + h * 32, v);
          tmem_ld_wait();
          if (h == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty0 + 8 * acc);
          }
          if (a.dbg & 4) continue;
          const int64_t hbase = base + (int64_t)(h * 32) * row_stride;
          if (!last || xin == nullptr) {
            if (accumulate) {
#pragma unroll
              for (int i = 0; i < 32; ++i) outp[hbase + (int64_t)i * row_stride] += alpha * __uint_as_float(v[i]);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) outp[hbase + (int64_t)i * row_stride] = alpha * __uint_as_float(v[i]);
            }
          } else {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              float xv[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) xv[i] = xin[hbase + (int64_t)(c * 16 + i) * row_stride];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int64_t o = hbase + (int64_t)(c * 16 + i) * row_stride;
                float y = alpha * __uint_as_float(v[c * 16 + i]);
                if (a.shift != 0.f) y += a.shift * xv[i];
                if (dg != nullptr) y += dg[((p * kD + h * 32 + c * 16 + i) << lsh) + l] * xv[i];
                if (accumulate) y += outp[o];
                dacc += (double)xv[i] * (double)y;     // of the stored value: earlier terms of a Sum included
                outp[o] = y;
              }
            }
          }
        }
      }
      if (dp != nullptr && dacc_r0 >= 0) atomicAdd(dp + dacc_r0 + lane, dacc);
      // make this warp's global stores visible device-wide (and to the TMA / async proxy) before the phase ends
      __threadfence();
      asm volatile("fence.proxy.async;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_phase);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(128));
  }
}

// =======================================================================================================
// Second-generation fused kernel (round 2).  Same formulation, tile shape, descriptors and phase walk as
// kron_fused_tc_kernel; what changed is everything around the MMAs, driven by profiles/r1_kron_tc_summary.md
// (per-tile critical path, 339 MB written for a 134 MB result, 85 us of the last phase spent in 64 scattered
// loads + stores per epilogue thread):
//   * the RAW tile is the hi operand.  kind::tf32 reads the upper 19 bits of an fp32 word, i.e. hi = trunc(x) for
//     free; the split warps only compute lo = rna_tf32(x - trunc(x)) (exact difference, rounded once) into ONE lo
//     buffer: half the split's shared-memory writes, 64 KB less shared memory, and a ring slot is simply held
//     until the MMAs that read it retire.  x = hi + lo to 2^-22 |x|, the same bound as the rounded split.
//   * epilogue through shared memory and TMA stores (UTMASTG): TMEM -> registers -> a 32 KB staging tile (one
//     128-byte row per warp instruction: conflict-free) -> four cp.async.bulk.tensor stores issued by a dedicated
//     warp.  In the last mode the operator's INPUT tile for the fused epilogue (alpha K x + (shift + diag) x,
//     <x, y>) is TMA-loaded into the same staging tile ahead of time and y is written over it in place, so the
//     epilogue threads issue no global memory instructions at all.
//   * 8 epilogue warps (two per TMEM lane quadrant, 32 columns each), 4 split warps.
//   * launched cooperatively (co-residency of the device-wide phase barrier is checked by the driver), and EVERY
//     phase after the first starts with that barrier: with several column chunks the workspace a phase overwrites
//     may still be read by a slower CTA's previous phase (write-after-read race for k > 128).
// Shared memory: factor hi|lo 32 KB, lo tile 32 KB, raw ring 3 x 32 KB, staging 2 x 32 KB = 224 KB.
// =======================================================================================================
constexpr int kF2Threads = 512;       // warps: 0 TMA loads, 1 MMA, 2 TMEM alloc, 3 staging (TMA stores + x tiles), 4-7 split, 8-15 epilogue
constexpr int kF2Ring = 3;
constexpr int kF2SplitWarps = 4;
constexpr int kF2EpiWarps = 8;
constexpr int kF2OffFacHi = 0;
constexpr int kF2OffFacLo = kFacBytes;
constexpr int kF2OffLo = 2 * kFacBytes;
constexpr int kF2OffRing = kF2OffLo + kTileBytes;
constexpr int kF2OffStg = kF2OffRing + kF2Ring * kTileBytes;
constexpr int kF2OffBars = kF2OffStg + 2 * kTileBytes;
constexpr int kF2Smem = kF2OffBars + 256 + 1024;
static_assert(kF2Smem <= 232448, "fused2: shared memory budget");

struct Fused2Maps {
  CUtensorMap in[kMaxFused];     // source of mode i (swizzled operand boxes)
  CUtensorMap fac[kMaxFused];
  CUtensorMap out[kMaxFused];    // destination of mode i (plain boxes): ws0 / ws1 / Y
  CUtensorMap xin;               // X with the destination geometry of the last mode (plain boxes): epilogue operand
};

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// lo = rna_tf32(x - trunc_tf32(x)) for `bytes` of a raw tile (layout-agnostic, elementwise)
__device__ __forceinline__ void split_lo(const unsigned char* raw, unsigned char* lo, int bytes, int tid, int nthreads) {
  for (int o = tid * 16; o < bytes; o += nthreads * 16) {
    const float4 v = *reinterpret_cast<const float4*>(raw + o);
    float4 l;
    l.x = tf32_rna(v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u));
    l.y = tf32_rna(v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u));
    l.z = tf32_rna(v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u));
    l.w = tf32_rna(v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u));
    *reinterpret_cast<float4*>(lo + o) = l;
  }
}

__global__ void __launch_bounds__(kF2Threads, 1)
    kron_fused2_tc_kernel(const __grid_constant__ Fused2Maps maps, FusedArgs a) {
  if (a.gate != nullptr && *a.gate != 0) return;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const int D = a.D;
  const uint32_t bar_full0 = sbase + kF2OffBars;               // [ring] raw tile (or factor) landed
  const uint32_t bar_slotfree0 = bar_full0 + 8 * kF2Ring;      // [ring] MMAs reading the slot retired / factor copied (count 1)
  const uint32_t bar_loready = bar_slotfree0 + 8 * kF2Ring;    // lo tile written (count kF2SplitWarps)
  const uint32_t bar_lofree = bar_loready + 8;                 // MMAs reading the lo tile retired
  const uint32_t bar_tfull0 = bar_lofree + 8;                  // [2] accumulator complete
  const uint32_t bar_tempty0 = bar_tfull0 + 16;                // [2] accumulator drained (count kF2EpiWarps)
  const uint32_t bar_stgfull0 = bar_tempty0 + 16;              // [2] staging tile ready for its next epilogue (free, or x tile landed)
  const uint32_t bar_outready0 = bar_stgfull0 + 16;            // [2] staging tile holds y (count kF2EpiWarps)
  const uint32_t bar_phase = bar_outready0 + 16;               // all stores of a phase complete (count 1)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kF2OffBars + 200);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kF2Ring; ++s) { mbar_init(bar_full0 + 8 * s, 1); mbar_init(bar_slotfree0 + 8 * s, 1); }
    mbar_init(bar_loready, kF2SplitWarps);
    mbar_init(bar_lofree, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_tfull0 + 8 * s, 1);
      mbar_init(bar_tempty0 + 8 * s, kF2EpiWarps);
      mbar_init(bar_stgfull0 + 8 * s, 1);
      mbar_init(bar_outready0 + 8 * s, kF2EpiWarps);
    }
    mbar_init(bar_phase, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  int64_t n = 1;
  for (int i = 0; i < D; ++i) n *= kD;
  const int64_t n_pos_tiles = n / kD / 4;
  const int pos_sh = 6 * (D - 1) - 2;
  const int64_t pos_mask = n_pos_tiles - 1;
  const int64_t n_tiles = n_pos_tiles * a.cpc;
  const int n_phases = (int)a.n_chunks * D;
  const int64_t first = blockIdx.x, step = gridDim.x;
  const int T = (int)((n_tiles - first + step - 1) / step);    // tiles of this CTA per phase (>= 1: grid <= n_tiles)

  if (warp == 0) {
    // ===== TMA loads: per phase the factor (through a ring slot), then the raw tiles =====
    if (lane == 0) {
      int rit = 0;
      unsigned int barriers_passed = 0;
      for (int ph = 0; ph < n_phases; ++ph) {
        const int chunk = ph / D, mode = ph - chunk * D;
        const int lsh = 6 * (D - 1 - mode);
        const int64_t lmask = ((int64_t)1 << lsh) - 1;
        {
          const int s = rit % kF2Ring;
          mbar_wait(bar_slotfree0 + 8 * s, ((rit / kF2Ring) & 1) ^ 1);
          mbar_expect_tx(bar_full0 + 8 * s, kFacBytes);
          const uint32_t dst = sbase + kF2OffRing + s * kTileBytes;
          tma_load_2d(dst, &maps.fac[mode], bar_full0 + 8 * s, 0, 0);
          tma_load_2d(dst + kFacBytes / 2, &maps.fac[mode], bar_full0 + 8 * s, 32, 0);
          ++rit;
        }
        if (ph > 0) {
          // the previous phase's stores are complete (this CTA), then device-wide: this phase reads what the others
          // wrote (mode > 0) or overwrites a workspace the others may still be reading (mode 0 of a later chunk)
          mbar_wait(bar_phase, (ph - 1) & 1);
          ++barriers_passed;
          if (!(a.dbg & 8)) grid_arrive_and_wait(a.sync_counter, barriers_passed * gridDim.x);
          asm volatile("fence.proxy.async;" ::: "memory");
        }
        const int in_base = (mode == 0) ? chunk * 32 * a.cpc : 0;
        for (int64_t t = first; t < n_tiles; t += step, ++rit) {
          const int64_t cc = t >> pos_sh, tp = t & pos_mask;
          const int in_r0 = in_base + (int)cc * 32;
          const int s = rit % kF2Ring;
          mbar_wait(bar_slotfree0 + 8 * s, ((rit / kF2Ring) & 1) ^ 1);
          mbar_expect_tx(bar_full0 + 8 * s, kTileBytes);
          const uint32_t dst = sbase + kF2OffRing + s * kTileBytes;
#pragma unroll
          for (int at = 0; at < 4; ++at) {
            const int64_t flat = tp * 4 + at;
            const int64_t p = flat >> lsh, l = flat & lmask;
            tma_load_3d(dst + at * kAtomBytes, &maps.in[mode], bar_full0 + 8 * s, in_r0, (int)l, (int)(p * kD));
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      int it = 0, rit = 0;
      const uint64_t ad_ring0 = make_desc(sbase + kF2OffRing, kAtomBytes, 512, kLayoutSw128Base32);
      const uint64_t ad_lo = make_desc(sbase + kF2OffLo, kAtomBytes, 512, kLayoutSw128Base32);
      const uint64_t bd_hi = make_desc(sbase + kF2OffFacHi, 16, 1024, kLayoutSw128);
      const uint64_t bd_lo = make_desc(sbase + kF2OffFacLo, 16, 1024, kLayoutSw128);
      for (int ph = 0; ph < n_phases; ++ph) {
        ++rit;                                                  // the factor's ring item
        for (int j = 0; j < T; ++j, ++it, ++rit) {
          const int acc = it & 1;
          const int s = rit % kF2Ring;
          mbar_wait(bar_tempty0 + 8 * acc, ((it >> 1) & 1) ^ 1);
          mbar_wait(bar_loready, it & 1);                       // lo tile of this tile written (=> raw tile landed, factor in place)
          tc_fence_after();
          const uint32_t d = tmem_base + acc * kD;
          const uint64_t ad_hi = ad_ring0 + (uint64_t)((s * kTileBytes) >> 4);
          uint32_t accum = 0;
          if (!(a.dbg & 2))
#pragma unroll
          for (int term = 0; term < 3; ++term) {               // small terms first: A_lo B_hi, A_hi B_lo, A_hi B_hi
            const uint64_t ad0 = (term == 0) ? ad_lo : ad_hi;
            const uint64_t bd0 = (term == 1) ? bd_lo : bd_hi;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              const uint64_t ad = ad0 + (uint64_t)((kk * 1024) >> 4);
              const uint64_t bd = bd0 + (uint64_t)(((kk / 4) * (kFacBytes / 2) + (kk % 4) * 32) >> 4);
              umma_tf32(d, ad, bd, kIdesc, accum);
              accum = 1;
            }
          }
          umma_commit(bar_slotfree0 + 8 * s);
          umma_commit(bar_lofree);
          umma_commit(bar_tfull0 + 8 * acc);
        }
      }
    }
  } else if (warp == 3) {
    // ===== staging manager: x tiles of the last mode in, y tiles out (TMA) =====
    if (lane == 0) {
      const int total = n_phases * T;
      // staging tile b is prepared for the epilogue of global tile j: free (plain arrive) or, for a last-mode tile,
      // filled with the operator's input at the destination coordinates
      auto prepare = [&](int j) {
        if (j >= total) return;
        const int b = j & 1;
        const int ph = j / T;
        const int chunk = ph / D, mode = ph - chunk * D;
        const bool need_x = (mode == D - 1) && (a.shift != 0.f || a.diag != nullptr || a.dots != nullptr);
        if (!need_x) { mbar_arrive(bar_stgfull0 + 8 * b); return; }
        const int64_t t = first + (int64_t)(j - ph * T) * step;
        const int64_t cc = t >> pos_sh, tp = t & pos_mask;
        const int r0 = chunk * 32 * a.cpc + (int)cc * 32;
        mbar_expect_tx(bar_stgfull0 + 8 * b, kTileBytes);
        const uint32_t dst = sbase + kF2OffStg + b * kTileBytes;
#pragma unroll
        for (int at = 0; at < 4; ++at)                         // last mode: L = 1, atom = one p
          tma_load_3d(dst + at * kAtomBytes, &maps.xin, bar_stgfull0 + 8 * b, r0, 0, (int)((tp * 4 + at) * kD));
      };
      prepare(0);
      prepare(1);
      int it = 0;
      for (int ph = 0; ph < n_phases; ++ph) {
        const int chunk = ph / D, mode = ph - chunk * D;
        const bool last = (mode == D - 1);
        const int lsh = 6 * (D - 1 - mode);
        const int64_t lmask = ((int64_t)1 << lsh) - 1;
        const int out_base = last ? chunk * 32 * a.cpc : 0;
        for (int64_t t = first; t < n_tiles; t += step, ++it) {
          const int b = it & 1;
          const int64_t cc = t >> pos_sh, tp = t & pos_mask;
          mbar_wait(bar_outready0 + 8 * b, (it >> 1) & 1);
          if (!(a.dbg & 4)) {
            const uint32_t src = sbase + kF2OffStg + b * kTileBytes;
#pragma unroll
            for (int at = 0; at < 4; ++at) {
              const int64_t flat = tp * 4 + at;
              const int64_t p = flat >> lsh, l = flat & lmask;
              tma_store_3d(&maps.out[mode], src + at * kAtomBytes, out_base + (int)cc * 32, (int)l, (int)(p * kD));
            }
            bulk_commit();
            bulk_wait_read0();                                  // the stores have read the staging tile
          }
          prepare(it + 2);
        }
        bulk_wait_all0();                                       // this phase's results are in global memory
        __threadfence();
        mbar_arrive(bar_phase);
      }
    }
  } else if (warp >= 4 && warp < 4 + kF2SplitWarps) {
    // ===== lo operand: raw ring slot -> lo tile (or, at a phase start, the factor pair) =====
    const int tid = threadIdx.x - 128;
    constexpr int kSplitThreads = kF2SplitWarps * 32;
    int it = 0, rit = 0;
    for (int ph = 0; ph < n_phases; ++ph) {
      {
        const int s = rit % kF2Ring;
        mbar_wait(bar_full0 + 8 * s, (rit / kF2Ring) & 1);
        if (it >= 1) mbar_wait(bar_lofree, (it - 1) & 1);      // every MMA of the previous phase retired (in order)
        const unsigned char* raw = smem + kF2OffRing + s * kTileBytes;
        for (int o = tid * 16; o < kFacBytes; o += kSplitThreads * 16)
          *reinterpret_cast<float4*>(smem + kF2OffFacHi + o) = *reinterpret_cast<const float4*>(raw + o);
        split_lo(raw, smem + kF2OffFacLo, kFacBytes, tid, kSplitThreads);
        fence_async_smem();
        // all split warps are done with the slot before it is handed back
        asm volatile("bar.sync 1, %0;" ::"n"(kF2SplitWarps * 32) : "memory");
        if (tid == 0) mbar_arrive(bar_slotfree0 + 8 * s);
        ++rit;
      }
      for (int j = 0; j < T; ++j, ++it, ++rit) {
        const int s = rit % kF2Ring;
        mbar_wait(bar_full0 + 8 * s, (rit / kF2Ring) & 1);
        if (it >= 1) mbar_wait(bar_lofree, (it - 1) & 1);      // MMAs of the previous tile no longer read the lo tile
        if (!(a.dbg & 1)) split_lo(smem + kF2OffRing + s * kTileBytes, smem + kF2OffLo, kTileBytes, tid, kSplitThreads);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_loready);
      }
    }
  } else if (warp >= 8) {
    // ===== epilogue: two warps per TMEM lane quadrant (= atom of the tile), 32 columns each =====
    const int q = warp & 3, h = (warp - 8) >> 2;
    int it = 0;
    for (int ph = 0; ph < n_phases; ++ph) {
      const int chunk = ph / D, mode = ph - chunk * D;
      const bool last = (mode == D - 1);
      const bool fused = last && (a.shift != 0.f || a.diag != nullptr || a.dots != nullptr);
      const float* __restrict__ dg = last ? a.diag : nullptr;
      const float alpha = last ? a.alpha : 1.f;
      const int lsh = 6 * (D - 1 - mode);
      const int64_t lmask = ((int64_t)1 << lsh) - 1;
      const int64_t out_base = last ? (int64_t)chunk * 32 * a.cpc : 0;
      double dacc = 0.0;
      int64_t dacc_r0 = -1;
      double* const dp = (last && a.dots != nullptr) ? a.dots + (a.dots_row ? (int64_t)(*a.dots_row) * a.k : 0) : nullptr;
      for (int64_t t = first; t < n_tiles; t += step, ++it) {
        const int acc = it & 1, b = it & 1;
        const uint32_t par = (it >> 1) & 1;
        const int64_t cc = t >> pos_sh, tp = t & pos_mask;
        const int64_t out_r0 = out_base + cc * 32;
        if (dp != nullptr && out_r0 != dacc_r0) {
          if (dacc_r0 >= 0) atomicAdd(dp + dacc_r0 + lane, dacc);
          dacc = 0.0;
          dacc_r0 = out_r0;
        }
        mbar_wait(bar_tfull0 + 8 * acc, par);
        tc_fence_after();
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * kD + h * 32;
        tmem_ld16_at<0>(taddr, v);
        tmem_ld16_at<1>(taddr, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty0 + 8 * acc);      // accumulator back to the MMA warp
        mbar_wait(bar_stgfull0 + 8 * b, par);                   // staging tile free / x tile landed
        // atom q of the staging tile: row a (= column of D') is 128 B: r = lane
        float* stg = reinterpret_cast<float*>(smem + kF2OffStg + b * kTileBytes + q * kAtomBytes) + (h * 32) * 32 + lane;
        if (!fused) {
#pragma unroll
          for (int i = 0; i < 32; ++i) stg[i * 32] = alpha * __uint_as_float(v[i]);
        } else {
          const int64_t flat = tp * 4 + q;
          const int64_t p = flat >> lsh, l = flat & lmask;
          float facc = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float x = stg[i * 32];
            float sd = a.shift;
            if (dg != nullptr) sd += dg[((p * kD + h * 32 + i) << lsh) + l];
            const float y = alpha * __uint_as_float(v[i]) + sd * x;
            facc += x * y;
            stg[i * 32] = y;
          }
          dacc += (double)facc;
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_outready0 + 8 * b);
      }
      if (dp != nullptr && dacc_r0 >= 0) atomicAdd(dp + dacc_r0 + lane, dacc);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(128));
  }
}

// ---- host side ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && p) fn = (EncodeTiledFn)p;
  return fn;
}

static int make_map_in(CUtensorMap* m, const float* in, int64_t pre, int64_t L, int64_t k) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(COLA_E_UNSUPPORTED, "kron_tc: cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[3] = {(cuuint64_t)k, (cuuint64_t)L, (cuuint64_t)(pre * kD)};
  cuuint64_t strides[2] = {(cuuint64_t)(k * 4), (cuuint64_t)(L * k * 4)};
  cuuint32_t box[3] = {32, 1, (cuuint32_t)kD};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)in, dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(COLA_E_BADARG, "kron_tc: cuTensorMapEncodeTiled(in) failed");
  return COLA_OK;
}

// same geometry as make_map_in, unswizzled: staging tiles of the epilogue (TMA stores of y, TMA loads of the x operand)
static int make_map_plain(CUtensorMap* m, const float* ptr, int64_t pre, int64_t L, int64_t k) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(COLA_E_UNSUPPORTED, "kron_tc: cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[3] = {(cuuint64_t)k, (cuuint64_t)L, (cuuint64_t)(pre * kD)};
  cuuint64_t strides[2] = {(cuuint64_t)(k * 4), (cuuint64_t)(L * k * 4)};
  cuuint32_t box[3] = {32, 1, (cuuint32_t)kD};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)ptr, dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(COLA_E_BADARG, "kron_tc: cuTensorMapEncodeTiled(plain) failed");
  return COLA_OK;
}

static int make_map_fac(CUtensorMap* m, const float* F, int64_t ldf) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(COLA_E_UNSUPPORTED, "kron_tc: cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)kD, (cuuint64_t)kD};
  cuuint64_t strides[1] = {(cuuint64_t)(ldf * 4)};
  cuuint32_t box[2] = {32, (cuuint32_t)kD};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)F, dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(COLA_E_BADARG, "kron_tc: cuTensorMapEncodeTiled(factor) failed");
  return COLA_OK;
}

}  // namespace cola

using namespace cola;

extern "C" {

int64_t cola_kron_tc_workspace_bytes(int64_t n, int64_t n_factors) {
  // two chunk-sized intermediates (up to 128 columns) + one cache line for the device-wide phase barrier
  return n_factors > 1 ? 2 * n * 128 * (int64_t)sizeof(float) + 128 : 0;
}

int cola_kron_tc_supported(int64_t n_factors, const int64_t* dims, int64_t k) {
  if (n_factors < 1 || n_factors > 8 || k < 32 || k % 32 != 0) return 0;
  int64_t n = 1;
  for (int64_t i = 0; i < n_factors; ++i) {
    if (dims[i] != kD) return 0;
    n *= kD;
  }
  // tiles are groups of 4 atoms: pre*L must be a multiple of 4 for every mode, i.e. n/64 % 4 == 0
  if (n_factors == 1) return 0;   // a single 64x64 dense factor: plain GEMM path
  return ((n / kD) % 4 == 0) ? 1 : 0;
}

int cola_kron_matmat_tc_f32(int64_t n_factors, const float* const* factors, const int64_t* ldf, const float* X,
                            float* Y, int64_t k, float* workspace, float alpha, float shift, const float* diag,
                            int accumulate, double* dots, const int32_t* dots_row, const int32_t* gate,
                            void* stream) {
  COLA_REQUIRE(factors && ldf && X && Y, "kron_tc: null pointer");
  COLA_REQUIRE(n_factors >= 2 && n_factors <= 8, "kron_tc: 2..8 factors of 64x64");
  COLA_REQUIRE(k >= 32 && k % 32 == 0, "kron_tc: k must be a multiple of 32");
  COLA_REQUIRE(workspace, "kron_tc: workspace required (cola_kron_tc_workspace_bytes)");
  COLA_REQUIRE(((uintptr_t)X % 16 == 0) && ((uintptr_t)Y % 16 == 0) && ((uintptr_t)workspace % 128 == 0),
               "kron_tc: X/Y must be 16-byte and workspace 128-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int64_t n = 1;
  for (int64_t i = 0; i < n_factors; ++i) n *= kD;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(kron_mode_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    attr_set = true;
  }
  float* ws0 = workspace;
  float* ws1 = workspace + n * 32;
  static const bool no_fused = getenv("COLA_KRON_NO_FUSED") != nullptr;   // A/B knob
  static const bool v1_only = getenv("COLA_KRON_V1") != nullptr;           // A/B knob: first-generation fused kernel
  if (n_factors <= kMaxFused && !no_fused && !v1_only && !accumulate) {
    // ---- second-generation fused kernel: one cooperative launch for the whole matmat
    static const int cpc_env2 = getenv("COLA_KRON_CPC") ? atoi(getenv("COLA_KRON_CPC")) : 0;
    int cpc = cpc_env2 > 0 ? cpc_env2 : 4;
    while (cpc > 1 && (k / 32) % cpc != 0) cpc >>= 1;
    ws1 = workspace + n * 32 * cpc;
    struct MapKey2 { const void* x; const void* y; const void* ws; const void* f[kMaxFused]; int64_t ldf[kMaxFused]; int64_t k, nf; int cpc; };
    static thread_local MapKey2 cached_key2 = {};
    static thread_local Fused2Maps cached_maps2;
    static thread_local bool cached_valid2 = false;
    MapKey2 key = {};
    key.x = X; key.y = Y; key.ws = workspace; key.k = k; key.nf = n_factors; key.cpc = cpc;
    for (int64_t i = 0; i < n_factors; ++i) { key.f[i] = factors[i]; key.ldf[i] = ldf[i]; }
    if (!(cached_valid2 && memcmp(&key, &cached_key2, sizeof(MapKey2)) == 0)) {
      for (int64_t i = 0; i < n_factors; ++i) {
        COLA_REQUIRE(((uintptr_t)factors[i] % 16 == 0) && (ldf[i] % 4 == 0), "kron_tc: factor alignment");
        int rc = make_map_fac(&cached_maps2.fac[i], factors[i], ldf[i]);
        if (rc) return rc;
        int64_t pre = 1, L = 1;
        for (int64_t j = 0; j < i; ++j) pre *= kD;
        for (int64_t j = i + 1; j < n_factors; ++j) L *= kD;
        const float* src = (i == 0) ? X : (((i - 1) % 2 == 0) ? ws0 : ws1);
        rc = make_map_in(&cached_maps2.in[i], src, pre, L, (i == 0) ? k : 32 * cpc);
        if (rc) return rc;
        const bool last = (i == n_factors - 1);
        float* dst = last ? Y : ((i % 2 == 0) ? ws0 : ws1);
        rc = make_map_plain(&cached_maps2.out[i], dst, pre, L, last ? k : 32 * cpc);
        if (rc) return rc;
      }
      int rc = make_map_plain(&cached_maps2.xin, X, n / kD, 1, k);
      if (rc) return rc;
      cached_key2 = key;
      cached_valid2 = true;
    }
    FusedArgs fa;
    fa.D = (int)n_factors;
    fa.cpc = cpc;
    fa.k = k; fa.n_chunks = k / (32 * cpc); fa.ws0 = ws0; fa.ws1 = ws1; fa.Y = Y; fa.X = X; fa.diag = diag; fa.alpha = alpha;
    fa.shift = shift; fa.accumulate = 0; fa.dots = dots; fa.dots_row = dots_row; fa.gate = gate;
    fa.sync_counter = reinterpret_cast<unsigned int*>(workspace + 2 * n * 128);
    static const int fdbg2 = getenv("COLA_KRON_DBG") ? atoi(getenv("COLA_KRON_DBG")) : 0;
    fa.dbg = fdbg2;
    static bool smem_set2 = false;
    if (!smem_set2) {
      cudaFuncSetAttribute(kron_fused2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kF2Smem);
      smem_set2 = true;
    }
    cudaMemsetAsync(fa.sync_counter, 0, sizeof(unsigned int), st);
    const int64_t n_tiles = n / kD / 4 * cpc;
    int64_t grid = sm_count();
    if (grid > n_tiles) grid = n_tiles;
    // cooperative launch: the device-wide phase barrier needs every CTA resident; the driver refuses the launch otherwise
    static const bool no_coop = getenv("COLA_KRON_NO_COOP") != nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kF2Threads); cfg.dynamicSmemBytes = kF2Smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative; attr[0].val.cooperative = 1;
    cfg.attrs = attr; cfg.numAttrs = no_coop ? 0 : 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, kron_fused2_tc_kernel, cached_maps2, fa);
    if (le != cudaSuccess) { cudaGetLastError(); return fail((int)le, cudaGetErrorString(le)); }
    return cuda_status("kron_fused2_tc");
  }
  if (n_factors <= kMaxFused && !no_fused) {
    // ---- one persistent launch for the whole matmat (first generation: kept for accumulate = 1 and A/B runs)
    FusedMaps maps;
    FusedArgs fa;
    // chunk width: whole RHS block up to 128 columns (fewest device-wide phases); COLA_KRON_CPC overrides (1, 2, 4)
    static const int cpc_env = getenv("COLA_KRON_CPC") ? atoi(getenv("COLA_KRON_CPC")) : 0;
    int cpc = cpc_env > 0 ? cpc_env : 4;
    while (cpc > 1 && (k / 32) % cpc != 0) cpc >>= 1;
    ws1 = workspace + n * 32 * cpc;
    // tensor maps depend only on (pointers, shapes): Krylov loops call with the same buffers every iteration, so the
    // last set is cached (cuTensorMapEncodeTiled costs several microseconds each on the host)
    struct MapKey { const void* x; const void* ws; const void* f[kMaxFused]; int64_t ldf[kMaxFused]; int64_t k, nf; int cpc; };
    static thread_local MapKey cached_key = {};
    static thread_local FusedMaps cached_maps;
    static thread_local bool cached_valid = false;
    MapKey key = {};
    key.x = X; key.ws = workspace; key.k = k; key.nf = n_factors; key.cpc = cpc;
    for (int64_t i = 0; i < n_factors; ++i) { key.f[i] = factors[i]; key.ldf[i] = ldf[i]; }
    if (!(cached_valid && memcmp(&key, &cached_key, sizeof(MapKey)) == 0)) {
      for (int64_t i = 0; i < n_factors; ++i) {
        COLA_REQUIRE(((uintptr_t)factors[i] % 16 == 0) && (ldf[i] % 4 == 0), "kron_tc: factor alignment");
        int rc = make_map_fac(&cached_maps.fac[i], factors[i], ldf[i]);
        if (rc) return rc;
        int64_t pre = 1, L = 1;
        for (int64_t j = 0; j < i; ++j) pre *= kD;
        for (int64_t j = i + 1; j < n_factors; ++j) L *= kD;
        const float* src = (i == 0) ? X : (((i - 1) % 2 == 0) ? ws0 : ws1);
        rc = make_map_in(&cached_maps.in[i], src, pre, L, (i == 0) ? k : 32 * cpc);
        if (rc) return rc;
      }
      cached_key = key;
      cached_valid = true;
    }
    maps = cached_maps;
    fa.D = (int)n_factors;
    fa.cpc = cpc;
    fa.k = k; fa.n_chunks = k / (32 * cpc); fa.ws0 = ws0; fa.ws1 = ws1; fa.Y = Y; fa.X = X; fa.diag = diag; fa.alpha = alpha;
    fa.shift = shift; fa.accumulate = accumulate; fa.dots = dots; fa.dots_row = dots_row; fa.gate = gate;
    fa.sync_counter = reinterpret_cast<unsigned int*>(workspace + 2 * n * 128);
    static const int fdbg = getenv("COLA_KRON_DBG") ? atoi(getenv("COLA_KRON_DBG")) : 0;
    fa.dbg = fdbg;
    const size_t smem = kFusedSmem;
    static bool smem_set = false;
    if (!smem_set) {
      cudaFuncSetAttribute(kron_fused_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      smem_set = true;
    }
    cudaMemsetAsync(fa.sync_counter, 0, sizeof(unsigned int), st);
    const int64_t n_tiles = n / kD / 4 * cpc;
    int64_t grid = sm_count();   // all CTAs must be co-resident for the device-wide barrier: 1 CTA per SM
    if (grid > n_tiles) grid = n_tiles;
    kron_fused_tc_kernel<<<(unsigned)grid, kFusedThreads, smem, st>>>(maps, fa);
    return cuda_status("kron_fused_tc");
  }
  CUtensorMap fac_maps[8];
  for (int64_t i = 0; i < n_factors; ++i) {
    COLA_REQUIRE(((uintptr_t)factors[i] % 16 == 0) && (ldf[i] % 4 == 0), "kron_tc: factor alignment");
    int rc = make_map_fac(&fac_maps[i], factors[i], ldf[i]);
    if (rc) return rc;
  }
  const int grid_max = sm_count();
  // One 32-column chunk of right-hand sides at a time through ALL modes: the two chunk-sized intermediates
  // (n*32 floats = 33.5 MB for n = 64^3) ping-pong inside L2, only X (in) and Y (out) stream through HBM.
  for (int64_t r0 = 0; r0 < k; r0 += 32) {
    const float* src = X;
    int64_t src_k = k, src_r0 = r0;       // the chunk inside X is strided (row length k); intermediates are dense (32)
    for (int64_t i = 0; i < n_factors; ++i) {
      const bool last = (i == n_factors - 1);
      int64_t pre = 1, L = 1;
      for (int64_t j = 0; j < i; ++j) pre *= kD;
      for (int64_t j = i + 1; j < n_factors; ++j) L *= kD;
      float* dst = last ? Y : ((i % 2 == 0) ? ws0 : ws1);
      CUtensorMap map_in;
      int rc = make_map_in(&map_in, src, pre, L, src_k);
      if (rc) return rc;
      TcArgs a;
      a.out = dst; a.epi_x = nullptr; a.diag = nullptr; a.L = L; a.pre = pre;
      a.in_r0 = src_r0; a.out_k = last ? k : 32; a.out_r0 = last ? r0 : 0;
      a.n_tiles = pre * L / 4; a.alpha = last ? alpha : 1.f; a.shift = 0.f; a.accumulate = 0;
      a.last_mode = last ? 1 : 0; a.dots = nullptr; a.dots_row = dots_row; a.gate = gate;
      static const int dbg = getenv("COLA_KRON_DBG") ? atoi(getenv("COLA_KRON_DBG")) : 0;
      a.dbg = dbg;
      if (last) {
        const bool epi = (shift != 0.f) || diag || dots;
        a.epi_x = epi ? X : nullptr; a.diag = diag; a.shift = shift; a.accumulate = accumulate; a.dots = dots;
      }
      const int64_t grid = a.n_tiles < grid_max ? a.n_tiles : grid_max;
      kron_mode_tc_kernel<<<(unsigned)grid, kTcThreads, kSmemBytes, st>>>(map_in, fac_maps[i], a);
      rc = cuda_status("kron_mode_tc");
      if (rc) return rc;
      src = dst; src_k = a.out_k; src_r0 = a.out_r0;
    }
  }
  return COLA_OK;
}

int cola_kronsum_matmat_tc_f32(int64_t n_factors, const float* const* factors, const int64_t* ldf, const float* X,
                               float* Y, int64_t k, float alpha, float shift, const float* diag, int accumulate,
                               double* dots, const int32_t* dots_row, const int32_t* gate, void* stream) {
  COLA_REQUIRE(factors && ldf && X && Y, "kronsum_tc: null pointer");
  COLA_REQUIRE(X != Y, "kronsum_tc: X and Y must not alias");
  COLA_REQUIRE(n_factors >= 2 && n_factors <= 8, "kronsum_tc: 2..8 factors of 64x64");
  COLA_REQUIRE(k >= 32 && k % 32 == 0, "kronsum_tc: k must be a multiple of 32");
  COLA_REQUIRE(((uintptr_t)X % 16 == 0) && ((uintptr_t)Y % 16 == 0), "kronsum_tc: X/Y must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int64_t n = 1;
  for (int64_t i = 0; i < n_factors; ++i) n *= kD;
  COLA_REQUIRE((n / kD) % 4 == 0, "kronsum_tc: n / 64 must be a multiple of 4");
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(kron_mode_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    attr_set = true;
  }
  CUtensorMap fac_maps[8], in_maps[8];
  int64_t pres[8], Ls[8];
  for (int64_t i = 0; i < n_factors; ++i) {
    COLA_REQUIRE(((uintptr_t)factors[i] % 16 == 0) && (ldf[i] % 4 == 0), "kronsum_tc: factor alignment");
    int rc = make_map_fac(&fac_maps[i], factors[i], ldf[i]);
    if (rc) return rc;
    pres[i] = 1; Ls[i] = 1;
    for (int64_t j = 0; j < i; ++j) pres[i] *= kD;
    for (int64_t j = i + 1; j < n_factors; ++j) Ls[i] *= kD;
    rc = make_map_in(&in_maps[i], X, pres[i], Ls[i], k);      // every mode contracts X itself
    if (rc) return rc;
  }
  const int grid_max = sm_count();
  const bool epi = (shift != 0.f) || diag || dots;
  // A 32-column chunk of right-hand sides goes through ALL modes before the next one starts: its slice of X
  // (n*32 floats) is fetched from HBM by the first mode and found in L2 by the others, its slice of Y is
  // read-modified in L2, so the sum of D contractions moves about what one does.
  for (int64_t r0 = 0; r0 < k; r0 += 32) {
    for (int64_t i = 0; i < n_factors; ++i) {
      const bool last = (i == n_factors - 1);
      TcArgs a;
      a.out = Y; a.L = Ls[i]; a.pre = pres[i];
      a.in_r0 = r0; a.out_k = k; a.out_r0 = r0;
      a.n_tiles = pres[i] * Ls[i] / 4; a.alpha = alpha;
      a.last_mode = 1;                                         // epilogue indexing is mode-independent
      a.accumulate = (i > 0 || accumulate) ? 1 : 0;
      a.epi_x = (last && epi) ? X : nullptr;
      a.shift = last ? shift : 0.f; a.diag = last ? diag : nullptr;
      a.dots = last ? dots : nullptr; a.dots_row = dots_row; a.gate = gate;
      static const int dbg = getenv("COLA_KRON_DBG") ? atoi(getenv("COLA_KRON_DBG")) : 0;
      a.dbg = dbg;
      const int64_t grid = a.n_tiles < grid_max ? a.n_tiles : grid_max;
      kron_mode_tc_kernel<<<(unsigned)grid, kTcThreads, kSmemBytes, st>>>(in_maps[i], fac_maps[i], a);
      int rc = cuda_status("kronsum_mode_tc");
      if (rc) return rc;
    }
  }
  return COLA_OK;
}

}  // extern "C"
