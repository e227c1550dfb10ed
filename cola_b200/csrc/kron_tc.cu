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
//   B' = F^T: K-major (F is row-major [a][j]), TMA-loaded with the same 128B swizzle.
//   D' = 128 TMEM lanes (positions) x 64 columns (a), fp32.
// fp32-grade accuracy from tf32 tensor cores: 3xTF32 error compensation, x = hi + lo with hi, lo in tf32 and
// D' = A_hi B_hi + A_hi B_lo + A_lo B_hi (the dropped lo x lo term is 2^-21 relative).
// Roofline: the contraction itself is HBM-bound at d = 64 (48 flop/B < tf32 ridge); what the B200 actually charges is
// measured in scripts/umma_rate.cu: 79 cycles per 128xNx8 tf32 MMA for N <= 128 (MN-major A in shared memory).
//
// Two kernels: kron_fused4_tc_kernel (2-3 factors: the whole matmat in ONE persistent cooperative launch, see its
// header) and mode_tc_kernel (one mode per launch, factor size 64 or 128: longer or mixed chains, KronSum, BlockDiag).
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace cola {

constexpr int kD = 64;                 // factor size of the fused path
constexpr int kAtomBytes = 64 * 128;   // one TMA box: 64 rows (j) x 32 floats
constexpr int kTileBytes = 4 * kAtomBytes;   // 32 KB: 128 positions x 64 j
constexpr int kFacBytes = kD * kD * 4;       // 16 KB

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
// ---- shared by the fused kernel ------------------------------------------------------------------------------
constexpr int kMaxFused = 3;
constexpr int kFusedMaxPhases = 256;
struct FusedArgs {
  int D, cpc, n_pairs, n_phases;
  int64_t k, n_chunks;
  float* out[2][kMaxFused];       // destination of mode i for workspace set s (the last mode: Y)
  const float* X; const float* diag;
  float alpha, shift; int accumulate;
  double* dots; const int32_t* dots_row; const int32_t* gate;
  unsigned int* counters;         // one per phase, zeroed by the host before the launch
  int dbg;                        // bring-up knob (COLA_KRON_DBG): 1 skip split math, 2 skip MMAs, 4 skip stores, 8 skip phase waits
  long long* prof;                // bring-up: per-CTA wait-time breakdown (COLA_KRON_PROF), else null
};
// Phase walk.  Column chunks are interleaved in pairs: (c0,m0) (c1,m0) (c0,m1) (c1,m1) ...  A phase depends on the
// phase two back, so a CTA arrives at the device-wide counter of its finished phase and waits on it one whole phase
// later: the barrier latency and the pipeline drain are hidden.  Each chunk of a pair has its own workspace pair.
// An odd last chunk (or chunks too wide to pair) runs mode after mode with dep = the previous phase.
struct PhaseInfo { int chunk, mode, set, dep; };
__device__ __forceinline__ PhaseInfo phase_info(int ph, int D, int n_pairs) {
  PhaseInfo pi;
  const int paired = n_pairs * 2 * D;
  if (ph < paired) {
    const int pair = ph / (2 * D), w = ph - pair * 2 * D;
    pi.mode = w >> 1; pi.set = w & 1; pi.chunk = pair * 2 + pi.set; pi.dep = ph - 2;
  } else {
    const int q = ph - paired;
    pi.chunk = n_pairs * 2 + q / D; pi.mode = q % D; pi.set = 0; pi.dep = ph - 1;
  }
  return pi;
}
// lo word of the 3xTF32 split: x - trunc_tf32(x) is exact in fp32; adding half a tf32 ulp to its bit pattern makes
// the tensor core's truncation a round-to-nearest (no cvt instruction on the operand path)
__device__ __forceinline__ uint32_t tf32_lo_bits(float x) {
  const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  return __float_as_uint(x - hi) + 0x1000u;
}
// bring-up instrumentation: cycles a role's leading thread spends in one class of wait
#define F3_TIMED(slot, stmt)                                  \
  do {                                                        \
    if (a.prof != nullptr) {                                  \
      const long long t0_ = clock64();                        \
      stmt;                                                   \
      pw[slot] += clock64() - t0_;                            \
    } else {                                                  \
      stmt;                                                   \
    }                                                         \
  } while (0)

// =======================================================================================================
// Fourth-generation fused kernel (round 2), shaped by scripts/umma_rate.cu (profiles/r2_umma_rate.log): on the B200 a
// tcgen05.mma kind::tf32 128xNx8 costs 79 cycles for N = 64 AND N = 128 with the MN-major A tile in shared memory, 131
// cycles for any N <= 128 with A in tensor memory (v3: 24 x 131 cycles per tile = 133 us per matmat, measured 138) and
// 135 with a K-major shared A.  So:
//   * A' stays in shared memory (raw TMA tile = hi operand, the tensor core reads the upper 19 bits; lo tile written by
//     the split warps, DOUBLE-buffered so that split(t+1) runs under MMA(t));
//   * the factor pair is STACKED along N: B' = [F_hi ; F_lo] (128 rows, K-major).  One pass of 8 MMAs with N = 128
//     gives A_hi F_hi (accumulator columns 0-63) and A_hi F_lo (columns 64-127) for the price of one N = 64 pass; a second
//     pass of 8 MMAs (N = 64) adds A_lo F_hi into columns 0-63; the epilogue adds the two halves in registers.
//     16 MMAs per tile instead of 24: 1264 cycles, 53 us per cfg3 matmat at the tensor-pipe floor.
//   * interleaved chunk pairs with per-phase device-wide counters (a phase waits for the phase two back, see
//     phase_info), epilogue straight from registers with the x operand of the fused last mode requested before the
//     accumulator wait, cooperative launch.  accumulate = 1 (earlier terms of a Sum already in Y) is supported.
// Shared memory: stacked factor pair 32 KB + lo tiles 2 x 32 KB + raw ring 4 x 32 KB = 224 KB.  TMEM: 2 x 128 columns.
// =======================================================================================================
constexpr int kF4Threads = 512;       // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 phase signal, 4-7 split, 8-15 epilogue
constexpr int kF4Ring = 4;
constexpr int kF4SplitWarps = 4;
constexpr int kF4EpiWarps = 8;
constexpr int kF4OffFac = 0;                                   // k-chunk c at c * 16 KB: [hi rows 0-63 | lo rows 64-127]
constexpr int kF4OffLo = 2 * kFacBytes;
constexpr int kF4OffRing = kF4OffLo + 2 * kTileBytes;
constexpr int kF4OffBars = kF4OffRing + kF4Ring * kTileBytes;
constexpr int kF4Smem = kF4OffBars + 256 + 1024;
static_assert(kF4Smem <= 232448, "fused4: shared memory budget");
constexpr uint32_t kIdescN128 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (0u << 16) | ((128u >> 3) << 17) |
                                ((128u >> 4) << 24);

struct Fused4Maps {
  CUtensorMap in[2][kMaxFused];   // [workspace set][mode]: swizzled operand boxes
  CUtensorMap fac[kMaxFused];
};

__global__ void __launch_bounds__(kF4Threads, 1)
    kron_fused4_tc_kernel(const __grid_constant__ Fused4Maps maps, FusedArgs a) {
  if (a.gate != nullptr && *a.gate != 0) return;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const int D = a.D;
  const uint32_t bar_full0 = sbase + kF4OffBars;               // [ring] raw tile (or factor) landed
  const uint32_t bar_slotfree0 = bar_full0 + 8 * kF4Ring;      // [ring] MMAs reading the slot retired / factor copied (count 1)
  const uint32_t bar_loready0 = bar_slotfree0 + 8 * kF4Ring;   // [2] lo tile written (count kF4SplitWarps)
  const uint32_t bar_lofree0 = bar_loready0 + 16;              // [2] MMAs reading the lo tile retired
  const uint32_t bar_tfull0 = bar_lofree0 + 16;                // [2] accumulator complete
  const uint32_t bar_tempty0 = bar_tfull0 + 16;                // [2] accumulator drained (count kF4EpiWarps)
  const uint32_t bar_phase = bar_tempty0 + 16;                 // epilogue warps finished a phase (count kF4EpiWarps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kF4OffBars + 200);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  long long pw[4] = {0, 0, 0, 0};
  const long long t_start = clock64();

  if (threadIdx.x == 0) {
    for (int s = 0; s < kF4Ring; ++s) { mbar_init(bar_full0 + 8 * s, 1); mbar_init(bar_slotfree0 + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_loready0 + 8 * s, kF4SplitWarps);
      mbar_init(bar_lofree0 + 8 * s, 1);
      mbar_init(bar_tfull0 + 8 * s, 1);
      mbar_init(bar_tempty0 + 8 * s, kF4EpiWarps);
    }
    mbar_init(bar_phase, kF4EpiWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(256));
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
  const int n_phases = a.n_phases;
  const int64_t first = blockIdx.x, step = gridDim.x;
  const int T = (int)((n_tiles - first + step - 1) / step);    // tiles of this CTA per phase (>= 1: grid <= n_tiles)

  if (warp == 0) {
    // ===== TMA loads: the factor when the mode changes (through a ring slot), then the raw tiles =====
    if (lane == 0) {
      int rit = 0, prev_mode = -1;
      for (int ph = 0; ph < n_phases; ++ph) {
        const PhaseInfo pi = phase_info(ph, D, a.n_pairs);
        const int lsh = 6 * (D - 1 - pi.mode);
        const int64_t lmask = ((int64_t)1 << lsh) - 1;
        if (pi.mode != prev_mode) {
          const int s = rit % kF4Ring;
          F3_TIMED(0, mbar_wait(bar_slotfree0 + 8 * s, ((rit / kF4Ring) & 1) ^ 1));
          mbar_expect_tx(bar_full0 + 8 * s, kFacBytes);
          const uint32_t dst = sbase + kF4OffRing + s * kTileBytes;
          tma_load_2d(dst, &maps.fac[pi.mode], bar_full0 + 8 * s, 0, 0);
          tma_load_2d(dst + kFacBytes / 2, &maps.fac[pi.mode], bar_full0 + 8 * s, 32, 0);
          ++rit;
          prev_mode = pi.mode;
        }
        if (pi.dep >= 0 && !(a.dbg & 8)) {
          // every CTA has finished phase `dep` (and, in order, everything before it): its results may be read and the
          // workspace this phase overwrites is no longer in use
          const unsigned int* c = a.counters + pi.dep;
          unsigned int v;
          const long long t0 = clock64();
          do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c) : "memory");
          } while (v < gridDim.x);
          pw[1] += clock64() - t0;
          asm volatile("fence.proxy.async;" ::: "memory");
        }
        const int in_base = (pi.mode == 0) ? pi.chunk * 32 * a.cpc : 0;
        const CUtensorMap* map = &maps.in[pi.set][pi.mode];
        for (int64_t t = first; t < n_tiles; t += step, ++rit) {
          const int64_t cc = t >> pos_sh, tp = t & pos_mask;
          const int in_r0 = in_base + (int)cc * 32;
          const int s = rit % kF4Ring;
          F3_TIMED(0, mbar_wait(bar_slotfree0 + 8 * s, ((rit / kF4Ring) & 1) ^ 1));
          mbar_expect_tx(bar_full0 + 8 * s, kTileBytes);
          const uint32_t dst = sbase + kF4OffRing + s * kTileBytes;
#pragma unroll
          for (int at = 0; at < 4; ++at) {
            const int64_t flat = tp * 4 + at;
            const int64_t p = flat >> lsh, l = flat & lmask;
            tma_load_3d(dst + at * kAtomBytes, map, bar_full0 + 8 * s, in_r0, (int)l, (int)(p * kD));
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: pass 1 raw tile x [F_hi ; F_lo] (N = 128), pass 2 lo tile x F_hi (N = 64) =====
    if (lane == 0) {
      int it = 0, rit = 0, prev_mode = -1;
      const uint64_t ad_ring0 = make_desc(sbase + kF4OffRing, kAtomBytes, 512, kLayoutSw128Base32);
      const uint64_t ad_lo0 = make_desc(sbase + kF4OffLo, kAtomBytes, 512, kLayoutSw128Base32);
      const uint64_t bd0 = make_desc(sbase + kF4OffFac, 16, 1024, kLayoutSw128);
      for (int ph = 0; ph < n_phases; ++ph) {
        const PhaseInfo pi = phase_info(ph, D, a.n_pairs);
        if (pi.mode != prev_mode) { ++rit; prev_mode = pi.mode; }   // the factor's ring item
        for (int j = 0; j < T; ++j, ++it, ++rit) {
          const int acc = it & 1;
          const uint32_t par = (it >> 1) & 1;
          const int s = rit % kF4Ring;
          F3_TIMED(0, mbar_wait(bar_tempty0 + 8 * acc, par ^ 1));
          F3_TIMED(1, mbar_wait(bar_loready0 + 8 * acc, par));      // lo tile written (=> raw tile landed, factor in place)
          tc_fence_after();
          const uint32_t d = tmem_base + acc * 128;
          const uint64_t ad_hi = ad_ring0 + (uint64_t)((s * kTileBytes) >> 4);
          const uint64_t ad_lo = ad_lo0 + (uint64_t)((acc * kTileBytes) >> 4);
          if (!(a.dbg & 2)) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              const uint64_t bd = bd0 + (uint64_t)(((kk / 4) * (2 * kFacBytes / 2) + (kk % 4) * 32) >> 4);
              umma_tf32(d, ad_hi + (uint64_t)((kk * 1024) >> 4), bd, kIdescN128, kk ? 1u : 0u);
            }
          }
          umma_commit(bar_slotfree0 + 8 * s);                       // the raw tile is only read by pass 1
          if (!(a.dbg & 2)) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              const uint64_t bd = bd0 + (uint64_t)(((kk / 4) * (2 * kFacBytes / 2) + (kk % 4) * 32) >> 4);
              umma_tf32(d, ad_lo + (uint64_t)((kk * 1024) >> 4), bd, kIdesc, 1u);
            }
          }
          umma_commit(bar_lofree0 + 8 * acc);
          umma_commit(bar_tfull0 + 8 * acc);
        }
      }
    }
  } else if (warp == 3) {
    // ===== phase signal: this CTA's results of a phase are in global memory -> device-wide counter =====
    if (lane == 0) {
      for (int ph = 0; ph < n_phases; ++ph) {
        mbar_wait(bar_phase, ph & 1);
        __threadfence();
        atomicAdd(a.counters + ph, 1u);
      }
    }
  } else if (warp >= 4 && warp < 4 + kF4SplitWarps) {
    // ===== split warps: lo tile of every raw tile; the stacked factor pair when the mode changes =====
    const int tid = threadIdx.x - 128;
    constexpr int kSplitThreads = kF4SplitWarps * 32;
    int it = 0, rit = 0, prev_mode = -1;
    for (int ph = 0; ph < n_phases; ++ph) {
      const PhaseInfo pi = phase_info(ph, D, a.n_pairs);
      if (pi.mode != prev_mode) {
        prev_mode = pi.mode;
        const int s = rit % kF4Ring;
        mbar_wait(bar_full0 + 8 * s, (rit / kF4Ring) & 1);
        // every MMA issued so far has retired (they complete in order) before the factor changes
        if (it >= 1) mbar_wait(bar_lofree0 + 8 * ((it - 1) & 1), ((it - 1) >> 1) & 1);
        const unsigned char* raw = smem + kF4OffRing + s * kTileBytes;
        for (int o = tid * 16; o < kFacBytes; o += kSplitThreads * 16) {
          // the TMA wrote k-chunk c (8 KB) at c * 8 KB; the stacked operand keeps it at c * 16 KB with lo 8 KB further
          const int c = o >> 13, w = o & 8191;
          // the factor is rounded to nearest (|lo| <= 2^-12 |F|, half of what truncation leaves): the dropped lo x lo
          // term of the product shrinks with it; the position operand cannot afford the extra shared-memory write
          const float4 v = *reinterpret_cast<const float4*>(raw + o);
          float4 hh, l;
          hh.x = tf32_rna(v.x); hh.y = tf32_rna(v.y); hh.z = tf32_rna(v.z); hh.w = tf32_rna(v.w);
          l.x = tf32_rna(v.x - hh.x); l.y = tf32_rna(v.y - hh.y); l.z = tf32_rna(v.z - hh.z); l.w = tf32_rna(v.w - hh.w);
          *reinterpret_cast<float4*>(smem + kF4OffFac + c * 16384 + w) = hh;
          *reinterpret_cast<float4*>(smem + kF4OffFac + c * 16384 + 8192 + w) = l;
        }
        fence_async_smem();
        asm volatile("bar.sync 1, %0;" ::"n"(kF4SplitWarps * 32) : "memory");   // all split warps are done with the slot
        if (tid == 0) mbar_arrive(bar_slotfree0 + 8 * s);
        ++rit;
      }
      for (int j = 0; j < T; ++j, ++it, ++rit) {
        const int s = rit % kF4Ring;
        const int buf = it & 1;
        F3_TIMED(0, mbar_wait(bar_full0 + 8 * s, (rit / kF4Ring) & 1));
        F3_TIMED(1, mbar_wait(bar_lofree0 + 8 * buf, ((it >> 1) & 1) ^ 1));   // MMAs of tile it-2 no longer read this lo tile
        if (!(a.dbg & 1)) {
          const unsigned char* raw = smem + kF4OffRing + s * kTileBytes;
          unsigned char* lo = smem + kF4OffLo + buf * kTileBytes;
#pragma unroll 4
          for (int o = tid * 16; o < kTileBytes; o += kSplitThreads * 16) {
            const float4 v = *reinterpret_cast<const float4*>(raw + o);
            uint4 l;
            l.x = tf32_lo_bits(v.x); l.y = tf32_lo_bits(v.y); l.z = tf32_lo_bits(v.z); l.w = tf32_lo_bits(v.w);
            *reinterpret_cast<uint4*>(lo + o) = l;
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_loready0 + 8 * buf);
      }
    }
  } else if (warp >= 8) {
    // ===== epilogue: two warps per TMEM lane quadrant (= atom of the tile), 32 output columns each, straight to global =====
    const int q = warp & 3, h = (warp - 8) >> 2;
    int it = 0;
    for (int ph = 0; ph < n_phases; ++ph) {
      const PhaseInfo pi = phase_info(ph, D, a.n_pairs);
      const bool last = (pi.mode == D - 1);
      const bool fused = last && (a.shift != 0.f || a.diag != nullptr || a.dots != nullptr);
      const bool accumulate = last && a.accumulate;
      float* __restrict__ outp = a.out[pi.set][pi.mode];
      const float* __restrict__ xin = a.X;
      const float* __restrict__ dg = last ? a.diag : nullptr;
      const float alpha = last ? a.alpha : 1.f;
      const int lsh = 6 * (D - 1 - pi.mode);
      const int64_t lmask = ((int64_t)1 << lsh) - 1;
      const int64_t out_k = last ? a.k : 32 * a.cpc;
      const int64_t out_base = last ? (int64_t)pi.chunk * 32 * a.cpc : 0;
      const int64_t row_stride = out_k << lsh;
      double dacc = 0.0;
      int64_t dacc_r0 = -1;
      double* const dp = (last && a.dots != nullptr) ? a.dots + (a.dots_row ? (int64_t)(*a.dots_row) * a.k : 0) : nullptr;
      for (int64_t t = first; t < n_tiles; t += step, ++it) {
        const int acc = it & 1;
        const uint32_t par = (it >> 1) & 1;
        const int64_t cc = t >> pos_sh, tp = t & pos_mask;
        const int64_t out_r0 = out_base + cc * 32;
        const int64_t flat = tp * 4 + q;
        const int64_t p = flat >> lsh, l = flat & lmask;
        // element (a = h*32 + i, r = lane) of this atom
        const int64_t base = ((((p * kD + h * 32) << lsh) + l) * out_k) + out_r0 + lane;
        if (dp != nullptr && out_r0 != dacc_r0) {
          if (dacc_r0 >= 0) atomicAdd(dp + dacc_r0 + lane, dacc);
          dacc = 0.0;
          dacc_r0 = out_r0;
        }
        // y starts as what the result is accumulated onto (earlier terms of a Sum already in Y) or zero; x (fused last
        // mode) and those old values are requested BEFORE the accumulator wait -- read in the store loop they would queue
        // behind every store (possible aliasing) at one exposed latency each
        float xv[32], y[32];
        if (fused) {
          const float* px = xin + base;
#pragma unroll
          for (int i = 0; i < 32; ++i, px += row_stride) xv[i] = *px;
        }
        if (accumulate) {
          const float* po_ = outp + base;
#pragma unroll
          for (int i = 0; i < 32; ++i, po_ += row_stride) y[i] = *po_;
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) y[i] = 0.f;
        }
        F3_TIMED(0, mbar_wait(bar_tfull0 + 8 * acc, par));
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 128 + h * 32;
#pragma unroll
        for (int half = 0; half < 2; ++half) {                 // A_hi F_hi + A_lo F_hi (columns 0-63), then A_hi F_lo (64-127)
          uint32_t v[32];
          tmem_ld16_at<0>(taddr + half * 64, v);
          tmem_ld16_at<1>(taddr + half * 64, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) y[i] += alpha * __uint_as_float(v[i]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty0 + 8 * acc);      // accumulator back to the MMA warp
        if (a.dbg & 4) continue;
        float* po = outp + base;
        if (!fused) {
#pragma unroll
          for (int i = 0; i < 32; ++i, po += row_stride) *po = y[i];
        } else {
          const float* pd = (dg != nullptr) ? dg + (((p * kD + h * 32) << lsh) + l) : nullptr;
          const int64_t dstride = (int64_t)1 << lsh;
          // <x, y> of the 32 stored values in fp32, folded into the fp64 accumulator once per tile: a per-element fp64
          // product costs two F2F conversions each (measured on the SpMM epilogue: 17 % of that kernel)
          float facc = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i, po += row_stride) {
            float sd = a.shift;
            if (pd != nullptr) sd += pd[i * dstride];
            const float yy = y[i] + sd * xv[i];
            facc += xv[i] * yy;                                 // of the stored value: earlier terms of a Sum included
            *po = yy;
          }
          dacc += (double)facc;
        }
      }
      if (dp != nullptr && dacc_r0 >= 0) atomicAdd(dp + dacc_r0 + lane, dacc);
      // this warp's stores visible device-wide (and to the TMA / async proxy of the reading CTAs) before the phase ends
      __threadfence();
      asm volatile("fence.proxy.async;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_phase);
    }
  }

  if (a.prof != nullptr && lane == 0 && (warp == 0 || warp == 1 || warp == 4 || warp == 8)) {
    // [cta][role: 0 producer, 1 mma, 2 split, 3 epilogue][wait class 0, wait class 1, -, total cycles]
    const int role = warp == 0 ? 0 : warp == 1 ? 1 : warp == 4 ? 2 : 3;
    long long* o = a.prof + ((int64_t)blockIdx.x * 4 + role) * 4;
    o[0] = pw[0]; o[1] = pw[1]; o[2] = pw[2]; o[3] = clock64() - t_start;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256));
  }
}

// =======================================================================================================
// One mode contraction per launch on the tensor cores, factor size KD = 64 or 128 (round 2):
//     out[p, a, l, r] = sum_j F[a, j] in[p, j, l, r],   F (KD x KD),  in (pre, KD, L, k) with k % 32 == 0
// Same tile formulation and 3xTF32 scheme as the fused kernel (positions = 4 atoms x 32 right-hand sides, raw TMA tile =
// hi operand, lo tile from the split warps, epilogue straight from registers), without phases: the tiles of one mode are
// independent, so a persistent grid walks them with no device-wide synchronisation.  It serves the modes the fused kernel
// does not take: 128-wide factors (BASELINE config 4: Kronecker(128, 128, 64)) and 64-wide modes next to them.
//   KD = 64 : factor pair stacked along N (8 + 8 MMAs per tile), ring of 4 raw tiles, 2 lo tiles.
//   KD = 128: the contraction index is taken in two halves of 64 (two raw half-tiles per output tile accumulate into one
//             128-column accumulator); per half  A_lo F_hi  first (the lo tile is then free for the next split while the
//             16 MMAs on the raw tile run), then  A_hi F_hi, A_hi F_lo  (N = 128 each: 24 MMAs of 79 cycles per half).
//             F_hi | F_lo resident (128 KB) + ring of 2 raw half-tiles + 1 lo tile = 224 KB.
// The epilogue of the operator (alpha, shift, diag, accumulate, <x, y> dots) applies when L == 1 (the last factor of a
// Kronecker chain), like cola_mode_contract_*'s.
// =======================================================================================================
constexpr int kMtThreads = 512;       // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-7 split, 8-15 epilogue
struct ModeTcArgs {
  float* out; const float* epi_x; const float* diag;
  int64_t L, pre, k;                  // in viewed as (pre, KD, L, k); out (pre, KD, L, k)
  int64_t n_pos_tiles;                // pre * L / 4
  float alpha, shift; int accumulate; int fused;
  double* dots; const int32_t* dots_row; const int32_t* gate;
};

template <int KD>
struct ModeTcCfg {
  static constexpr int kHalves = KD / 64;
  static constexpr int kFacBytes1 = KD * KD * 4;                  // one of hi / lo
  static constexpr int kRing = KD == 64 ? 4 : 2;
  static constexpr int kLoBufs = KD == 64 ? 2 : 1;
  static constexpr int kOffFac = 0;
  static constexpr int kOffLo = 2 * kFacBytes1;
  static constexpr int kOffRing = kOffLo + kLoBufs * kTileBytes;
  static constexpr int kOffBars = kOffRing + kRing * kTileBytes;
  static constexpr int kSmem = kOffBars + 256 + 1024;
  static_assert(kSmem <= 232448, "mode_tc: shared memory budget");
};

template <int KD>
__global__ void __launch_bounds__(kMtThreads, 1)
    mode_tc_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_fac, ModeTcArgs a) {
  using C = ModeTcCfg<KD>;
  if (a.gate != nullptr && *a.gate != 0) return;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_full0 = sbase + C::kOffBars;                 // [ring] raw half-tile landed
  const uint32_t bar_slotfree0 = bar_full0 + 8 * C::kRing;        // [ring] MMAs reading the slot retired
  const uint32_t bar_loready0 = bar_slotfree0 + 8 * C::kRing;     // [lo bufs] lo tile written (count 4)
  const uint32_t bar_lofree0 = bar_loready0 + 16;                 // [lo bufs] MMAs reading the lo tile retired
  const uint32_t bar_tfull0 = bar_lofree0 + 16;                   // [2] accumulator complete
  const uint32_t bar_tempty0 = bar_tfull0 + 16;                   // [2] accumulator drained (count 8)
  const uint32_t bar_fac = bar_tempty0 + 16;                      // factor landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::kOffBars + 200);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::kRing; ++s) { mbar_init(bar_full0 + 8 * s, 1); mbar_init(bar_slotfree0 + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_loready0 + 8 * s, 4);
      mbar_init(bar_lofree0 + 8 * s, 1);
      mbar_init(bar_tfull0 + 8 * s, 1);
      mbar_init(bar_tempty0 + 8 * s, 8);
    }
    mbar_init(bar_fac, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- factor: TMA boxes of 32 k-columns x KD rows, then hi = rna(F) in place and lo beside it -------------------
  // KD = 64 : k-chunk c at c * 16 KB as [hi rows | lo rows] (the stacked N = 128 operand);  KD = 128: hi chunks c * 16 KB,
  // lo chunks 64 KB further.
  constexpr int kChunkBytes = KD * 128;                           // one k-chunk of one of hi / lo
  constexpr int kChunkStride = KD == 64 ? 2 * kChunkBytes : kChunkBytes;
  constexpr int kLoOffset = KD == 64 ? kChunkBytes : C::kFacBytes1;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar_fac, C::kFacBytes1);
    for (int c = 0; c < KD / 32; ++c) tma_load_2d(sbase + C::kOffFac + c * kChunkStride, &map_fac, bar_fac, c * 32, 0);
  }
  mbar_wait(bar_fac, 0);
  for (int o = threadIdx.x * 16; o < C::kFacBytes1; o += kMtThreads * 16) {
    const int c = o / kChunkBytes, w = o - c * kChunkBytes;
    unsigned char* hp = smem + C::kOffFac + c * kChunkStride + w;
    const float4 v = *reinterpret_cast<const float4*>(hp);
    float4 hh, l;
    hh.x = tf32_rna(v.x); hh.y = tf32_rna(v.y); hh.z = tf32_rna(v.z); hh.w = tf32_rna(v.w);
    l.x = tf32_rna(v.x - hh.x); l.y = tf32_rna(v.y - hh.y); l.z = tf32_rna(v.z - hh.z); l.w = tf32_rna(v.w - hh.w);
    *reinterpret_cast<float4*>(hp) = hh;
    *reinterpret_cast<float4*>(hp + kLoOffset) = l;
  }
  fence_async_smem();
  __syncthreads();

  const int64_t ncb = a.k / 32;                                   // 32-column blocks
  const int64_t n_tiles = a.n_pos_tiles * ncb;                    // position tile major, column block minor
  const int64_t first = blockIdx.x, step = gridDim.x;

  if (warp == 0) {
    // ===== TMA loads: kHalves raw half-tiles per output tile =====
    if (lane == 0) {
      int rit = 0;
      for (int64_t t = first; t < n_tiles; t += step) {
        const int64_t tp = t / ncb, cc = t - tp * ncb;
        for (int h = 0; h < C::kHalves; ++h, ++rit) {
          const int s = rit % C::kRing;
          mbar_wait(bar_slotfree0 + 8 * s, ((rit / C::kRing) & 1) ^ 1);
          mbar_expect_tx(bar_full0 + 8 * s, kTileBytes);
          const uint32_t dst = sbase + C::kOffRing + s * kTileBytes;
#pragma unroll
          for (int at = 0; at < 4; ++at) {
            const int64_t flat = tp * 4 + at;
            const int64_t p = flat / a.L, l = flat - p * a.L;
            tma_load_3d(dst + at * kAtomBytes, &map_in, bar_full0 + 8 * s, (int)(cc * 32), (int)l, (int)(p * KD + h * 64));
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      int it = 0, rit = 0;
      const uint64_t ad_ring0 = make_desc(sbase + C::kOffRing, kAtomBytes, 512, kLayoutSw128Base32);
      const uint64_t ad_lo0 = make_desc(sbase + C::kOffLo, kAtomBytes, 512, kLayoutSw128Base32);
      const uint64_t bd0 = make_desc(sbase + C::kOffFac, 16, 1024, kLayoutSw128);
      for (int64_t t = first; t < n_tiles; t += step, ++it) {
        const int acc = it & 1;
        mbar_wait(bar_tempty0 + 8 * acc, (((it >> 1) & 1) ^ 1));
        const uint32_t d = tmem_base + acc * 128;
        for (int h = 0; h < C::kHalves; ++h, ++rit) {
          const int s = rit % C::kRing;
          const int lb = rit % C::kLoBufs;
          mbar_wait(bar_loready0 + 8 * lb, (rit / C::kLoBufs) & 1);      // lo tile written (=> raw half-tile landed)
          tc_fence_after();
          const uint64_t ad_hi = ad_ring0 + (uint64_t)((s * kTileBytes) >> 4);
          const uint64_t ad_lo = ad_lo0 + (uint64_t)((lb * kTileBytes) >> 4);
          if constexpr (KD == 64) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {                              // A_hi x [F_hi ; F_lo]
              const uint64_t bd = bd0 + (uint64_t)(((kk / 4) * kChunkStride + (kk % 4) * 32) >> 4);
              umma_tf32(d, ad_hi + (uint64_t)((kk * 1024) >> 4), bd, kIdescN128, kk ? 1u : 0u);
            }
            umma_commit(bar_slotfree0 + 8 * s);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {                              // A_lo x F_hi
              const uint64_t bd = bd0 + (uint64_t)(((kk / 4) * kChunkStride + (kk % 4) * 32) >> 4);
              umma_tf32(d, ad_lo + (uint64_t)((kk * 1024) >> 4), bd, kIdesc, 1u);
            }
            umma_commit(bar_lofree0 + 8 * lb);
          } else {
            // k-steps of this half: global k index 8 * h + kk
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {                              // A_lo x F_hi first: frees the single lo tile early
              const int kg = h * 8 + kk;
              const uint64_t bd = bd0 + (uint64_t)(((kg / 4) * kChunkStride + (kg % 4) * 32) >> 4);
              umma_tf32(d, ad_lo + (uint64_t)((kk * 1024) >> 4), bd, kIdescN128, (h | kk) ? 1u : 0u);
            }
            umma_commit(bar_lofree0 + 8 * lb);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {                              // A_hi x F_hi
              const int kg = h * 8 + kk;
              const uint64_t bd = bd0 + (uint64_t)(((kg / 4) * kChunkStride + (kg % 4) * 32) >> 4);
              umma_tf32(d, ad_hi + (uint64_t)((kk * 1024) >> 4), bd, kIdescN128, 1u);
            }
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {                              // A_hi x F_lo
              const int kg = h * 8 + kk;
              const uint64_t bd = bd0 + (uint64_t)((kLoOffset + (kg / 4) * kChunkStride + (kg % 4) * 32) >> 4);
              umma_tf32(d, ad_hi + (uint64_t)((kk * 1024) >> 4), bd, kIdescN128, 1u);
            }
            umma_commit(bar_slotfree0 + 8 * s);
          }
        }
        umma_commit(bar_tfull0 + 8 * acc);
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===== split warps: lo tile of every raw half-tile =====
    const int tid = threadIdx.x - 128;
    int rit = 0;
    for (int64_t t = first; t < n_tiles; t += step) {
      for (int h = 0; h < C::kHalves; ++h, ++rit) {
        const int s = rit % C::kRing;
        const int lb = rit % C::kLoBufs;
        mbar_wait(bar_full0 + 8 * s, (rit / C::kRing) & 1);
        mbar_wait(bar_lofree0 + 8 * lb, ((rit / C::kLoBufs) & 1) ^ 1);
        const unsigned char* raw = smem + C::kOffRing + s * kTileBytes;
        unsigned char* lo = smem + C::kOffLo + lb * kTileBytes;
#pragma unroll 4
        for (int o = tid * 16; o < kTileBytes; o += 128 * 16) {
          const float4 v = *reinterpret_cast<const float4*>(raw + o);
          uint4 l;
          l.x = tf32_lo_bits(v.x); l.y = tf32_lo_bits(v.y); l.z = tf32_lo_bits(v.z); l.w = tf32_lo_bits(v.w);
          *reinterpret_cast<uint4*>(lo + o) = l;
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_loready0 + 8 * lb);
      }
    }
  } else if (warp >= 8) {
    // ===== epilogue: two warps per TMEM lane quadrant (= atom), KD / 2 output columns each, in groups of 32 =====
    const int q = warp & 3, hw = (warp - 8) >> 2;
    const bool fused = a.fused != 0;
    const float* __restrict__ xin = a.epi_x;
    const float* __restrict__ dg = a.diag;
    float* __restrict__ outp = a.out;
    const int64_t row_stride = a.L * a.k;
    double dacc = 0.0;
    int64_t dacc_c = -1;
    double* const dp = (fused && a.dots != nullptr) ? a.dots + (a.dots_row ? (int64_t)(*a.dots_row) * a.k : 0) : nullptr;
    int it = 0;
    for (int64_t t = first; t < n_tiles; t += step, ++it) {
      const int acc = it & 1;
      const uint32_t par = (it >> 1) & 1;
      const int64_t tp = t / ncb, cc = t - tp * ncb;
      const int64_t flat = tp * 4 + q;
      const int64_t p = flat / a.L, l = flat - p * a.L;
      if (dp != nullptr && cc != dacc_c) {
        if (dacc_c >= 0) atomicAdd(dp + dacc_c * 32 + lane, dacc);
        dacc = 0.0;
        dacc_c = cc;
      }
      constexpr int kGroups = KD / 64;                           // groups of 32 output columns per warp
#pragma unroll
      for (int g = 0; g < kGroups; ++g) {
        const int a0 = hw * (KD / 2) + g * 32;                   // first output index of the group
        const int64_t base = ((p * KD + a0) * a.L + l) * a.k + cc * 32 + lane;
        // as in the fused kernel: y starts as the values accumulated onto (or zero), x and those values are requested
        // before the accumulator wait
        float xv[32], y[32];
        if (fused) {
          const float* px = xin + base;
#pragma unroll
          for (int i = 0; i < 32; ++i, px += row_stride) xv[i] = *px;
        }
        if (a.accumulate) {
          const float* po_ = outp + base;
#pragma unroll
          for (int i = 0; i < 32; ++i, po_ += row_stride) y[i] = *po_;
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) y[i] = 0.f;
        }
        if (g == 0) {
          mbar_wait(bar_tfull0 + 8 * acc, par);
          tc_fence_after();
        }
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 128 + (KD == 64 ? hw * 32 : a0);
#pragma unroll
        for (int half = 0; half < (KD == 64 ? 2 : 1); ++half) {  // KD = 64: the stacked accumulator halves add up
          uint32_t v[32];
          tmem_ld16_at<0>(taddr + half * 64, v);
          tmem_ld16_at<1>(taddr + half * 64, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) y[i] += a.alpha * __uint_as_float(v[i]);
        }
        if (g == kGroups - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty0 + 8 * acc);     // accumulator back to the MMA warp
        }
        float* po = outp + base;
        if (!fused) {
#pragma unroll
          for (int i = 0; i < 32; ++i, po += row_stride) *po = y[i];
        } else {
          const float* pd = (dg != nullptr) ? dg + (p * KD + a0) : nullptr;     // L == 1 when fused
          float facc = 0.f;                                      // fp32 over the 32 values, fp64 across (see the fused kernel)
#pragma unroll
          for (int i = 0; i < 32; ++i, po += row_stride) {
            float sd = a.shift;
            if (pd != nullptr) sd += pd[i];
            const float yy = y[i] + sd * xv[i];
            facc += xv[i] * yy;
            *po = yy;
          }
          dacc += (double)facc;
        }
      }
    }
    if (dp != nullptr && dacc_c >= 0) atomicAdd(dp + dacc_c * 32 + lane, dacc);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256));
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

static int make_map_in(CUtensorMap* m, const float* in, int64_t pre, int64_t L, int64_t k, int64_t d = kD) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(COLA_E_UNSUPPORTED, "kron_tc: cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[3] = {(cuuint64_t)k, (cuuint64_t)L, (cuuint64_t)(pre * d)};
  cuuint64_t strides[2] = {(cuuint64_t)(k * 4), (cuuint64_t)(L * k * 4)};
  cuuint32_t box[3] = {32, 1, (cuuint32_t)kD};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)in, dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(COLA_E_BADARG, "kron_tc: cuTensorMapEncodeTiled(in) failed");
  return COLA_OK;
}

static int make_map_fac(CUtensorMap* m, const float* F, int64_t ldf, int64_t d = kD) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(COLA_E_UNSUPPORTED, "kron_tc: cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)d};
  cuuint64_t strides[1] = {(cuuint64_t)(ldf * 4)};
  cuuint32_t box[2] = {32, (cuuint32_t)d};
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
  // chunk-sized intermediates (2 x 128 columns, or 2 x 2 x up to 64 for interleaved chunk pairs) + the device-wide
  // phase counters
  return n_factors > 1 ? 2 * n * 128 * (int64_t)sizeof(float) + 4096 : 0;
}

int cola_kron_tc_supported(int64_t n_factors, const int64_t* dims, int64_t k) {
  if (n_factors < 2 || n_factors > kMaxFused || k < 32 || k % 32 != 0) return 0;   // longer chains: per-mode kernel
  int64_t n = 1;
  for (int64_t i = 0; i < n_factors; ++i) {
    if (dims[i] != kD) return 0;
    n *= kD;
  }
  // tiles are groups of 4 atoms: pre*L must be a multiple of 4 for every mode, i.e. n/64 % 4 == 0
  return ((n / kD) % 4 == 0 && k / 32 * n_factors <= kFusedMaxPhases) ? 1 : 0;
}

int cola_kron_matmat_tc_f32(int64_t n_factors, const float* const* factors, const int64_t* ldf, const float* X,
                            float* Y, int64_t k, float* workspace, float alpha, float shift, const float* diag,
                            int accumulate, double* dots, const int32_t* dots_row, const int32_t* gate,
                            void* stream) {
  COLA_REQUIRE(factors && ldf && X && Y, "kron_tc: null pointer");
  COLA_REQUIRE(n_factors >= 2 && n_factors <= kMaxFused, "kron_tc: 2..3 factors of 64x64");
  COLA_REQUIRE(k >= 32 && k % 32 == 0, "kron_tc: k must be a multiple of 32");
  COLA_REQUIRE(workspace, "kron_tc: workspace required (cola_kron_tc_workspace_bytes)");
  COLA_REQUIRE(((uintptr_t)X % 16 == 0) && ((uintptr_t)Y % 16 == 0) && ((uintptr_t)workspace % 128 == 0),
               "kron_tc: X/Y must be 16-byte and workspace 128-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int64_t n = 1;
  for (int64_t i = 0; i < n_factors; ++i) n *= kD;
  {
    // ---- fused kernel: one cooperative launch for the whole matmat
    // chunk width: 64 columns when the block splits into pairs of them (measured on cfg3: 145 us vs 151 us for 32),
    // else 32; COLA_KRON_CPC overrides (1, 2, 4 blocks of 32 columns)
    static const int cpc_env = getenv("COLA_KRON_CPC") ? atoi(getenv("COLA_KRON_CPC")) : 0;
    int cpc = cpc_env > 0 ? cpc_env : (((k / 32) % 4 == 0) ? 2 : 1);
    while (cpc > 1 && (k / 32) % cpc != 0) cpc >>= 1;
    const int64_t n_chunks = k / (32 * cpc);
    static const bool no_pairs = getenv("COLA_KRON_NO_PAIRS") != nullptr;
    // a pair needs four intermediates of n * 32 * cpc floats: they fit the workspace up to cpc = 2
    const int n_pairs = (cpc <= 2 && !no_pairs) ? (int)(n_chunks / 2) : 0;
    const int64_t n_phases = n_chunks * n_factors;
    if (n_phases <= kFusedMaxPhases) {
      const int64_t wsz = n * 32 * cpc;                         // floats per intermediate
      auto wsp = [&](int set, int b) { return workspace + (int64_t)(set * 2 + b) * wsz; };
      // tensor maps depend only on (pointers, shapes): Krylov loops call with the same buffers every iteration, so the
      // last set is cached (cuTensorMapEncodeTiled costs several microseconds each on the host)
      struct MapKey { const void* x; const void* ws; const void* f[kMaxFused]; int64_t ldf[kMaxFused]; int64_t k, nf; int cpc; };
      static thread_local MapKey cached_key = {};
      static thread_local Fused4Maps cached_maps;
      static thread_local bool cached_valid = false;
      MapKey key = {};
      key.x = X; key.ws = workspace; key.k = k; key.nf = n_factors; key.cpc = cpc;
      for (int64_t i = 0; i < n_factors; ++i) { key.f[i] = factors[i]; key.ldf[i] = ldf[i]; }
      if (!(cached_valid && memcmp(&key, &cached_key, sizeof(MapKey)) == 0)) {
        cached_valid = false;
        for (int64_t i = 0; i < n_factors; ++i) {
          COLA_REQUIRE(((uintptr_t)factors[i] % 16 == 0) && (ldf[i] % 4 == 0), "kron_tc: factor alignment");
          int rc = make_map_fac(&cached_maps.fac[i], factors[i], ldf[i]);
          if (rc) return rc;
          int64_t pre = 1, L = 1;
          for (int64_t j = 0; j < i; ++j) pre *= kD;
          for (int64_t j = i + 1; j < n_factors; ++j) L *= kD;
          for (int set = 0; set < 2; ++set) {
            const float* src = (i == 0) ? X : wsp(set, (int)((i - 1) % 2));
            rc = make_map_in(&cached_maps.in[set][i], src, pre, L, (i == 0) ? k : 32 * cpc);
            if (rc) return rc;
          }
        }
        cached_key = key;
        cached_valid = true;
      }
      FusedArgs fa = {};
      fa.D = (int)n_factors; fa.cpc = cpc; fa.n_pairs = n_pairs; fa.n_phases = (int)n_phases;
      fa.k = k; fa.n_chunks = n_chunks;
      for (int set = 0; set < 2; ++set)
        for (int64_t i = 0; i < n_factors; ++i) fa.out[set][i] = (i == n_factors - 1) ? Y : wsp(set, (int)(i % 2));
      fa.X = X; fa.diag = diag; fa.alpha = alpha; fa.shift = shift; fa.dots = dots; fa.dots_row = dots_row; fa.gate = gate;
      fa.accumulate = accumulate;
      fa.counters = reinterpret_cast<unsigned int*>(workspace + 2 * n * 128);
      static const int fdbg = getenv("COLA_KRON_DBG") ? atoi(getenv("COLA_KRON_DBG")) : 0;
      fa.dbg = fdbg;
      static long long* prof_buf = nullptr;
      static const bool want_prof = getenv("COLA_KRON_PROF") != nullptr;
      if (want_prof && !prof_buf) cudaMalloc(&prof_buf, sizeof(long long) * 16 * 256);
      fa.prof = want_prof ? prof_buf : nullptr;
      static bool smem_set = false;
      if (!smem_set) {
        cudaFuncSetAttribute(kron_fused4_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kF4Smem);
        smem_set = true;
      }
      // (measured: letting the kernel's last CTA reset the counters instead of this memset changes nothing, 0.403 vs
      // 0.407 ms per cfg3 CG iteration, and would rely on the workspace's contents surviving between calls)
      cudaMemsetAsync(fa.counters, 0, sizeof(unsigned int) * n_phases, st);
      const int64_t n_tiles = n / kD / 4 * cpc;
      int64_t grid = sm_count();
      if (grid > n_tiles) grid = n_tiles;
      // cooperative launch: the device-wide phase counters need every CTA resident; the driver refuses the launch
      // otherwise (two such kernels on different streams are serialised by it instead of deadlocking)
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kF4Threads); cfg.stream = st;
      cfg.dynamicSmemBytes = kF4Smem;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeCooperative; attr[0].val.cooperative = 1;
      static const bool no_coop = getenv("COLA_KRON_NO_COOP") != nullptr;   // A/B knob (launch overhead measurement only)
      cfg.attrs = attr; cfg.numAttrs = no_coop ? 0 : 1;
      cudaError_t le = cudaLaunchKernelEx(&cfg, kron_fused4_tc_kernel, cached_maps, fa);
      if (le != cudaSuccess) { cudaGetLastError(); return fail((int)le, cudaGetErrorString(le)); }
      if (want_prof) {   // bring-up only: synchronous dump of the per-role wait breakdown (mean over CTAs)
        static long long host[16 * 256];
        cudaStreamSynchronize(st);
        cudaMemcpy(host, prof_buf, sizeof(long long) * 16 * grid, cudaMemcpyDeviceToHost);
        const char* names[4] = {"producer (slotfree, dep)", "mma (tempty, loready)", "split (full, lofree)", "epilogue (tfull, -)"};
        for (int r = 0; r < 4; ++r) {
          double w0 = 0, w1 = 0, tot = 0;
          for (int64_t c = 0; c < grid; ++c) { w0 += host[(c * 4 + r) * 4]; w1 += host[(c * 4 + r) * 4 + 1]; tot += host[(c * 4 + r) * 4 + 3]; }
          fprintf(stderr, "[kron prof] %-26s wait0 %8.0f wait1 %8.0f total %8.0f cycles (mean over %d CTAs)\n", names[r], w0 / grid, w1 / grid, tot / grid, (int)grid);
        }
      }
      return cuda_status("kron_fused4_tc");
    }
  }
  return fail(COLA_E_UNSUPPORTED, "kron_tc: too many column chunks for one launch (use cola_mode_contract_tc_f32 per mode)");
}

int cola_mode_contract_tc_supported(int64_t d, int64_t pre, int64_t L, int64_t k) {
  return ((d == 64 || d == 128) && pre >= 1 && L >= 1 && (pre * L) % 4 == 0 && k >= 32 && k % 32 == 0 &&
          pre * d < (int64_t)1 << 31 && L < (int64_t)1 << 31) ? 1 : 0;
}

int cola_mode_contract_tc_f32(const float* M, int64_t ldm, int64_t d, int64_t pre, int64_t L, int64_t k, const float* in,
                              float* out, float alpha, float shift, const float* diag, const float* epi_x, int accumulate,
                              double* dots, const int32_t* dots_row, const int32_t* gate, void* stream) {
  COLA_REQUIRE(M && in && out, "mode_contract_tc: null pointer");
  COLA_REQUIRE(in != out, "mode_contract_tc: in and out must not alias");
  COLA_REQUIRE(cola_mode_contract_tc_supported(d, pre, L, k), "mode_contract_tc: d in {64, 128}, k % 32 == 0, pre * L % 4 == 0");
  COLA_REQUIRE(((uintptr_t)M % 16 == 0) && (ldm % 4 == 0) && ((uintptr_t)in % 16 == 0) && ((uintptr_t)out % 16 == 0),
               "mode_contract_tc: 16-byte alignment");
  const bool fused = (shift != 0.f) || diag || dots;
  COLA_REQUIRE(!fused || (L == 1 && epi_x), "mode_contract_tc: shift / diag / dots need L == 1 and epi_x");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // tensor maps depend only on (pointers, shapes): Krylov loops call with the same buffers every step (two or three modes
  // alternate), so a few recent pairs are kept
  struct Entry { const void* in; const void* M; int64_t ldm, d, pre, L, k; CUtensorMap map_in, map_fac; bool valid; };
  static thread_local Entry cache[8] = {};
  static thread_local int next_slot = 0;
  Entry* e = nullptr;
  for (auto& c : cache)
    if (c.valid && c.in == in && c.M == M && c.ldm == ldm && c.d == d && c.pre == pre && c.L == L && c.k == k) { e = &c; break; }
  if (e == nullptr) {
    e = &cache[next_slot];
    next_slot = (next_slot + 1) % 8;
    e->valid = false;
    int rc = make_map_in(&e->map_in, in, pre, L, k, d);
    if (rc) return rc;
    rc = make_map_fac(&e->map_fac, M, ldm, d);
    if (rc) return rc;
    e->in = in; e->M = M; e->ldm = ldm; e->d = d; e->pre = pre; e->L = L; e->k = k; e->valid = true;
  }
  ModeTcArgs a;
  a.out = out; a.epi_x = fused ? epi_x : nullptr; a.diag = diag; a.L = L; a.pre = pre; a.k = k;
  a.n_pos_tiles = pre * L / 4; a.alpha = alpha; a.shift = shift; a.accumulate = accumulate; a.fused = fused ? 1 : 0;
  a.dots = dots; a.dots_row = dots_row; a.gate = gate;
  const int64_t n_tiles = a.n_pos_tiles * (k / 32);
  int64_t grid = sm_count();
  if (grid > n_tiles) grid = n_tiles;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(mode_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, ModeTcCfg<64>::kSmem);
    cudaFuncSetAttribute(mode_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, ModeTcCfg<128>::kSmem);
    attr_set = true;
  }
  if (d == 64) mode_tc_kernel<64><<<(unsigned)grid, kMtThreads, ModeTcCfg<64>::kSmem, st>>>(e->map_in, e->map_fac, a);
  else mode_tc_kernel<128><<<(unsigned)grid, kMtThreads, ModeTcCfg<128>::kSmem, st>>>(e->map_in, e->map_fac, a);
  return cuda_status("mode_contract_tc");
}

}  // extern "C"
