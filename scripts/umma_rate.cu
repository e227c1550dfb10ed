// Micro-benchmark (not part of the product): sustained cycles per tcgen05.mma kind::tf32 (M=128, K=8) for the operand
// sources / N values the Kronecker kernel can choose between, alone and with tcgen05.ld traffic from 4 or 8 other warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/umma_rate scripts/umma_rate.cu && scripts/umma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
               ::"r"(bar), "r"(parity) : "memory");
}
// mode: 0 SS (A MN-major SW128_BASE32B), 1 TS (A in TMEM), 2 SS (A K-major SW128)
// side: 0 nothing, 1 = warps 4-7 run tcgen05.ld loops, 2 = warps 4-11, 3 = warps 4-7 run LDS loops, 4 = warps 4-7 tcgen05.st loops
__global__ void __launch_bounds__(384, 1) rate(int mode, int N, int side, int reps, long long* out) {
  extern __shared__ unsigned char raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  const uint32_t sb = smem_u32(smem);
  const uint32_t sA = sb, sB = sb + 65536;                 // A: 64 KB region, B: 64 KB region
  uint64_t* bar = (uint64_t*)(smem + 131072);
  uint32_t* slot = (uint32_t*)(smem + 131072 + 64);
  volatile int* stop = (volatile int*)(smem + 131072 + 128);
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  for (int i = tid; i < 131072 / 4; i += blockDim.x) ((float*)smem)[i] = 1.0f;
  if (tid == 0) {
    *stop = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = *slot;
  if (warp == 1 && lane == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((mode == 0 ? 1u : 0u) << 15) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t ad_mn = make_desc(sA, 8192, 512, 1);
    const uint64_t ad_k = make_desc(sA, 16, 1024, 2);
    const uint64_t bd0 = make_desc(sB, 16, 1024, 2);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint64_t bd = bd0 + (uint64_t)(((kk / 4) * 32768 + (kk % 4) * 32) >> 4);
        if (mode == 0) mma_ss(tm, ad_mn + (uint64_t)((kk * 1024) >> 4), bd, idesc, kk ? 1u : (r ? 1u : 0u));
        else if (mode == 1) mma_ts(tm, tm + 256 + kk * 8, bd, idesc, kk ? 1u : (r ? 1u : 0u));
        else mma_ss(tm, ad_k + (uint64_t)(((kk / 4) * 16384 + (kk % 4) * 32) >> 4), bd, idesc, kk ? 1u : (r ? 1u : 0u));
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    mbar_wait(smem_u32(bar), 0);
    const long long t1 = clock64();
    *stop = 1;
    out[blockIdx.x * 4 + 0] = t1 - t0;
  } else if (warp >= 4 && side != 0 && (side == 2 || warp < 8)) {
    // background traffic until the MMA thread is done
    long long n = 0;
    const long long t0 = clock64();
    const uint32_t taddr = tm + ((uint32_t)((warp & 3) * 32) << 16) + 320 + ((warp >= 8) ? 64 : 0);
    float sink = 0.f;
    while (!*stop) {
      if (side == 1 || side == 2) {
        uint32_t v[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                       : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                         "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                       : "r"(taddr + c * 16) : "memory");
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        sink += __uint_as_float(v[0]);
        n += 64 * 4 * 32;        // bytes per warp iteration
      } else if (side == 3) {
        const float* p = (const float*)(smem + 32768) + lane;
#pragma unroll
        for (int i = 0; i < 64; ++i) sink += p[i * 32];
        n += 64 * 128;
      } else {
        uint32_t v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = lane + i;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                       ::"r"(taddr + c * 16), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                         "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        n += 64 * 4 * 32;
      }
    }
    const long long t1 = clock64();
    if (lane == 0 && warp == 4) { out[blockIdx.x * 4 + 1] = n; out[blockIdx.x * 4 + 2] = t1 - t0; out[blockIdx.x * 4 + 3] = (long long)sink; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(512));
}

int main() {
  long long* d;
  cudaMalloc(&d, 4 * 148 * sizeof(long long));
  const int smem = 131072 + 1024 + 1024;
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 400;
  const char* mname[3] = {"SS A MN-major", "TS A in TMEM ", "SS A K-major  "};
  const char* sname[5] = {"alone", "+4 warps LDTM", "+8 warps LDTM", "+4 warps LDS", "+4 warps STTM"};
  for (int grid : {1, 148}) {
    for (int mode = 0; mode < 3; ++mode) {
      for (int N : {64, 128, 256}) {
        for (int side = 0; side < 5; ++side) {
          cudaMemset(d, 0, 4 * 148 * sizeof(long long));
          rate<<<grid, 384, smem>>>(mode, N, side, reps, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("mode %d N %d side %d: %s\n", mode, N, side, cudaGetErrorString(e)); return 1; }
          long long h[4];
          cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
          printf("grid %3d  %s N=%3d %-14s: %6.1f cycles/MMA", grid, mname[mode], N, sname[side], (double)h[0] / (reps * 8));
          if (side) printf("   side traffic %.1f B/cycle/warp", (double)h[1] / (double)h[2]);
          printf("\n");
        }
      }
    }
  }
  return 0;
}
