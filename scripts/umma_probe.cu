// Bring-up probe for tcgen05.mma kind::tf32 descriptors (not part of the product).
#include <cstdio>
#include <cstdint>
#include <cstring>
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
__host__ __device__ inline float Aval(int m, int k) { return (float)((m + 3 * k) % 7 - 3); }
__host__ __device__ inline float Bval(int n, int k) { return (float)((2 * n + k) % 5 - 2); }
// variant: 0 = A K-major noswz; 1 = A MN-major noswz; 2 = A MN-major SW128 + B K-major SW128
__global__ void probe(int variant, float* out, uint32_t* info) {
  extern __shared__ unsigned char raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = smem;            // up to 16 KB
  unsigned char* sB = smem + 16384;    // up to 8 KB
  uint64_t* bar = (uint64_t*)(smem + 32768);
  uint32_t* slot = (uint32_t*)(smem + 32768 + 64);
  const int tid = threadIdx.x, warp = tid / 32;
  for (int i = tid; i < 32768 / 4; i += blockDim.x) ((float*)smem)[i] = 0.f;
  __syncthreads();
  // fill A (128 x 8) and B (64 x 8)
  for (int idx = tid; idx < 128 * 8; idx += blockDim.x) {
    int m = idx / 8, k = idx % 8;
    uint32_t off;
    if (variant == 0) {          // K-major interleave: core = 8 rows x 16 B; k-chunk stride 128 B (LBO), row-group stride 256 B (SBO)
      off = (m % 8) * 16 + (m / 8) * 256 + (k / 4) * 128 + (k % 4) * 4;
    } else if (variant == 1) {   // MN-major interleave: core = 8 k-rows x 16 B (4 m); m-chunk stride 128 B (SBO)
      off = (m % 4) * 4 + k * 16 + (m / 4) * 128;
    } else {                     // MN-major SW128: atom = 8 k-rows x 128 B (32 m); atoms 8192 B apart; swizzle 16B-chunk ^= k
      int atom = m / 32, mm = m % 32;
      uint32_t chunk = (mm / 8) ^ (k % 4);          // Swizzle<2,5,2>: 32-byte chunk ^= row % 4
      off = atom * 8192 + k * 128 + chunk * 32 + (mm % 8) * 4;
    }
    *(float*)(sA + off) = Aval(m, k);
  }
  for (int idx = tid; idx < 64 * 8; idx += blockDim.x) {
    int n = idx / 8, k = idx % 8;
    uint32_t off;
    if (variant < 2) off = (n % 8) * 16 + (n / 8) * 256 + (k / 4) * 128 + (k % 4) * 4;
    else { uint32_t chunk = (k / 4) ^ (n % 8); off = n * 128 + chunk * 16 + (k % 4) * 4; }   // K-major SW128, row = 128 B
    *(float*)(sB + off) = Bval(n, k);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (tid == 0) {
    info[0] = tmem;
    uint64_t ad, bd;
    uint32_t amajor;
    if (variant == 0) { ad = make_desc(smem_u32(sA), 128, 256, 0); amajor = 0; }
    else if (variant == 1) { ad = make_desc(smem_u32(sA), 128 /*LBO: k-group stride (unused)*/, 128 /*SBO: m-chunk stride*/, 0); amajor = 1; }
    else { ad = make_desc(smem_u32(sA), 8192, 512, 1); amajor = 1; }
    if (variant < 2) bd = make_desc(smem_u32(sB), 128, 256, 0);
    else bd = make_desc(smem_u32(sB), 16, 1024, 2);
    uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (amajor << 15) | (0u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    info[1] = idesc;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  // everyone waits for the MMA
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
               ::"r"(smem_u32(bar)) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[16];
  for (int c = 0; c < 4; ++c) {
    uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c * 16;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) out[tid * 64 + c * 16 + i] = __uint_as_float(v[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(64));
}
int main() {
  float* d; uint32_t* info;
  cudaMalloc(&d, 128 * 64 * 4); cudaMalloc(&info, 64);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  static float h[128 * 64];
  for (int variant = 0; variant < 3; ++variant) {
    cudaMemset(d, 0xff, 128 * 64 * 4);
    probe<<<1, 128, 40000>>>(variant, d, info);
    cudaError_t e = cudaDeviceSynchronize();
    uint32_t hi[2]; cudaMemcpy(hi, info, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double maxerr = 0; int nz = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
      float r = 0; for (int k = 0; k < 8; ++k) r += Aval(m, k) * Bval(n, k);
      double err = fabs((double)h[m * 64 + n] - r); if (err > maxerr) maxerr = err; if (h[m * 64 + n] != 0) nz++;
    }
    { int badm[128] = {0}, badn[64] = {0};
      for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) { float r = 0; for (int k = 0; k < 8; ++k) r += Aval(m, k) * Bval(n, k);
        if (h[m * 64 + n] != r) { badm[m]++; badn[n]++; } }
      printf("  bad rows m:"); for (int m = 0; m < 128; ++m) if (badm[m]) printf(" %d(%d)", m, badm[m]); printf("\n  bad cols n:");
      for (int n = 0; n < 64; ++n) if (badn[n]) printf(" %d(%d)", n, badn[n]); printf("\n"); }
    printf("variant %d: %s tmem=0x%x idesc=0x%x maxerr=%g nonzero=%d  D[0][0..3]=%g %g %g %g  D[1][0]=%g D[33][5]=%g\n", variant,
           cudaGetErrorString(e), hi[0], hi[1], maxerr, nz, h[0], h[1], h[2], h[3], h[64], h[33 * 64 + 5]);
    float r00 = 0, r10 = 0; for (int k = 0; k < 8; ++k) { r00 += Aval(0, k) * Bval(0, k); r10 += Aval(1, k) * Bval(0, k); }
    printf("   expected D[0][0]=%g D[1][0]=%g\n", r00, r10);
  }
  return 0;
}
