// Shared device/host helpers for libcola_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>

#include "../../include/cola_b200.h"

namespace cola {

constexpr int kSMsFallback = 148;  // B200: 2 dies x 74 SMs

// ---- host-side status plumbing ------------------------------------------------------------------
extern thread_local char g_err[512];
extern std::atomic<int64_t> g_launches;
int sm_count();

inline int fail(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
inline int cuda_status(const char* where) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return (int)e;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return COLA_OK;
}
#define COLA_REQUIRE(cond, msg) \
  do {                          \
    if (!(cond)) return cola::fail(COLA_E_BADARG, msg); \
  } while (0)

// ---- vector access ------------------------------------------------------------------------------
// Vec<T,N>: N contiguous elements moved with one 4/8/16-byte access.
template <typename T, int N>
struct alignas(sizeof(T) * N) Vec {
  T v[N];
};

template <typename T, int N>
__device__ __forceinline__ Vec<T, N> ldg(const T* p) {
  return *reinterpret_cast<const Vec<T, N>*>(p);
}
template <typename T, int N>
__device__ __forceinline__ void stg(T* p, const Vec<T, N>& x) {
  *reinterpret_cast<Vec<T, N>*>(p) = x;
}
// streaming variants: data touched once per sweep should not displace reusable lines in L1
template <typename T, int N>
__device__ __forceinline__ Vec<T, N> ldg_stream(const T* p) {
  Vec<T, N> r;
  if constexpr (sizeof(T) * N == 16) {
    float4 t = __ldcs(reinterpret_cast<const float4*>(p));
    r = *reinterpret_cast<Vec<T, N>*>(&t);
  } else if constexpr (sizeof(T) * N == 8) {
    float2 t = __ldcs(reinterpret_cast<const float2*>(p));
    r = *reinterpret_cast<Vec<T, N>*>(&t);
  } else {
    r = *reinterpret_cast<const Vec<T, N>*>(p);
  }
  return r;
}
// L2-only load (ld.global.cg): data another SM, or an earlier pass of this kernel, may have rewritten
template <typename T, int N>
__device__ __forceinline__ Vec<T, N> ldg_cg(const T* p) {
  Vec<T, N> r;
  if constexpr (sizeof(T) * N == 16) {
    float4 t = __ldcg(reinterpret_cast<const float4*>(p));
    r = *reinterpret_cast<Vec<T, N>*>(&t);
  } else if constexpr (sizeof(T) * N == 8) {
    float2 t = __ldcg(reinterpret_cast<const float2*>(p));
    r = *reinterpret_cast<Vec<T, N>*>(&t);
  } else if constexpr (sizeof(T) * N == 4) {
    float t = __ldcg(reinterpret_cast<const float*>(p));
    r = *reinterpret_cast<Vec<T, N>*>(&t);
  } else {
    r = *reinterpret_cast<const Vec<T, N>*>(p);
  }
  return r;
}
// 16-byte accesses carrying an L2 eviction policy (createpolicy): lines a kernel re-reads pass after pass are marked
// evict_last so that streamed operands do not push them out of L2
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
template <typename T, int N>
__device__ __forceinline__ Vec<T, N> ldg_hint(const T* p, uint64_t pol) {
  static_assert(sizeof(T) * N == 16, "16-byte vectors only");
  uint4 t;
  asm volatile("ld.global.cg.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
               : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "l"(p), "l"(pol));
  return *reinterpret_cast<Vec<T, N>*>(&t);
}
template <typename T, int N>
__device__ __forceinline__ void stg_hint(T* p, const Vec<T, N>& x, uint64_t pol) {
  static_assert(sizeof(T) * N == 16, "16-byte vectors only");
  const uint4 t = *reinterpret_cast<const uint4*>(&x);
  asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;"
               :: "l"(p), "r"(t.x), "r"(t.y), "r"(t.z), "r"(t.w), "l"(pol) : "memory");
}
template <typename T, int N>
__device__ __forceinline__ void stg_stream(T* p, const Vec<T, N>& x) {
  if constexpr (sizeof(T) * N == 16) {
    __stcs(reinterpret_cast<float4*>(p), *reinterpret_cast<const float4*>(&x));
  } else if constexpr (sizeof(T) * N == 8) {
    __stcs(reinterpret_cast<float2*>(p), *reinterpret_cast<const float2*>(&x));
  } else {
    *reinterpret_cast<Vec<T, N>*>(p) = x;
  }
}

// widest vector (in elements) usable for rows of length k starting at `base` with leading dim ld
template <typename T>
inline int pick_vec(int64_t k, int64_t ld, const void* a, const void* b = nullptr, const void* c = nullptr,
                    const void* d = nullptr) {
  int maxv = 16 / (int)sizeof(T);
  for (int v = maxv; v > 1; v >>= 1) {
    uintptr_t al = (uintptr_t)(v * sizeof(T));
    bool ok = (k % v == 0) && (ld % v == 0);
    const void* ps[4] = {a, b, c, d};
    for (auto p : ps)
      if (p && ((uintptr_t)p % al)) ok = false;
    if (ok) return v;
  }
  return 1;
}

// ---- reductions ---------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// safe divide of the reference (cola/linalg/inverse/cg.py:173-178): |den| < 1e-40 -> den = 1e-40
template <typename T>
__device__ __forceinline__ T safe_div(T num, T den) {
  const T tiny = (T)1e-40;
  T d = (fabs((double)den) < (double)tiny) ? tiny : den;
  return num / d;
}

// Row/lane decomposition used by every (n,k) sweep: a block of NT threads covers `rows_per_pass`
// consecutive rows; thread (r,l) owns columns [l*VEC, l*VEC+VEC) (+ stride lanes*VEC for wide k).
struct RowMap {
  int lanes;          // threads per row
  int rows_per_pass;  // rows covered by one block pass
};
inline RowMap row_map(int64_t k, int vec, int nthreads) {
  int64_t need = (k + vec - 1) / vec;
  int lanes = (int)(need < nthreads ? need : nthreads);
  RowMap m;
  m.lanes = lanes;
  m.rows_per_pass = nthreads / lanes;
  return m;
}

}  // namespace cola
