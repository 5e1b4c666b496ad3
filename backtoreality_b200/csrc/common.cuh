// common.cuh -- shared helpers for libb2r.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b2r.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libb2r targets sm_100a (B200) only"
#endif

namespace b2r {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// thread-local detail of the last failing call (b2r_last_error)
void set_error(const char *fmt, ...);

#define B2R_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      b2r::set_error(__VA_ARGS__);      \
      return B2R_ERR_INVALID_ARG;       \
    }                                   \
  } while (0)

#define B2R_CUDA(call)                                                                 \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess) {                                                          \
      b2r::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                     __LINE__);                                                        \
      return B2R_ERR_CUDA;                                                             \
    }                                                                                  \
  } while (0)

#define B2R_CHECK_LAUNCH()                                                                 \
  do {                                                                                     \
    cudaError_t e__ = cudaGetLastError();                                                  \
    if (e__ != cudaSuccess) {                                                              \
      b2r::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, \
                     __LINE__);                                                            \
      return B2R_ERR_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

// The reference's squared-norm contraction: the y term is a rounded
// product, x then z are fused.  Spelled with round-to-nearest intrinsics so that nvcc can
// neither re-associate nor re-contract it.
__device__ __forceinline__ float sumsq_ref(float a, float b, float c) {
  return __fmaf_rn(c, c, __fmaf_rn(a, a, __fmul_rn(b, b)));
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// streaming (read-once) 128-bit load / store that do not pollute L1
__device__ __forceinline__ int4 ldg_stream_v4(const int4 *p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream_v4(float4 *p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

inline int ceil_div(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }

}  // namespace b2r
