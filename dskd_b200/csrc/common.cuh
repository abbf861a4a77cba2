// Shared helpers for the dskd_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string.h>

#include <algorithm>

#include "../../include/dskd_b200.h"

namespace dskd {

void set_error(const char* fmt, ...);
void count_launch();

#define DSKD_REQUIRE(cond, ...)                \
  do {                                         \
    if (!(cond)) {                             \
      ::dskd::set_error(__VA_ARGS__);          \
      return DSKD_EINVAL;                      \
    }                                          \
  } while (0)

#define DSKD_CUDA_OK(expr)                                                               \
  do {                                                                                   \
    cudaError_t e__ = (expr);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      ::dskd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return DSKD_ECUDA;                                                                 \
    }                                                                                    \
  } while (0)

#define DSKD_LAUNCH_OK(name)                                                             \
  do {                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess) {                                                            \
      ::dskd::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));       \
      return DSKD_ECUDA;                                                                 \
    }                                                                                    \
    ::dskd::count_launch();                                                              \
  } while (0)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of it

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

#ifdef __CUDACC__
// 128-bit streaming accesses: the features are read once and the gradient is written once, so
// they bypass L1 allocation (non-volatile asm: the compiler may batch them); small reused tables
// use the default cached path.
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
  float4 v;
  asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_stream_f1(const float* p) {
  float v;
  asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream_f4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_stream_f1(float* p, float v) {
  asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum for blockDim.x <= 1024 (result valid in thread 0).
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* smem /* >= 32 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? smem[threadIdx.x] : T(0);
  if (warp == 0) v = warp_sum(v);
  __syncthreads();
  return v;
}
#endif  // __CUDACC__

}  // namespace dskd
