// Shared helpers for libsga_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/sga_b200.h"

namespace sga {

void set_error(const char* fmt, ...);

#define SGA_REQUIRE(cond, ...)                      \
  do {                                              \
    if (!(cond)) {                                  \
      sga::set_error(__VA_ARGS__);                  \
      return SGA_EINVAL;                            \
    }                                               \
  } while (0)

#define SGA_CUDA(expr)                                                                  \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      sga::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return (int)_e;                                                                   \
    }                                                                                   \
  } while (0)

#define SGA_LAUNCH_CHECK() SGA_CUDA(cudaGetLastError())

int sm_count();
int persistent_ctas();              // sm_count() capped by sga_pointnet_set_max_ctas
void set_persistent_cta_cap(int n);

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <typename T>
__device__ __forceinline__ float load_as_float(const void* p, int64_t i, int is_f64) {
  return is_f64 ? (float)reinterpret_cast<const double*>(p)[i] : reinterpret_cast<const float*>(p)[i];
}

}  // namespace sga
