// Output helpers shared by the tcgen05 GEMM epilogues (gemm_tc.cu, gram_ts.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sga {
namespace {

// A thread's 32 consecutive output values -> its row (dst = address of the slab's first column).  The row pitch is a
// multiple of 4 floats but a segment may start at any column, so the 16-byte alignment phase is warp-uniform: head
// scalars up to the next aligned column, seven 128-bit stores, tail scalars.
__device__ __forceinline__ void st4(float* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  *reinterpret_cast<uint4*>(p) = make_uint4(a, b, c, d);
}
template <int PH>
__device__ __forceinline__ void store_row32_phase(float* dst, const uint32_t (&v)[32]) {
  constexpr int H = (4 - PH) & 3;      // head scalars
#pragma unroll
  for (int e = 0; e < H; ++e) dst[e] = __uint_as_float(v[e]);
#pragma unroll
  for (int i = 0; i < (PH == 0 ? 8 : 7); ++i) st4(dst + H + 4 * i, v[H + 4 * i], v[H + 4 * i + 1], v[H + 4 * i + 2], v[H + 4 * i + 3]);
  if (PH != 0) {
#pragma unroll
    for (int e = H + 28; e < 32; ++e) dst[e] = __uint_as_float(v[e]);
  }
}
__device__ __forceinline__ void store_row32(float* dst, const uint32_t (&v)[32], int nv) {
  if (nv >= 32) {
    switch ((int)((reinterpret_cast<uintptr_t>(dst) >> 2) & 3)) {
      case 0: store_row32_phase<0>(dst, v); break;
      case 1: store_row32_phase<1>(dst, v); break;
      case 2: store_row32_phase<2>(dst, v); break;
      default: store_row32_phase<3>(dst, v); break;
    }
  } else {
#pragma unroll
    for (int e = 0; e < 32; ++e)
      if (e < nv) dst[e] = __uint_as_float(v[e]);
  }
}

}  // namespace
}  // namespace sga
