// Shared building blocks of the fp32-faithful tensor-core GEMMs ("tf32x3"):
//   a*b ~= hi(a)*hi(b) + lo(a)*hi(b) + hi(a)*lo(b),  hi = rn_tf32(a), lo = rn_tf32(a - hi),
// fp32 accumulation in TMEM.  Both parts are rounded to nearest (cvt.rna.tf32.f32), so the dropped
// lo*lo term and the rounding of lo are ~2^-22 with random sign (truncation would make them one-signed
// and ~8x larger on a Gram diagonal): fp32-class results at 1/3 of the tf32 tensor rate -- what the
// matching head needs for bit-exact ranks and the loss Grams need under exp(x / 0.1).
//
// Operand tiles are [128 rows x 32 k] fp32, K-major, 128-byte rows with the 128B swizzle: one tile is
// 16 KiB; a pipeline stage holds {A_hi, A_lo, B_hi, B_lo}.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace sga {
namespace tf32x3 {

__device__ __forceinline__ uint32_t rn_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

constexpr int kTileRows = 128;
constexpr int kTileK = 32;                 // fp32 elements per 128-byte row
constexpr uint32_t kTileBytes = 16384;
constexpr uint32_t kStageBytes = 4 * kTileBytes;   // A_hi | A_lo | B_hi | B_lo

// Row `t` (0..127) of a [128 x 32] operand tile: gather (optional index), scale (optional divisor),
// split and store swizzled.  Rows >= nrows and k >= K are zero-filled.
__device__ __forceinline__ void load_rows(unsigned char* hi, unsigned char* lo, const float* __restrict__ src, int64_t ld,
                                          const int32_t* __restrict__ idx, const float* __restrict__ div, int row0,
                                          int nrows, int k0, int K, int t, bool vec_ok) {
  const bool valid = t < nrows;
  int64_t srow = 0;
  float d = 1.f;
  if (valid) {
    srow = idx ? (int64_t)idx[row0 + t] : (int64_t)(row0 + t);
    if (div) d = div[srow];
  }
  const float* p = src + srow * ld + k0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (valid) {
      if (vec_ok && k0 + 4 * j + 3 < K) {
        float4 q = *reinterpret_cast<const float4*>(p + 4 * j);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (k0 + 4 * j + e < K) v[e] = p[4 * j + e];
      }
    }
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float x = div ? v[e] / d : v[e];
      h[e] = rn_tf32(x);
      l[e] = rn_tf32(x - __uint_as_float(h[e]));
    }
    const uint32_t off = ptx::sw128_offset(t, j);
    *reinterpret_cast<uint4*>(hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// One pipeline stage: D[128 x N] (+)= A[128 x 32] * B[N x 32]^T in three passes x four k-steps.
// `stage_addr` = shared-memory address of the stage; issued by the elected lane.
__device__ __forceinline__ void issue_stage(uint32_t d_tmem, uint32_t stage_addr, uint32_t idesc, bool first) {
  const uint64_t dAhi = ptx::smem_desc_sw128(stage_addr);
  const uint64_t dAlo = ptx::smem_desc_sw128(stage_addr + kTileBytes);
  const uint64_t dBhi = ptx::smem_desc_sw128(stage_addr + 2 * kTileBytes);
  const uint64_t dBlo = ptx::smem_desc_sw128(stage_addr + 3 * kTileBytes);
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
    const uint64_t a = (pass == 1) ? dAlo : dAhi;
    const uint64_t b = (pass == 2) ? dBlo : dBhi;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
      ptx::umma_tf32(d_tmem, a + (uint64_t)(ks * 2), b + (uint64_t)(ks * 2), idesc, (first && pass == 0 && ks == 0) ? 0u : 1u);
  }
}

// Same with independent A / B stage addresses ({hi, lo} tiles each): a Gram block on the diagonal passes the
// same tiles for both operands.
__device__ __forceinline__ void issue_stage_ab(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr, uint32_t idesc, bool first) {
  const uint64_t dAhi = ptx::smem_desc_sw128(a_addr);
  const uint64_t dAlo = ptx::smem_desc_sw128(a_addr + kTileBytes);
  const uint64_t dBhi = ptx::smem_desc_sw128(b_addr);
  const uint64_t dBlo = ptx::smem_desc_sw128(b_addr + kTileBytes);
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
    const uint64_t a = (pass == 1) ? dAlo : dAhi;
    const uint64_t b = (pass == 2) ? dBlo : dBhi;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
      ptx::umma_tf32(d_tmem, a + (uint64_t)(ks * 2), b + (uint64_t)(ks * 2), idesc, (first && pass == 0 && ks == 0) ? 0u : 1u);
  }
}

}  // namespace tf32x3
}  // namespace sga
